"""CPU tier: the C-ABI library loads, exports every symbol include/opencloth.h declares, its
parameter struct matches the ctypes mirror, host-only entry points work, and compute entry points
fail loudly (never fall back) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

import helpers
from opencloth_b200 import _abi


def declared_symbols():
    hdr = open(os.path.join(helpers.ROOT, "include", "opencloth.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(oc_[a-z_]+)\s*\(", hdr)))


def test_header_symbols_all_exported_and_bound():
    names = declared_symbols()
    assert "oc_create" in names and "oc_step" in names and "oc_download" in names and "oc_set_params" in names
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/opencloth.h but not exported"
    assert sorted(_abi.SYMBOLS) == names, "ctypes binding table and header disagree"


def test_params_struct_mirror():
    lib = _abi.load()
    assert lib.oc_sizeof_params() == ctypes.sizeof(_abi.OcParams)
    p = _abi.OcParams()
    assert lib.oc_default_params(ctypes.byref(p), 21, 21) == 0
    assert (p.nx, p.ny, p.batch, p.exact) == (21, 21, 1, 1)
    assert p.dt == ctypes.c_float(1 / 60.0).value and p.radius == 1.0 and p.ks_bend == ctypes.c_float(50.95).value
    assert p.ellipsoid[13] == 2.0 and p.inv_ellipsoid[15] == 1.0


def test_version_string():
    v = _abi.load().oc_version().decode()
    assert "opencloth_b200" in v and "sm_100a" in v


def test_argument_validation_without_device():
    """Bad arguments are rejected with OC_ERR_INVALID before any device work."""
    lib = _abi.load()
    h = ctypes.c_void_p()
    p = _abi.OcParams()
    lib.oc_default_params(ctypes.byref(p), 2, 21)          # bend springs need >= 3
    assert lib.oc_create(ctypes.byref(h), ctypes.byref(p)) == _abi.OC_ERR_INVALID
    assert b">= 3" in lib.oc_last_error()
    lib.oc_default_params(ctypes.byref(p), 21, 21)
    p.row_begin, p.row_end, p.halo_rows = 4, 12, 3         # odd halo
    assert lib.oc_create(ctypes.byref(h), ctypes.byref(p)) == _abi.OC_ERR_INVALID
    assert lib.oc_step(None, 1) == _abi.OC_ERR_INVALID


@pytest.mark.skipif(b"devices=0" not in _abi.load().oc_version(), reason="a CUDA device is present")
def test_no_cpu_fallback():
    """Without a GPU the product refuses to run: OC_ERR_NO_DEVICE, and the Python layer raises."""
    import opencloth_b200 as oc
    with pytest.raises(oc.OpenClothError) as e:
        oc.Cloth(21, 21)
    assert e.value.code == _abi.OC_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under opencloth_b200/ or include/ may mention it."""
    bad = []
    for base in ("opencloth_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(helpers.ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"liboc_oracle|libocref|oc_oracle\.c|oracle/", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
