"""ctypes binding of include/opencloth.h (libopencloth_b200.so).

The shared library is built in-tree by ``__graft_entry__.build()`` (``make -C opencloth_b200/csrc``).
There is no CPU fallback: if the library is missing, loading fails loudly; if no CUDA device is
present, ``oc_create`` returns ``OC_ERR_NO_DEVICE`` and the Python layer raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OC_LIB") or os.path.join(_HERE, "libopencloth_b200.so")      # OC_LIB: development builds (tools/)

OC_OK = 0
OC_ERR_INVALID = -1
OC_ERR_NO_DEVICE = -2
OC_ERR_CUDA = -3
OC_ERR_NOMEM = -4
OC_ERR_UNSUPPORTED = -5

OC_KERNEL_AUTO = 0
OC_KERNEL_GATHER = 1
OC_KERNEL_MARCH = 2
OC_KERNEL_MARCH2 = 3
OC_KERNEL_RESIDENT = 4
OC_KERNEL_TWIN = 5
OC_KERNEL_STREAM = 6
OC_KERNEL_STREAM2 = 7
OC_KERNEL_BANDRES = 8

OC_BAND_ENDPOINT_BYTES = 512

OC_INTEGRATOR_VERLET = 0
OC_INTEGRATOR_EULER = 1
OC_INTEGRATOR_SEMI_IMPLICIT = 2


class OcParams(ctypes.Structure):
    """Mirror of ``oc_params`` (include/opencloth.h). Field order and types must match exactly."""
    _fields_ = [
        ("nx", ctypes.c_int), ("ny", ctypes.c_int),
        ("batch", ctypes.c_int),
        ("row_begin", ctypes.c_int), ("row_end", ctypes.c_int),
        ("halo_rows", ctypes.c_int),
        ("device", ctypes.c_int),
        ("fullsize", ctypes.c_float),
        ("substeps_per_launch", ctypes.c_int),
        ("exact", ctypes.c_int),
        ("kernel", ctypes.c_int),
        ("ks_struct", ctypes.c_float), ("kd_struct", ctypes.c_float),
        ("ks_shear", ctypes.c_float), ("kd_shear", ctypes.c_float),
        ("ks_bend", ctypes.c_float), ("kd_bend", ctypes.c_float),
        ("damping", ctypes.c_float),
        ("gravity", ctypes.c_float * 3),
        ("mass", ctypes.c_float),
        ("dt", ctypes.c_float),
        ("ellipsoid", ctypes.c_float * 16),
        ("inv_ellipsoid", ctypes.c_float * 16),
        ("center", ctypes.c_float * 3),
        ("radius", ctypes.c_float),
        ("integrator", ctypes.c_int),
        ("provot", ctypes.c_int),
    ]


# every symbol include/opencloth.h declares: name -> (restype, argtypes)
_P = ctypes.POINTER
SYMBOLS = {
    "oc_default_params": (ctypes.c_int, [_P(OcParams), ctypes.c_int, ctypes.c_int]),
    "oc_default_params_for": (ctypes.c_int, [_P(OcParams), ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "oc_create": (ctypes.c_int, [_P(ctypes.c_void_p), _P(OcParams)]),
    "oc_set_params": (ctypes.c_int, [ctypes.c_void_p, _P(OcParams)]),
    "oc_get_params": (ctypes.c_int, [ctypes.c_void_p, _P(OcParams)]),
    "oc_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "oc_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "oc_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "oc_download": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "oc_download_normals": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "oc_destroy": (None, [ctypes.c_void_p]),
    "oc_last_error": (ctypes.c_char_p, []),
    "oc_set_particle": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _P(ctypes.c_float)]),
    "oc_set_particles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _P(ctypes.c_int), _P(ctypes.c_int), _P(ctypes.c_float)]),
    "oc_upload_cloth": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "oc_download_cloth": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "oc_set_pins": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _P(ctypes.c_int), ctypes.c_int]),
    "oc_reset_pins": (ctypes.c_int, [ctypes.c_void_p]),
    "oc_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "oc_reset_stream": (ctypes.c_int, [ctypes.c_void_p]),
    "oc_band_endpoint": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "oc_band_link": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "oc_band_pull_halo": (ctypes.c_int, [ctypes.c_void_p]),
    "oc_band_unlink": (ctypes.c_int, [ctypes.c_void_p]),
    "oc_band_link_local": (ctypes.c_int, [_P(ctypes.c_void_p), ctypes.c_int]),
    "oc_launch_count": (ctypes.c_longlong, [ctypes.c_void_p]),
    "oc_step_timed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _P(ctypes.c_float)]),
    "oc_halo_send_region": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _P(ctypes.c_void_p), _P(ctypes.c_size_t)]),
    "oc_halo_recv_region": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _P(ctypes.c_void_p), _P(ctypes.c_size_t)]),
    "oc_halo_refreshed": (ctypes.c_int, [ctypes.c_void_p]),
    "oc_halo_budget": (ctypes.c_int, [ctypes.c_void_p]),
    "oc_step_split": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, _P(ctypes.c_int)]),
    "oc_halo_exchange": (ctypes.c_int, [_P(ctypes.c_void_p), ctypes.c_int]),
    "oc_debug_counters": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_ulonglong)]),
    "oc_debug_pipeline": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_float), ctypes.c_int]),
    "oc_debug_timeline": (ctypes.c_int, [ctypes.c_void_p, _P(ctypes.c_ulonglong), ctypes.c_size_t]),
    "oc_sizeof_params": (ctypes.c_size_t, []),
    "oc_selftest_math": (ctypes.c_int, [ctypes.c_ulonglong, ctypes.c_uint, _P(ctypes.c_ulonglong)]),
    "oc_spring_energy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _P(ctypes.c_double)]),
    "oc_version": (ctypes.c_char_p, []),
}

_lib = None


class OpenClothError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"opencloth_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load libopencloth_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C opencloth_b200/csrc`. opencloth_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.oc_sizeof_params() != ctypes.sizeof(OcParams):
        raise ImportError(f"oc_params mirror out of date: library {lib.oc_sizeof_params()} B, ctypes {ctypes.sizeof(OcParams)} B")
    _lib = lib
    return lib


def check(rc):
    if rc != OC_OK:
        msg = load().oc_last_error()
        raise OpenClothError(rc, msg.decode() if msg else "")
