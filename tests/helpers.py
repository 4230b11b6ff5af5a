"""Test helpers: ctypes wrappers of the CHECKERS (oracle restatement, verbatim reference build,
CPU kernel emulator).  Test infrastructure only — nothing here is imported by the product."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboc_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libocref.so")
EMU_SO = os.path.join(ROOT, "tests", "emu", "liboc_emu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_emulator():
    src = os.path.join(ROOT, "tests", "emu", "oc_emu.cu")
    csrc = os.path.join(ROOT, "opencloth_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("oc_core.cuh", "oc_host.h", "oc_gather.cuh", "oc_provot.cuh", "oc_normals.cuh", "oc_march.cuh", "oc_march2.cuh", "oc_twin.cuh", "oc_stream.cuh", "oc_stream2.cuh", "oc_resident.cuh")]
    if _newer(EMU_SO, *deps):
        return
    # six parts compiled in parallel (tests/emu/oc_emu.cu, EMU_PART): the kernel-body instantiations dominate the compile time
    objdir = os.path.join(ROOT, "tests", "emu", "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
             "-diag-suppress", "20011,20014"]
    procs, objs = [], []
    for part in range(6):
        obj = os.path.join(objdir, f"oc_emu_{part}.o")
        objs.append(obj)
        procs.append(subprocess.Popen([NVCC] + flags + [f"-DEMU_PART={part}", "-c", src, "-o", obj]))
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("building the CPU kernel emulator failed")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared"] + objs + ["-o", EMU_SO])



def ensure_built():
    env = dict(os.environ, CC="gcc", CXX="g++")
    if not _newer(ORACLE_SO, os.path.join(ROOT, "oracle", "oc_oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboc_oracle.so"], env=env)
    if not os.path.exists(REF_SO) and os.path.exists("/root/reference"):
        subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")], env=env)
    from opencloth_b200 import _abi
    if not os.path.exists(_abi.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "opencloth_b200", "csrc"), "-j8"], env=env)
    build_emulator()


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def bitwise_equal(a, b):
    return a.shape == b.shape and bool((bits(a) == bits(b)).all())


# ---------------------------------------------------------------------------------------------
# oracle restatement (oracle/oc_oracle.c)
# ---------------------------------------------------------------------------------------------
class OcoParams(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("fullsize", ctypes.c_float),
                ("ks_struct", ctypes.c_float), ("kd_struct", ctypes.c_float),
                ("ks_shear", ctypes.c_float), ("kd_shear", ctypes.c_float),
                ("ks_bend", ctypes.c_float), ("kd_bend", ctypes.c_float),
                ("damping", ctypes.c_float), ("gravity", ctypes.c_float * 3),
                ("mass", ctypes.c_float), ("dt", ctypes.c_float),
                ("ellipsoid", ctypes.c_float * 16), ("inv_ellipsoid", ctypes.c_float * 16),
                ("center", ctypes.c_float * 3), ("radius", ctypes.c_float),
                ("integrator", ctypes.c_int), ("provot", ctypes.c_int)]


VERLET, EULER, SEMI = 0, 1, 2
_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        L = ctypes.CDLL(ORACLE_SO)
        L.oco_create.restype = ctypes.c_void_p
        L.oco_create.argtypes = [ctypes.POINTER(OcoParams)]
        L.oco_default_params.argtypes = [ctypes.POINTER(OcoParams), ctypes.c_int, ctypes.c_int]
        L.oco_default_params_for.argtypes = [ctypes.POINTER(OcoParams), ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oco_provot.argtypes = [ctypes.c_void_p]
        L.oco_set_pins.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        L.oco_destroy.argtypes = [ctypes.c_void_p]
        L.oco_set_params.argtypes = [ctypes.c_void_p, ctypes.POINTER(OcoParams)]
        L.oco_step.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oco_step_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.oco_get_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oco_set_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oco_get_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.oco_set_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.oco_spring_energy.restype = ctypes.c_double
        L.oco_spring_energy.argtypes = [ctypes.c_void_p]
        L.oco_get_tables.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 6
        _oracle = L
    return _oracle


class Oracle:
    """CPU restatement of StepPhysics (gather order). The checker."""

    def __init__(self, nx, ny, integrator=VERLET, **overrides):
        """integrator: VERLET (the reference's Verlet demo defaults), EULER / SEMI (its sibling demos' own defaults:
        spring constants, mass 0.5, Provot pass on).  For the Euler variants `state()[1]` is V, not X_last."""
        L = oracle_lib()
        self.L = L
        self.p = OcoParams()
        L.oco_default_params_for(ctypes.byref(self.p), nx, ny, integrator)
        self._apply(overrides)
        self.h = ctypes.c_void_p(L.oco_create(ctypes.byref(self.p)))
        assert self.h, "oco_create failed"
        self.nx, self.ny, self.n = nx, ny, nx * ny

    def _apply(self, overrides):
        for k, v in overrides.items():
            cur = getattr(self.p, k)
            if hasattr(cur, "__len__"):
                for i, x in enumerate(v):
                    cur[i] = x
            else:
                setattr(self.p, k, v)

    def set_params(self, **overrides):
        self._apply(overrides)
        assert self.L.oco_set_params(self.h, ctypes.byref(self.p)) == 0

    def step(self, n=1):
        self.L.oco_step(self.h, n)

    def step_rows(self, j0, j1):
        self.L.oco_step_rows(self.h, j0, j1)

    def state(self):
        x = np.empty((self.n, 3), np.float32)
        xl = np.empty((self.n, 3), np.float32)
        self.L.oco_get_state(self.h, vp(x), vp(xl))
        return x, xl

    def set_state(self, x, xl):
        x = np.ascontiguousarray(x, np.float32)
        xl = np.ascontiguousarray(xl, np.float32)
        self.L.oco_set_state(self.h, vp(x), vp(xl))

    def get_rows(self, j0, j1):
        x = np.empty(((j1 - j0) * self.nx, 3), np.float32)
        xl = np.empty_like(x)
        self.L.oco_get_rows(self.h, j0, j1, vp(x), vp(xl))
        return x, xl

    def set_rows(self, j0, j1, x, xl):
        self.L.oco_set_rows(self.h, j0, j1, vp(np.ascontiguousarray(x, np.float32)), vp(np.ascontiguousarray(xl, np.float32)))

    def energy(self):
        return self.L.oco_spring_energy(self.h)

    def provot(self):
        self.L.oco_provot(self.h)

    def set_pins(self, indices):
        """None: the reference's literals (0 and numX)."""
        if indices is None:
            self.L.oco_set_pins(self.h, None, -1)
        else:
            idx = [int(i) for i in indices]
            self.L.oco_set_pins(self.h, (ctypes.c_int * max(1, len(idx)))(*idx), len(idx))

    def tables(self):
        t = [np.empty(self.nx, np.float32) for _ in range(2)] + [np.empty(self.ny, np.float32) for _ in range(2)] + \
            [np.empty(self.nx, np.float32), np.empty(self.ny, np.float32)]
        self.L.oco_get_tables(self.h, *[vp(a) for a in t])
        return dict(rh1=t[0], rh2=t[1], rv1=t[2], rv2=t[3], dx2=t[4], dz2=t[5])

    def close(self):
        if self.h:
            self.L.oco_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# verbatim reference build (oracle/_ref/libocref.so) — single global simulation, like the reference
# ---------------------------------------------------------------------------------------------
def have_ref():
    return os.path.exists(REF_SO)


_ref = None
_ref_variants = {}
REF_VARIANT_SO = {EULER: os.path.join(ROOT, "oracle", "_ref", "libocref_euler.so"), SEMI: os.path.join(ROOT, "oracle", "_ref", "libocref_semi.so")}


def have_ref_variant(integrator):
    return os.path.exists(REF_VARIANT_SO[integrator])


class RefVariant:
    """Verbatim build of a sibling demo (explicit Euler / semi-implicit Euler): one global simulation; state = (X, V)."""

    def __init__(self, integrator, nx, ny, provot=1):
        if integrator not in _ref_variants:
            L = ctypes.CDLL(REF_VARIANT_SO[integrator])
            L.ref_num_particles.restype = ctypes.c_size_t
            L.ref_get_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            L.ref_set_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            L.ref_get_params.argtypes = [ctypes.c_void_p]
            _ref_variants[integrator] = L
        self.L = _ref_variants[integrator]
        assert self.L.ref_init(nx, ny) == 0
        self.L.ref_set_provot(provot)
        self.n = self.L.ref_num_particles()

    def step(self, n=1):
        self.L.ref_step(n)

    def state(self):
        x = np.empty((self.n, 3), np.float32)
        v = np.empty((self.n, 3), np.float32)
        self.L.ref_get_state(vp(x), vp(v))
        return x, v

    def set_state(self, x, v):
        self.L.ref_set_state(vp(np.ascontiguousarray(x, np.float32)), vp(np.ascontiguousarray(v, np.float32)))

    def params(self):
        p = np.empty(16, np.float32)
        self.L.ref_get_params(vp(p))
        return p


def ref_lib():
    global _ref
    if _ref is None:
        L = ctypes.CDLL(REF_SO)
        L.ref_num_particles.restype = ctypes.c_size_t
        L.ref_num_springs.restype = ctypes.c_size_t
        L.ref_spring_energy.restype = ctypes.c_double
        L.ref_get_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_set_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_get_springs.argtypes = [ctypes.c_void_p] * 6
        L.ref_get_ellipsoid.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_get_params.argtypes = [ctypes.c_void_p]
        _ref = L
    return _ref


class Ref:
    def __init__(self, nx, ny):
        self.L = ref_lib()
        assert self.L.ref_init(nx, ny) == 0
        self.n = self.L.ref_num_particles()
        self.nx, self.ny = nx, ny

    def step(self, n=1):
        self.L.ref_step(n)

    def step_provot(self, n=1):
        """StepPhysics with the ApplyProvotDynamicInverse call of V:561 enabled."""
        self.L.ref_step_provot(n)

    def state(self):
        x = np.empty((self.n, 3), np.float32)
        xl = np.empty((self.n, 3), np.float32)
        self.L.ref_get_state(vp(x), vp(xl))
        return x, xl

    def set_state(self, x, xl):
        self.L.ref_set_state(vp(np.ascontiguousarray(x, np.float32)), vp(np.ascontiguousarray(xl, np.float32)))

    def energy(self):
        return self.L.ref_spring_energy()

    def springs(self):
        s = self.L.ref_num_springs()
        p1 = np.empty(s, np.int32); p2 = np.empty(s, np.int32); rest = np.empty(s, np.float32)
        ks = np.empty(s, np.float32); kd = np.empty(s, np.float32); ty = np.empty(s, np.int32)
        self.L.ref_get_springs(vp(p1), vp(p2), vp(rest), vp(ks), vp(kd), vp(ty))
        return dict(p1=p1, p2=p2, rest=rest, ks=ks, kd=kd, type=ty)

    def ellipsoid(self):
        m = np.empty(16, np.float32); mi = np.empty(16, np.float32)
        self.L.ref_get_ellipsoid(vp(m), vp(mi))
        return m, mi

    def params(self):
        p = np.empty(16, np.float32)
        self.L.ref_get_params(vp(p))
        return p


# ---------------------------------------------------------------------------------------------
# CPU emulation of the product's kernel bodies (tests/emu/oc_emu.cu)
# ---------------------------------------------------------------------------------------------
_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        from opencloth_b200._abi import OcParams
        L = ctypes.CDLL(EMU_SO)
        L.emu_create.restype = ctypes.c_void_p
        L.emu_create.argtypes = [ctypes.POINTER(OcParams)]
        L.emu_destroy.argtypes = [ctypes.c_void_p]
        L.emu_set_params.argtypes = [ctypes.c_void_p, ctypes.POINTER(OcParams)]
        L.emu_upload.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.emu_download.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.emu_step.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6
        L.emu_halo_copy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.emu_halo_refreshed.argtypes = [ctypes.c_void_p]
        L.emu_halo_budget.argtypes = [ctypes.c_void_p]
        L.emu_halo_region.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
        L.emu_set_order.argtypes = [ctypes.c_int]
        L.emu_set_provot_threads.argtypes = [ctypes.c_int]
        L.emu_normals.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.emu_set_pins.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        L.emu_band_link.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        L.emu_bounding_sphere.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.emu_check_tiling.argtypes = [ctypes.c_int] * 10
        L.emu_bandres_plan.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int)] * 2 + [ctypes.POINTER(ctypes.c_ulonglong)]
        L.emu_bandres_rows.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int)] * 2
        _emu = L
    return _emu


class Emu:
    """The product's kernels run on the CPU (fibers as CUDA threads). Same host sequencing as the C-ABI."""

    def __init__(self, nx, ny, **overrides):
        from opencloth_b200._abi import OcParams, load
        self.L = emu_lib()
        p = OcParams()
        load().oc_default_params(ctypes.byref(p), nx, ny)       # host-only call, no GPU needed
        for k, v in overrides.items():
            cur = getattr(p, k)
            if hasattr(cur, "__len__"):
                for i, x in enumerate(v):
                    cur[i] = x
            else:
                setattr(p, k, v)
        self.p = p
        self.h = ctypes.c_void_p(self.L.emu_create(ctypes.byref(p)))
        assert self.h, "emu_create failed"
        rb, re = p.row_begin, p.row_end
        if rb == 0 and re == 0:
            re = ny
        self.nx, self.ny, self.rows = nx, ny, re - rb
        self.n_local = p.batch * self.rows * nx

    def step(self, n, kernel=2, exact=1, k=1, TW=32, RS=0):
        rc = self.L.emu_step(self.h, n, kernel, exact, k, TW, RS)
        assert rc == 0, f"emu_step rc={rc} (-1 barrier mismatch, -2 unsupported variant, -3 halo exhausted, -4 last segment under 2 rows)"

    def upload(self, x, xl):
        self.L.emu_upload(self.h, vp(np.ascontiguousarray(x, np.float32)), vp(np.ascontiguousarray(xl, np.float32)))

    def download(self):
        x = np.empty((self.n_local, 3), np.float32)
        xl = np.empty((self.n_local, 3), np.float32)
        self.L.emu_download(self.h, vp(x), vp(xl))
        return x, xl

    def normals(self):
        n = np.empty((self.p.batch * self.ny * self.nx, 3), np.float32)
        assert self.L.emu_normals(self.h, vp(n)) == 0
        return n

    def set_pins(self, indices, cloth=-1):
        idx = [int(i) for i in indices]
        assert self.L.emu_set_pins(self.h, cloth, (ctypes.c_int * max(1, len(idx)))(*idx), len(idx)) == 0

    def bounding_sphere(self):
        out = np.zeros(4, np.float32)
        self.L.emu_bounding_sphere(self.h, vp(out))
        return out

    def halo_region(self, side, which, send):
        ptr = ctypes.c_void_p(); cnt = ctypes.c_size_t()
        assert self.L.emu_halo_region(self.h, side, which, 1 if send else 0, ctypes.byref(ptr), ctypes.byref(cnt)) == 0
        return (ptr.value or 0), cnt.value

    def halo_refreshed(self):
        self.L.emu_halo_refreshed(self.h)

    @property
    def halo_budget(self):
        return self.L.emu_halo_budget(self.h)

    def close(self):
        if self.h:
            self.L.emu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def emu_link_bands(bands):
    """oc_band_link_local for emulated bands (rows in order)."""
    arr = (ctypes.c_void_p * len(bands))(*[b.h for b in bands])
    rc = emu_lib().emu_band_link(arr, len(bands))
    assert rc == 0, f"emu_band_link rc={rc}"


def developed_state(nx, ny, steps):
    """(X, X_last) of the default cloth after `steps` oracle steps: a non-trivial start state."""
    o = Oracle(nx, ny)
    o.step(steps)
    s = o.state()
    o.close()
    return s


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


REF_NORMALS_SO = os.path.join(ROOT, "oracle", "_ref", "libocref_normals.so")


def verbatim_normals(x, nx, ny, calls=1):
    """The reference's own UpdateNormals (oracle/_ref/libocref_normals.so)."""
    L = ctypes.CDLL(REF_NORMALS_SO)
    L.ref_normals.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    n = np.empty((nx * ny, 3), np.float32)
    assert L.ref_normals(nx, ny, vp(np.ascontiguousarray(x, np.float32)), vp(n), calls) == 0
    return n


def reference_normals(x, nx, ny):
    """UpdateNormals of the reference's lit demo (OpenCloth_ExplicitEuler_TextureMapped_Lit/.../main.cpp:684-707) on its
    triangle list (:313-327), restated literally in float32 numpy scalars: scatter loop over the triangles in list
    order, accumulators starting at zero (the reference's first call)."""
    f = np.float32
    X = np.ascontiguousarray(x, np.float32).reshape(-1, 3)
    idx = []
    numX, numY = nx - 1, ny - 1
    for i in range(numY):
        for j in range(numX):
            i0 = i * (numX + 1) + j; i1 = i0 + 1; i2 = i0 + (numX + 1); i3 = i2 + 1
            if (j + i) % 2:
                idx += [i0, i2, i1, i1, i2, i3]
            else:
                idx += [i0, i2, i3, i0, i3, i1]
    n = np.zeros_like(X)
    three = f(3.0)
    for t in range(0, len(idx), 3):
        p1, p2, p3 = X[idx[t]], X[idx[t + 1]], X[idx[t + 2]]
        a = p2 - p1; b = p3 - p1
        c = np.array([f(f(a[1] * b[2]) - f(b[1] * a[2])), f(f(a[2] * b[0]) - f(b[2] * a[0])), f(f(a[0] * b[1]) - f(b[0] * a[1]))], np.float32)
        for v in idx[t:t + 3]:
            n[v] = n[v] + c / three
    with np.errstate(all="ignore"):
        for v in range(len(n)):
            d = f(f(f(n[v][0] * n[v][0]) + f(n[v][1] * n[v][1])) + f(n[v][2] * n[v][2]))
            n[v] = n[v] * (f(1.0) / np.sqrt(d, dtype=np.float32))
    return n
