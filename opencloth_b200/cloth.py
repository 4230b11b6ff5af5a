"""Host-side mirror of the reference's cloth step interface, over the C-ABI.

The reference (mmmovania/opencloth, OpenCloth_Verlet/OpenCloth_Verlet/main.cpp, "V:") has no class:
``InitGL`` (V:232-328) builds the state, ``StepPhysics(dt)`` (V:557-562) advances it, the render
loop reads the global ``X`` (V:78).  Its own GPU back ends use Init / Upload / Verlet / ReadBuffer /
Shutdown (…/OpenCloth_Verlet_CUDA/verlet.cu:16-100, verlet_cl.cpp:57-222).  ``Cloth`` keeps those
names' meaning:

    cloth = Cloth(nx=21, ny=21)        # InitGL + InitCUDA + UploadCUDA: flat sheet, springs, collider
    cloth.step(1000)                   # 1000 x StepPhysics(timeStep)
    X, X_last = cloth.download()       # ReadBuffer

Everything runs in libopencloth_b200.so (CUDA, sm_100a).  numpy is only used for host buffers.
"""
import ctypes

import numpy as np

from . import _abi
from ._abi import OcParams, check


def default_params(nx=21, ny=21, integrator=0, **overrides):
    """``oc_params`` filled with the reference's values (V:59-62, V:97-104, V:123-130, V:324-327); with
    ``integrator`` = OC_INTEGRATOR_EULER / OC_INTEGRATOR_SEMI_IMPLICIT the values of that sibling demo."""
    p = OcParams()
    check(_abi.load().oc_default_params_for(ctypes.byref(p), nx, ny, integrator))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(f"oc_params has no field {k!r}")
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(p, k, v)
    return p


class Cloth:
    """One simulation (or a batch of independent ones, or one row band of a large one)."""

    def __init__(self, nx=21, ny=21, params=None, **overrides):
        self._lib = _abi.load()
        self._h = ctypes.c_void_p()
        self.params = params if params is not None else default_params(nx, ny, **overrides)
        h = ctypes.c_void_p()
        check(self._lib.oc_create(ctypes.byref(h), ctypes.byref(self.params)))
        self._h = h
        q = OcParams()
        check(self._lib.oc_get_params(self._h, ctypes.byref(q)))
        self.params = q
        self.nx, self.ny, self.batch = q.nx, q.ny, q.batch
        self.rows = q.row_end - q.row_begin
        self.n_local = self.batch * self.rows * self.nx     # particles held by this handle

    # ---- reference surface -------------------------------------------------------------------
    def step(self, n=1):
        """n x StepPhysics(dt) (V:557-562). Asynchronous."""
        check(self._lib.oc_step(self._h, int(n)))

    def step_split(self, n, exchange_stream_ptr):
        """oc_step_split: the last substep's boundary rows first; returns True if `exchange_stream` now waits for them."""
        did = ctypes.c_int(0)
        check(self._lib.oc_step_split(self._h, int(n), ctypes.c_void_p(exchange_stream_ptr), ctypes.byref(did)))
        return bool(did.value)

    def step_timed(self, n=1):
        """n substeps timed with CUDA events on the handle's stream; returns milliseconds."""
        ms = ctypes.c_float()
        check(self._lib.oc_step_timed(self._h, int(n), ctypes.byref(ms)))
        return ms.value

    def sync(self):
        check(self._lib.oc_sync(self._h))

    def download(self, stride=3, out=None):
        """Returns (X, X_last) as float32 arrays of shape (n_local, stride)."""
        if out is None:
            x = np.empty((self.n_local, stride), np.float32)
            xl = np.empty((self.n_local, stride), np.float32)
        else:
            x, xl = out
        check(self._lib.oc_download(self._h, x.ctypes.data_as(ctypes.c_void_p),
                                    xl.ctypes.data_as(ctypes.c_void_p), stride))
        return x, xl

    def download_normals(self, stride=3):
        """Per-vertex normals as the reference's lit demo computes them (UpdateNormals), shape (batch*ny*nx, stride)."""
        n = np.empty((self.batch * self.ny * self.nx, stride), np.float32)
        check(self._lib.oc_download_normals(self._h, n.ctypes.data_as(ctypes.c_void_p), stride))
        return n

    def download_into(self, x_ptr, xl_ptr, stride=3):
        """Download into raw host pointers (ints), e.g. pinned torch tensors' data_ptr()."""
        check(self._lib.oc_download(self._h, ctypes.c_void_p(x_ptr), ctypes.c_void_p(xl_ptr) if xl_ptr else None, stride))

    def upload(self, x, x_last):
        x = np.ascontiguousarray(x, np.float32)
        xl = np.ascontiguousarray(x_last, np.float32)
        stride = x.shape[-1]
        if x.size != self.n_local * stride or xl.size != x.size:
            raise ValueError(f"expected {self.n_local} x {stride} floats")
        check(self._lib.oc_upload(self._h, x.ctypes.data_as(ctypes.c_void_p), xl.ctypes.data_as(ctypes.c_void_p), stride))
        # (pageable arrays: cudaMemcpyAsync has staged them before returning; nothing to wait for here)

    def upload_from(self, x_ptr, xl_ptr, stride=3):
        check(self._lib.oc_upload(self._h, ctypes.c_void_p(x_ptr), ctypes.c_void_p(xl_ptr), stride))

    def set_params(self, **overrides):
        """Edit run-time scalars (V:97-104, V:123-130): spring constants, damping, gravity, dt, collider,
        substeps_per_launch, exact, kernel."""
        p = self.params
        for k, v in overrides.items():
            cur = getattr(p, k)
            if hasattr(cur, "__len__"):
                for i, x in enumerate(v):
                    cur[i] = x
            else:
                setattr(p, k, v)
        check(self._lib.oc_set_params(self._h, ctypes.byref(p)))

    def set_particle(self, idx, xyz, cloth=0):
        """Mouse-drag write-back of the reference (V:203-208): X[idx] = X_last[idx] = xyz."""
        v = (ctypes.c_float * 3)(*[float(t) for t in xyz])
        check(self._lib.oc_set_particle(self._h, int(cloth), int(idx), v))

    def set_particles(self, cloths, indices, xyz):
        """Many write-backs in one launch (per-environment actions of a batch)."""
        n = len(indices)
        a = (ctypes.c_int * max(1, n))(*[int(t) for t in cloths])
        b = (ctypes.c_int * max(1, n))(*[int(t) for t in indices])
        v = np.ascontiguousarray(xyz, np.float32).reshape(-1)
        assert v.size == 3 * n
        check(self._lib.oc_set_particles(self._h, n, a, b, v.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))

    def upload_cloth(self, cloth, x, x_last):
        x = np.ascontiguousarray(x, np.float32); xl = np.ascontiguousarray(x_last, np.float32)
        check(self._lib.oc_upload_cloth(self._h, int(cloth), x.ctypes.data_as(ctypes.c_void_p), xl.ctypes.data_as(ctypes.c_void_p), x.shape[-1]))
        self.sync()          # x, xl are pageable temporaries

    def download_cloth(self, cloth, stride=3):
        n1 = self.rows * self.nx
        x = np.empty((n1, stride), np.float32); xl = np.empty((n1, stride), np.float32)
        check(self._lib.oc_download_cloth(self._h, int(cloth), x.ctypes.data_as(ctypes.c_void_p), xl.ctypes.data_as(ctypes.c_void_p), stride))
        return x, xl

    def set_pins(self, indices, cloth=-1):
        """Replace the pinned set (the reference's literals 0 and numX, V:455, V:479-482) of one cloth (-1: all)."""
        idx = [int(i) for i in indices]
        arr = (ctypes.c_int * max(1, len(idx)))(*idx)
        check(self._lib.oc_set_pins(self._h, int(cloth), arr, len(idx)))

    def reset_pins(self):
        check(self._lib.oc_reset_pins(self._h))

    def spring_energy(self, cloth=0):
        e = ctypes.c_double()
        check(self._lib.oc_spring_energy(self._h, int(cloth), ctypes.byref(e)))
        return e.value

    # ---- plumbing ----------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        """Launch on this cudaStream_t (0 / None = CUDA's legacy default stream, e.g. torch's default stream)."""
        check(self._lib.oc_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def reset_stream(self):
        """Back to the handle's own non-blocking stream."""
        check(self._lib.oc_reset_stream(self._h))

    # ---- linked row bands (include/opencloth.h, "linked row bands") -----------------------------
    def band_endpoint(self):
        blob = ctypes.create_string_buffer(_abi.OC_BAND_ENDPOINT_BYTES)
        check(self._lib.oc_band_endpoint(self._h, blob, len(blob)))
        return blob.raw

    def band_link(self, upper, lower):
        check(self._lib.oc_band_link(self._h, upper, lower))      # bytes or None

    def band_pull_halo(self):
        check(self._lib.oc_band_pull_halo(self._h))

    def band_unlink(self):
        check(self._lib.oc_band_unlink(self._h))

    @property
    def launch_count(self):
        return int(self._lib.oc_launch_count(self._h))

    def halo_region(self, side, which, send):
        ptr = ctypes.c_void_p()
        cnt = ctypes.c_size_t()
        fn = self._lib.oc_halo_send_region if send else self._lib.oc_halo_recv_region
        check(fn(self._h, side, which, ctypes.byref(ptr), ctypes.byref(cnt)))
        return (ptr.value or 0), cnt.value

    def halo_refreshed(self):
        check(self._lib.oc_halo_refreshed(self._h))

    @property
    def halo_budget(self):
        return int(self._lib.oc_halo_budget(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.oc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def version():
    return _abi.load().oc_version().decode()


def link_bands_local(cloths):
    """Link band handles that live in this process (rows in order): oc_band_link_local."""
    arr = (ctypes.c_void_p * len(cloths))(*[c._h for c in cloths])
    check(_abi.load().oc_band_link_local(arr, len(cloths)))
