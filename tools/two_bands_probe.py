import sys, time, ctypes, os
sys.path.insert(0, ".")
import opencloth_b200 as oc
from opencloth_b200 import _abi
n = 8192
for halo in (24, 40, 48):
    bands = [oc.Cloth(n, n, row_begin=0, row_end=n // 2, halo_rows=halo, kernel=3), oc.Cloth(n, n, row_begin=n // 2, row_end=n, halo_rows=halo, kernel=3)]
    arr = (ctypes.c_void_p * 2)(*[b._h for b in bands])
    per = halo // 2
    lib = _abi.load()
    def group():
        _abi.check(lib.oc_halo_exchange(arr, 2))
        for b in bands:
            b.step(per)
    group()
    for b in bands: b.sync()
    t0 = time.perf_counter()
    G = 6
    for _ in range(G): group()
    for b in bands: b.sync()
    dt = time.perf_counter() - t0
    print("halo", halo, "per-step (both bands on one GPU) %.1f us" % (dt / (G * per) * 1e6), flush=True)
    # per-step timing inside one group for band 0
    _abi.check(lib.oc_halo_exchange(arr, 2))
    ts = [round(bands[0].step_timed(1) * 1e3) for _ in range(per)]
    bands[1].step(per)
    print("   band 0 single steps (us):", ts)
    for b in bands: b.close()
