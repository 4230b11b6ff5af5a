#!/usr/bin/env python
"""Small runs of every kernel for compute-sanitizer (memcheck / racecheck / initcheck); development tool.
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc  # noqa: E402

only = [int(t) for t in os.environ.get("OC_SAN_KERNELS", "").split(",") if t]      # e.g. OC_SAN_KERNELS=8: that kernel only
for kernel, k in ((1, 1), (2, 1), (2, 4), (3, 1), (4, 1), (5, 1), (6, 1), (7, 1), (8, 1)):
    if only and kernel not in only:
        continue
    for exact in (1, 0):
        for nx, ny, batch in ((150, 70, 1), (37, 23, 3), (260, 40, 1), (21, 21, 2)):
            c = oc.Cloth(nx, ny, batch=batch, kernel=kernel, exact=exact, substeps_per_launch=k)
            c.step(12)
            x, xl = c.download()
            assert np.isfinite(x).all()
            c.close()
            print("ok", kernel, k, exact, nx, ny, batch, flush=True)
if only:
    sys.exit(0)
# row bands in one process (halo exchange copies, band launches)
import ctypes  # noqa: E402
from opencloth_b200 import _abi  # noqa: E402
bands = [oc.Cloth(150, 64, row_begin=0, row_end=32, halo_rows=8, kernel=3), oc.Cloth(150, 64, row_begin=32, row_end=64, halo_rows=8, kernel=3)]
arr = (ctypes.c_void_p * 2)(*[b._h for b in bands])
for _ in range(3):
    _abi.check(_abi.load().oc_halo_exchange(arr, 2))
    for b in bands:
        b.step(4)
for b in bands:
    b.download()
print("bands ok")
# linked row bands on one device (peer stores into the neighbour's halo, flag words; odd bands launch bottom to top)
for kernel in (3, 5, 6):
    lb = [oc.Cloth(150, 96, row_begin=32 * b, row_end=32 * (b + 1), halo_rows=2, kernel=kernel) for b in range(3)]
    oc.link_bands_local(lb)
    for _ in range(6):
        for b in lb:
            b.step(1)
    for b in lb:
        b.download(); b.close()
    print("linked bands ok", kernel)
