#!/usr/bin/env python
"""CTA time line of one oc_k_march2 launch (development tool; needs OC_DEBUG=8 in the environment).
Prints when CTAs start / finish set-up / finish relative to the first CTA entry, per-SM occupancy of the
launch and the idle tail, so that launch overhead, ramp and imbalance can be told apart."""
import argparse
import ctypes
import os
import sys

import numpy as np

os.environ.setdefault("OC_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=2048)
ap.add_argument("--ny", type=int, default=2048)
ap.add_argument("--exact", type=int, default=1)
ap.add_argument("--warm", type=int, default=300)
a = ap.parse_args()
c = oc.Cloth(a.nx, a.ny, kernel=3, exact=a.exact)
c.step(a.warm)
ms = c.step_timed(50) / 50
c.step(3)
lib = c._lib
buf = (ctypes.c_ulonglong * (8 * 4096))()
rc = lib.oc_debug_timeline(c._h, buf, 8 * 4096)
assert rc == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(4096, 8).astype(np.int64)
t = t[t[:, 0] > 0]
n = len(t)
t0 = t[:, 0].min()
rel = (t[:, :5] - t0) / 1000.0          # us
names = ["entry", "setup", "lead-in", "steady", "exit"]
print(f"grid {a.nx}x{a.ny} exact={a.exact}: {n} CTAs, {ms * 1e3:.1f} us/step (timed), span entry->last exit {rel[:, 4].max():.1f} us")
for k, nm in enumerate(names):
    v = rel[:, k]
    print(f"  {nm:8s} min {v.min():7.1f}  p10 {np.percentile(v, 10):7.1f}  median {np.median(v):7.1f}  p90 {np.percentile(v, 90):7.1f}  max {v.max():7.1f}")
d = rel[:, 1:] - rel[:, :-1]
for k, nm in enumerate(["set-up", "lead-in", "steady", "lead-out"]):
    print(f"  phase {nm:8s} median {np.median(d[:, k]):7.1f} us   max {d[:, k].max():7.1f}")
sm = t[:, 5]
per = np.bincount(sm, minlength=148)
print("  CTAs per SM: ", dict(zip(*np.unique(per, return_counts=True))))
for cnt in np.unique(per):
    sel = np.isin(sm, np.where(per == cnt)[0])
    print(f"    SMs with {cnt} CTAs: last exit median {np.median(rel[sel, 4]):.1f} us, max {rel[sel, 4].max():.1f}")
# map of steady-loop durations (us): rows = row segment (blockIdx.y), columns = strip (blockIdx.x)
wout = 124 if a.nx > 128 else a.nx
nstrips = (a.nx + wout - 1) // wout
if n % nstrips == 0:
    dur = rel[:, 4].reshape(n // nstrips, nstrips)
    print("  exit time map (us after the first CTA entry), one line per row segment:")
    for r in range(dur.shape[0]):
        print("   ", " ".join(f"{int(v):3d}" for v in dur[r]), f"  | SMs {' '.join(str(int(x)) for x in sm[r * nstrips:(r + 1) * nstrips][:6])} ...")
