#!/usr/bin/env python
"""Rate of ONE row band of a big cloth on one GPU, without any exchange (development tool): what a rank of the
N-GPU run can reach at best.  The halo is declared fresh without being exchanged, so the band's edge rows drift away
from the real cloth: only the first ~300 steps are meaningful (later the stale edges stretch, operands leave the
exact range and the few CTAs at the band edge run the IEEE fallback every iteration, ~3x slower).  usage: band_rate.py [--n 8192] [--bands 8] [--halo 16] [--exact 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--bands", type=int, default=8)
ap.add_argument("--halo", type=int, default=16)
ap.add_argument("--exact", type=int, default=1)
a = ap.parse_args()
rows = a.n // a.bands
b = a.bands // 2
c = oc.Cloth(a.n, a.n, row_begin=b * rows, row_end=(b + 1) * rows, halo_rows=a.halo, exact=a.exact, kernel=3)
per = a.halo // 2
for _ in range(5):
    c.halo_refreshed(); c.step(per)
ms = 0.0
groups = max(4, min(40, 300 // per - 5))
for _ in range(groups):
    c.halo_refreshed()
    ms += c.step_timed(per)
owned = a.n * rows * per * groups
print(f"band {rows} rows of {a.n}x{a.n}, halo {a.halo}, exact {a.exact}: {ms / (groups * per) * 1e3:.1f} us/step, "
      f"{owned / (ms * 1e-3) / 1e9:.2f} G owned updates/s  (x{a.bands} = {a.bands * owned / (ms * 1e-3) / 1e9:.1f})")
