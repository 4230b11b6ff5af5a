#!/usr/bin/env python
"""Kernel sweep on one GPU: particle-updates/s for kernel x mode x k x grid (development tool).
Timing: CUDA events on the handle's stream (oc_step_timed), warm-up first, state larger than L2
for grids >= 2048^2."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc  # noqa: E402

PEAK = 6534.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def run(nx, ny, batch, kernel, exact, k, steps, warm=3, reps=3):
    c = oc.Cloth(nx, ny, batch=batch, kernel=kernel, exact=exact, substeps_per_launch=k)
    c.step(200)            # leave the perfectly flat start
    for _ in range(warm):
        c.step_timed(steps)
    best = min(c.step_timed(steps) for _ in range(reps))
    c.close()
    ups = nx * ny * batch * steps / (best * 1e-3)
    return best / steps * 1e3, ups


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="2048x2048x1")
    ap.add_argument("--ks", default="1,2,4")
    ap.add_argument("--modes", default="1,0")
    ap.add_argument("--kernels", default="2,1")
    ap.add_argument("--steps", type=int, default=64)
    a = ap.parse_args()
    print(oc.version())
    for g in a.grids.split(","):
        nx, ny, b = [int(t) for t in g.split("x")]
        for kern in [int(t) for t in a.kernels.split(",")]:
            for exact in [int(t) for t in a.modes.split(",")]:
                for k in ([int(t) for t in a.ks.split(",")] if kern == 2 else [1]):
                    us, ups = run(nx, ny, b, kern, exact, k, a.steps)
                    print(json.dumps(dict(grid=g, kernel={1: "gather", 2: "march", 3: "march2", 5: "twin", 6: "stream"}[kern], exact=exact, k=k,
                                          us_per_step=round(us, 2), gupdates_s=round(ups / 1e9, 3),
                                          roofline_frac_48B=round(ups * 48 / 1e9 / PEAK, 4),
                                          rs=os.environ.get("OC_MARCH_RS", "auto"))), flush=True)
