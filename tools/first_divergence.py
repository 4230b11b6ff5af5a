#!/usr/bin/env python
"""Development tool: step the march kernel and the gather kernel side by side on the GPU and
report the first step / particles at which they differ (both are bit-exact when correct)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=21); ap.add_argument("--ny", type=int, default=21)
ap.add_argument("--steps", type=int, default=200); ap.add_argument("--k", type=int, default=1)
a = ap.parse_args()
m = oc.Cloth(a.nx, a.ny, kernel=oc.OC_KERNEL_MARCH, substeps_per_launch=a.k)
g = oc.Cloth(a.nx, a.ny, kernel=oc.OC_KERNEL_GATHER)
for st in range(1, a.steps + 1):
    m.step(1); g.step(1)
    mx, ml = m.download(); gx, gl = g.download()
    bad = np.argwhere((mx.view(np.uint32) != gx.view(np.uint32)).any(1)).ravel()
    if len(bad):
        print(f"first divergence at step {st}: {len(bad)} particles, e.g.")
        for p in bad[:8]:
            print(f"  idx {p} (i={p % a.nx}, j={p // a.nx}) march {mx[p]} {[hex(v) for v in mx[p].view(np.uint32)]} gather {gx[p]} {[hex(v) for v in gx[p].view(np.uint32)]}")
        # resync march to gather and continue to see how often it happens
        m.upload(gx, gl)
        if st > 150: break
print("done")
