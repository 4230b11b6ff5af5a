#!/usr/bin/env python
"""Run a handful of launches of one kernel configuration (for ncu)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--ny", type=int, default=0)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--kernel", type=int, default=2)
ap.add_argument("--exact", type=int, default=1)
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--warm", type=int, default=40)
ap.add_argument("--launches", type=int, default=3)
a = ap.parse_args()
c = oc.Cloth(a.n, a.ny or a.n, batch=a.batch, kernel=a.kernel, exact=a.exact, substeps_per_launch=a.k)
c.step(a.warm * a.k)
c.sync()
c.step(a.launches * a.k)
c.sync()
print("done", c.launch_count)
