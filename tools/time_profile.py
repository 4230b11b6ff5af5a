#!/usr/bin/env python
"""Step time and slow-path counters of one cloth over the course of a run (development tool).
usage: [OC_DEBUG=4] time_profile.py [--n 2048] [--exact 1] [--steps 2000] [--every 100]"""
import argparse
import ctypes
import os
import sys

os.environ.setdefault("OC_DEBUG", "4")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--exact", type=int, default=1)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--every", type=int, default=100)
a = ap.parse_args()
c = oc.Cloth(a.n, a.n, kernel=3, exact=a.exact)
out = (ctypes.c_ulonglong * 4)()
c._lib.oc_debug_counters(c._h, out)
warp_iters = a.n * a.n / 64
done = 0
while done < a.steps:
    ms = c.step_timed(a.every)
    done += a.every
    c._lib.oc_debug_counters(c._h, out)
    e = a.every
    print(f"steps {done:5d}: {ms / e * 1e3:7.1f} us/step  redo warps {out[1] / e / warp_iters * 100:6.3f} %  vel fallback lanes {out[2] / e:9.0f}  "
          f"collider warps {(out[3] >> 32) / e / warp_iters * 100:6.2f} %", flush=True)
