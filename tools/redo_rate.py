#!/usr/bin/env python
"""Development: how often does the exact-mode spring phase fall back to the IEEE intrinsics?"""
import ctypes, os, sys
os.environ["OC_DEBUG"] = "4"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc
from opencloth_b200 import _abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
c = oc.Cloth(n, n, exact=1)
out = (ctypes.c_ulonglong * 4)()
done = 0
for chunk in (20, 80, 100, 300, 500, 1000, 1000):
    ms = c.step_timed(chunk); done += chunk
    _abi.check(_abi.load().oc_debug_counters(c._h, out))
    lanes, warps, vel = out[0], out[1], out[2]
    tot_warps = n * n / 32 * chunk
    cls = out[3]; sq = out[2] >> 32; vel = out[2] & 0xffffffff
    print(f"   classes: -0 numerators {cls & 0xfffff}, tiny {(cls >> 20) & 0xfffff}, huge {cls >> 40}, sqr {sq}")
    print(f"steps {done-chunk:5d}-{done:5d}: {n*n*chunk/ms/1e6:7.2f} G upd/s  fallback lanes/particle-step {lanes/(n*n*chunk):.3e}  warps {warps/tot_warps:.3e}  velocity lanes {vel/(n*n*chunk):.3e}")
