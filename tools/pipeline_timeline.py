#!/usr/bin/env python
"""Per-chunk time line of one oc_upload + oc_step(1) + oc_download (OC_DEBUG=64): when each chunk's H2D copy, unpack,
step, pack and D2H copy finished, in ms after the start of the upload."""
import ctypes, os, sys, time
os.environ["OC_DEBUG"] = "64"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opencloth_b200 as m
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
c = m.Cloth(nx, nx); c.step(5)
hx = torch.empty((nx * nx, 3), dtype=torch.float32).pin_memory(); hl = torch.empty((nx * nx, 3), dtype=torch.float32).pin_memory()
c.download_into(hx.data_ptr(), hl.data_ptr(), 3)
for i in range(4):
    t0 = time.perf_counter()
    c.upload_from(hx.data_ptr(), hl.data_ptr(), 3); t1 = time.perf_counter()
    c.step(1); t2 = time.perf_counter()
    c.download_into(hx.data_ptr(), hl.data_ptr(), 3); t3 = time.perf_counter()
print("host: upload call %.3f ms, step call %.3f ms, download call (incl. sync) %.3f ms, total %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
out = (ctypes.c_float * (5 * 32))()
n = c._lib.oc_debug_pipeline(c._h, out, 32)
print("chunk   h2d   unpack   step    pack    d2h   (ms after upload start)")
for k in range(n):
    print("%5d " % k + " ".join("%7.3f" % out[5 * k + q] for q in range(5)))
