#!/usr/bin/env python
"""End-to-end rate of oc_upload + oc_step(1) + oc_download with pinned host buffers (bench.py's e2e leg) for several
chunk counts of the library's host <-> device pipeline (OC_PIPE_CHUNKS; 1 = no pipelining)."""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time, torch
sys.path.insert(0, %r)
import opencloth_b200 as m
nx = int(sys.argv[1])
c = m.Cloth(nx, nx); c.step(5)
hx = torch.empty((nx * nx, 3), dtype=torch.float32).pin_memory(); hl = torch.empty((nx * nx, 3), dtype=torch.float32).pin_memory()
c.download_into(hx.data_ptr(), hl.data_ptr(), 3)
def it():
    c.upload_from(hx.data_ptr(), hl.data_ptr(), 3); c.step(1); c.download_into(hx.data_ptr(), hl.data_ptr(), 3)
for i in range(3): it()
t0 = time.perf_counter()
for i in range(20): it()
dt = (time.perf_counter() - t0) / 20
print("n", nx, "chunks", sys.argv[2], "ms/iter %%.3f" %% (dt * 1e3), "G updates/s %%.3f" %% (nx * nx / dt / 1e9))
''' % ROOT
for nx in (2048,):
    for ch in (1, 4, 8, 16, 32):
        subprocess.run([sys.executable, "-c", code, str(nx), str(ch)], env=dict(os.environ, OC_PIPE_CHUNKS=str(ch)))
