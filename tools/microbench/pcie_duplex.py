#!/usr/bin/env python
"""PCIe copy rates of the box: H2D alone, D2H alone, both directions at once (pinned host memory, 100 MB buffers,
two streams).  Decides how much the upload/step/download pipeline of oc_upload can gain from full duplex."""
import time
import torch
n = 100 * (1 << 20)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=1, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step = n // chunks
        for k in range(chunks):
            sl = slice(k * step, (k + 1) * step)
            if h2d:
                with torch.cuda.stream(s1): d_in[sl].copy_(h_in[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for chunks in (1, 16):
    a = run(True, False, chunks); b = run(False, True, chunks); c = run(True, True, chunks)
    print(f"chunks {chunks:2d}: H2D {n / a / 1e9:5.1f} GB/s  D2H {n / b / 1e9:5.1f} GB/s  both at once {n / c / 1e9:5.1f} GB/s per direction ({2 * n / c / 1e9:5.1f} aggregate), {c * 1e3:.2f} ms for 100 MB each way")

# the library's pattern: X and X_last are separate host arrays, each chunk is two copies per direction
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); h3 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda"); d3 = torch.empty(n, dtype=torch.uint8, device="cuda")
def run2(chunks, reps=10, lag=2):
    half = n // 2
    step = half // chunks
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        evs = []
        for k in range(chunks):
            sl = slice(k * step, (k + 1) * step)
            with torch.cuda.stream(s1):
                d_in[sl].copy_(h_in[sl], non_blocking=True); d2[sl].copy_(h2[sl], non_blocking=True)
                e = torch.cuda.Event(); e.record(s1); evs.append(e)
        for k in range(chunks):
            sl = slice(k * step, (k + 1) * step)
            with torch.cuda.stream(s2):
                s2.wait_event(evs[min(k + lag, chunks - 1)])
                h_out[sl].copy_(d_out[sl], non_blocking=True); h3[sl].copy_(d3[sl], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for chunks in (8, 16, 32):
    c = run2(chunks)
    print(f"library pattern, {chunks:2d} chunks (2 x {n // 2 // chunks / 1e6:.1f} MB copies per chunk and direction, D2H of chunk k after H2D of chunk k+2): {c * 1e3:.2f} ms for 100 MB each way = {n / c / 1e9:.1f} GB/s per direction")
