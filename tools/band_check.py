#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): a cloth cut into row bands with NCCL halo exchange must
equal the same cloth stepped on one GPU, bit for bit.  Prints 'BAND_CHECK OK' on rank 0."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencloth_b200 as oc
from opencloth_b200 import bands as B

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=1024); ap.add_argument("--ny", type=int, default=1024)
ap.add_argument("--steps", type=int, default=50); ap.add_argument("--halo", type=int, default=8)
ap.add_argument("--k", type=int, default=2); ap.add_argument("--exact", type=int, default=1)
ap.add_argument("--overlap", type=int, default=0)
ap.add_argument("--linked", type=int, default=0, help="1: linked bands (in-kernel peer stores + flags), 0: NCCL halo exchange")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
band = B.CudaBand(a.nx, a.ny, world, rank, 2 if a.linked else a.halo, local, exact=a.exact, substeps_per_launch=1 if a.linked else a.k)
drv = B.LinkedBandDriver(band, rank, world) if a.linked else B.BandDriver(band, rank, world, overlap=bool(a.overlap))
# non-trivial start: whole cloth advanced on every rank's own GPU first (deterministic), then cut
whole = oc.Cloth(a.nx, a.ny, device=local, exact=a.exact, substeps_per_launch=4)
whole.step(300)
x0, xl0 = whole.download()
sl = slice(band.begin * a.nx, band.end * a.nx)
band.cloth.upload(x0[sl], xl0[sl])
if a.linked:
    drv.link()
drv.step(a.steps)
drv.finish()
x, xl = band.cloth.download()
whole.step(a.steps)
wx, wxl = whole.download()
ok = bool((x.view(np.uint32) == wx[sl].view(np.uint32)).all() and (xl.view(np.uint32) == wxl[sl].view(np.uint32)).all())
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("BAND_CHECK", "OK" if int(t.item()) == 1 else "MISMATCH", f"world={world} grid={a.nx}x{a.ny} steps={a.steps} linked={a.linked} halo={a.halo} k={a.k} overlap={a.overlap} exchanges={drv.exchanges}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
