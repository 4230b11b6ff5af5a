"""GPU tier, needs >= 2 GPUs (skipped otherwise): the multi-process row-band path over NCCL
(torchrun, one rank per GPU) equals the single-GPU result bit for bit."""
import os
import subprocess
import sys

import pytest

import helpers

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,k,overlap", [(2, 2, 0), (2, 1, 1), (4, 1, 1), (8, 1, 1), (8, 2, 0)])
def test_row_bands_over_nccl_equal_single_gpu(world, k, overlap):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world + 10 * k + overlap),
           os.path.join(helpers.ROOT, "tools", "band_check.py"), "--nx", "1024", "--ny", "1024", "--steps", "37", "--halo", "8", "--k", str(k), "--overlap", str(overlap)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "BAND_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
