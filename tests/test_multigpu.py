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


@pytest.mark.parametrize("world,nx,ny,steps", [(2, 1024, 1024, 50), (2, 4100, 3000, 150), (4, 2048, 2048, 80), (8, 8192, 4096, 60)])
def test_linked_row_bands_across_gpus_equal_single_gpu(world, nx, ny, steps):
    """Linked row bands, one process per GPU over CUDA IPC: in-kernel peer stores of the boundary rows over NVLink and
    cross-GPU tile flags; the process group only carries the endpoints and the barriers of the link protocol."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world),
           os.path.join(helpers.ROOT, "tools", "band_check.py"), "--nx", str(nx), "--ny", str(ny), "--steps", str(steps), "--linked", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "BAND_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
