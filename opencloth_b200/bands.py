"""Row-band decomposition of one large cloth across processes (one process per GPU).

SURVEY.md section 8(e): the step's only data dependence is the reach-2 stencil of the bend springs
(reference OpenCloth_Verlet/OpenCloth_Verlet/main.cpp:311, :317), so the rows are cut into `world`
contiguous bands.  Each band stores `halo_rows` rows of its neighbours' state either side; after an
exchange it can take `halo_rows / 2` substeps on its own, recomputing a shrinking part of the halo
(communication avoiding: one exchange per group of substeps, and the exchange is 2 x halo_rows x nx
float4 per neighbour — 1 MiB at nx = 8192, halo_rows = 8).  No collective is needed: the exchange is
pairwise send/recv with the band above and below.

The host logic here is backend agnostic so that the CPU test tier can run it with world_size 2 over
gloo (tests/test_bands_gloo.py drives it with the kernel emulator as the band object); in production
the band object is an ``opencloth_b200.Cloth`` created with a row range and the tensors alias the
library's device buffers (NCCL send/recv over NVLink).
"""
import numpy as np
import torch
import torch.distributed as dist


def band_rows(ny, world, rank):
    """Rows [begin, end) owned by `rank`: contiguous, sizes differ by at most one row."""
    begin = (ny * rank) // world
    end = (ny * (rank + 1)) // world
    return begin, end


class _CudaAlias:
    """Zero-copy view of library-owned device memory as a torch tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count, 4), "typestr": "<f4", "data": (ptr, False), "version": 2}


def alias_cuda(ptr, count, device):
    if count == 0 or ptr == 0:
        return None
    return torch.as_tensor(_CudaAlias(ptr, count), device=device)


class CudaBand:
    """Band object of the product: a Cloth handle owning rows [begin, end) on the current device."""

    def __init__(self, nx, ny, world, rank, halo_rows, device, **params):
        from .cloth import Cloth
        self.begin, self.end = band_rows(ny, world, rank)
        self.device = torch.device("cuda", device)
        self.cloth = Cloth(nx, ny, row_begin=self.begin, row_end=self.end, halo_rows=halo_rows, device=device, **params)
        self.halo_rows = halo_rows

    def regions(self, side, send):
        """(X(t) rows, X(t-1) rows) to send to / receive from the neighbour on `side` (0 = up, 1 = down)."""
        out = []
        for which in (0, 1):
            ptr, cnt = self.cloth.halo_region(side, which, send)
            out.append(alias_cuda(ptr, cnt, self.device))
        return out

    def step(self, n):
        self.cloth.step(n)

    def refreshed(self):
        self.cloth.halo_refreshed()

    @property
    def budget(self):
        return self.cloth.halo_budget

    def use_current_stream(self):
        """Run the band's kernels on torch's current stream so that torch.distributed orders against them."""
        self.cloth.set_stream(torch.cuda.current_stream(self.device).cuda_stream)


class BandDriver:
    """Steps one band in lock step with its neighbours.

    exchange(): post the receives of both halos and the sends of the owned boundary rows as one
    batch of point-to-point operations (NCCL groups them; over gloo they are plain isend/irecv), wait,
    mark the halo current.  step(n): as many groups of `halo_rows/2` substeps as needed, one exchange
    before each group.
    """

    def __init__(self, band, rank, world, group=None):
        self.band, self.rank, self.world, self.group = band, rank, world, group
        self.exchanges = 0

    def exchange(self):
        ops = []
        for side, peer in ((0, self.rank - 1), (1, self.rank + 1)):
            if peer < 0 or peer >= self.world:
                continue
            for t in self.band.regions(side, send=False):
                ops.append(dist.P2POp(dist.irecv, t, peer, self.group))
            for t in self.band.regions(side, send=True):
                ops.append(dist.P2POp(dist.isend, t, peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self.band.refreshed()
        self.exchanges += 1

    def step(self, n):
        while n > 0:
            if self.band.budget == 0:
                self.exchange()
            m = min(n, self.band.budget)
            self.band.step(m)
            n -= m


def gather_rows(local_x, ny, nx, world, rank, group=None):
    """All ranks' owned rows concatenated on every rank (diagnostics / tests only; not on the step path)."""
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local_x), group=group)
    return np.concatenate(parts, axis=0)
