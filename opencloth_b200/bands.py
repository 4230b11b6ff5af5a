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
import contextlib

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

    def step_split(self, n, side_stream):
        """Advance n substeps; when this uses up the halo budget the last substep publishes its boundary rows
        first and `side_stream` is made to wait for exactly those — returns True in that case."""
        return self.cloth.step_split(n, side_stream.cuda_stream)

    def refreshed(self):
        self.cloth.halo_refreshed()

    @property
    def budget(self):
        return self.cloth.halo_budget

    def use_current_stream(self):
        """Run the band's kernels on torch's current stream so that torch.distributed orders against them
        (torch's default stream has handle 0 = CUDA's legacy default stream, which oc_set_stream(NULL) now means)."""
        self.cloth.set_stream(torch.cuda.current_stream(self.device).cuda_stream)


class LinkedBandDriver:
    """One row band per process / GPU, linked to its neighbours through peer memory (include/opencloth.h, "linked
    row bands").  The process group is used ONCE, to hand the neighbours' endpoints round and for the barriers of the
    link protocol; after that ``step(n)`` is just ``oc_step(n)`` on every rank — the kernel itself stores the two
    boundary rows into the neighbour's halo over NVLink and orders the steps of different GPUs with flag words, tile by
    tile.  No exchange step, no recomputed halo rows, no host synchronisation while stepping.

    Every rank must call step() with the same n in the same order.
    """

    def __init__(self, band, rank, world, group=None):
        self.band, self.rank, self.world, self.group = band, rank, world, group
        self.cloth = band.cloth
        self.exchanges = 0                  # host-side exchanges on the step path: none

    def link(self):
        c = self.cloth
        c.sync()
        blobs = [None] * self.world
        dist.all_gather_object(blobs, c.band_endpoint(), group=self.group)      # doubles as the barrier after sync
        c.band_link(blobs[self.rank - 1] if self.rank > 0 else None,
                    blobs[self.rank + 1] if self.rank + 1 < self.world else None)
        self.resync(barrier_before=False)

    def resync(self, barrier_before=True):
        """After oc_upload / oc_set_particle on the bands: refresh the halo rows from the neighbours' owned rows."""
        c = self.cloth
        if barrier_before:
            c.sync()
            dist.barrier(group=self.group)
        c.band_pull_halo()
        c.sync()
        dist.barrier(group=self.group)

    def exchange(self):
        self.resync()

    def step(self, n):
        self.cloth.step(n)

    def finish(self):
        pass


class BandDriver:
    """Steps one band in lock step with its neighbours.

    exchange(): post the receives of both halos and the sends of the owned boundary rows as one
    batch of point-to-point operations (NCCL groups them; over gloo they are plain isend/irecv), wait,
    mark the halo current.  step(n): as many groups of `halo_rows/2` substeps as needed, one exchange
    before each group.  With overlap=True the exchange that follows a full group is started on a side
    stream as soon as the group's last substep has produced its boundary rows (oc_step_split) and runs
    concurrently with the interior of that substep; call finish() before reading the state.
    """

    def __init__(self, band, rank, world, group=None, overlap=False):
        self.band, self.rank, self.world, self.group = band, rank, world, group
        self.exchanges = 0
        # torch.distributed orders its send/recv against torch's CURRENT stream: the band's kernels must run on that
        # very stream, or the exchange could read boundary rows before the step has written them.  CUDA bands are
        # bound to a dedicated torch stream here; step() and exchange() run inside it.
        self.stream = None
        if hasattr(band, "cloth") and torch.cuda.is_available():
            self.stream = torch.cuda.Stream(band.device)
            band.cloth.set_stream(self.stream.cuda_stream)
        # overlap: start the exchange as soon as the boundary rows of the last substep of a group exist, on a
        # side stream, while the interior of that substep is still being computed (CUDA bands only)
        self.overlap = overlap and hasattr(band, "step_split") and torch.cuda.is_available()
        self.side = torch.cuda.Stream(band.device) if self.overlap else None
        self.pending = None

    # Largest single point-to-point message, in float4 elements (3 MiB).  Measured on 2 and 8 B200s: NCCL send/recv of
    # one 5 MiB or 6 MiB tensor (halo_rows 40 / 48 at nx = 8192) is ~100x slower than of a 3 MiB one (halo_rows 24),
    # so long halo regions are sent as several messages.
    MAX_MSG = 3 * (1 << 20) // 16

    def _ops(self):
        ops = []
        for side, peer in ((0, self.rank - 1), (1, self.rank + 1)):
            if peer < 0 or peer >= self.world:
                continue
            for fn, send in ((dist.irecv, False), (dist.isend, True)):
                for t in self.band.regions(side, send=send):
                    if t is None:
                        continue
                    for lo in range(0, t.shape[0], self.MAX_MSG):
                        ops.append(dist.P2POp(fn, t[lo:lo + self.MAX_MSG], peer, self.group))
        return ops

    def _finish_pending(self):
        if self.pending is not None:
            for w in self.pending:
                w.wait()                      # the current (compute) stream waits for the exchange
            self.pending = None
            self.band.refreshed()
            self.exchanges += 1

    def _in_stream(self):
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def exchange(self):
        with self._in_stream():
            self._exchange()

    def step(self, n):
        with self._in_stream():
            self._step(n)

    def finish(self):
        """Complete an exchange that is still in flight (call before reading the state)."""
        with self._in_stream():
            self._finish_pending()

    def _exchange(self):
        if self.pending is not None:
            return self._finish_pending()
        ops = self._ops()
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self.band.refreshed()
        self.exchanges += 1

    def _step(self, n):
        while n > 0:
            if self.band.budget == 0:
                self._exchange()
            m = min(n, self.band.budget)
            if self.overlap and m == self.band.budget:
                # this group ends at an exchange point: boundary rows first, exchange on the side stream
                if self.band.step_split(m, self.side):
                    ops = self._ops()
                    if ops:
                        with torch.cuda.stream(self.side):
                            self.pending = dist.batch_isend_irecv(ops)
                    else:
                        self.pending = []
            else:
                self.band.step(m)
            n -= m

def gather_rows(local_x, ny, nx, world, rank, group=None):
    """All ranks' owned rows concatenated on every rank (diagnostics / tests only; not on the step path)."""
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local_x), group=group)
    return np.concatenate(parts, axis=0)
