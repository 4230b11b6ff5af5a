"""CPU tier, world_size 2 (and 3) over gloo: the multi-process row-band driver
(opencloth_b200/bands.py — the code bench.py runs under torchrun with NCCL) exchanging halos between
processes.  The band object here is the CPU kernel emulator, so the very same marching-kernel body,
host sequencing and halo-region arithmetic as on the GPU are exercised; the result must equal the
undivided cloth stepped by the oracle, bit for bit."""
import ctypes
import os
import socket
import sys
import time

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class EmuBand:
    """Same surface as bands.CudaBand, backed by tests/emu (host memory)."""

    def __init__(self, nx, ny, world, rank, halo_rows, k):
        import helpers
        from opencloth_b200.bands import band_rows
        self.begin, self.end = band_rows(ny, world, rank)
        self.e = helpers.Emu(nx, ny, row_begin=self.begin, row_end=self.end, halo_rows=halo_rows)
        self.k = k

    def regions(self, side, send):
        out = []
        for which in (0, 1):
            ptr, cnt = self.e.halo_region(side, which, send)
            if cnt == 0:
                out.append(None)
                continue
            a = np.ctypeslib.as_array((ctypes.c_float * (cnt * 4)).from_address(ptr)).reshape(cnt, 4)
            out.append(torch.from_numpy(a))
        return [t for t in out if t is not None]

    def step(self, n):
        self.e.step(n, kernel=2, exact=1, k=self.k, TW=32, RS=5)

    def refreshed(self):
        self.e.halo_refreshed()

    @property
    def budget(self):
        return self.e.halo_budget


def _worker(rank, world, port, nx, ny, halo, k, steps, x0, xl0, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import helpers
        from opencloth_b200.bands import BandDriver, gather_rows
        helpers.ensure_built()
        band = EmuBand(nx, ny, world, rank, halo, k)
        sl = slice(band.begin * nx, band.end * nx)
        band.e.upload(x0[sl], xl0[sl])
        drv = BandDriver(band, rank, world)
        for n in steps:
            drv.step(n)
        x, xl = band.e.download()
        X = gather_rows(x, ny, nx, world, rank)
        XL = gather_rows(xl, ny, nx, world, rank)
        if rank == 0:
            ret["x"], ret["xl"], ret["exchanges"] = X, XL, drv.exchanges
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo,k,steps", [(2, 4, 1, (3, 4)), (2, 8, 2, (9,)), (3, 4, 2, (5, 2))])
def test_multiprocess_row_bands_equal_whole_cloth(world, halo, k, steps):
    import helpers
    nx, ny = 19, 36
    x0, xl0 = helpers.developed_state(nx, ny, 1500)
    o = helpers.Oracle(nx, ny)
    o.set_state(x0, xl0)
    o.step(sum(steps))
    ox, oxl = o.state()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), nx, ny, halo, k, steps, x0, xl0, ret), nprocs=world, join=True)
    assert helpers.bitwise_equal(ret["x"], ox), "row-band result differs from the undivided cloth"
    assert helpers.bitwise_equal(ret["xl"], oxl)
    per = halo // 2
    assert ret["exchanges"] >= -(-sum(steps) // per)      # at least one exchange per halo_rows/2 substeps


class _FakeLinkedCloth:
    """Records the calls LinkedBandDriver makes (the link protocol of include/opencloth.h) into a shared log."""

    def __init__(self, rank, log):
        self.rank, self.log = rank, log

    def _rec(self, what):
        self.log.append((time.monotonic(), self.rank, what))

    def sync(self):
        self._rec("sync")

    def band_endpoint(self):
        self._rec("endpoint")
        return b"endpoint-of-%d" % self.rank

    def band_link(self, upper, lower):
        self._rec("link")
        self.linked = (upper, lower)

    def band_pull_halo(self):
        self._rec("pull")

    def step(self, n):
        self._rec("step%d" % n)


class _FakeBand:
    def __init__(self, cloth):
        self.cloth = cloth


def _linked_worker(rank, world, port, log, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from opencloth_b200.bands import LinkedBandDriver
        c = _FakeLinkedCloth(rank, log)
        time.sleep(0.05 * rank)                   # skew the ranks: the barriers must still separate the phases
        drv = LinkedBandDriver(_FakeBand(c), rank, world)
        drv.link()
        drv.step(5)
        drv.resync()
        drv.step(2)
        ret[rank] = c.linked
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_linked_band_driver_protocol_over_gloo(world):
    """Host side of the linked-band path (the one bench.py runs under torchrun): every rank receives exactly its
    neighbours' endpoints; nobody links before every band has synchronised, nobody steps before every band has pulled
    its halos, and a resync pulls only after every band has stopped stepping."""
    mgr = mp.Manager()
    log, ret = mgr.list(), mgr.dict()
    mp.spawn(_linked_worker, args=(world, _free_port(), log, ret), nprocs=world, join=True)
    for r in range(world):
        up, lo = ret[r]
        assert up == (b"endpoint-of-%d" % (r - 1) if r > 0 else None)
        assert lo == (b"endpoint-of-%d" % (r + 1) if r + 1 < world else None)
    ev = sorted(log)
    def times(what):
        return [t for t, _, w in ev if w == what]
    assert len(times("link")) == world and len(times("pull")) == 2 * world
    assert max(times("endpoint")) <= min(times("link"))                  # all_gather of the endpoints is the barrier
    first_pulls = sorted(times("pull"))[:world]
    assert max(first_pulls) <= min(times("step5"))                       # halos current everywhere before the first step
    second_pulls = sorted(times("pull"))[world:]
    assert max(times("step5")) <= min(second_pulls)                      # resync: everybody stopped stepping first
    assert max(second_pulls) <= min(times("step2"))


def test_band_rows_partition():
    from opencloth_b200.bands import band_rows
    for ny, world in ((8192, 8), (36, 3), (21, 4), (2048, 5)):
        cuts = [band_rows(ny, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == ny
        for a, b in zip(cuts, cuts[1:]):
            assert a[1] == b[0]
        sizes = [e - b for b, e in cuts]
        assert max(sizes) - min(sizes) <= 1
