"""CPU tier: the product's kernel BODIES (opencloth_b200/csrc/*.cuh) executed on the CPU by
tests/emu (fibers as CUDA threads, cooperative barrier as __syncthreads) against the oracle.

This checks the logic a GPU run would otherwise be needed for — index arithmetic of strips /
segments / stages, the shared-memory pipeline hazards, the accumulation order, the k-substep
temporal blocking, batches and the row-band halo protocol — bit for bit.  The arithmetic on the
host (IEEE fp32, no contraction) is what MathExact's device intrinsics compute."""
import numpy as np
import pytest

import helpers
from helpers import Emu, Oracle, bitwise_equal

GATHER, MARCH, MARCH2, RESIDENT, TWIN, STREAM, STREAM2 = 1, 2, 3, 4, 5, 6, 7


def run_pair(nx, ny, pre, steps, **kw):
    x0, xl0 = helpers.developed_state(nx, ny, pre) if pre else Oracle(nx, ny).state()
    o = Oracle(nx, ny); o.set_state(x0, xl0)
    e = Emu(nx, ny)
    if pre:
        e.upload(x0, xl0)
    e.step(steps, **kw)
    o.step(steps)
    ex, exl = e.download(); ox, oxl = o.state()
    assert bitwise_equal(ex, ox), f"X differs in {(helpers.bits(ex) != helpers.bits(ox)).any(1).sum()} particles"
    assert bitwise_equal(exl, oxl), "X_last differs"
    return int((ox == oxl).all(1).sum())


@pytest.mark.parametrize("nx,ny,pre,steps", [(21, 21, 0, 60), (21, 21, 1800, 150), (37, 23, 1900, 40), (3, 3, 3, 50), (5, 4, 3, 50)])
def test_gather_kernel_body(nx, ny, pre, steps):
    hits = run_pair(nx, ny, pre, steps, kernel=GATHER)
    if pre >= 1800:
        assert hits > 2          # the collider is active in this window


# resident small-cloth kernel: (nx, ny, pre, steps, threads of the CTA); every case also split into several calls
@pytest.mark.parametrize("nx,ny,pre,steps,threads", [(21, 21, 0, 40, 0), (21, 21, 1800, 120, 0), (21, 21, 1800, 60, 96), (37, 23, 1900, 30, 0),
                                                     (3, 3, 3, 60, 0), (5, 4, 3, 60, 32), (39, 39, 1200, 12, 256)])
def test_resident_kernel_body(nx, ny, pre, steps, threads):
    hits = run_pair(nx, ny, pre, steps, kernel=RESIDENT, TW=threads)
    if pre >= 1800:
        assert hits > 2
    # the same in calls of 1, 2, 1, rest substeps (one substep per launch writes one buffer, more write two)
    x0, xl0 = helpers.developed_state(nx, ny, pre) if pre else Oracle(nx, ny).state()
    o = Oracle(nx, ny); o.set_state(x0, xl0); o.step(steps)
    e = Emu(nx, ny)
    if pre:
        e.upload(x0, xl0)
    for n in (1, 2, 1, steps - 4):
        e.step(n, kernel=RESIDENT, TW=threads)
    ex, exl = e.download(); ox, oxl = o.state()
    assert bitwise_equal(ex, ox) and bitwise_equal(exl, oxl)


@pytest.mark.parametrize("order", [1, 2])
def test_resident_kernel_is_independent_of_thread_schedule(order):
    L = helpers.emu_lib()
    L.emu_set_order(order)
    try:
        run_pair(21, 21, 1800, 40, kernel=RESIDENT, TW=0)
        run_pair(37, 23, 1900, 10, kernel=RESIDENT, TW=160)
    finally:
        L.emu_set_order(0)


# (nx, ny, pre-steps, steps, k, TW, RS): single strip / multi strip, one / many segments, every k
MARCH_CASES = [
    (21, 21, 0, 20, 1, 32, 0), (21, 21, 0, 20, 1, 32, 7), (21, 21, 0, 20, 2, 32, 0), (21, 21, 0, 20, 4, 32, 5),
    (21, 21, 0, 16, 8, 32, 0), (21, 21, 1800, 100, 4, 32, 0), (21, 21, 1800, 40, 8, 32, 6),
    (37, 23, 1900, 12, 1, 16, 5), (37, 23, 1900, 12, 2, 16, 6), (37, 23, 1900, 12, 3, 16, 4),
    (64, 64, 2000, 8, 1, 32, 10), (64, 64, 2000, 8, 2, 32, 9), (64, 64, 2000, 8, 4, 32, 0),
    (70, 40, 1500, 8, 4, 64, 13), (70, 40, 1500, 8, 8, 64, 0),
    (130, 20, 500, 4, 4, 128, 7), (130, 20, 500, 4, 1, 128, 0), (128, 24, 500, 4, 2, 128, 0),
    (3, 3, 5, 40, 1, 32, 0), (3, 3, 5, 40, 8, 32, 0), (5, 4, 5, 40, 2, 16, 2), (4, 9, 5, 23, 3, 16, 3),
]


@pytest.mark.parametrize("nx,ny,pre,steps,k,TW,RS", MARCH_CASES)
def test_march_kernel_body(nx, ny, pre, steps, k, TW, RS):
    run_pair(nx, ny, pre, steps, kernel=MARCH, k=k, TW=TW, RS=RS)


# two-columns-per-thread kernel: (nx, ny, pre, steps, window columns WC, rows per segment)
MARCH2_CASES = [
    (21, 21, 0, 20, 32, 0), (21, 21, 1800, 60, 32, 7), (21, 21, 1800, 30, 16, 5), (37, 23, 1900, 12, 16, 5), (37, 23, 1900, 12, 64, 0),
    (64, 64, 2000, 8, 32, 10), (64, 64, 2000, 8, 64, 9), (70, 40, 1500, 8, 64, 13), (130, 20, 500, 4, 128, 7), (128, 24, 500, 4, 128, 0),
    (3, 3, 5, 40, 16, 0), (5, 4, 5, 40, 16, 2), (4, 9, 5, 23, 16, 3), (33, 30, 1700, 10, 32, 11),
    # shorter segments in the edge strips (OcSeg2): RS = rows per segment | edge rows per segment << 8
    (70, 40, 1500, 8, 64, 9), (70, 40, 1500, 8, 32, 12 | 9 << 8), (37, 23, 1900, 12, 16, 7 | 4 << 8), (37, 23, 1900, 12, 16, 6 | 5 << 8),
]


@pytest.mark.parametrize("nx,ny,pre,steps,WC,RS", MARCH2_CASES)
def test_march2_kernel_body(nx, ny, pre, steps, WC, RS):
    run_pair(nx, ny, pre, steps, kernel=MARCH2, k=1, TW=WC, RS=RS)


@pytest.mark.parametrize("order", [1, 2])
def test_march2_is_independent_of_thread_schedule(order):
    L = helpers.emu_lib()
    L.emu_set_order(order)
    try:
        run_pair(37, 23, 1900, 9, kernel=MARCH2, k=1, TW=16, RS=4)
        run_pair(70, 40, 1500, 6, kernel=MARCH2, k=1, TW=64, RS=11)
    finally:
        L.emu_set_order(0)


# twin-tile kernel (oc_twin.cuh): (nx, ny, pre, steps, window columns WC = threads, rows per segment); the number of
# segments is even (a CTA takes segments 2k and 2k+1); RS = 0: two segments
TWIN_CASES = [
    (21, 21, 0, 20, 32, 0), (21, 21, 1800, 60, 32, 6), (21, 21, 1800, 30, 16, 4), (37, 23, 1900, 12, 16, 6), (37, 23, 1900, 12, 64, 0),
    (37, 23, 1900, 12, 8, 3), (64, 64, 2000, 8, 32, 11), (64, 64, 2000, 8, 64, 9), (70, 40, 1500, 8, 64, 13), (130, 20, 500, 4, 128, 5),
    (128, 24, 500, 4, 128, 0), (3, 3, 5, 40, 16, 0), (5, 4, 5, 40, 16, 2), (4, 9, 5, 23, 16, 5), (33, 30, 1700, 10, 32, 8), (33, 31, 1700, 10, 32, 4),
]


@pytest.mark.parametrize("kernel", [TWIN, STREAM, STREAM2])
@pytest.mark.parametrize("nx,ny,pre,steps,WC,RS", TWIN_CASES)
def test_twin_kernel_body(nx, ny, pre, steps, WC, RS, kernel):
    """oc_k_twin and oc_k_stream (the streaming gather kernel runs on the same twin tiles)"""
    run_pair(nx, ny, pre, steps, kernel=kernel, k=1, TW=WC, RS=RS)


def test_twin_refuses_an_odd_number_of_segments():
    e = Emu(21, 21)
    assert helpers.emu_lib().emu_step(e.h, 1, TWIN, 1, 1, 32, 7) == -2      # 3 segments of 7 rows


@pytest.mark.parametrize("kernel", [TWIN, STREAM, STREAM2])
@pytest.mark.parametrize("order", [1, 2])
def test_twin_is_independent_of_thread_schedule(order, kernel):
    L = helpers.emu_lib()
    L.emu_set_order(order)
    try:
        run_pair(37, 23, 1900, 9, kernel=kernel, k=1, TW=16, RS=3)
        run_pair(70, 40, 1500, 6, kernel=kernel, k=1, TW=64, RS=10)
    finally:
        L.emu_set_order(0)


@pytest.mark.parametrize("kernel", [TWIN, STREAM, STREAM2])
@pytest.mark.parametrize("pair_cloths,B,RS", [(1, 4, 0), (1, 2, 6), (0, 3, 3), (0, 2, 9)])
def test_twin_batched_cloths_are_independent(pair_cloths, B, RS, kernel):
    """The twins of a CTA are two segments of one cloth, or (bit 16 of RS) the same tile of two cloths of the batch."""
    nx, ny = 20, 17
    rng = np.random.RandomState(11)
    x0, xl0 = helpers.developed_state(nx, ny, 1750)
    starts = []
    for b in range(B):
        x = x0.copy()
        x[:, 1] += (1e-3 * rng.uniform(-1, 1, len(x))).astype(np.float32)
        starts.append(x)
    e = Emu(nx, ny, batch=B)
    e.upload(np.concatenate(starts), np.concatenate([xl0] * B))
    e.step(25, kernel=kernel, TW=16, RS=RS | pair_cloths << 16)
    ex, exl = e.download()
    for b in range(B):
        o = Oracle(nx, ny); o.set_state(starts[b], xl0); o.step(25)
        ox, oxl = o.state()
        sl = slice(b * nx * ny, (b + 1) * nx * ny)
        assert bitwise_equal(ex[sl], ox) and bitwise_equal(exl[sl], oxl), f"cloth {b}"


def test_stream_fast_mode_does_not_depend_on_the_tiling():
    """oc_k_stream's tolerance mode rounds alike on its steady and generic paths (one explicit fused multiply-add per
    term on both), so a result is independent of window width, rows per tile and of which rows are tile edges —
    which is what lets bench.py compare row bands with the whole cloth bit for bit in fast mode too."""
    nx, ny = 37, 30
    x0, xl0 = helpers.developed_state(nx, ny, 1700)
    ref = None
    for WC, RS in ((64, 0), (16, 5), (8, 3), (32, 15), (16, 8)):     # (steady loop: register window, 5 rows per trip + remainder)
        e = Emu(nx, ny); e.upload(x0, xl0); e.step(12, kernel=STREAM, exact=0, TW=WC, RS=RS)
        got = e.download()
        ref = ref or got
        assert bitwise_equal(got[0], ref[0]) and bitwise_equal(got[1], ref[1]), (WC, RS)
    # oc_k_stream2 (two columns per thread) does the same arithmetic per particle in the same order: the same bits
    for WC, RS in ((64, 0), (16, 5)):
        e = Emu(nx, ny); e.upload(x0, xl0); e.step(12, kernel=STREAM2, exact=0, TW=WC, RS=RS)
        got = e.download()
        assert bitwise_equal(got[0], ref[0]) and bitwise_equal(got[1], ref[1]), ("stream2", WC, RS)
    # long tiles: several trips of the unrolled steady loop and every remainder, against short tiles and oc_k_stream2
    nx, ny = 20, 66
    x0, xl0 = helpers.developed_state(nx, ny, 900)
    ref = None
    for kern, WC, RS in ((STREAM2, 32, 0), (STREAM, 32, 0), (STREAM, 16, 11), (STREAM, 32, 17), (STREAM, 8, 3), (STREAM, 32, 19 | (1 << 17))):
        e = Emu(nx, ny); e.upload(x0, xl0); e.step(6, kernel=kern, exact=0, TW=WC, RS=RS)
        got = e.download()
        ref = ref or got
        assert bitwise_equal(got[0], ref[0]) and bitwise_equal(got[1], ref[1]), (kern, WC, RS)


@pytest.mark.parametrize("order", [1, 2])
def test_march_is_independent_of_thread_schedule(order):
    """No intra-phase data race: resuming the threads in reverse / pseudo-random order between
    barriers must not change a bit."""
    L = helpers.emu_lib()
    L.emu_set_order(order)
    try:
        run_pair(37, 23, 1900, 9, kernel=MARCH, k=3, TW=16, RS=4)
        run_pair(64, 64, 2000, 8, kernel=MARCH, k=4, TW=32, RS=11)
        run_pair(70, 40, 1500, 8, kernel=MARCH, k=8, TW=64, RS=0)
    finally:
        L.emu_set_order(0)


def test_temporal_blocking_equals_single_steps():
    """k substeps per launch == k launches of one substep ('temporal blocking')."""
    x0, xl0 = helpers.developed_state(40, 33, 1700)
    ref = Emu(40, 33); ref.upload(x0, xl0); ref.step(24, kernel=MARCH, k=1, TW=64)
    rx, rxl = ref.download()
    for k in (2, 4, 8):
        e = Emu(40, 33); e.upload(x0, xl0); e.step(24, kernel=MARCH, k=k, TW=64, RS=9)
        x, xl = e.download()
        assert bitwise_equal(x, rx) and bitwise_equal(xl, rxl), f"k={k}"


def test_fast_mode_body_within_tolerance():
    """MathFast on the host (FMA-free, 1/sqrt) still re-associates: must stay within the north-star
    tolerance of the oracle: <= 1e-5 of the cloth extent (fullsize = 4) after 100 steps."""
    o = Oracle(21, 21); o.step(100)
    e = Emu(21, 21); e.step(100, kernel=MARCH, exact=0, k=4, TW=32)
    err = np.abs(e.download()[0].astype(np.float64) - o.state()[0]).max() / 4.0
    assert err <= 1e-5, err


def test_batched_cloths_are_independent():
    """batch > 1: every cloth of the batch evolves exactly like a single cloth from the same start."""
    nx, ny, B = 20, 17, 3
    rng = np.random.RandomState(7)
    base = Oracle(nx, ny).state()[0]
    starts = []
    for b in range(B):
        x = base.copy()
        x[:, 1] += (1e-3 * rng.uniform(-1, 1, len(x))).astype(np.float32)
        starts.append(x)
    e = Emu(nx, ny, batch=B)
    X0 = np.concatenate(starts)
    e.upload(X0, X0)
    e.step(30, kernel=MARCH, k=2, TW=32, RS=6)
    ex, exl = e.download()
    for b in range(B):
        o = Oracle(nx, ny); o.set_state(starts[b], starts[b]); o.step(30)
        ox, oxl = o.state()
        sl = slice(b * nx * ny, (b + 1) * nx * ny)
        assert bitwise_equal(ex[sl], ox) and bitwise_equal(exl[sl], oxl), f"cloth {b}"


@pytest.mark.parametrize("nbands,halo,k,kernel", [(2, 4, 1, MARCH), (3, 8, 2, MARCH), (4, 4, 2, MARCH), (2, 6, 1, GATHER), (3, 8, 4, MARCH), (3, 8, 1, TWIN), (2, 4, 1, TWIN), (3, 8, 1, STREAM), (2, 4, 1, STREAM), (3, 8, 1, STREAM2)])
def test_row_bands_with_halo_exchange_equal_single_domain(nbands, halo, k, kernel):
    """Row-band decomposition: g bands with halo_rows rows of neighbour state, one
    exchange per halo_rows/2 substeps, redundant recomputation of the shrinking halo in between.
    Result must equal the undivided cloth bit for bit."""
    nx, ny = 23, 48
    x0, xl0 = helpers.developed_state(nx, ny, 1500)
    whole = Oracle(nx, ny); whole.set_state(x0, xl0)
    cuts = [round(ny * b / nbands) for b in range(nbands + 1)]
    bands = []
    for b in range(nbands):
        e = Emu(nx, ny, row_begin=cuts[b], row_end=cuts[b + 1], halo_rows=halo)
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        e.upload(x0[sl], xl0[sl])
        assert e.halo_budget == 0           # an uploaded band must exchange before it may step
        bands.append(e)

    def exchange():
        L = helpers.emu_lib()
        for b in range(nbands):
            if b + 1 < nbands:
                assert L.emu_halo_copy(bands[b].h, 1, bands[b + 1].h) == 0
                assert L.emu_halo_copy(bands[b + 1].h, 0, bands[b].h) == 0
        for e in bands:
            e.halo_refreshed()

    total = 0
    per = halo // 2
    for rnd in range(3):
        exchange()
        n = per if rnd < 2 else max(1, per - 1)     # last round: a partial group
        for e in bands:
            e.step(n, kernel=kernel, k=k, TW=32, RS=0 if kernel in (TWIN, STREAM, STREAM2) else 5)
        total += n
    whole.step(total)
    wx, wxl = whole.state()
    for b, e in enumerate(bands):
        x, xl = e.download()
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        assert bitwise_equal(x, wx[sl]) and bitwise_equal(xl, wxl[sl]), f"band {b}"


@pytest.mark.parametrize("nbands,WC,RS,nx,kernel", [(2, 16, 7, 23, MARCH2), (3, 16, 6, 37, MARCH2), (4, 32, 12, 23, MARCH2), (3, 16, 0, 30, MARCH2),
                                                    (2, 16, 6, 23, TWIN), (3, 16, 4, 37, TWIN), (4, 32, 6, 23, TWIN), (3, 16, 0, 30, TWIN),
                                                    (2, 16, 6, 23, STREAM), (3, 16, 4, 37, STREAM), (4, 32, 6, 23, STREAM), (3, 16, 0, 30, STREAM), (2, 16, 6, 23, STREAM2), (3, 16, 4, 37, STREAM2), (4, 32, 6, 23, STREAM2)])
def test_linked_row_bands_push_their_boundary_rows(nbands, WC, RS, nx, kernel):
    """Linked row bands (OcPeer2): no exchange step — every band computes exactly its owned rows, and the tiles at a
    band edge store the two rows the neighbour's stencil reaches straight into the neighbour's halo.  Same kernel
    body as the GPU; must equal the undivided cloth bit for bit, through collider contact."""
    ny = 48
    x0, xl0 = helpers.developed_state(nx, ny, 1600)
    whole = Oracle(nx, ny); whole.set_state(x0, xl0)
    cuts = [round(ny * b / nbands) for b in range(nbands + 1)]
    bands = []
    for b in range(nbands):
        e = Emu(nx, ny, row_begin=cuts[b], row_end=cuts[b + 1], halo_rows=2)
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        e.upload(x0[sl], xl0[sl])
        bands.append(e)
    helpers.emu_link_bands(bands)
    steps = 9
    for s in range(steps):
        for e in (bands if s % 2 == 0 else bands[::-1]):       # the order of the bands within a step does not matter
            # odd bands launch their segments bottom to top, as the library does (OcPeer2::rev; bit 17 of RS here)
            e.step(1, kernel=kernel, TW=WC, RS=RS | (bands.index(e) & 1) << 17)
    whole.step(steps)
    wx, wxl = whole.state()
    for b, e in enumerate(bands):
        x, xl = e.download()
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        assert bitwise_equal(x, wx[sl]) and bitwise_equal(xl, wxl[sl]), f"band {b}"


def test_band_refuses_to_step_past_its_halo():
    e = Emu(21, 40, row_begin=10, row_end=30, halo_rows=4)
    assert e.halo_budget == 2
    e.step(2, kernel=MARCH, k=1, TW=32)
    assert e.halo_budget == 0
    rc = helpers.emu_lib().emu_step(e.h, 1, MARCH, 1, 1, 32, 0)
    assert rc == -3


def _mat4(rot_axis_angle, scale, translate):
    """column-major 4x4 of translate * rotate * scale and its inverse, float32 like the reference's GLM matrices"""
    ax, ang = rot_axis_angle
    ax = np.asarray(ax, np.float64); ax /= np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    E = np.eye(4); E[:3, :3] = R @ np.diag(scale); E[:3, 3] = translate
    return E.T.astype(np.float32).ravel(), np.linalg.inv(E).T.astype(np.float32).ravel()


@pytest.mark.parametrize("case", range(6))
def test_collider_bounding_sphere_is_conservative(case):
    """oc_k_march2 skips the collider transform (V:511-513) for particles outside a bounding sphere derived on the
    host from inverse_ellipsoid: no point that the reference's own fp32 test puts inside (or within 2 % of the
    surface) may lie outside that sphere; degenerate colliders switch the shortcut off."""
    rng = np.random.RandomState(100 + case)
    if case == 0:
        kw = {}                                                     # the reference's collider (V:324-327)
    elif case < 4:
        e, ie = _mat4((rng.normal(size=3), rng.uniform(0, 3)), rng.uniform(0.3, 2.5, 3), rng.uniform(-3, 3, 3))
        kw = dict(ellipsoid=e, inv_ellipsoid=ie, center=tuple(rng.uniform(-0.5, 0.5, 3)))
    elif case == 4:
        e, ie = _mat4(((1, 0, 0), 0.3), (1, 1, 1), (500.0, 0, 0))   # large offsets: rounding of V:511 is no longer negligible
        kw = dict(ellipsoid=e, inv_ellipsoid=ie)
    else:
        e, ie = _mat4(((0, 1, 0), 0.0), (1, 1, 1), (0, 0, 0))
        ie = ie.copy(); ie[0:3] = 0.0                               # singular linear part
        kw = dict(ellipsoid=e, inv_ellipsoid=ie)
    emu = Emu(5, 5, **kw)
    bs = emu.bounding_sphere()
    if case >= 4:
        assert np.isinf(bs[3])
        return
    assert np.isfinite(bs).all()
    o = Oracle(5, 5, **kw)
    p = o.params if hasattr(o, "params") else None
    ie = np.asarray(kw.get("inv_ellipsoid", emu_default_inv()), np.float32).reshape(4, 4).T       # row-major A | t
    c = np.asarray(kw.get("center", (0, 0, 0)), np.float32)
    X = (bs[:3] + rng.uniform(-1, 1, (400000, 3)) * np.sqrt(bs[3]) * 1.6).astype(np.float32)
    f = np.float32
    rows = []
    for r in range(3):                                              # products, then left-to-right sums (type_mat4x4.inl:567-571)
        acc = f(ie[r, 0]) * X[:, 0] + f(ie[r, 1]) * X[:, 1]
        acc = acc + f(ie[r, 2]) * X[:, 2]
        acc = acc + f(ie[r, 3])
        rows.append(acc - c[r])
    sq = rows[0] * rows[0] + rows[1] * rows[1] + rows[2] * rows[2]
    d2 = ((X - bs[:3]) ** 2).sum(1)
    inside = sq < f(1.04)
    assert inside.sum() > 1000, "the sample does not reach the collider"
    assert (d2[inside] <= bs[3]).all(), "a point inside the collider lies outside its bounding sphere"


def emu_default_inv():
    from opencloth_b200._abi import OcParams, load
    import ctypes
    p = OcParams()
    load().oc_default_params(ctypes.byref(p), 5, 5)
    return np.array(list(p.inv_ellipsoid), np.float32)


def test_tile_dependencies_cover_every_producer():
    """OcSeg2 / OcDep2 host arithmetic, brute force over random launch pairs: whole cloths (same rows), row bands whose
    range shrinks by two rows either side per substep, different segment heights in the two launches, shorter
    segments in the edge strips, one / two / many strips.  Every row is computed exactly once, and a tile waits for
    every tile of the previous launch that wrote what it reads."""
    L = helpers.emu_lib()
    rng = np.random.RandomState(7)
    chained = 0
    for trial in range(3000):
        nstrips = int(rng.choice([1, 2, 3, 4, 17, 34, 67]))
        prows = int(rng.randint(40, 1400))
        pra = int(rng.randint(0, 50))
        prb = pra + prows
        shrink = int(rng.choice([0, 2]))
        ra, rb = pra + shrink, prb - shrink
        prs = int(rng.randint(8, 300)); prs_e = max(8, int(prs / rng.uniform(1.0, 1.4)))
        if rng.rand() < 0.5:
            rs, rs_e = prs, prs_e
        else:
            rs = int(rng.randint(8, 300)); rs_e = max(8, int(rs / rng.uniform(1.0, 1.4)))
        if trial % 3 == 0:                       # what the planner does: a fixed number of segments, height = ceil(rows / nseg)
            nseg = int(rng.randint(1, 45))
            prs = prs_e = max(8, -(-prows // nseg)); rs = rs_e = max(8, -(-(rb - ra) // nseg))
        rc = L.emu_check_tiling(nstrips, pra, prb, prs, prs_e, ra, rb, rs, rs_e, 1)
        assert rc >= 0, f"violation {rc}: strips {nstrips}, prev [{pra},{prb}) {prs}|{prs_e}, now [{ra},{rb}) {rs}|{rs_e}"
        chained += rc == 0
    assert chained > 1000


def test_bandres_plan_cuts_every_cloth_into_bands_that_fit():
    """Host logic of kernel 8 (oc_k_bandres; the kernel itself needs concurrently running CTAs and is covered on the GPU):
    at most one band per SM, the bands tile the rows exactly, every band has at least two rows - so that a band's two halo
    rows come from its direct neighbours - and at most `rmax` rows, the tallest band's state fits one SM's shared memory,
    and a cloth whose bands would not fit is refused (it goes to the marching kernel)."""
    import ctypes
    L = helpers.emu_lib()
    nb, rmax, smem = ctypes.c_int(), ctypes.c_int(), ctypes.c_ulonglong()
    fits = {}
    for U, V in ((21, 21), (64, 48), (37, 23), (100, 61), (256, 256), (300, 200), (513, 301), (100, 1000), (700, 90), (512, 512),
                 (512, 576), (512, 640), (1024, 1024), (2048, 2048), (4096, 16), (8, 4), (5, 3)):
        ok = L.emu_bandres_plan(U, V, 148, ctypes.byref(nb), ctypes.byref(rmax), ctypes.byref(smem))
        fits[(U, V)] = bool(ok)
        if not ok:
            continue
        n = nb.value
        assert 1 <= n <= 148 and n <= max(1, V // 2) and smem.value <= 224 * 1024
        assert n == 1 or (n - 1) * 512 < U * V                      # no band without a particle per thread (but at least one band)
        prev = 0
        for b in range(n):
            r0, r1 = ctypes.c_int(), ctypes.c_int()
            L.emu_bandres_rows(V, n, b, ctypes.byref(r0), ctypes.byref(r1))
            assert r0.value == prev and (n == 1 or r1.value - r0.value >= 2) and r1.value - r0.value <= rmax.value
            prev = r1.value
        assert prev == V
    assert fits[(256, 256)] and fits[(512, 512)] and fits[(512, 576)] and fits[(100, 1000)] and fits[(64, 48)]
    assert not fits[(1024, 1024)] and not fits[(2048, 2048)] and not fits[(4096, 16)] and not fits[(5, 3)]


BANDRES = 8


@pytest.mark.parametrize("nx,ny,pre,calls,T,sms", [(37, 23, 1900, (1, 2, 5, 1, 9), 32, 6), (64, 48, 1700, (7, 1, 12), 64, 5), (20, 66, 900, (3, 4, 1, 1, 6), 32, 8),
                                                   (21, 21, 1800, (25,), 32, 3), (5, 4, 3, (30,), 32, 2), (70, 9, 400, (4, 4), 32, 4)])
def test_bandres_kernel_body_and_exchange(nx, ny, pre, calls, T, sms):
    """Kernel 8 (oc_k_bandres) on the CPU: all the CTAs of the launch run concurrently as fibers of one scheduler, so the
    bands really wait for each other's tagged boundary words inside the launch.  Bitwise against the oracle over several
    oc_step calls (the tags count on from launch to launch; n = 1 writes one buffer, n > 1 two), band heights that do
    not divide the cloth, a single band, cloths as narrow as the stencil."""
    x0, xl0 = helpers.developed_state(nx, ny, pre)
    o = Oracle(nx, ny); o.set_state(x0, xl0)
    e = Emu(nx, ny); e.upload(x0, xl0)
    for n in calls:
        e.step(n, kernel=BANDRES, exact=1, TW=T, RS=sms)
        o.step(n)
        x, xl = e.download()
        ox, oxl = o.state()
        assert bitwise_equal(x, ox) and bitwise_equal(xl, oxl), (nx, ny, n)


@pytest.mark.parametrize("order", [1, 2])
def test_bandres_is_independent_of_thread_and_cta_schedule(order):
    """No race inside a phase and none in the exchange protocol: CTAs and threads resumed in reverse / pseudo-random order."""
    L = helpers.emu_lib()
    x0, xl0 = helpers.developed_state(40, 33, 1700)
    ref = Emu(40, 33); ref.upload(x0, xl0); ref.step(14, kernel=BANDRES, exact=1, TW=32, RS=6)
    rx, rxl = ref.download()
    L.emu_set_order(order)
    try:
        e = Emu(40, 33); e.upload(x0, xl0); e.step(14, kernel=BANDRES, exact=1, TW=32, RS=6)
        x, xl = e.download()
    finally:
        L.emu_set_order(0)
    assert bitwise_equal(x, rx) and bitwise_equal(xl, rxl)
    # tolerance mode: the same bits whatever the number of bands (one arithmetic path)
    a = Emu(40, 33); a.upload(x0, xl0); a.step(10, kernel=BANDRES, exact=0, TW=32, RS=6)
    b = Emu(40, 33); b.upload(x0, xl0); b.step(10, kernel=BANDRES, exact=0, TW=32, RS=2)
    assert bitwise_equal(a.download()[0], b.download()[0])
