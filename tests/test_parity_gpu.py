"""GPU tier (-m gpu): the CUDA path, called through the C-ABI, against the oracle and the committed
golden vectors.  Bar: BIT-EXACT in exact mode (the kernels evaluate the reference's fp32 operations
in the reference's order); in fast mode the north-star tolerance: max position error <= 1e-5 of the
cloth extent (fullsize = 4) after 100 steps and <= 1e-3 after 1000 steps."""
import hashlib
import json

import numpy as np
import pytest

import helpers
from helpers import Oracle, bitwise_equal

pytestmark = pytest.mark.gpu

EXTENT = 4.0


def oc():
    import opencloth_b200
    return opencloth_b200


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()


def golden(name):
    g = helpers.load_golden(name)
    return g, json.loads(bytes(g["meta"]).decode())


def kid(m, kernel):
    return {"march": m.OC_KERNEL_MARCH, "gather": m.OC_KERNEL_GATHER, "march2": m.OC_KERNEL_MARCH2, "resident": m.OC_KERNEL_RESIDENT, "twin": m.OC_KERNEL_TWIN, "stream": m.OC_KERNEL_STREAM, "stream2": m.OC_KERNEL_STREAM2, "bandres": m.OC_KERNEL_BANDRES}[kernel]


def nbad(a, b):
    return int((helpers.bits(a) != helpers.bits(b)).any(1).sum())


# ---------------------------------------------------------------------------------------------
# golden vectors of the verbatim reference
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["grid_21x21.npz", "grid_37x23.npz", "grid_64x64.npz", "grid_256x256.npz"])
@pytest.mark.parametrize("kernel,k", [("march", 1), ("march", 4), ("gather", 1), ("march2", 1), ("resident", 1), ("twin", 1), ("stream", 1), ("stream2", 1), ("bandres", 1)])
def test_cuda_matches_reference_golden(name, kernel, k):
    g, meta = golden(name)
    nx, ny = meta["nx"], meta["ny"]
    m = oc()
    c = m.Cloth(nx, ny, kernel=kid(m, kernel), substeps_per_launch=k)
    step = 0
    for cp in meta["checkpoints"]:
        c.step(cp - step)
        step = cp
        x, xl = c.download()
        assert sha(x) == meta["sha_x"][str(cp)], f"X differs from the reference at step {cp}"
        assert sha(xl) == meta["sha_xl"][str(cp)], f"X_last differs from the reference at step {cp}"
        assert int((x == xl).all(1).sum()) == meta["hits"][str(cp)]
        assert c.spring_energy() == pytest.approx(meta["energy"][str(cp)], rel=1e-9)
    c.close()


def test_energy_trajectory_matches_reference():
    """Spring energy every 10 steps over 1000 steps of the 256x256 cloth (north star)."""
    g, meta = golden("grid_256x256.npz")
    m = oc()
    c = m.Cloth(256, 256, substeps_per_launch=2)
    e = []
    for _ in range(100):
        c.step(10)
        e.append(c.spring_energy())
    np.testing.assert_allclose(np.asarray(e), g["energy_traj"][:100], rtol=1e-9)
    c.close()


# ---------------------------------------------------------------------------------------------
# oracle on the same seeded inputs
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nx,ny,pre,steps", [(3, 3, 0, 300), (4, 7, 0, 300), (21, 21, 1650, 400), (37, 23, 1800, 300),
                                             (100, 61, 1500, 200), (129, 40, 800, 100), (300, 200, 600, 100), (1000, 37, 300, 50)])
@pytest.mark.parametrize("kernel,k", [("march", 1), ("march", 2), ("march", 8), ("gather", 1), ("march2", 1), ("resident", 1), ("twin", 1), ("stream", 1), ("stream2", 1), ("bandres", 1)])
def test_cuda_matches_oracle_bitwise(nx, ny, pre, steps, kernel, k):
    m = oc()
    x0, xl0 = helpers.developed_state(nx, ny, pre)
    o = Oracle(nx, ny); o.set_state(x0, xl0); o.step(steps)
    c = m.Cloth(nx, ny, kernel=kid(m, kernel), substeps_per_launch=k)
    c.upload(x0, xl0)
    c.step(steps)
    x, xl = c.download()
    ox, oxl = o.state()
    assert bitwise_equal(x, ox), f"{nbad(x, ox)} particles differ"
    assert bitwise_equal(xl, oxl)
    c.close()


def test_resident_kernel_small_cloths():
    """Kernel 4 (state resident in shared memory, all substeps of a call in one launch): what AUTO picks for the
    reference's own 21 x 21 cloth.  One launch per oc_step call whatever n; odd call patterns (n = 1 writes one
    buffer, n > 1 two), a particle edit in between, a batch of cloths, the largest size (39 x 39) and the
    fall-back for cloths that do not fit."""
    m = oc()
    c = m.Cloth(21, 21)                                      # AUTO
    l0 = c.launch_count
    c.step(1000)
    assert c.launch_count - l0 == 1
    g, meta = golden("grid_21x21.npz")
    x, xl = c.download()
    assert sha(x) == meta["sha_x"]["1000"] and sha(xl) == meta["sha_xl"]["1000"]
    c.close()
    for nx, ny, batch in ((21, 21, 7), (39, 39, 2), (33, 40, 1), (5, 4, 3)):
        x0, xl0 = helpers.developed_state(nx, ny, 900)
        o = Oracle(nx, ny); o.set_state(x0, xl0)
        c = m.Cloth(nx, ny, batch=batch, kernel=m.OC_KERNEL_RESIDENT)
        c.upload(np.tile(x0, (batch, 1)), np.tile(xl0, (batch, 1)))
        idx = (ny // 2) * nx + nx // 2
        for n in (1, 1, 2, 7, 1, 300):
            c.step(n); o.step(n)
            if n == 7:
                c.set_particle(idx, (0.1, 2.5, 0.2), cloth=batch - 1)
                o2 = Oracle(nx, ny); ox, oxl = o.state(); ox = ox.copy(); oxl = oxl.copy()
                ox[idx] = (0.1, 2.5, 0.2); oxl[idx] = (0.1, 2.5, 0.2)
                o2.set_state(ox, oxl)
        o2.step(1 + 300)
        x, xl = c.download()
        ox, oxl = o.state(); px, pxl = o2.state()
        n1 = nx * ny
        for b in range(batch):
            ex, exl = (px, pxl) if b == batch - 1 else (ox, oxl)
            assert bitwise_equal(x[b * n1:(b + 1) * n1], ex) and bitwise_equal(xl[b * n1:(b + 1) * n1], exl), f"{nx}x{ny} cloth {b}"
        c.close()
    big = m.Cloth(100, 61, kernel=m.OC_KERNEL_RESIDENT)      # 6100 particles do not fit one CTA's shared memory: served by oc_k_march2
    big.step(3)
    o = Oracle(100, 61); o.step(3)
    assert bitwise_equal(big.download()[0], o.state()[0])
    big.close()


def test_temporal_blocking_equals_single_steps():
    """k in {1,2,4,8} substeps per launch, and odd step counts that split into mixed launches."""
    m = oc()
    nx, ny = 517, 260
    x0, xl0 = helpers.developed_state(nx, ny, 400)
    outs = []
    for k in (1, 2, 4, 8, 3, 7):
        c = m.Cloth(nx, ny, kernel=m.OC_KERNEL_MARCH, substeps_per_launch=k)
        c.upload(x0, xl0)
        c.step(29)
        outs.append(c.download())
        c.close()
    for x, xl in outs[1:]:
        assert bitwise_equal(x, outs[0][0]) and bitwise_equal(xl, outs[0][1])


def test_large_grid_matches_oracle_2048():
    """BASELINE config 3 size: 2048 x 2048, 20 steps, bitwise, plus the stride-4 host layout."""
    m = oc()
    n = 2048
    o = Oracle(n, n); o.step(20)
    ox, oxl = o.state()
    for k, kern in ((1, m.OC_KERNEL_MARCH), (4, m.OC_KERNEL_MARCH), (1, m.OC_KERNEL_MARCH2)):
        c = m.Cloth(n, n, substeps_per_launch=k, kernel=kern)
        c.step(20)
        x, xl = c.download()
        assert bitwise_equal(x, ox) and bitwise_equal(xl, oxl), f"k={k}: {nbad(x, ox)} particles differ"
        x4, xl4 = c.download(stride=4)
        assert bitwise_equal(x4[:, :3], ox) and (x4[:, 3] == 1.0).all()
        c.close()


@pytest.mark.parametrize("nx,ny,steps", [(512, 512, 2400), (256, 256, 2400), (513, 301, 900), (100, 1000, 600), (700, 90, 600), (64, 40, 600)])
def test_bandres_kernel_mid_size_cloths(nx, ny, steps):
    """Kernel 8 (oc_k_bandres: one row band per CTA resident in shared memory, cooperative launch, two boundary rows
    per band and substep through global memory): what AUTO picks between the one-CTA resident kernel and the marching
    kernels.  One launch per oc_step call whatever n; bitwise equal to the gather kernel (plain launches, one thread
    per particle) over thousands of substeps in odd call patterns - n = 1 writes one buffer, n > 1 two; the flag words
    count on from launch to launch - across a particle edit, an upload and a change of time step; band heights that do
    not divide the cloth; no flag wait may time out."""
    import ctypes
    m = oc()
    a = m.Cloth(nx, ny)                                       # AUTO
    g = m.Cloth(nx, ny, kernel=m.OC_KERNEL_GATHER)
    l0 = a.launch_count
    a.step(5); g.step(5)
    assert a.launch_count - l0 == 1, "AUTO did not pick the band-resident kernel"
    done = 5
    plan = [1, 2, 7, 50, 3, 1, 1, 200]
    i = 0
    while done < steps:
        n = min(plan[i % len(plan)] if i < 24 else 400, steps - done)
        a.step(n); g.step(n); done += n
        if i % 4 == 1:
            xa, xla = a.download(); xg, xlg = g.download()
            assert bitwise_equal(xa, xg) and bitwise_equal(xla, xlg), f"after {done} steps: {nbad(xa, xg)} particles differ"
        if i == 5:
            for c in (a, g):
                c.set_particle((ny // 2) * nx + nx // 3, (0.25, 3.0, -0.5))
        if i == 9:
            xa, xla = a.download()
            a.upload(xa, xla); g.upload(xa, xla)
        if i == 13:
            for c in (a, g):
                c.set_params(dt=1.0 / 90.0)
        i += 1
    xa, xla = a.download(); xg, xlg = g.download()
    assert bitwise_equal(xa, xg) and bitwise_equal(xla, xlg), f"{nbad(xa, xg)} particles differ"
    out = (ctypes.c_ulonglong * 4)()
    a._lib.oc_debug_counters(a._h, out)
    assert (out[2] >> 40) == 0, "a flag wait timed out"
    a.close(); g.close()


def test_bandres_kernel_fast_mode_and_fallbacks():
    """Tolerance mode of the band-resident kernel against the reference CPU path (north-star bound), and the cases it
    hands to oc_k_march2: batches, cloths whose bands do not fit shared memory."""
    m = oc()
    nx = ny = 256
    o = Oracle(nx, ny); o.step(100)
    c = m.Cloth(nx, ny, kernel=m.OC_KERNEL_BANDRES, exact=0)
    c.step(100)
    err = np.abs(c.download()[0].astype(np.float64) - o.state()[0].astype(np.float64)).max()
    assert err <= 1e-5 * EXTENT, err
    c.close()
    for kw in (dict(nx=64, ny=48, batch=3), dict(nx=1024, ny=1024, batch=1)):
        b = m.Cloth(kw["nx"], kw["ny"], batch=kw["batch"], kernel=m.OC_KERNEL_BANDRES)
        gk = m.Cloth(kw["nx"], kw["ny"], batch=kw["batch"], kernel=m.OC_KERNEL_GATHER)
        l0 = b.launch_count
        b.step(3); gk.step(3)
        assert b.launch_count - l0 == 3                    # one launch per substep: the marching kernel
        assert bitwise_equal(b.download()[0], gk.download()[0])
        b.close(); gk.close()


@pytest.mark.parametrize("nx,ny,steps,kernel", [(2048, 2048, 3000, "march2"), (4096, 1200, 600, "march2"), (1100, 5000, 600, "march2"),
                                                (2048, 2048, 2400, "twin"), (1100, 5000, 600, "twin"), (2048, 2048, 2400, "stream"), (1100, 5000, 600, "stream"), (1100, 5000, 600, "stream2")])
def test_chained_launches_match_gather_kernel_full_size(nx, ny, steps, kernel):
    """Consecutive oc_k_march2 launches are chained by programmatic dependent launch and per-tile flags instead
    of a barrier between steps (OcDep2; active for tiles of >= 32 rows, i.e. only at full size).  The gather
    kernel (one thread per particle, plain launches) is the independent on-device reference: bitwise equal
    after thousands of chained steps, across everything that breaks and restarts the chain (download,
    oc_set_particle, upload, a change of time step), and no dependency wait may ever time out."""
    import ctypes
    m = oc()
    a = m.Cloth(nx, ny, kernel=kid(m, kernel))
    g = m.Cloth(nx, ny, kernel=m.OC_KERNEL_GATHER)
    done = 0
    plan = [1, 2, 7, 50, 3, 1, 1, 200]
    i = 0
    while done < steps:
        n = min(plan[i % len(plan)] if i < 24 else 400, steps - done)
        a.step(n); g.step(n); done += n
        if i % 4 == 1:
            xa, xla = a.download(); xg, xlg = g.download()
            assert bitwise_equal(xa, xg) and bitwise_equal(xla, xlg), f"after {done} steps: {nbad(xa, xg)} particles differ"
        if i == 5:
            for c in (a, g):
                c.set_particle((ny // 2) * nx + nx // 3, (0.25, 3.0, -0.5))
        if i == 9:
            xa, xla = a.download()
            a.upload(xa, xla); g.upload(xa, xla)
        if i == 13:
            for c in (a, g):
                c.set_params(dt=1.0 / 90.0)
        i += 1
    xa, xla = a.download(); xg, xlg = g.download()
    assert bitwise_equal(xa, xg) and bitwise_equal(xla, xlg), f"{nbad(xa, xg)} particles differ"
    out = (ctypes.c_ulonglong * 4)()
    a._lib.oc_debug_counters(a._h, out)
    assert (out[2] >> 40) == 0, "a tile-dependency wait timed out"
    a.close(); g.close()


@pytest.mark.parametrize("kernel,batch", [("march2", 96), ("twin", 96), ("twin", 33), ("stream", 96), ("stream", 33), ("stream2", 96)])
def test_chained_launches_batched_cloths_match_gather_kernel(kernel, batch):
    """The same for a batch (BASELINE config 5 shape): the tile flags are per cloth, every cloth one 128-row tile
    (oc_k_twin: a CTA takes the same tile of two cloths, or with an odd batch two 64-row tiles of one cloth)."""
    import ctypes
    m = oc()
    a = m.Cloth(128, 128, batch=batch, kernel=kid(m, kernel))
    g = m.Cloth(128, 128, batch=batch, kernel=m.OC_KERNEL_GATHER)
    for c in (a, g):
        c.set_particle(40 * 128 + 17, (0.5, 2.0, 0.25), cloth=5)
    for n in (1, 3, 300, 1, 1500):
        a.step(n); g.step(n)
        xa, xla = a.download(); xg, xlg = g.download()
        assert bitwise_equal(xa, xg) and bitwise_equal(xla, xlg), f"{nbad(xa, xg)} particles differ"
    out = (ctypes.c_ulonglong * 4)()
    a._lib.oc_debug_counters(a._h, out)
    assert (out[2] >> 40) == 0, "a tile-dependency wait timed out"
    a.close(); g.close()


def test_upload_download_round_trip_and_strides():
    m = oc()
    rng = np.random.RandomState(0)
    c = m.Cloth(50, 30)
    x = rng.randn(1500, 3).astype(np.float32); xl = rng.randn(1500, 3).astype(np.float32)
    c.upload(x, xl)
    gx, gxl = c.download()
    assert bitwise_equal(gx, x) and bitwise_equal(gxl, xl)
    x4 = np.concatenate([x, np.ones((1500, 1), np.float32)], 1); xl4 = np.concatenate([xl, np.ones((1500, 1), np.float32)], 1)
    c.upload(x4, xl4)
    gx, gxl = c.download(stride=4)
    assert bitwise_equal(gx, x4) and bitwise_equal(gxl, xl4)
    c.close()


# ---------------------------------------------------------------------------------------------
# fast mode: tolerance of the north star
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nx,ny", [(21, 21), (256, 256)])
@pytest.mark.parametrize("kernel,k", [("march", 4), ("march2", 1), ("twin", 1), ("stream", 1)])
def test_fast_mode_within_tolerance(nx, ny, kernel, k):
    """Tolerance mode of every marching kernel (oc_k_twin's fast arithmetic differs from the others': X - X_last
    instead of V in the ring, forces summed in production order, s = nks + rinv (kd dot rinv - nks rest))."""
    m = oc()
    o = Oracle(nx, ny)
    c = m.Cloth(nx, ny, exact=0, substeps_per_launch=k, kernel=kid(m, kernel))
    o.step(100); c.step(100)
    err100 = np.abs(c.download()[0].astype(np.float64) - o.state()[0]).max() / EXTENT
    assert err100 <= 1e-5, f"fast mode: {err100:.3e} of extent after 100 steps (tolerance 1e-5)"
    o.step(900); c.step(900)
    err1000 = np.abs(c.download()[0].astype(np.float64) - o.state()[0]).max() / EXTENT
    assert err1000 <= 1e-3, f"fast mode: {err1000:.3e} of extent after 1000 steps (tolerance 1e-3)"
    e_o, e_c = o.energy(), c.spring_energy()
    assert e_c == pytest.approx(e_o, rel=2e-2)
    c.close()


def test_fast_mode_within_tolerance_at_bench_size():
    """The tolerance mode at the size bench.py runs (2048 x 2048, AUTO = oc_k_stream) against the oracle (OpenMP restatement,
    pinned to the verbatim reference at this size by tests/test_oracle.py): north-star bound 1e-5 of the extent after 100
    steps."""
    m = oc()
    n = 2048
    o = Oracle(n, n)
    c = m.Cloth(n, n, exact=0)
    o.step(100); c.step(100)
    err = np.abs(c.download()[0].astype(np.float64) - o.state()[0]).max() / EXTENT
    assert err <= 1e-5, f"fast mode at 2048^2: {err:.3e} of extent after 100 steps (tolerance 1e-5)"
    c.close(); o.close()


def test_fast_mode_linked_bands_equal_whole_cloth_bitwise():
    """Tolerance mode of oc_k_stream (AUTO's choice for large handles): the kernel rounds alike on its steady and generic
    paths, so three linked row bands and a batch of two give exactly the bits of the whole single cloth."""
    m = oc()
    nx, ny, steps = 700, 384, 80
    K = m.OC_KERNEL_STREAM
    whole = m.Cloth(nx, ny, exact=0, kernel=K)
    whole.step(40)
    wx, wxl = whole.download()
    cuts = [0, 128, 256, 384]
    bands = []
    for b in range(3):
        c = m.Cloth(nx, ny, row_begin=cuts[b], row_end=cuts[b + 1], halo_rows=2, exact=0, kernel=K)
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        c.upload(wx[sl], wxl[sl])
        bands.append(c)
    m.link_bands_local(bands)
    pair = m.Cloth(nx, ny, batch=2, exact=0, kernel=K)
    pair.upload(np.concatenate([wx, wx]), np.concatenate([wxl, wxl]))
    for s in range(steps):
        for c in bands:
            c.step(1)
    whole.step(steps); pair.step(steps)
    wx, wxl = whole.download()
    for b, c in enumerate(bands):
        x, xl = c.download()
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        assert bitwise_equal(x, wx[sl]) and bitwise_equal(xl, wxl[sl]), f"band {b}: {nbad(x, wx[sl])} particles differ"
        c.close()
    px, pxl = pair.download()
    n = nx * ny
    assert bitwise_equal(px[:n], wx) and bitwise_equal(px[n:], wx) and bitwise_equal(pxl[n:], wxl)
    # oc_k_stream2 (two columns per thread) does the same arithmetic per particle in the same order
    two = m.Cloth(nx, ny, exact=0, kernel=m.OC_KERNEL_STREAM2)
    two.step(40 + steps)
    tx, txl = two.download()
    assert bitwise_equal(tx, wx) and bitwise_equal(txl, wxl)
    whole.close(); pair.close(); two.close()


# ---------------------------------------------------------------------------------------------
# parameters, interaction, batches, bands
# ---------------------------------------------------------------------------------------------
def test_set_params_and_set_particle_follow_the_oracle():
    m = oc()
    nx, ny = 40, 40
    o = Oracle(nx, ny); c = m.Cloth(nx, ny, substeps_per_launch=2)
    o.step(50); c.step(50)
    kw = dict(ks_struct=80.5, kd_shear=-0.5, damping=-0.05, gravity=(0.001, -0.02, 0.0005), mass=1.5, dt=1 / 90.0, radius=1.25)
    o.set_params(**kw); c.set_params(**kw)
    o.step(60); c.step(60)
    # mouse drag (V:203-208): X[idx] = X_last[idx] = p
    idx, p = 20 * nx + 20, (0.1, 3.0, 2.2)
    x, xl = o.state(); x[idx] = p; xl[idx] = p; o.set_state(x, xl)
    c.set_particle(idx, p)
    o.step(200); c.step(200)
    x, xl = c.download(); ox, oxl = o.state()
    assert bitwise_equal(x, ox) and bitwise_equal(xl, oxl)
    c.close()


def test_floor_clamp_and_collider_are_exercised():
    """Cloth dropped with strong gravity: reaches the ellipsoid and the floor (y < 0 -> 0, V:440-442)."""
    m = oc()
    nx, ny = 48, 48
    kw = dict(gravity=(0.0, -0.5, 0.0))
    o = Oracle(nx, ny, **kw)
    o.step(1200)
    ox, oxl = o.state()
    assert (ox[:, 1] == 0.0).sum() > 0, "test does not reach the floor"
    assert int((ox == oxl).all(1).sum()) > 10, "test does not reach the collider"
    for kern, k in ((m.OC_KERNEL_MARCH, 4), (m.OC_KERNEL_MARCH2, 1), (m.OC_KERNEL_TWIN, 1), (m.OC_KERNEL_STREAM, 1)):
        c = m.Cloth(nx, ny, substeps_per_launch=k, kernel=kern, **kw)
        c.step(1200)
        x, xl = c.download()
        assert bitwise_equal(x, ox) and bitwise_equal(xl, oxl), f"kernel {kern}"
        c.close()


def test_batched_cloths_match_per_cloth_oracle():
    """BASELINE config 5 shape (128x128 cloths, per-cloth perturbation X.y += 1e-3 u, velocities 0),
    a 24-cloth batch; every cloth checked against its own oracle run."""
    m = oc()
    nx = ny = 128
    B = 24
    base = Oracle(nx, ny).state()[0]
    X0 = np.empty((B, nx * ny, 3), np.float32)
    for b in range(B):
        rng = np.random.RandomState(1234 + b)
        X0[b] = base
        X0[b, :, 1] += (1e-3 * rng.uniform(-1, 1, nx * ny)).astype(np.float32)
    for k, kern in ((1, m.OC_KERNEL_MARCH), (4, m.OC_KERNEL_MARCH), (1, m.OC_KERNEL_MARCH2), (1, m.OC_KERNEL_TWIN), (1, m.OC_KERNEL_STREAM)):
        c = m.Cloth(nx, ny, batch=B, substeps_per_launch=k, kernel=kern)
        c.upload(X0.reshape(-1, 3), X0.reshape(-1, 3))
        c.step(60)
        x, xl = c.download()
        x = x.reshape(B, -1, 3); xl = xl.reshape(B, -1, 3)
        for b in (0, 1, 7, 13, 23):
            o = Oracle(nx, ny); o.set_state(X0[b], X0[b]); o.step(60)
            ox, oxl = o.state()
            assert bitwise_equal(x[b], ox) and bitwise_equal(xl[b], oxl), f"k={k} cloth {b}"
        c.close()


@pytest.mark.parametrize("nbands,halo,k,kern", [(2, 4, 1, 2), (4, 8, 4, 2), (3, 16, 8, 2), (8, 8, 2, 2), (4, 8, 1, 3), (4, 8, 1, 5), (4, 8, 1, 6)])
def test_row_bands_single_process_equal_whole_cloth(nbands, halo, k, kern):
    """SURVEY.md 8(e) on one device: g band handles exchanging halos with oc_halo_exchange
    (device-to-device copies ordered by events) must equal the undivided cloth bitwise."""
    import ctypes
    m = oc()
    from opencloth_b200 import _abi
    nx, ny = 200, 256
    x0, xl0 = helpers.developed_state(nx, ny, 500)
    whole = m.Cloth(nx, ny, substeps_per_launch=k); whole.upload(x0, xl0)
    cuts = [round(ny * b / nbands) for b in range(nbands + 1)]
    bands = []
    for b in range(nbands):
        c = m.Cloth(nx, ny, row_begin=cuts[b], row_end=cuts[b + 1], halo_rows=halo, substeps_per_launch=k, kernel=kern)
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        c.upload(x0[sl], xl0[sl])
        assert c.halo_budget == 0
        bands.append(c)
    arr = (ctypes.c_void_p * nbands)(*[c._h for c in bands])
    per = halo // 2
    total = 0
    for rnd in range(4):
        _abi.check(_abi.load().oc_halo_exchange(arr, nbands))
        n = per if rnd != 2 else max(1, per - 1)
        for c in bands:
            c.step(n)
        total += n
    whole.step(total)
    wx, wxl = whole.download()
    for b, c in enumerate(bands):
        x, xl = c.download()
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        assert bitwise_equal(x, wx[sl]) and bitwise_equal(xl, wxl[sl]), f"band {b}"
        with pytest.raises(m.OpenClothError):
            c.step(per + 1)          # more substeps than the halo allows
        c.close()
    whole.close()


@pytest.mark.parametrize("nx,ny,nbands,halo,kernel", [(2304, 4096, 2, 16, "march2"), (4100, 3000, 3, 24, "march2"), (2304, 4096, 2, 16, "twin"), (2304, 4096, 2, 16, "stream")])
def test_row_bands_chained_full_size(nx, ny, nbands, halo, kernel):
    """Row bands at a size where the launches of a band are chained tile by tile (OcDep2) although the row range,
    and with it the tiling, shrinks with every substep of a group: bitwise equal to the undivided cloth stepped
    by the gather kernel, through several exchanges, and no dependency wait times out."""
    import ctypes
    m = oc()
    from opencloth_b200 import _abi
    whole = m.Cloth(nx, ny, kernel=m.OC_KERNEL_GATHER)
    whole.step(40)                                   # leave the flat start
    wx, wxl = whole.download()
    cuts = [round(ny * b / nbands) for b in range(nbands + 1)]
    bands = []
    for b in range(nbands):
        c = m.Cloth(nx, ny, row_begin=cuts[b], row_end=cuts[b + 1], halo_rows=halo, kernel=kid(m, kernel))
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        c.upload(wx[sl], wxl[sl])
        bands.append(c)
    arr = (ctypes.c_void_p * nbands)(*[c._h for c in bands])
    total = 0
    for rnd in range(3):
        _abi.check(_abi.load().oc_halo_exchange(arr, nbands))
        n = halo // 2 if rnd != 1 else halo // 2 - 3
        for c in bands:
            c.step(n)
        total += n
    whole.step(total)
    wx, wxl = whole.download()
    out = (ctypes.c_ulonglong * 4)()
    for b, c in enumerate(bands):
        x, xl = c.download()
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        assert bitwise_equal(x, wx[sl]) and bitwise_equal(xl, wxl[sl]), f"band {b}: {nbad(x, wx[sl])} particles differ"
        c._lib.oc_debug_counters(c._h, out)
        assert (out[2] >> 40) == 0, "a tile-dependency wait timed out"
        c.close()
    whole.close()


@pytest.mark.parametrize("nx,ny,nbands,steps,kernel", [(512, 384, 3, 60, "auto"), (200, 64, 4, 40, "auto"), (2304, 2048, 2, 120, "auto"), (1100, 1536, 4, 90, "auto"),
                                                       (512, 384, 3, 60, "twin"), (2304, 2048, 2, 120, "twin"), (512, 384, 3, 60, "stream"), (2304, 2048, 2, 120, "stream"), (512, 384, 3, 60, "stream2")])
def test_linked_row_bands_equal_whole_cloth(nx, ny, nbands, steps, kernel):
    """Linked row bands (the multi-GPU path: in-kernel peer stores of the boundary rows + flag words between the
    bands' tiles, no exchange step), here with all bands on ONE device in one process — the same kernel path and the
    same protocol as across GPUs, only the peer pointers are local.  Bitwise equal to the undivided cloth stepped by
    the independent gather kernel, from a developed state through an upload and a resynchronisation."""
    m = oc()
    whole = m.Cloth(nx, ny, kernel=m.OC_KERNEL_GATHER)
    whole.step(40)
    wx, wxl = whole.download()
    cuts = [round(ny * b / nbands) for b in range(nbands + 1)]
    bands = []
    for b in range(nbands):
        c = m.Cloth(nx, ny, row_begin=cuts[b], row_end=cuts[b + 1], halo_rows=2, kernel=m.OC_KERNEL_AUTO if kernel == "auto" else kid(m, kernel))
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        c.upload(wx[sl], wxl[sl])
        bands.append(c)
    m.link_bands_local(bands)
    n0 = bands[0].launch_count
    for s in range(steps):
        for c in (bands if s % 2 == 0 else bands[::-1]):
            c.step(1)
    assert bands[0].launch_count - n0 == steps          # one launch per substep, nothing else on the step path
    whole.step(steps)
    wx, wxl = whole.download()
    for b, c in enumerate(bands):
        x, xl = c.download()
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        assert bitwise_equal(x, wx[sl]) and bitwise_equal(xl, wxl[sl]), f"band {b}: {nbad(x, wx[sl])} particles differ"
    # interaction on linked bands: oc_set_particle on every band (each updates the copies it stores), then step on
    idx = cuts[1] * nx + nx // 2                          # first row of band 1: also a halo row of band 0
    for c in bands:
        c.set_particle(idx, (0.3, 4.2, 1.1))
    whole.set_particle(idx, (0.3, 4.2, 1.1))
    for s in range(7):
        for c in bands:
            c.step(1)
    whole.step(7)
    wx, wxl = whole.download()
    for b, c in enumerate(bands):
        x, xl = c.download()
        sl = slice(cuts[b] * nx, cuts[b + 1] * nx)
        assert bitwise_equal(x, wx[sl]) and bitwise_equal(xl, wxl[sl]), f"after set_particle, band {b}: {nbad(x, wx[sl])} particles differ"
        c.close()
    whole.close()


def test_linked_band_validation():
    m = oc()
    a = m.Cloth(64, 64, row_begin=0, row_end=32, halo_rows=2)
    b = m.Cloth(64, 64, row_begin=32, row_end=64, halo_rows=2)
    w = m.Cloth(64, 64)
    with pytest.raises(m.OpenClothError):
        w.band_endpoint()                                # a whole cloth has no neighbours
    ea, eb = a.band_endpoint(), b.band_endpoint()
    with pytest.raises(m.OpenClothError):
        a.band_link(None, None)                          # rows [0,32) of 64 have a lower neighbour
    with pytest.raises(m.OpenClothError):
        a.band_link(None, ea)                            # not adjacent
    a.step(1)
    a.sync()
    with pytest.raises(m.OpenClothError):
        a.band_link(None, eb)                            # different step counts (buffer rotation)
    for c in (a, b, w):
        c.close()


def test_tile_dependency_timeout_is_an_error():
    """A tile dependency that never arrives must fail the step (OC_ERR_CUDA from oc_sync / oc_download), not compute
    from stale rows.  OC_DEBUG=32 makes the chained tiles wait for a launch that does not exist and give up after
    50 ms; run in a subprocess because the kernel traps (the CUDA context is lost, as after any device fault)."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import opencloth_b200 as m\n"
        "c = m.Cloth(2048, 2048, kernel=m.OC_KERNEL_MARCH2)\n"
        "c.step(3)\n"
        "try:\n"
        "    c.sync()\n"
        "    print('NO ERROR')\n"
        "except m.OpenClothError as e:\n"
        "    print('ERR', e.code, e)\n" % helpers.ROOT)
    import os
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, OC_DEBUG="32"))
    assert "ERR -3" in r.stdout and "timed out" in r.stdout, r.stdout + r.stderr


def test_branch_free_math_is_ieee():
    """The MUFU+FFMA sequences used instead of sqrt.rn / rcp.rn / div.rn (no slow-path branch) give
    the correctly rounded result on 2^30 random operands in their accepted exponent ranges."""
    import ctypes
    from opencloth_b200 import _abi
    bad = ctypes.c_ulonglong(123)
    _abi.check(_abi.load().oc_selftest_math(1 << 30, 20261017, ctypes.byref(bad)))
    assert bad.value == 0, f"{bad.value} mismatches against the IEEE intrinsics"


def test_launch_counter_and_timed_step():
    m = oc()
    c = m.Cloth(256, 256, substeps_per_launch=4)
    n0 = c.launch_count
    ms = c.step_timed(16)
    assert c.launch_count - n0 == 4 and ms > 0
    c.close()
