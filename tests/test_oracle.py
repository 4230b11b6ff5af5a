"""CPU tier: the oracle restatement (oracle/oc_oracle.c) against the golden vectors produced by the
verbatim reference build, and — where oracle/_ref/libocref.so exists — against that build itself.
Bar: bit-exact (the restatement only re-orders the reference's scatter loop into a gather)."""
import hashlib
import json

import numpy as np
import pytest

import helpers
from helpers import Oracle, bitwise_equal


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()


def golden(name):
    g = helpers.load_golden(name)
    meta = json.loads(bytes(g["meta"]).decode())
    return g, meta


@pytest.mark.parametrize("name", ["grid_21x21.npz", "grid_37x23.npz", "grid_64x64.npz", "grid_256x256.npz"])
def test_oracle_matches_golden(name):
    g, meta = golden(name)
    nx, ny = meta["nx"], meta["ny"]
    o = Oracle(nx, ny)
    step = 0
    traj = []
    ee = meta["energy_every"]
    for cp in meta["checkpoints"]:
        while step < cp:
            o.step(1)
            step += 1
            if step % ee == 0:
                traj.append(o.energy())
        x, xl = o.state()
        assert sha(x) == meta["sha_x"][str(cp)], f"X differs from the reference at step {cp}"
        assert sha(xl) == meta["sha_xl"][str(cp)], f"X_last differs from the reference at step {cp}"
        if meta["full"]:
            assert bitwise_equal(x, g[f"x_{cp}"]) and bitwise_equal(xl, g[f"xl_{cp}"])
        else:
            rows = np.arange(0, ny, meta["row_stride"])
            assert bitwise_equal(x.reshape(ny, nx, 3)[rows], g[f"x_{cp}"])
        assert int((x == xl).all(1).sum()) == meta["hits"][str(cp)]
        assert o.energy() == pytest.approx(meta["energy"][str(cp)], rel=1e-12)
    # spring-energy trajectory, every 10 steps (north star: "matching total spring energy trajectories")
    np.testing.assert_allclose(np.asarray(traj), g["energy_traj"][:len(traj)], rtol=1e-12)


def test_default_golden_values_of_survey():
    """The numbers SURVEY.md 8(c) quotes from the verbatim build (21x21 default, dt = 1/60)."""
    g, meta = golden("grid_21x21.npz")
    x1 = g["x_1"]; x100 = g["x_100"]; x1000 = g["x_1000"]; x2000 = g["x_2000"]
    assert x1[220].tolist() == [0.0, np.float32(4.99999714), 2.0]
    assert x100[220][1] == np.float32(4.98555183)
    np.testing.assert_array_equal(x1000[220], np.array([0.00197683298, 3.74753594, 1.57872772], np.float32))
    np.testing.assert_array_equal(x1000[440], np.array([2.00133777, 3.73311472, 3.58235264], np.float32))
    np.testing.assert_array_equal(x2000[440], np.array([1.90491867, 1.17931032, -0.242099196], np.float32))
    assert meta["energy"]["1000"] == pytest.approx(0.0904504509, rel=1e-8)
    # first contact with the ellipsoid is at step 1672: only the two pinned corners have X == X_last before
    assert meta["hits"]["1671"] == 2 and meta["hits"]["1672"] == 3


@pytest.mark.parametrize("nx,ny", [(21, 21), (37, 23)])
def test_rest_length_tables_match_reference_spring_list(nx, ny):
    """The implicit spring net: count, order-free multiset of (p1,p2), and every rest length (V:134-144, V:286-320)."""
    s = np.load(f"{helpers.GOLDEN}/springs_{nx}x{ny}.npz")
    u, v = nx, ny
    # structural V(U-1)+U(V-1); shear 2(U-1)(V-1); bend V(U-2)+U(V-2) plus one duplicate per row and per column
    assert len(s["p1"]) == 2 * (v * (u - 1) + u * (v - 1)) + 2 * (u - 1) * (v - 1)
    t = Oracle(nx, ny).tables()
    for p1, p2, rest, ty in zip(s["p1"], s["p2"], s["rest"], s["type"]):
        i1, j1, i2, j2 = p1 % u, p1 // u, p2 % u, p2 // u
        di, dj = i2 - i1, j2 - j1
        if ty == 0 and dj == 0: exp = t["rh1"][i1]
        elif ty == 0: exp = t["rv1"][j1]
        elif ty == 2 and dj == 0: exp = t["rh2"][i1]
        elif ty == 2: exp = t["rv2"][j1]
        else:
            ci, cj = min(i1, i2), min(j1, j2)
            exp = np.sqrt(np.float32(t["dx2"][ci] + t["dz2"][cj]), dtype=np.float32)
        assert np.float32(rest).view(np.uint32) == np.float32(exp).view(np.uint32), (p1, p2, ty)


def test_setup_constants_match_reference():
    """Collider matrices (V:324-327), scalars (V:97-104) and the initial sheet (V:254-260)."""
    g = np.load(f"{helpers.GOLDEN}/setup.npz")
    o = Oracle(21, 21)
    assert bitwise_equal(np.ctypeslib.as_array(o.p.ellipsoid), g["ellipsoid"])
    assert bitwise_equal(np.ctypeslib.as_array(o.p.inv_ellipsoid), g["inv_ellipsoid"])
    p = g["params"]
    got = [o.p.damping, o.p.ks_struct, o.p.kd_struct, o.p.ks_shear, o.p.kd_shear, o.p.ks_bend, o.p.kd_bend,
           o.p.gravity[0], o.p.gravity[1], o.p.gravity[2], o.p.mass, o.p.dt, o.p.fullsize, o.p.radius]
    assert bitwise_equal(np.asarray(got, np.float32), p[:14])
    assert bitwise_equal(o.state()[0], g["x0"])
    # the product's defaults are the same numbers (host-only call, no GPU needed)
    from opencloth_b200 import default_params
    q = default_params(21, 21)
    assert bitwise_equal(np.ctypeslib.as_array(q.ellipsoid), g["ellipsoid"])
    assert bitwise_equal(np.ctypeslib.as_array(q.inv_ellipsoid), g["inv_ellipsoid"])
    got = [q.damping, q.ks_struct, q.kd_struct, q.ks_shear, q.kd_shear, q.ks_bend, q.kd_bend,
           q.gravity[0], q.gravity[1], q.gravity[2], q.mass, q.dt, q.fullsize, q.radius]
    assert bitwise_equal(np.asarray(got, np.float32), p[:14])


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref/libocref.so not built (no /root/reference here)")
@pytest.mark.parametrize("nx,ny,steps", [(21, 21, 3000), (37, 23, 2500), (3, 3, 400), (4, 7, 400), (64, 64, 2200), (128, 96, 300)])
def test_oracle_matches_verbatim_reference(nx, ny, steps):
    """Restatement == the reference's own code, every 100 steps, collisions included."""
    r = helpers.Ref(nx, ny)
    o = Oracle(nx, ny)
    done = 0
    while done < steps:
        k = min(100, steps - done)
        r.step(k); o.step(k); done += k
        rx, rxl = r.state(); ox, oxl = o.state()
        assert bitwise_equal(rx, ox) and bitwise_equal(rxl, oxl), f"diverged by step {done}"
    assert o.energy() == r.energy()


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref/libocref.so not built (no /root/reference here)")
def test_oracle_matches_verbatim_reference_at_full_size():
    """The size the bench runs (2048 x 2048, BASELINE config 3): the restatement against the reference's own code with
    its 25 million-entry spring list, 3 steps from the flat sheet and 3 from a rumpled one.  Everything larger rests on
    the same per-particle code, whose only size-dependent inputs are the rest-length tables checked here as well."""
    n = 2048
    r = helpers.Ref(n, n)
    o = Oracle(n, n)
    r.step(3); o.step(3)
    rx, rl = r.state(); ox, ol = o.state()
    assert bitwise_equal(rx, ox) and bitwise_equal(rl, ol)
    rng = np.random.RandomState(5)
    x = rx.copy(); x += (2e-3 * rng.uniform(-1, 1, x.shape)).astype(np.float32)
    r.set_state(x, rx); o.set_state(x, rx)
    r.step(3); o.step(3)
    rx, rl = r.state(); ox, ol = o.state()
    assert bitwise_equal(rx, ox) and bitwise_equal(rl, ol)
    # every rest length of the reference's list against the six 1-D tables (sampled: the list has 25 149 442 entries)
    s = r.springs()
    assert len(s["p1"]) == 2 * (n * (n - 1) + n * (n - 1)) + 2 * (n - 1) * (n - 1)
    t = o.tables()
    p1, p2, rest, ty = s["p1"], s["p2"], s["rest"], s["type"]
    i1, j1, i2, j2 = p1 % n, p1 // n, p2 % n, p2 // n
    exp = np.empty(len(p1), np.float32)
    h = (j1 == j2)
    m = (ty == 0) & h;  exp[m] = t["rh1"][i1[m]]
    m = (ty == 0) & ~h; exp[m] = t["rv1"][j1[m]]
    m = (ty == 2) & h;  exp[m] = t["rh2"][i1[m]]
    m = (ty == 2) & ~h; exp[m] = t["rv2"][j1[m]]
    m = (ty == 1)
    exp[m] = np.sqrt(t["dx2"][np.minimum(i1[m], i2[m])] + t["dz2"][np.minimum(j1[m], j2[m])], dtype=np.float32)
    assert (exp.view(np.uint32) == rest.view(np.uint32)).all()


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref/libocref.so not built")
def test_oracle_matches_reference_from_perturbed_state():
    """Batched-mode style start: the sheet with a per-particle y perturbation, velocities zero."""
    rng = np.random.RandomState(1234)
    r = helpers.Ref(33, 29); o = Oracle(33, 29)
    x, _ = o.state()
    x[:, 1] += (1e-3 * rng.uniform(-1, 1, len(x))).astype(np.float32)
    r.set_state(x, x); o.set_state(x, x)
    r.step(500); o.step(500)
    assert bitwise_equal(r.state()[0], o.state()[0]) and bitwise_equal(r.state()[1], o.state()[1])


def test_row_range_stepping_is_consistent():
    """oco_step_rows (used by the band tests) restricted to a row range equals the full step on those rows."""
    x0, xl0 = helpers.developed_state(24, 30, 700)
    full = Oracle(24, 30); full.set_state(x0, xl0); full.step(1)
    part = Oracle(24, 30); part.set_state(x0, xl0); part.step_rows(7, 19)
    fx, fxl = full.state(); px, pxl = part.state()
    sl = slice(7 * 24, 19 * 24)
    assert bitwise_equal(fx[sl], px[sl]) and bitwise_equal(fxl[sl], pxl[sl])
    assert bitwise_equal(px[:7 * 24], x0[:7 * 24]) and bitwise_equal(px[19 * 24:], x0[19 * 24:])
