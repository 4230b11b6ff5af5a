"""The reference's explicit-integrator siblings of the Verlet step and its Provot pass.

  EULER / SEMI   OpenCloth_ExplicitEuler / OpenCloth_SemiImplicit main.cpp: state (X, V), same spring net and force.
  provot         ApplyProvotDynamicInverse after EllipsoidCollision (V:486-508; E:554-577 / S:437-462).

CPU tier: the oracle restatement against the golden vectors of the verbatim builds (tests/golden/make_golden.py
variants) and, where oracle/_ref exists, against those builds themselves; the product's kernel bodies in the CPU
emulator against the oracle.  GPU tier (-m gpu): the CUDA path through the C-ABI against the golden vectors and the
oracle.  Bar everywhere: bit-exact.
"""
import hashlib
import json

import numpy as np
import pytest

import helpers
from helpers import EULER, SEMI, VERLET, Emu, Oracle, bitwise_equal

VARIANTS = {"euler": (EULER, 1), "semi": (SEMI, 1), "euler_noprovot": (EULER, 0), "semi_noprovot": (SEMI, 0), "verlet_provot": (VERLET, 1)}
FIXTURES = [f"{v}_{g}.npz" for v in VARIANTS for g in ("21x21", "64x64")] + ["euler_256x256.npz", "verlet_provot_256x256.npz"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()


def golden(name):
    g = helpers.load_golden(name)
    return g, json.loads(bytes(g["meta"]).decode())


def check_against_golden(sim_step, sim_state, g, meta):
    nx, ny = meta["nx"], meta["ny"]
    step = 0
    for cp in meta["checkpoints"]:
        sim_step(cp - step)
        step = cp
        x, xl = sim_state()
        assert sha(x) == meta["sha_x"][str(cp)], f"X differs from the reference at step {cp}"
        assert sha(xl) == meta["sha_xl"][str(cp)], f"V / X_last differs from the reference at step {cp}"
        if meta["full"]:
            assert bitwise_equal(x, g[f"x_{cp}"]) and bitwise_equal(xl, g[f"xl_{cp}"])
        else:
            rows = np.arange(0, ny, meta["row_stride"])
            assert bitwise_equal(x.reshape(ny, nx, 3)[rows], g[f"x_{cp}"])


# ---------------------------------------------------------------------------------------------
# CPU tier
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_variant_golden(name):
    g, meta = golden(name)
    integ, provot = VARIANTS[meta["variant"]]
    o = Oracle(meta["nx"], meta["ny"], integ, provot=provot)
    check_against_golden(o.step, o.state, g, meta)


@pytest.mark.skipif(not (helpers.have_ref() and helpers.have_ref_variant(EULER)), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("nx,ny,steps", [(21, 21, 3000), (37, 23, 2500), (3, 3, 400), (4, 7, 400), (64, 64, 1200)])
def test_oracle_matches_verbatim_variant(variant, nx, ny, steps):
    integ, provot = VARIANTS[variant]
    if integ == VERLET:
        r = helpers.Ref(nx, ny)
        rstep = r.step_provot
    else:
        r = helpers.RefVariant(integ, nx, ny, provot)
        rstep = r.step
    o = Oracle(nx, ny, integ, provot=provot)
    done = 0
    while done < steps:
        k = min(200, steps - done)
        rstep(k); o.step(k); done += k
        rx, rs = r.state(); ox, os_ = o.state()
        assert bitwise_equal(rx, ox) and bitwise_equal(rs, os_), f"diverged by step {done}"


@pytest.mark.skipif(not helpers.have_ref_variant(EULER), reason="oracle/_ref not built")
def test_variant_defaults_match_reference_globals():
    """oc_default_params_for == the globals of the sibling demos (E:97-102, S:80-85)."""
    from opencloth_b200 import default_params
    for integ in (EULER, SEMI):
        p = helpers.RefVariant(integ, 21, 21).params()
        q = default_params(21, 21, integrator=integ)
        got = [q.damping, q.ks_struct, q.kd_struct, q.ks_shear, q.kd_shear, q.ks_bend, q.kd_bend,
               q.gravity[0], q.gravity[1], q.gravity[2], q.mass, q.dt, q.fullsize, q.radius]
        assert bitwise_equal(np.asarray(got, np.float32), p[:14])
        assert q.provot == 1 and q.integrator == integ
        o = Oracle(21, 21, integ)
        got = [o.p.damping, o.p.ks_struct, o.p.kd_struct, o.p.ks_shear, o.p.kd_shear, o.p.ks_bend, o.p.kd_bend,
               o.p.gravity[0], o.p.gravity[1], o.p.gravity[2], o.p.mass, o.p.dt, o.p.fullsize, o.p.radius]
        assert bitwise_equal(np.asarray(got, np.float32), p[:14])


def _variant_overrides(integ, provot):
    o = Oracle(5, 5, integ)
    return dict(integrator=integ, provot=provot, ks_struct=o.p.ks_struct, ks_shear=o.p.ks_shear, ks_bend=o.p.ks_bend, mass=o.p.mass)


def _contact_state(nx, ny, integ):
    """A state in collider contact: developed WITHOUT the Provot pass (with it the sheet hangs above the ellipsoid)."""
    o = Oracle(nx, ny, integ, provot=0)
    o.step(1900 if integ == VERLET else 2000)
    s = o.state()
    o.close()
    return s


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("nx,ny,steps,threads", [(21, 21, 40, 32), (9, 13, 30, 4), (3, 3, 30, 32), (37, 23, 12, 8)])
def test_emulated_kernel_bodies_match_oracle(variant, nx, ny, steps, threads):
    """The library's kernel bodies (gather with (X, V) state, Provot gather into V, Provot sweep of X by rows, columns and
    the skewed shear wavefront) on the CPU emulator, under three thread schedules, from a state in collider contact."""
    integ, provot = VARIANTS[variant]
    x0, s0 = _contact_state(nx, ny, integ)
    o = Oracle(nx, ny, integ, provot=provot); o.set_state(x0, s0); o.step(steps)
    ox, os_ = o.state()
    L = helpers.emu_lib()
    try:
        for order in ((0, 1, 2) if (integ == VERLET and nx > 3) else (0,)):
            L.emu_set_order(order)
            L.emu_set_provot_threads(threads)
            e = Emu(nx, ny, **_variant_overrides(integ, provot))
            e.upload(x0, s0)
            e.step(steps, kernel=3 if integ == VERLET else 1, TW=16, RS=7)
            ex, es = e.download()
            assert bitwise_equal(ex, ox) and bitwise_equal(es, os_), f"schedule {order}"
    finally:
        L.emu_set_order(0)
        L.emu_set_provot_threads(32)


# ---------------------------------------------------------------------------------------------
# GPU tier
# ---------------------------------------------------------------------------------------------
def oc():
    import opencloth_b200
    return opencloth_b200


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_matches_variant_golden(name):
    g, meta = golden(name)
    integ, provot = VARIANTS[meta["variant"]]
    m = oc()
    c = m.Cloth(meta["nx"], meta["ny"], integrator=integ, provot=provot)
    check_against_golden(c.step, c.download, g, meta)
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("nx,ny,steps", [(3, 3, 60), (4, 7, 60), (21, 21, 300), (100, 61, 120), (300, 200, 40), (1300, 1100, 6)])
def test_cuda_variants_match_oracle_bitwise(variant, nx, ny, steps):
    integ, provot = VARIANTS[variant]
    m = oc()
    x0, s0 = _contact_state(nx, ny, integ) if nx * ny <= 60000 else Oracle(nx, ny, integ).state()
    o = Oracle(nx, ny, integ, provot=provot); o.set_state(x0, s0); o.step(steps)
    ox, os_ = o.state()
    for kernel in ((m.OC_KERNEL_AUTO, m.OC_KERNEL_GATHER, m.OC_KERNEL_MARCH) if integ == VERLET else (m.OC_KERNEL_AUTO,)):
        c = m.Cloth(nx, ny, integrator=integ, provot=provot, kernel=kernel)
        c.upload(x0, s0)
        c.step(steps)
        x, s = c.download()
        assert bitwise_equal(x, ox), f"kernel {kernel}: {int((helpers.bits(x) != helpers.bits(ox)).any(1).sum())} particles differ"
        assert bitwise_equal(s, os_), f"kernel {kernel}: V / X_last differs"
        c.close()


@pytest.mark.gpu
def test_cuda_variants_batched_and_runtime_provot_switch():
    """A batch of cloths under the Euler integrator, Provot toggled at run time with oc_set_params; set_particle zeroes V."""
    m = oc()
    nx, ny, B = 33, 29, 3
    o = [Oracle(nx, ny, EULER) for _ in range(B)]
    c = m.Cloth(nx, ny, batch=B, integrator=EULER)
    rng = np.random.RandomState(7)
    xs, vs = [], []
    for b in range(B):
        x, v = o[b].state()
        x[:, 1] += (1e-3 * rng.uniform(-1, 1, len(x))).astype(np.float32)
        o[b].set_state(x, v); xs.append(x); vs.append(v)
    c.upload(np.concatenate(xs), np.concatenate(vs))
    for b in range(B):
        o[b].step(50)
    c.step(50)
    c.set_params(provot=0)
    for b in range(B):
        o[b].set_params(provot=0); o[b].step(30)
    c.step(30)
    c.set_particle(5 * nx + 7, (0.1, 4.5, 0.9), cloth=1)
    x1, v1 = o[1].state(); x1[5 * nx + 7] = (0.1, 4.5, 0.9); v1[5 * nx + 7] = 0; o[1].set_state(x1, v1)
    c.set_params(provot=1)
    for b in range(B):
        o[b].set_params(provot=1); o[b].step(20)
    c.step(20)
    x, v = c.download()
    n = nx * ny
    for b in range(B):
        ox, ov = o[b].state()
        assert bitwise_equal(x[b * n:(b + 1) * n], ox) and bitwise_equal(v[b * n:(b + 1) * n], ov), f"cloth {b}"
    c.close()


@pytest.mark.gpu
def test_variants_need_whole_cloth_handles():
    m = oc()
    with pytest.raises(m.OpenClothError):
        m.Cloth(64, 64, row_begin=0, row_end=32, halo_rows=2, integrator=EULER)
    with pytest.raises(m.OpenClothError):
        m.Cloth(64, 64, row_begin=0, row_end=32, halo_rows=2, provot=1)
    c = m.Cloth(21, 21)
    with pytest.raises(m.OpenClothError):
        c.set_params(integrator=EULER)          # fixed at create
    c.close()
