"""Run-time control of the simulation: pin sets, write-backs, per-cloth transfers — and the
host <-> device pipeline of upload / step / download.

The reference pins particles 0 and numX by literal index tests (V:455, V:479-482, V:498-501) and drags a particle by
writing X and X_last (V:203-208).  oc_set_pins generalises the set; the oracle applies the same rule to the same set
(for the default set it is pinned to the verbatim reference, tests/test_oracle.py).  Bar: bit-exact.
"""
import numpy as np
import pytest

import helpers
from helpers import EULER, VERLET, Emu, Oracle, bitwise_equal

GATHER, MARCH, MARCH2, RESIDENT = 1, 2, 3, 4


def pin_sets(nx, ny):
    return [
        [],                                                        # nothing pinned: the sheet falls freely
        [0, nx - 1, (ny - 1) * nx, ny * nx - 1],                   # four corners
        [0, nx - 1, (ny // 2) * nx + nx // 2, (ny // 2) * nx + nx // 2 + 1, 5 * nx + 3],     # interior rows too
    ]


# ---------------------------------------------------------------------------------------------
# CPU tier: the library's kernel bodies in the emulator
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", [0, 1, 2])
@pytest.mark.parametrize("kernel,k,provot", [(GATHER, 1, 0), (MARCH, 1, 0), (MARCH, 2, 0), (MARCH2, 1, 0), (MARCH2, 1, 1), (RESIDENT, 1, 0)])
def test_emulated_kernels_with_custom_pins(which, kernel, k, provot):
    nx, ny = 23, 19
    pins = pin_sets(nx, ny)[which]
    x0, xl0 = helpers.developed_state(nx, ny, 300)
    o = Oracle(nx, ny, provot=provot); o.set_state(x0, xl0); o.set_pins(pins); o.step(25)
    ox, oxl = o.state()
    e = Emu(nx, ny, provot=provot); e.upload(x0, xl0); e.set_pins(pins)
    e.step(25, kernel=kernel, k=k, TW=16 if kernel != RESIDENT else 0, RS=7 if kernel != RESIDENT else 0)
    ex, exl = e.download()
    assert bitwise_equal(ex, ox) and bitwise_equal(exl, oxl)
    # (a particle pinned while moving keeps coasting: the reference's "pinned" only removes gravity and spring forces)


def test_default_pin_set_is_the_reference_set():
    nx, ny = 21, 21
    a = Oracle(nx, ny); b = Oracle(nx, ny); b.set_pins([0, nx - 1])
    a.step(200); b.step(200)
    assert bitwise_equal(a.state()[0], b.state()[0])


# ---------------------------------------------------------------------------------------------
# GPU tier
# ---------------------------------------------------------------------------------------------
def oc():
    import opencloth_b200
    return opencloth_b200


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny,steps", [(23, 19, 120), (200, 150, 60), (700, 600, 25)])
@pytest.mark.parametrize("which", [0, 1, 2])
def test_cuda_custom_pins_match_oracle(nx, ny, steps, which):
    m = oc()
    pins = pin_sets(nx, ny)[which]
    x0, xl0 = helpers.developed_state(nx, ny, 200)
    o = Oracle(nx, ny); o.set_state(x0, xl0); o.set_pins(pins); o.step(steps)
    ox, oxl = o.state()
    for kernel, k in ((m.OC_KERNEL_MARCH2, 1), (m.OC_KERNEL_TWIN, 1), (m.OC_KERNEL_STREAM, 1), (m.OC_KERNEL_MARCH, 2), (m.OC_KERNEL_GATHER, 1), (m.OC_KERNEL_AUTO, 1), (m.OC_KERNEL_BANDRES, 1)):
        c = m.Cloth(nx, ny, kernel=kernel, substeps_per_launch=k)
        c.upload(x0, xl0)
        c.set_pins(pins)
        c.step(steps)
        x, xl = c.download()
        assert bitwise_equal(x, ox) and bitwise_equal(xl, oxl), f"kernel {kernel}"
        c.close()


@pytest.mark.gpu
def test_cuda_pins_per_cloth_actions_and_single_cloth_transfers():
    """RL-style batch: every cloth has its own pin set, takes its own write-backs in one launch, and single cloths are
    reset / observed on their own; the Provot pass and the Euler integrator honour the pins as well."""
    m = oc()
    nx, ny, B = 40, 33, 4
    for integ, provot in ((VERLET, 0), (VERLET, 1), (EULER, 1)):
        o = [Oracle(nx, ny, integ, provot=provot) for _ in range(B)]
        over = dict(ks_struct=o[0].p.ks_struct, ks_shear=o[0].p.ks_shear, ks_bend=o[0].p.ks_bend, mass=o[0].p.mass)
        c = m.Cloth(nx, ny, batch=B, integrator=integ, provot=provot, **over)
        sets = pin_sets(nx, ny) + [None]
        for b in range(B):
            if sets[b] is not None:
                c.set_pins(sets[b], cloth=b); o[b].set_pins(sets[b])
            else:
                c.set_pins([0, nx - 1], cloth=b)              # the reference's own set, spelled out
        c.step(40)
        for b in range(B):
            o[b].step(40)
        # actions: one particle per cloth moved in one launch
        idx = [7 * nx + 5 + b for b in range(B)]
        xyz = np.array([[0.1 * b, 4.6, 1.0 + 0.1 * b] for b in range(B)], np.float32)
        c.set_particles(list(range(B)), idx, xyz)
        for b in range(B):
            x, s = o[b].state()
            x[idx[b]] = xyz[b]; s[idx[b]] = xyz[b] if integ == VERLET else 0
            o[b].set_state(x, s)
        c.step(15)
        for b in range(B):
            o[b].step(15)
        # reset environment 2 from the host, observe environment 1 alone
        x2, s2 = Oracle(nx, ny, integ, provot=provot).state()
        c.upload_cloth(2, x2, s2); o[2].set_state(x2, s2)
        c.step(10)
        for b in range(B):
            o[b].step(10)
        x1, s1 = c.download_cloth(1)
        assert bitwise_equal(x1, o[1].state()[0]) and bitwise_equal(s1, o[1].state()[1])
        x, s = c.download()
        n = nx * ny
        for b in range(B):
            ox, os_ = o[b].state()
            assert bitwise_equal(x[b * n:(b + 1) * n], ox) and bitwise_equal(s[b * n:(b + 1) * n], os_), f"integrator {integ} provot {provot} cloth {b}"
        c.reset_pins()
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny", [(2048, 2048), (1500, 1100), (300, 200)])
def test_upload_step_download_pipeline_is_exact(nx, ny):
    """oc_upload, oc_step(1), oc_download on a whole cloth run as a row-chunked pipeline (copy streams, per-chunk
    launches): the state that comes back must equal the plain sequence (upload, sync, step, sync, download) bit for
    bit, for several rounds through the host, and through other call orders that fall back to the plain path."""
    import torch
    m = oc()
    ref = m.Cloth(nx, ny, kernel=m.OC_KERNEL_GATHER)
    ref.step(30)
    x0, xl0 = ref.download()
    hx = torch.from_numpy(x0.copy()).pin_memory(); hl = torch.from_numpy(xl0.copy()).pin_memory()
    c = m.Cloth(nx, ny)
    for rnd in range(4):
        c.upload_from(hx.data_ptr(), hl.data_ptr(), 3)
        c.step(1)
        c.download_into(hx.data_ptr(), hl.data_ptr(), 3)
    ref.step(4)
    rx, rl = ref.download()
    assert bitwise_equal(hx.numpy(), rx) and bitwise_equal(hl.numpy(), rl)
    # upload followed by several steps, by a write-back, by a download without a step
    c.upload_from(hx.data_ptr(), hl.data_ptr(), 3)
    c.step(3)
    c.upload_from(hx.data_ptr(), hl.data_ptr(), 3)
    x, xl = c.download()
    assert bitwise_equal(x, rx) and bitwise_equal(xl, rl)
    c.upload_from(hx.data_ptr(), hl.data_ptr(), 3)
    c.set_particle(3 * nx + 4, (0.0, 4.0, 1.0)); ref.set_particle(3 * nx + 4, (0.0, 4.0, 1.0))
    c.step(2); ref.step(2)
    x, xl = c.download(); rx, rl = ref.download()
    assert bitwise_equal(x, rx) and bitwise_equal(xl, rl)
    c.close(); ref.close()
