"""Render hand-off: vertex normals as the reference's lit demo computes them (UpdateNormals,
OpenCloth_ExplicitEuler_TextureMapped_Lit/.../main.cpp:684-707 over the triangle list :313-327), the float4 vertex
buffer, and the harness's self-describing .npy frames.  Bar: bit-exact against the verbatim UpdateNormals."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from helpers import Emu, bitwise_equal


def test_normals_kernel_body_matches_golden():
    g = helpers.load_golden("grid_21x21.npz"); n = helpers.load_golden("normals_21x21.npz")
    for cp in (100, 2000, 3000):
        e = Emu(21, 21); e.upload(g[f"x_{cp}"], g[f"xl_{cp}"])
        assert bitwise_equal(e.normals(), n[f"n_{cp}"]), cp
        assert bitwise_equal(helpers.reference_normals(g[f"x_{cp}"], 21, 21), n[f"n_{cp}"])      # the numpy restatement used on the GPU tier


@pytest.mark.skipif(not os.path.exists(helpers.REF_NORMALS_SO), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("nx,ny,pre", [(21, 21, 1800), (8, 5, 300), (3, 3, 50), (12, 17, 900), (64, 64, 700), (2, 2, 0)])
def test_normals_match_verbatim_update_normals(nx, ny, pre):
    if nx < 3:
        x0 = np.array([[0, 0, 0], [1, 0, 0], [0, 0.5, 1], [1, 0.2, 1]], np.float32)
        assert bitwise_equal(helpers.reference_normals(x0, 2, 2), helpers.verbatim_normals(x0, 2, 2))
        return
    x0, xl0 = helpers.developed_state(nx, ny, pre)
    v = helpers.verbatim_normals(x0, nx, ny)
    e = Emu(nx, ny); e.upload(x0, xl0)
    assert bitwise_equal(e.normals(), v)
    assert bitwise_equal(helpers.reference_normals(x0, nx, ny), v)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-6)


@pytest.mark.gpu
def test_cuda_normals_and_vertex_buffer():
    import opencloth_b200 as m
    for nx, ny, steps, B in ((21, 21, 1900, 1), (37, 23, 800, 3), (150, 90, 120, 1)):
        c = m.Cloth(nx, ny, batch=B)
        c.step(steps)
        x4, _ = c.download(stride=4)
        assert (x4[:, 3] == 1.0).all()
        for stride in (3, 4):
            n = c.download_normals(stride)
            for b in range(B):
                sl = slice(b * nx * ny, (b + 1) * nx * ny)
                assert bitwise_equal(n[sl, :3], helpers.reference_normals(x4[sl, :3], nx, ny)), (nx, ny, b)
            if stride == 4:
                assert (n[:, 3] == 0.0).all()
        c.close()
    g = helpers.load_golden("grid_21x21.npz"); gn = helpers.load_golden("normals_21x21.npz")
    c = m.Cloth(21, 21); c.step(2000)
    assert bitwise_equal(c.download()[0], g["x_2000"]) and bitwise_equal(c.download_normals(), gn["n_2000"])
    c.close()
    band = m.Cloth(64, 64, row_begin=0, row_end=32, halo_rows=2)
    with pytest.raises(m.OpenClothError):
        band.download_normals()
    band.close()


@pytest.mark.gpu
def test_harness_npy_frames(tmp_path):
    harness = os.path.join(helpers.ROOT, "opencloth_b200", "harness", "oc_harness")
    if not os.path.exists(harness):
        subprocess.check_call(["make", "-C", os.path.dirname(harness)], env=dict(os.environ, CC="gcc", CXX="g++"))
    for extra in ((), ("--gpus", "3", "--devices", "1")):
        prefix = str(tmp_path / ("f" + str(len(extra))))
        r = subprocess.run([harness, "--nx", "40", "--ny", "30", "--frames", "2", "--substeps", "50", "--energy", "0", "--dump-npy", prefix, *extra],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        X = np.load(prefix + "_X.npy"); N = np.load(prefix + "_N.npy")
        assert X.shape == (1200, 4) and X.dtype == np.float32 and N.shape == (1200, 3)
        o = helpers.Oracle(40, 30); o.step(100)
        assert bitwise_equal(X[:, :3], o.state()[0]) and (X[:, 3] == 1.0).all()
        assert bitwise_equal(N, helpers.reference_normals(X[:, :3], 40, 30))
