"""GPU tier: the headless C++ harness (the host program that replaces the reference's GLUT main loop)
runs over the C-ABI and reports the same spring energy as the oracle."""
import json
import os
import subprocess

import pytest

import helpers

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(helpers.ROOT, "opencloth_b200", "harness", "oc_harness")


def _build():
    if not os.path.exists(HARNESS):
        subprocess.check_call(["make", "-C", os.path.dirname(HARNESS)], env=dict(os.environ, CC="gcc", CXX="g++"))


def _run(*args):
    _build()
    r = subprocess.run([HARNESS, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]


def test_harness_default_cloth_energy_matches_oracle():
    lines = _run("--nx", "21", "--ny", "21", "--frames", "4", "--substeps", "500")
    o = helpers.Oracle(21, 21)
    for ln in lines:
        o.step(500)
        assert ln["spring_energy"] == pytest.approx(o.energy(), rel=1e-8), ln
    assert lines[-1]["step"] == 2000


def test_harness_row_bands_in_one_process(tmp_path):
    """--gpus g cuts the cloth into g bands; with one device they all live on it (oc_halo_exchange path)."""
    import numpy as np
    import torch
    g = min(4, max(1, torch.cuda.device_count()))
    d1, d2 = str(tmp_path / "a.f32"), str(tmp_path / "b.f32")
    _run("--nx", "300", "--ny", "256", "--frames", "2", "--substeps", "40", "--energy", "0", "--dump", d1)
    # four bands spread over the devices that exist (all on one device on a 1-GPU box): linked bands, then the
    # host-driven exchange
    for extra in (("--link", "1"), ("--link", "0", "--halo", "8")):
        _run("--nx", "300", "--ny", "256", "--frames", "2", "--substeps", "40", "--gpus", "4", "--devices", str(g), *extra, "--energy", "0", "--dump", d2)
        a, b = np.fromfile(d1, np.uint32), np.fromfile(d2, np.uint32)
        assert a.shape == b.shape and (a == b).all(), extra
    x = np.fromfile(d1, np.float32).reshape(-1, 4)
    o = helpers.Oracle(300, 256); o.step(80)
    assert helpers.bitwise_equal(x[:, :3], o.state()[0]) and (x[:, 3] == 1.0).all()
