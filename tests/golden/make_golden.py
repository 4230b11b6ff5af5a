#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ from the VERBATIM reference physics.

Run in the build container (where /root/reference is mounted):
    ./oracle/build_ref.sh && python tests/golden/make_golden.py

The generator is oracle/_ref/libocref.so, i.e. the reference's own StepPhysics / InitGL text
(/root/reference/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp line slices, see oracle/build_ref.sh)
compiled against its vendored GLM 0.9.0.0 with g++ -O2 -ffp-contract=off.  The reference ships no
tests or golden data of its own, so these files are what pins parity on
machines where /root/reference does not exist (the GPU box).

Per grid and checkpoint: full fp32 state for small grids; for larger ones the SHA-256 of the raw
X / X_last bytes plus every `row_stride`-th row, so that the fixtures stay small while a bit-exact
implementation can still be verified completely (hash) and a tolerance-mode one meaningfully (rows).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.float32).tobytes()).hexdigest()


def run(nx, ny, checkpoints, full, row_stride=16, energy_every=10):
    r = helpers.Ref(nx, ny)
    out = {}
    meta = {"nx": nx, "ny": ny, "checkpoints": list(checkpoints), "full": full, "row_stride": row_stride,
            "energy_every": energy_every, "sha_x": {}, "sha_xl": {}, "energy": {}, "hits": {}}
    traj = []
    step = 0
    last = max(checkpoints)
    while step < last:
        r.step(1)
        step += 1
        if step % energy_every == 0:
            traj.append(r.energy())
        if step in checkpoints:
            x, xl = r.state()
            meta["sha_x"][str(step)] = sha(x)
            meta["sha_xl"][str(step)] = sha(xl)
            meta["energy"][str(step)] = r.energy()
            meta["hits"][str(step)] = int((x == xl).all(1).sum())
            if full:
                out[f"x_{step}"] = x
                out[f"xl_{step}"] = xl
            else:
                rows = np.arange(0, ny, row_stride)
                out[f"x_{step}"] = x.reshape(ny, nx, 3)[rows].copy()
                out[f"xl_{step}"] = xl.reshape(ny, nx, 3)[rows].copy()
    out["energy_traj"] = np.asarray(traj, np.float64)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    path = os.path.join(HERE, f"grid_{nx}x{ny}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


def run_variant(name, nx, ny, checkpoints, full, row_stride=16):
    """The sibling integrators and the Provot pass, same fixture format.
    name: euler | semi (verbatim builds of OpenCloth_ExplicitEuler / OpenCloth_SemiImplicit, Provot on as they ship;
    the second array is V), euler_noprovot | semi_noprovot (ApplyProvotDynamicInverse switched off: reaches the
    collider), verlet_provot (the Verlet demo with the call of V:561 enabled; second array is X_last)."""
    if name == "verlet_provot":
        r = helpers.Ref(nx, ny)
        step_fn = r.step_provot
    else:
        integ = helpers.EULER if name.startswith("euler") else helpers.SEMI
        r = helpers.RefVariant(integ, nx, ny, provot=0 if name.endswith("noprovot") else 1)
        step_fn = r.step
    out = {}
    meta = {"variant": name, "nx": nx, "ny": ny, "checkpoints": list(checkpoints), "full": full, "row_stride": row_stride,
            "sha_x": {}, "sha_xl": {}, "hits": {}}
    step = 0
    for cp in checkpoints:
        step_fn(cp - step)
        step = cp
        x, xl = r.state()
        meta["sha_x"][str(cp)] = sha(x)
        meta["sha_xl"][str(cp)] = sha(xl)
        meta["hits"][str(cp)] = int((x == xl).all(1).sum()) if name == "verlet_provot" else int((xl == 0).all(1).sum())
        if full:
            out[f"x_{cp}"] = x
            out[f"xl_{cp}"] = xl
        else:
            rows = np.arange(0, ny, row_stride)
            out[f"x_{cp}"] = x.reshape(ny, nx, 3)[rows].copy()
            out[f"xl_{cp}"] = xl.reshape(ny, nx, 3)[rows].copy()
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    path = os.path.join(HERE, f"{name}_{nx}x{ny}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", meta["hits"])


def main():
    assert helpers.have_ref(), "build oracle/_ref/libocref.so first (./oracle/build_ref.sh)"
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        for name in ("euler", "semi", "euler_noprovot", "semi_noprovot", "verlet_provot"):
            run_variant(name, 21, 21, [1, 100, 1000, 2000], full=True)
            run_variant(name, 64, 64, [100, 1000] + ([2000] if "noprovot" in name else []), full=False, row_stride=16)
        for name in ("euler", "verlet_provot"):
            run_variant(name, 256, 256, [100, 400], full=False, row_stride=64)
        # vertex normals of the lit demo's UpdateNormals (first call) on states of the Verlet golden run
        g = np.load(os.path.join(HERE, "grid_21x21.npz"))
        out = {}
        for cp in (100, 2000, 3000):
            out[f"n_{cp}"] = helpers.verbatim_normals(g[f"x_{cp}"], 21, 21)
        np.savez_compressed(os.path.join(HERE, "normals_21x21.npz"), **out)
        return
    # 1. the reference's default configuration; 1672 is the first step at which the collider acts
    run(21, 21, [1, 10, 100, 1000, 1671, 1672, 2000, 3000], full=True)
    run(37, 23, [100, 1000, 2500], full=True)
    run(64, 64, [100, 1000, 2500], full=False, row_stride=8)
    run(256, 256, [100, 1000], full=False, row_stride=32)
    # 2. set-up data: spring list of the default grid and a non-square one, collider matrices, scalars
    for nx, ny in ((21, 21), (37, 23)):
        r = helpers.Ref(nx, ny)
        s = r.springs()
        np.savez_compressed(os.path.join(HERE, f"springs_{nx}x{ny}.npz"), **s)
    r = helpers.Ref(21, 21)
    m, mi = r.ellipsoid()
    np.savez(os.path.join(HERE, "setup.npz"), ellipsoid=m, inv_ellipsoid=mi, params=r.params(),
             x0=r.state()[0])
    print("done")


if __name__ == "__main__":
    main()
