#!/usr/bin/env python
"""Development probe for oc_k_twin on one GPU: bitwise check against oc_k_march2 / oc_k_gather, then rates of the
variants (window width, register cap) against oc_k_march2.  CUDA events inside oc_step_timed."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rate(nx, ny, batch, kernel, exact, steps, pre=200):
    import opencloth_b200 as oc
    c = oc.Cloth(nx, ny, batch=batch, kernel=kernel, exact=exact)
    c.step(pre)
    for _ in range(2):
        c.step_timed(steps)
    best = min(c.step_timed(steps) for _ in range(3))
    c.close()
    return nx * ny * batch * steps / (best * 1e-3)


def sha(nx, ny, batch, kernel, exact, steps):
    import opencloth_b200 as oc
    c = oc.Cloth(nx, ny, batch=batch, kernel=kernel, exact=exact)
    c.step(steps)
    x, xl = c.download()
    c.close()
    return hashlib.sha256(x.tobytes() + xl.tobytes()).hexdigest()[:16]


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        nx, ny, b, kern, exact, steps = [int(t) for t in sys.argv[2:8]]
        print(json.dumps(dict(grid=f"{nx}x{ny}x{b}", kernel=kern, exact=exact, wc=os.environ.get("OC_TWIN_WC", "-"), occ=os.environ.get("OC_TWIN_OCC", "-"),
                              rs=os.environ.get("OC_MARCH_RS", "auto"), gups=round(rate(nx, ny, b, kern, exact, steps) / 1e9, 2))), flush=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sha":
        nx, ny, b, kern, exact, steps = [int(t) for t in sys.argv[2:8]]
        print(sha(nx, ny, b, kern, exact, steps), flush=True)
        sys.exit(0)

    def sub(args, **env):
        e = dict(os.environ); e.update({k: str(v) for k, v in env.items()})
        r = subprocess.run([sys.executable, __file__] + [str(a) for a in args], env=e, capture_output=True, text=True)
        return (r.stdout.strip() or r.stderr.strip()[-400:])

    # parity first: 2300 steps cover free fall and collider contact
    for (nx, ny, b, steps) in ((2048, 2048, 1, 2300), (128, 128, 64, 2300), (1000, 777, 1, 600)):
        ref = sub(["sha", nx, ny, b, 3, 1, steps])
        for env in (dict(OC_TWIN_WC=128), dict(OC_TWIN_WC=64), dict(OC_TWIN_WC=128, OC_TWIN_OCC=3), dict(OC_TWIN_WC=128, OC_TWIN_PAIR=0)):
            got = sub(["sha", nx, ny, b, 5, 1, steps], **env)
            print(f"parity {nx}x{ny}x{b} {steps} steps {env}: {'OK' if got == ref else 'MISMATCH ' + got + ' vs ' + ref}", flush=True)
    for grid in ((2048, 2048, 1, 400), (8192, 8192, 1, 60), (128, 128, 512, 400)):
        nx, ny, b, steps = grid
        for exact in (1, 0):
            print(sub(["one", nx, ny, b, 3, exact, steps]), flush=True)
            for env in (dict(OC_TWIN_WC=128), dict(OC_TWIN_WC=128, OC_TWIN_OCC=3), dict(OC_TWIN_WC=64), dict(OC_TWIN_WC=64, OC_TWIN_OCC=5), dict(OC_TWIN_WC=64, OC_TWIN_OCC=6)):
                print(sub(["one", nx, ny, b, 5, exact, steps], **env), flush=True)
