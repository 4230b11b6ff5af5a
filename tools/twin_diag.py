import os, sys, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
mode = sys.argv[1]
kern = int(sys.argv[2]); exact = int(sys.argv[3])
os.environ["OC_DEBUG"] = "4" if mode == "cnt" else "8"
import opencloth_b200 as oc
c = oc.Cloth(2048, 2048, kernel=kern, exact=exact)
c.step(300)
if mode == "cnt":
    out=(ctypes.c_ulonglong*4)()
    c._lib.oc_debug_counters(c._h,out)
    a=[int(v) for v in out]
    c.step(10)
    c._lib.oc_debug_counters(c._h,out)
    b=[int(v) for v in out]
    n=2048*2048/2
    print("kern",kern,"exact",exact,"per step: redo lanes %.0f (%.3f%% of thread-iters) warps %.0f badv %.0f hits %.0f" % ((b[0]-a[0])/10, 100*(b[0]-a[0])/10/n, (b[1]-a[1])/10, ((b[2]-a[2])&0xffffffffff)/10, ((b[3]-a[3])&0xffffffff)/10))
else:
    ms = c.step_timed(50) / 50
    c.step(3)
    buf = (ctypes.c_ulonglong * (8 * 4096))()
    assert c._lib.oc_debug_timeline(c._h, buf, 8 * 4096) == 0
    t = np.frombuffer(buf, dtype=np.uint64).reshape(4096, 8).astype(np.int64)
    t = t[t[:, 0] > 0]
    rel = (t[:, :5] - t[:, 0].min()) / 1000.0
    d = rel[:, 1:] - rel[:, :-1]
    print("kern",kern,"exact",exact,f"{len(t)} CTAs, {ms*1e3:.1f} us/step")
    for k, nm in enumerate(["set-up", "lead-in", "steady", "lead-out"]):
        print(f"  phase {nm:8s} median {np.median(d[:, k]):7.1f} us   p10 {np.percentile(d[:,k],10):7.1f} p90 {np.percentile(d[:,k],90):7.1f} max {d[:, k].max():7.1f}")
