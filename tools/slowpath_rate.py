import os, sys, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OC_DEBUG"]="4"
import opencloth_b200 as oc
for warm in (300, 1000, 3000):
    c = oc.Cloth(2048,2048,kernel=3,exact=1)
    c.step(warm)
    out=(ctypes.c_ulonglong*4)()
    c._lib.oc_debug_counters(c._h,out)
    c.step(10)
    c._lib.oc_debug_counters(c._h,out)
    n=2048*2048/2
    print("warm",warm,"per step: redo lanes %.0f (%.3f%% of threads-iters) warps %.0f (%.2f%% of warp-iters) badv %.0f hit lanes %.0f hit warps %.0f (%.2f%%)" % (
        out[0]/10, 100*out[0]/10/n, out[1]/10, 100*out[1]/10/(n/32), out[2]/10, (out[3]&0xffffffff)/10, (out[3]>>32)/10, 100*(out[3]>>32)/10/(n/32)))
    X, XL = c.download()
    hit = np.all(X==XL, axis=-1).reshape(2048,2048)
    print("  X==X_last rows profile (per 128 rows):", [int(hit[r:r+128].sum()) for r in range(0,2048,128)])
    c.close()
