import sys, time
sys.path.insert(0, '/root/repo')
import opencloth_b200 as oc
def run(pre_scratch, warm, steps=20, exact=0):
    if pre_scratch:
        s = oc.Cloth(2048, 2048, exact=exact); s.step(pre_scratch); s.sync(); s.close()
    c = oc.Cloth(2048, 2048, exact=exact)
    c.step(warm); c.sync()
    ms = c.step_timed(steps)
    ms2 = c.step_timed(steps)
    c.close()
    return 2048*2048*steps/(ms*1e-3)/1e9, 2048*2048*steps/(ms2*1e-3)/1e9
for args in ((0,5),(0,5),(300,5),(0,300),(0,1000)):
    print("pre_scratch %4d warm %4d: first 20 timed steps %.2f G, next 20: %.2f G" % (args + run(*args)), flush=True)
