import sys, time; sys.path.insert(0, ".")
import opencloth_b200 as oc
for kern, name in ((4, "resident"), (1, "gather"), (3, "march2")):
    for nx, ny, batch in ((21, 21, 1), (21, 21, 4096), (32, 32, 1), (32, 32, 1024)):
        c = oc.Cloth(nx, ny, batch=batch, kernel=kern)
        c.step(50); c.sync()
        t0 = time.perf_counter(); c.step(1000); c.sync(); dt = time.perf_counter() - t0
        print(f"{name:9s} {nx}x{ny} x{batch}: {dt*1e3:8.2f} ms per 1000 steps, {nx*ny*batch*1000/dt/1e9:7.3f} G updates/s", flush=True)
        c.close()
