# development: e2e round trip against chunk count, tapered and equal chunks
for taper in 1 0; do
  for ch in 8 10 12 16; do
    echo "taper=$taper $(OC_PIPE_TAPER=$taper OC_PIPE_CHUNKS=$ch python - <<PY
import sys, time, torch
sys.path.insert(0, ".")
import opencloth_b200 as m
nx = 2048
c = m.Cloth(nx, nx); c.step(5)
hx = torch.empty((nx * nx, 3), dtype=torch.float32).pin_memory(); hl = torch.empty((nx * nx, 3), dtype=torch.float32).pin_memory()
c.download_into(hx.data_ptr(), hl.data_ptr(), 3)
def it():
    c.upload_from(hx.data_ptr(), hl.data_ptr(), 3); c.step(1); c.download_into(hx.data_ptr(), hl.data_ptr(), 3)
for i in range(3): it()
t0 = time.perf_counter()
for i in range(20): it()
dt = (time.perf_counter() - t0) / 20
print("chunks", $ch, "ms/iter %.3f" % (dt * 1e3), "G updates/s %.3f" % (nx * nx / dt / 1e9))
PY
)"
  done
done
