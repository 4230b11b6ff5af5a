# development: the driver's N-GPU command (N = $1), 20 steps after 5 warm-up steps
N=$1
mkdir -p gpurun_out/r2e
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2e/bench_n${N}_20.json 2> gpurun_out/r2e/bench_n${N}_20.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2e/bench_n${N}_20.json").read().strip().splitlines()[-1])
print("N=$N", round(d["value"] / 1e9, 1), "G  eff", round(d["parallel_efficiency"], 4), [round(x, 4) for x in d["detail"]["ms_per_rank"]], "base", round(d["scaling_base"]["value"] / 1e9, 1), "parity", (d.get("parity") or {}).get("bitwise"), "e2e", round(d["e2e"]["value"] / 1e9, 2))
PY
