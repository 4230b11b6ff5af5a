# development: the driver's N = 8 command in the default (fast) mode, in exact mode, and the batch workload
mkdir -p gpurun_out/r2i
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
$TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2i/bench_n8_fast_20.json 2> gpurun_out/r2i/bench_n8_fast_20.err
$TR bench.py --gpus 8 --steps 20 --warmup 5 --mode exact --e2e-steps 3 > gpurun_out/r2i/bench_n8_exact_20.json 2> gpurun_out/r2i/bench_n8_exact_20.err
$TR bench.py --gpus 8 --steps 200 --warmup 20 --workload batch --e2e-steps 3 > gpurun_out/r2i/bench_n8_batch.json 2> gpurun_out/r2i/bench_n8_batch.err
python - <<'PY'
import json
for f in ("bench_n8_fast_20", "bench_n8_exact_20", "bench_n8_batch"):
    try:
        d = json.loads(open(f"gpurun_out/r2i/{f}.json").read().strip().splitlines()[-1])
        print(f, d["config"]["mode"], round(d["value"] / 1e9, 1), "G  eff", d.get("parallel_efficiency"), [round(x, 4) for x in d["detail"]["ms_per_rank"]], "base", (d.get("scaling_base") or {}).get("value"), "parity", (d.get("parity") or {}).get("bitwise"), "e2e", round(d["e2e"]["value"] / 1e9, 2))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2i/*.err
