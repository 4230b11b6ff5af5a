# development: the driver's N = 8 command, then the same with one launch direction on every band (A/B) and a long run
mkdir -p gpurun_out/r2e
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2e/bench_n8_20.json 2> gpurun_out/r2e/bench_n8_20.err
OC_LINK_REV=0 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-parity --e2e-steps 3 > gpurun_out/r2e/bench_n8_20_rev0.json 2> gpurun_out/r2e/bench_n8_20_rev0.err
$TR bench.py --gpus 8 --steps 200 --warmup 20 --no-parity --e2e-steps 3 > gpurun_out/r2e/bench_n8_200.json 2> gpurun_out/r2e/bench_n8_200.err
python - <<'PY'
import json
for f in ("bench_n8_20", "bench_n8_20_rev0", "bench_n8_200"):
    try:
        d = json.loads(open(f"gpurun_out/r2e/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"] / 1e9, 1), "G  eff", round(d["parallel_efficiency"], 4), [round(x, 4) for x in d["detail"]["ms_per_rank"]], "base", round(d["scaling_base"]["value"] / 1e9, 1), "parity", (d.get("parity") or {}).get("bitwise"), "e2e", round(d["e2e"]["value"] / 1e9, 2))
    except Exception as e:
        print(f, "failed", e)
PY
