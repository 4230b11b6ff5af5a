# single-GPU validation of the final build (round 2, with oc_k_bandres): GPU tests, smoke, bench lines, sanitizer, ncu of kernel 8
R=${1:-r2h}
mkdir -p gpurun_out/$R
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/$R/pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/$R/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/$R/bench_n1_20.json 2> gpurun_out/$R/bench_n1_20.err
python bench.py > gpurun_out/$R/bench_n1_default.json 2> gpurun_out/$R/bench_n1_default.err
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/$R/sanitizer_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/$R/sanitizer_racecheck.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:oc_k_bandres -s 1 -c 1 -f -o gpurun_out/$R/bandres_exact_256 python tools/prof_one.py --n 256 --kernel 8 --exact 1 --warm 100 --launches 200 > gpurun_out/$R/ncu_bandres.log 2>&1
tail -3 gpurun_out/$R/pytest_gpu.log; cat gpurun_out/$R/smoke.log; tail -2 gpurun_out/$R/sanitizer_memcheck.log; tail -2 gpurun_out/$R/sanitizer_racecheck.log; tail -2 gpurun_out/$R/ncu_bandres.log
R=$R python - <<'P'
import json, os
for f in ("bench_n1_20", "bench_n1_default"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{os.environ['R']}/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"] / 1e9, 2), round(d["roofline"]["frac"], 3), round(d["e2e"]["value"] / 1e9, 3), d.get("mid_size"))
    except Exception as e:
        print(f, "unreadable", e)
P
