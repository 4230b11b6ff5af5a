# the single-GPU evidence of a round: GPU tests, the bench at the driver's step counts and at the default, ncu launch list of
# the bench command, ncu --set full of the default kernel in both modes, sanitizer, PCIe duplex rates, e2e time line
R=${1:-r2f}
mkdir -p gpurun_out/$R
(timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/$R/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/$R/bench_n1_20.json 2> gpurun_out/$R/bench_n1_20.err
python bench.py > gpurun_out/$R/bench_n1_default.json 2> gpurun_out/$R/bench_n1_default.err
python bench.py --mode exact --no-cpu-baseline > gpurun_out/$R/bench_n1_exact.json 2> gpurun_out/$R/bench_n1_exact.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/$R/bench_reference.json 2> gpurun_out/$R/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$R/launches_bench_2048.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/$R/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:oc_k_stream -s 45 -c 1 -f -o gpurun_out/$R/stream_fast_2048 python tools/prof_one.py --kernel 6 --exact 0 --warm 40 --launches 8 > gpurun_out/$R/ncu_stream_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:oc_k_march2 -s 45 -c 1 -f -o gpurun_out/$R/march2_exact_2048 python tools/prof_one.py --kernel 3 --exact 1 --warm 40 --launches 8 > gpurun_out/$R/ncu_march2_exact.log 2>&1
compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/$R/sanitizer_memcheck.log 2>&1
compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/$R/sanitizer_racecheck.log 2>&1
python tools/microbench/pcie_duplex.py > gpurun_out/$R/pcie_duplex.log 2>&1
python tools/pipeline_timeline.py > gpurun_out/$R/pipeline_timeline_2048.txt 2>&1
tail -3 gpurun_out/$R/pytest_gpu.log; tail -2 gpurun_out/$R/sanitizer_memcheck.log; tail -2 gpurun_out/$R/sanitizer_racecheck.log; cat gpurun_out/$R/pcie_duplex.log
