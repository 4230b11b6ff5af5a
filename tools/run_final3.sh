# last check of the final build: the GPU tier, smoke(), the bench at the driver's step counts
R=${1:-r2k}
mkdir -p gpurun_out/$R
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/$R/pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2) > gpurun_out/$R/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/$R/bench_n1_20.json 2> gpurun_out/$R/bench_n1_20.err
tail -2 gpurun_out/$R/pytest_gpu.log; tail -1 gpurun_out/$R/smoke.log; cut -c1-300 gpurun_out/$R/bench_n1_20.json
