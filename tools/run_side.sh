# development: oc_k_stream with side-column loads (non-overlapping windows): rates, then the GPU tests
D=${1:-gpurun_out/r2x}; mkdir -p $D
{
for g in "2048 2048 1 6 0 400" "8192 8192 1 6 0 60" "128 128 512 6 0 400" "2048 2048 1 6 1 400" "1024 1024 1 6 0 400" "1000 777 1 6 0 400"; do
  echo "stream $(python tools/twin_probe.py one $g)"
done
echo "march2 $(python tools/twin_probe.py one 2048 2048 1 3 1 400)"
} 2>&1 | tee $D/stream_side_rates.log
[ -n "$TESTS" ] && { timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest_gpu.log 2>&1; tail -3 $D/pytest_gpu.log; }
