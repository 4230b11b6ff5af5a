# development: the band-resident kernel (kernel 8): parity tests, then rates against the marching kernels at mid sizes
D=${1:-gpurun_out/r2z}; mkdir -p $D
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "bandres" > $D/pytest_bandres.log 2>&1; tail -5 $D/pytest_bandres.log
{
for g in "128 128" "256 256" "384 384" "512 512" "512 576" "300 200" "100 1000"; do
  for e in 1 0; do
    echo "bandres $(timeout 120 python tools/twin_probe.py one $g 1 8 $e 2000)"
    echo "march2  $(timeout 120 python tools/twin_probe.py one $g 1 3 $e 2000)"
  done
done
} 2>&1 | tee $D/bandres_rates.log
