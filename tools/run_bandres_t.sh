# development: oc_k_bandres with 768 instead of 512 threads per CTA
D=${1:-gpurun_out/r2z3}; mkdir -p $D
{
for g in "256 256" "384 384" "512 512" "512 576" "100 1000"; do
  for e in 1 0; do
    echo "T=512 $(timeout 120 python tools/twin_probe.py one $g 1 8 $e 2000)"
    echo "T=768 $(OC_BANDRES_T=768 timeout 120 python tools/twin_probe.py one $g 1 8 $e 2000)"
  done
done
echo "sha 512 $(python tools/twin_probe.py sha 512 512 1 8 1 500) 768 $(OC_BANDRES_T=768 python tools/twin_probe.py sha 512 512 1 8 1 500) gather $(python tools/twin_probe.py sha 512 512 1 1 1 500)"
} 2>&1 | tee $D/bandres_threads.log
