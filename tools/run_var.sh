# development: rates of oc_k_twin variants (window width, register cap) against oc_k_march2
mkdir -p gpurun_out/r2c
for e in 0 1; do
  echo "march2 $(python tools/twin_probe.py one 2048 2048 1 3 $e 400)"
  echo "march2 $(python tools/twin_probe.py one 8192 8192 1 3 $e 60)"
  for v in "128 2" "128 3" "64 4" "64 5" "64 6"; do
    set -- $v
    [ $e = 1 ] && [ $2 != 2 ] && [ $2 != 4 ] && continue
    echo "twin $(OC_TWIN_WC=$1 OC_TWIN_OCC=$2 python tools/twin_probe.py one 2048 2048 1 5 $e 400)"
    echo "twin $(OC_TWIN_WC=$1 OC_TWIN_OCC=$2 python tools/twin_probe.py one 8192 8192 1 5 $e 60)"
  done
done
