# development: oc_k_march2 with one more row in flight (OC_M2_DEPTH = 1): parity against the gather kernel, rates in both modes
for g in "2048 2048 1 2300" "128 128 64 2300" "1000 777 1 600"; do
  set -- $g
  ref=$(python tools/twin_probe.py sha $1 $2 $3 1 1 $4)
  got=$(python tools/twin_probe.py sha $1 $2 $3 3 1 $4)
  [ "$got" = "$ref" ] && echo "parity march2 vs gather $1x$2x$3 $4 steps OK" || echo "parity $1x$2x$3 MISMATCH $got vs $ref"
done
for e in 0 1; do
  echo "march2 $(python tools/twin_probe.py one 2048 2048 1 3 $e 400)"
  echo "march2 $(python tools/twin_probe.py one 8192 8192 1 3 $e 60)"
  echo "march2 $(python tools/twin_probe.py one 128 128 512 3 $e 400)"
done
echo "stream $(python tools/twin_probe.py one 2048 2048 1 6 0 400)"
