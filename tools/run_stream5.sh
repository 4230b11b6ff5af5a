# development: parity + rates of the current oc_k_stream build against oc_k_march2 (fast mode)
for g in "2048 2048 1 2300" "128 128 64 2300" "1000 777 1 600"; do
  set -- $g
  ref=$(python tools/twin_probe.py sha $1 $2 $3 3 1 $4)
  got=$(python tools/twin_probe.py sha $1 $2 $3 6 1 $4)
  [ "$got" = "$ref" ] && echo "parity exact $1x$2x$3 $4 steps OK" || echo "parity $1x$2x$3 MISMATCH $got vs $ref"
done
echo "march2 $(python tools/twin_probe.py one 2048 2048 1 3 0 400)"
echo "march2 $(python tools/twin_probe.py one 8192 8192 1 3 0 60)"
for v in "128 2" "128 3" "128 4" "64 6"; do
  set -- $v
  echo "stream wc=$1 occ=$2 $(OC_STREAM_WC=$1 OC_STREAM_OCC=$2 python tools/twin_probe.py one 2048 2048 1 6 0 400)"
  echo "stream wc=$1 occ=$2 $(OC_STREAM_WC=$1 OC_STREAM_OCC=$2 python tools/twin_probe.py one 8192 8192 1 6 0 60)"
done
echo "batch stream $(python tools/twin_probe.py one 128 128 512 6 0 400)"
