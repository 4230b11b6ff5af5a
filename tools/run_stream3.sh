# development: oc_k_stream (kernel 6) against window width and register cap, both modes, + parity
for g in "2048 2048 1 2300" "128 128 64 2300" "1000 777 1 600"; do
  set -- $g
  ref=$(python tools/twin_probe.py sha $1 $2 $3 3 1 $4)
  for v in "128 0" "128 4" "64 0"; do
    set -- $g $v
    got=$(OC_STREAM_WC=$5 OC_STREAM_OCC=$6 python tools/twin_probe.py sha $1 $2 $3 6 1 $4)
    [ "$got" = "$ref" ] && echo "parity $1x$2x$3 $4 steps wc=$5 occ=$6 OK" || echo "parity $1x$2x$3 wc=$5 occ=$6 MISMATCH $got vs $ref"
  done
done
echo "march2 $(python tools/twin_probe.py one 2048 2048 1 3 0 400)"
echo "march2 $(python tools/twin_probe.py one 8192 8192 1 3 0 60)"
for v in "128 2" "128 3" "128 4" "64 4" "64 6" "64 8"; do
  set -- $v
  echo "stream wc=$1 occ=$2 $(OC_STREAM_WC=$1 OC_STREAM_OCC=$2 OC_DEBUG=16 python tools/twin_probe.py one 2048 2048 1 6 0 400 2>&1 | tail -2 | tr '\n' ' ')"
  echo "stream wc=$1 occ=$2 $(OC_STREAM_WC=$1 OC_STREAM_OCC=$2 python tools/twin_probe.py one 8192 8192 1 6 0 60)"
done
echo "batch march2 $(python tools/twin_probe.py one 128 128 512 3 0 400)"
echo "batch stream $(python tools/twin_probe.py one 128 128 512 6 0 400)"
echo "batch stream occ4 $(OC_STREAM_OCC=4 python tools/twin_probe.py one 128 128 512 6 0 400)"
