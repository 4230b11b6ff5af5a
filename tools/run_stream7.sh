# development: oc_k_stream2 (kernel 7, two columns per thread): parity, rates against oc_k_stream and oc_k_march2 (fast mode)
for g in "2048 2048 1 2300" "128 128 64 2300" "1000 777 1 600"; do
  set -- $g
  ref=$(python tools/twin_probe.py sha $1 $2 $3 3 1 $4)
  got=$(python tools/twin_probe.py sha $1 $2 $3 7 1 $4)
  [ "$got" = "$ref" ] && echo "parity exact $1x$2x$3 $4 steps OK" || echo "parity $1x$2x$3 MISMATCH $got vs $ref"
  a=$(python tools/twin_probe.py sha $1 $2 $3 6 0 $4); b=$(python tools/twin_probe.py sha $1 $2 $3 7 0 $4)
  [ "$a" = "$b" ] && echo "fast mode stream2 == stream $1x$2x$3 OK" || echo "fast mode $1x$2x$3 MISMATCH"
done
echo "march2 $(python tools/twin_probe.py one 2048 2048 1 3 0 400)"
echo "stream $(python tools/twin_probe.py one 2048 2048 1 6 0 400)"
echo "stream $(python tools/twin_probe.py one 8192 8192 1 6 0 60)"
for occ in 4 5 6; do
  echo "stream2 occ=$occ $(OC_STREAM2_OCC=$occ OC_DEBUG=16 python tools/twin_probe.py one 2048 2048 1 7 0 400 2>&1 | tail -2 | tr '\n' ' ')"
  echo "stream2 occ=$occ $(OC_STREAM2_OCC=$occ python tools/twin_probe.py one 8192 8192 1 7 0 60)"
  echo "stream2 occ=$occ $(OC_STREAM2_OCC=$occ python tools/twin_probe.py one 128 128 512 7 0 400)"
done
echo "stream2 exact $(python tools/twin_probe.py one 8192 8192 1 7 1 60)"
