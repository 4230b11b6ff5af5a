# development: oc_k_stream against the prefetch depth (rows in flight), window width and register cap
for lib in d0 d1 default d4; do
  if [ $lib = default ]; then unset OC_LIB; else export OC_LIB=$PWD/opencloth_b200/libvar_$lib.so; fi
  for v in "128 2" "128 3" "64 4" "64 6"; do
    set -- $v
    echo "depth=$lib wc=$1 occ=$2 $(OC_STREAM_WC=$1 OC_STREAM_OCC=$2 OC_DEBUG=16 python tools/twin_probe.py one 2048 2048 1 6 0 400 2>&1 | tail -2 | tr '\n' ' ')"
    echo "depth=$lib wc=$1 occ=$2 $(OC_STREAM_WC=$1 OC_STREAM_OCC=$2 python tools/twin_probe.py one 8192 8192 1 6 0 60)"
  done
done
