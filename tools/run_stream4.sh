for lib in r8p0 r8p1 r6p0 default; do
  if [ $lib = default ]; then unset OC_LIB; else export OC_LIB=$PWD/opencloth_b200/libvar_$lib.so; fi
  for v in "128 0" "128 4"; do
    set -- $v
    echo "lib=$lib wc=$1 occ=$2 $(OC_STREAM_WC=$1 OC_STREAM_OCC=$2 python tools/twin_probe.py one 2048 2048 1 6 0 400)"
    echo "lib=$lib wc=$1 occ=$2 $(OC_STREAM_WC=$1 OC_STREAM_OCC=$2 python tools/twin_probe.py one 8192 8192 1 6 0 60)"
  done
  echo "lib=$lib batch $(python tools/twin_probe.py one 128 128 512 6 0 400)"
done
