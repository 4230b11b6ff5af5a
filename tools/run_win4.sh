# development: oc_k_stream with the register window AND a 6-row ring (4 CTAs of 128 threads per SM at 128 registers)
D=${2:-gpurun_out/r2y}; mkdir -p $D
DEV=$1
{
echo "sha old  $(python tools/twin_probe.py sha 2048 2048 1 6 0 300)"
echo "sha new  $(OC_LIB=$DEV OC_STREAM_WC=128 OC_STREAM_OCC=4 python tools/twin_probe.py sha 2048 2048 1 6 0 300)"
for g in "2048 2048 1 6 0 400" "8192 8192 1 6 0 60" "128 128 512 6 0 400"; do
  echo "old     $(python tools/twin_probe.py one $g)"
  for v in "128 4" "128 3" "64 8" "64 6"; do
    set -- $v
    echo "win+ring6 $1/$2 $(OC_LIB=$DEV OC_STREAM_WC=$1 OC_STREAM_OCC=$2 OC_DEBUG=16 python tools/twin_probe.py one $g 2>&1 | grep -o 'CTAs/SM [0-9]*\|{.*}' | tr '\n' ' ')"
  done
done
} 2>&1 | tee $D/stream_window_ring6_rates.log
