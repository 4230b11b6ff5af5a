# development: oc_k_stream with the register window (dev build in $1, all variants) against the installed library
D=${2:-gpurun_out/r2w}; mkdir -p $D
DEV=$1
{
echo "sha old  $(python tools/twin_probe.py sha 2048 2048 1 6 0 300)"
echo "sha new  $(OC_LIB=$DEV python tools/twin_probe.py sha 2048 2048 1 6 0 300)"
echo "sha new2 $(OC_LIB=$DEV OC_STREAM_WC=64 OC_STREAM_OCC=5 python tools/twin_probe.py sha 2048 2048 1 6 0 300)"
for g in "2048 2048 1 6 0 400" "8192 8192 1 6 0 60" "128 128 512 6 0 400"; do
  echo "old     $(python tools/twin_probe.py one $g)"
  for v in "128 3" "128 2" "64 6" "64 5" "64 4"; do
    set -- $v
    echo "win $1/$2 $(OC_LIB=$DEV OC_STREAM_WC=$1 OC_STREAM_OCC=$2 python tools/twin_probe.py one $g)"
  done
done
} 2>&1 | grep -v "^old     $" | tee $D/stream_window_rates.log
if [ -n "$NCU" ]; then
  OC_LIB=$DEV ncu --set full --clock-control none --import-source on -k regex:oc_k_stream -s 45 -c 1 -f -o $D/stream_win_2048 python tools/prof_one.py --kernel 6 --exact 0 --warm 40 --launches 8 > $D/ncu_stream_win.log 2>&1
fi
