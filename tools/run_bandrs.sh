# development: an 8192-wide cloth of as many rows as one band of the N-GPU run has, against the rows per tile
for rows in 1024 2048; do
  echo "auto   $(OC_DEBUG=16 python tools/twin_probe.py one 8192 $rows 1 3 1 200 2>&1 | tail -1)"
  for rs in 86 94 103 114 128 147 171 205 256; do
    echo "rs=$rs $(OC_MARCH_RS=$rs python tools/twin_probe.py one 8192 $rows 1 3 1 200)"
  done
done
