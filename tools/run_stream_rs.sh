# development: oc_k_stream against the rows per tile (2048^2: even segment counts), fast mode
for rs in 26 30 32 35 38 43 47 52 57 64 79 86 103 128; do
  echo "2048 rs=$rs $(OC_MARCH_RS=$rs OC_DEBUG=16 python tools/twin_probe.py one 2048 2048 1 6 0 400 2>&1 | tail -2 | tr '\n' ' ' | sed 's/pair_cloths.*exact 0//')"
done
for rs in 256 342 410 512 683 1024; do
  echo "8192 rs=$rs $(OC_MARCH_RS=$rs python tools/twin_probe.py one 8192 8192 1 6 0 60)"
done
