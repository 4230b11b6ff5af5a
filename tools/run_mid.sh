# development: mid-size cloths, window width 64 against 128 (oc_k_march2 and oc_k_twin)
for g in "512 512" "1024 1024" "1536 1536"; do
  set -- $g
  for e in 1 0; do
    echo "march2 wc128 $(OC_MARCH2_WC=128 python tools/twin_probe.py one $1 $2 1 3 $e 600)"
    echo "march2 wc64  $(OC_MARCH2_WC=64 python tools/twin_probe.py one $1 $2 1 3 $e 600)"
    echo "twin wc128 $(OC_TWIN_WC=128 python tools/twin_probe.py one $1 $2 1 5 $e 600)"
    echo "twin wc64  $(OC_TWIN_WC=64 python tools/twin_probe.py one $1 $2 1 5 $e 600)"
  done
done
