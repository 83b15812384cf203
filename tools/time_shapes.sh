# quick stage timings of the named shapes on sub-blocks (device events inside the library)
for w in "burgers1d 441" "burgers2d 64" "vlasov1d1v 16" "rising_bubble 8"; do set -- $w; echo "== $1 P=$2"; GLASDI_WORKLOAD=$1 python tools/profile_case.py $2 3 3 2>&1 | tail -1 | cut -c1-200; done
