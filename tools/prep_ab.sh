# prepare() with the default CTA size against forced warps-per-CTA counts (python tools/prep_time.py n dim order k knowns)
for cfg in "1000000 2 4 30 0" "1000000 2 3 24 1" "1000000 2 2 12 1" "2000000 1 3 8 1" "500000 3 3 40 0" "1000000 3 2 20 1" "2000000 1 4 9 0" "2000000 2 1 6 0" "2000000 3 1 8 0"; do
  for w in 16 20 0; do
    if [ $w = 0 ]; then unset WLSQM_PREP_WARPS; else export WLSQM_PREP_WARPS=$w; fi
    echo -n "warps=$w: "; python tools/prep_time.py $cfg
  done
done
unset WLSQM_PREP_WARPS
