for F in ${SW_FLAGS_LIST:-6 4}; do
  echo "== PLF_SW_FLAGS=$F"
  PLF_SW_FLAGS=$F timeout 300 python tools/parity_sweep.py ${SW_SWEEP_N:-32} 30000 rect 0 752 480 1 2>&1 | grep -v "^sw img" | tail -2
done
