# experiments with the streaming small-batch grower (PLF_SW_FLAGS: 1 owner reads from L2, 2 no parking, 4 statistics, 16 no retry)
for F in ${SW_FLAGS_LIST:-4 6 20 22}; do
  echo "== PLF_SW_FLAGS=$F"
  PLF_SW_FLAGS=$F timeout 200 python tools/parity_sweep.py ${SW_SWEEP_N:-32} 30000 rect 0 752 480 1 2>&1 | grep -v "^sw img" | tail -3
  PLF_SW_FLAGS=$F timeout 100 python tools/latency_stages.py 1 2>&1 | grep "^sw img" | tail -2
  PLF_SW_FLAGS=$F timeout 100 python tools/latency_stages.py 1 2>&1 | grep -o "lsd_grow [0-9.]*"
done
