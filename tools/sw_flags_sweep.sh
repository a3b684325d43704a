# exactness of the streaming grower does not depend on its heuristics: the same sweep with every experiment switch
# (1 owner reads through L1, 2 parking on, 16 no second try, 64 slow commit path only; see lsd_sw.cuh)
for F in 1 2 16 64 19 83; do
  echo "== PLF_SW_FLAGS=$F"
  PLF_SW_FLAGS=$F python tools/parity_sweep.py 64 97000 rect 0 752 480 2 2>&1 | tail -3
  PLF_SW_FLAGS=$F python tools/parity_sweep.py 24 98000 curvy 0 752 480 3 2>&1 | tail -3
done
