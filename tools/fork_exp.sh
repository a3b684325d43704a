# A/B of the forked line path in plf_batch_run (PLF_NO_FORK=1: sequential) at several numbers of calls in flight
for C in 8 6 4; do
  for NF in 0 1; do
    if [ $NF = 1 ]; then export PLF_NO_FORK=1; else unset PLF_NO_FORK; fi
    python bench.py --steps 6 --warmup 3 --contexts $C --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('contexts $C nofork $NF value %.0f e2e %.0f ms_per_step %.1f lat1 %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d.get('latency_ms_single_pair')))"
  done
done
