# calls in flight x call size at constant device memory (8 x 512 pairs = 16 x 256 pairs)
for cfg in "8 8" "4 16" "4 12" "6 10"; do set -- $cfg
  python bench.py --steps 6 --warmup 3 --streams $1 --contexts $2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $1 contexts $2 value %.0f e2e %.0f ms_per_step %.1f pairs/step %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['pairs_per_step']))"
done
