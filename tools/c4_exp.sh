for cfg in "2 4" "2 8" "4 6" "3 8"; do set -- $cfg
  python bench.py --config c4 --steps 5 --warmup 3 --streams $1 --contexts $2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 streams $1 contexts $2 value %.0f e2e %.0f ms_per_step %.1f pairs/step %d grow_alone %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['pairs_per_step'],d['ms_per_stage']['lsd_grow']))"
done
