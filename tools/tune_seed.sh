# development aid: seed-kernel shared-memory variants
for cfg in "8 0" "12 0" "16 0"; do set -- $cfg
  echo "== lcap $1 shared_read $2"
  B200_SEED_LCAP=$1 B200_SEED_SHARED_READ=$2 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), [(k['kernel'],k['ms']) for k in d['kernels'] if 'seed' in k['kernel'] or 'extend' in k['kernel']])"; done
