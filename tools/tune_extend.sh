# development aid: extension routing variants
for env in "X=1" "B200_TEST_ALL_HEAVY=1" "B200_LIGHT_SEEDS=4 B200_LIGHT_CHAINS=2" "B200_LIGHT_SEEDS=2 B200_LIGHT_CHAINS=1"; do
  echo "== $env"
  env $env python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), [(k['kernel'],k['ms']) for k in d['kernels'] if 'extend' in k['kernel'] or 'heavy_prep' in k['kernel']])"; done
