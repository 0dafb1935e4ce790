# development aid: warp-per-item kernels beside / before their thread-per-item neighbours
for env in "X=1" "B200_SEED_NO_OVERLAP=1" "B200_CHAIN_NO_OVERLAP=1" "B200_SEED_NO_OVERLAP=1 B200_CHAIN_NO_OVERLAP=1" "X=2"; do
  echo "== $env"
  env $env python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), [(k['kernel'],k['ms']) for k in d['kernels'] if 'seed' in k['kernel'] or 'chain' in k['kernel']])"; done
