# development aid: extension variants (cell width; cells of a DP row per visit of the row site)
for env in "X=1" "B200_EXT_WIDE_CELLS=1" "B200_TX_CHUNK=0" "B200_TX_CHUNK=64"; do
  echo "== $env"
  env $env python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), [(k['kernel'],k['ms']) for k in d['kernels'] if 'extend' in k['kernel'] or 'heavy_prep' in k['kernel']])"; done
