for cfg in "12 4" "32 8" "64 16" "200 50" "100000 100000"; do set -- $cfg; echo "== seeds $1 chains $2"; B200_LIGHT_SEEDS=$1 B200_LIGHT_CHAINS=$2 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), [(k['kernel'],k['ms']) for k in d['kernels'] if 'extend' in k['kernel']])"; done
