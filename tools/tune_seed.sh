# development aid: seed-kernel occupancy variants (libraries built with -DSEED_MIN_BLOCKS=n)
for mb in 5 6 8; do for cfg in "8 0" "8 1" "16 1" "4 0"; do set -- $cfg
  echo "== min_blocks $mb lcap $1 shared_read $2"
  LIBBWA_PATH=$PWD/gatk-bwamem-jni_b200/lib/libbwa_b200_mb$mb.so B200_SEED_LCAP=$1 B200_SEED_SHARED_READ=$2 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), [(k['kernel'],k['ms']) for k in d['kernels'] if 'seed' in k['kernel'] or 'sal' in k['kernel']])"; done; done
