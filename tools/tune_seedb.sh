# development aid: candidates in flight per lane in the backward sweeps (libraries built with -DSEED_BWD_WAYS=n)
for lib in libbwa_b200.so libbwa_b200_h12.so libbwa_b200_h48.so; do
  echo "== $lib"
  LIBBWA_PATH=$PWD/gatk-bwamem-jni_b200/lib/$lib python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), [(k['kernel'],k['ms']) for k in d['kernels'] if 'seedB' in k['kernel']])"; done
