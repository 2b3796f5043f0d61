#!/bin/bash
# bench every library variant (unfused step; PD=0 kernel) and the fused step of a few
T=${1:-r2d}
mkdir -p gpurun_out
A="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
for f in skeletor_b200/lib/variants/lib_*.so; do
  n=$(basename $f .so)
  ( SKELETOR_B200_LIB=$PWD/$f SKELETOR_B200_FUSE=0 timeout 300 python bench.py $A ) > gpurun_out/${T}_${n}_unfused.json 2> gpurun_out/${T}_${n}_unfused.err
  ( SKELETOR_B200_LIB=$PWD/$f timeout 300 python bench.py $A ) > gpurun_out/${T}_${n}_fused.json 2> gpurun_out/${T}_${n}_fused.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_*json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; nm=[x for x in k if x.startswith('push')][0]
        print(f.split('/')[-1], 'step %.2f'%d['ms_per_step'], nm, k[nm]['ms'], k[nm].get('live_ms'))
    except Exception as e:
        print(f,'ERR',e)
PY
