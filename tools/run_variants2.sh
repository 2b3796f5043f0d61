#!/bin/bash
T=${1:-r2v}
mkdir -p gpurun_out
A="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
for f in skeletor_b200/lib/variants/lib_*.so; do
  n=$(basename $f .so)
  ( SKELETOR_B200_LIB=$PWD/$f timeout 300 python bench.py $A ) > gpurun_out/${T}_${n}.json 2> gpurun_out/${T}_${n}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_lib_*json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d['kernels']
        print(f.split('/')[-1], 'step %.2f'%d['ms_per_step'], {a:(b.get('live_ms') or b.get('ms')) for a,b in k.items()})
    except Exception as e:
        print(f,'ERR',e)
PY
