#!/bin/bash
# push variants: gapped tests + first-launch time + sustained bench
T=${1:-r2v3}
mkdir -p gpurun_out
for f in skeletor_b200/lib/variants/lib_*.so; do
  n=$(basename $f .so)
  ( SKELETOR_B200_LIB=$PWD/$f timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gapped" -p no:cacheprovider ) > gpurun_out/${T}_${n}_tests.log 2>&1
  tail -1 gpurun_out/${T}_${n}_tests.log
  ( SKELETOR_B200_LIB=$PWD/$f REPEAT=8 timeout 300 python tools/one_push.py ) > gpurun_out/${T}_${n}_onepush.log 2>&1
  tail -2 gpurun_out/${T}_${n}_onepush.log
  ( SKELETOR_B200_LIB=$PWD/$f timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/${T}_${n}_bench.json 2> gpurun_out/${T}_${n}_bench.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_${n}_bench.json").read().strip().splitlines()[-1])
    k=d["kernels"]; c=d.get("checks",{})
    print("$n: step", round(d["ms_per_step"],3), {a:(b.get("live_ms") or b.get("ms")) for a,b in k.items() if isinstance(b,dict)}, c.get("particles_bitexact"), c.get("sources_rel"))
except Exception as e:
    print("$n failed", e)
PY
done
( REPEAT=8 timeout 300 python tools/one_push.py ) > gpurun_out/${T}_default_onepush.log 2>&1
tail -2 gpurun_out/${T}_default_onepush.log
