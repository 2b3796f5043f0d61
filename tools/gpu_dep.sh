#!/bin/bash
# ring deposit: tests + bench (TSC default ring vs old; CIC ring vs old)
T=${1:-r2dep}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "deposit or tsc or scenario or golden" -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log
A="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for v in "2 1" "2 0" "1 1" "1 0"; do
  set -- $v
  ( SKB_DEP_RING=$2 timeout 600 python bench.py $A --order $1 ) > gpurun_out/${T}_o$1_r$2.json 2> gpurun_out/${T}_o$1_r$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_o$1_r$2.json").read().strip().splitlines()[-1])
    k=d["kernels"]; print("order $1 ring $2: step", round(d["ms_per_step"],2), "push", k["push"]["live_ms"], "deposit", k["deposit"]["live_ms"], k["deposit"]["live_frac"], d.get("checks"))
except Exception as e:
    print("order $1 ring $2 failed", e)
PY
done
