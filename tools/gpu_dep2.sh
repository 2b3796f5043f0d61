#!/bin/bash
# ring deposit variants (TSC) + ncu of the default ring kernels at 1024^2
T=${1:-r2dep2}
mkdir -p gpurun_out
A="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
for f in default skeletor_b200/lib/variants/lib_D_nst2.so skeletor_b200/lib/variants/lib_D_nst4.so; do
  n=$(basename $f .so)
  for o in 2 1; do
    if [ $f = default ]; then L=""; else L=$PWD/$f; fi
    ( SKELETOR_B200_LIB=$L SKB_DEP_RING=1 timeout 600 python bench.py $A --order $o ) > gpurun_out/${T}_${n}_o$o.json 2> gpurun_out/${T}_${n}_o$o.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_${n}_o$o.json").read().strip().splitlines()[-1])
    k=d["kernels"]; print("$n order $o: step", round(d["ms_per_step"],2), "push", k["push"]["live_ms"], "deposit", k["deposit"]["live_ms"], k["deposit"]["live_frac"])
except Exception as e:
    print("$n order $o failed", e)
PY
  done
done
B="--nx 1024 --ny 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
( SKB_DEP_RING=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deposit_cells" -s 3 -c 1 -f -o gpurun_out/${T}_o2 python bench.py $B --order 2 ) > gpurun_out/${T}_ncu_o2.log 2>&1
( SKB_DEP_RING=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deposit_cells" -s 3 -c 1 -f -o gpurun_out/${T}_o1 python bench.py $B --order 1 ) > gpurun_out/${T}_ncu_o1.log 2>&1
ls -la gpurun_out/${T}*ncu-rep
