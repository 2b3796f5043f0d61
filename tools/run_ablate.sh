#!/bin/bash
T=${1:-r2e}
mkdir -p gpurun_out
for f in skeletor_b200/lib/variants/lib_*.so; do
  ( SKELETOR_B200_LIB=$PWD/$f timeout 200 python tools/one_push.py ) >> gpurun_out/${T}_ablate.log 2>&1
done
( SKB_GAP_GENERIC=1 timeout 200 python tools/one_push.py ) >> gpurun_out/${T}_ablate.log 2>&1
cat gpurun_out/${T}_ablate.log | grep -v Warn
