#!/bin/bash
# gap_insert_counted variants: kernel time at 1024^2 x 256 ppc under ncu (time only) + gapped tests
T=${1:-r2ins}
mkdir -p gpurun_out
A="--nx 1024 --ny 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
for f in default skeletor_b200/lib/variants/lib_*.so; do
  n=$(basename $f .so)
  if [ $f = default ]; then L=""; else L=$PWD/$f; fi
  ( SKELETOR_B200_LIB=$L timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gapped" -p no:cacheprovider ) > gpurun_out/${T}_${n}_tests.log 2>&1
  ( SKELETOR_B200_LIB=$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gap_insert_counted -s 6 -c 4 --csv --log-file gpurun_out/${T}_${n}.csv python bench.py $A ) > gpurun_out/${T}_${n}.log 2>&1
  echo "$n: $(tail -1 gpurun_out/${T}_${n}_tests.log) :: $(grep time_duration gpurun_out/${T}_${n}.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
