#!/bin/bash
# ncu evidence at FULL size (config 5, 1 GPU): launch list of two steps, full-set captures
# of the push (cell_stream_kernel), deposit and insertion kernels
T=${1:-r02}
mkdir -p gpurun_out
A="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${T}.csv python bench.py $A ) > gpurun_out/launches_${T}.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:cell_stream -s 3 -c 1 -f -o gpurun_out/ncu_${T}_push python bench.py $A ) > gpurun_out/ncu_${T}_push.log 2>&1
( timeout 900 ncu --set full --clock-control none -k regex:deposit_cells -s 3 -c 1 -f -o gpurun_out/ncu_${T}_deposit python bench.py $A ) > gpurun_out/ncu_${T}_deposit.log 2>&1
( timeout 900 ncu --set full --clock-control none -k regex:gap_insert -s 6 -c 2 -f -o gpurun_out/ncu_${T}_insert python bench.py $A ) > gpurun_out/ncu_${T}_insert.log 2>&1
ls -la gpurun_out/*${T}*
