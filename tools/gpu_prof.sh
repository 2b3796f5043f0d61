#!/bin/bash
# ncu captures of the cell-stream kernel at 1024^2 x 256 ppc (PD=3 fused and PD=0)
T=${1:-r2b}
mkdir -p gpurun_out
A="--nx 1024 --ny 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:cell_stream -s 3 -c 1 -f -o gpurun_out/${T}_pd3 python bench.py $A ) > gpurun_out/${T}_pd3.log 2>&1
( SKELETOR_B200_FUSE=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:cell_stream -s 3 -c 1 -f -o gpurun_out/${T}_pd0 python bench.py $A ) > gpurun_out/${T}_pd0.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log; ls -la gpurun_out/${T}_*
