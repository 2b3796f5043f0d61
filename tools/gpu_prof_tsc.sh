#!/bin/bash
# ncu captures of the TSC kernels (order 2) at 1024^2 x 256 ppc: gapped push and deposit
T=${1:-r2tsc}
mkdir -p gpurun_out
A="--nx 1024 --ny 1024 --order 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cell_stream|deposit_cells" -s 6 -c 2 -f -o gpurun_out/${T} python bench.py $A ) > gpurun_out/${T}.log 2>&1
tail -3 gpurun_out/${T}.log; ls -la gpurun_out/${T}*
