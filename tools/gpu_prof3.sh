#!/bin/bash
# ncu: gap_insert at 1024^2 x 256 ppc; config 2 (1024^2 x 64 ppc) push + deposit
T=${1:-r2p3}
mkdir -p gpurun_out
A="--nx 1024 --ny 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gap_insert" -s 6 -c 1 -f -o gpurun_out/${T}_ins python bench.py $A ) > gpurun_out/${T}_ins.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cell_stream|deposit_cells" -s 6 -c 2 -f -o gpurun_out/${T}_c2 python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-parity ) > gpurun_out/${T}_c2.log 2>&1
ls -la gpurun_out/${T}*
