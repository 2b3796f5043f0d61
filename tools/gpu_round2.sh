#!/bin/bash
# tests + bench (fused / unfused) + ncu of the PD=0 and PD=3 kernels at 1024^2
T=${1:-r2c}
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
( timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${T}_tests.log
( SKELETOR_B200_FUSE=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline ) > gpurun_out/${T}_bench_unfused.json 2> gpurun_out/${T}_bench_unfused.err
( timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/${T}_bench_fused.json 2> gpurun_out/${T}_bench_fused.err
A="--nx 1024 --ny 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
( SKELETOR_B200_FUSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_stream -s 3 -c 1 -f -o gpurun_out/${T}_pd0 python bench.py $A ) > gpurun_out/${T}_pd0.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_stream -s 3 -c 1 -f -o gpurun_out/${T}_pd3 python bench.py $A ) > gpurun_out/${T}_pd3.log 2>&1
tail -3 gpurun_out/${T}_smoke.log; tail -4 gpurun_out/${T}_tests.log; head -c 300 gpurun_out/${T}_bench_unfused.json
