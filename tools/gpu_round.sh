#!/bin/bash
# one GPU session: smoke, GPU tests, bench (default = cell-stream kernel, fused deposit), A/B
mkdir -p gpurun_out
T=${1:-r2a}
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
( timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${T}_tests.log
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
( SKELETOR_B200_FUSE=0 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity ) > gpurun_out/${T}_bench_unfused.json 2> gpurun_out/${T}_bench_unfused.err
( SKB_GAP_GENERIC=1 SKELETOR_B200_FUSE=0 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity ) > gpurun_out/${T}_bench_generic.json 2> gpurun_out/${T}_bench_generic.err
tail -3 gpurun_out/${T}_smoke.log; tail -5 gpurun_out/${T}_tests.log; head -c 600 gpurun_out/${T}_bench.json
