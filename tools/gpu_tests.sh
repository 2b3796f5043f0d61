#!/bin/bash
T=${1:-r2p}
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${T}_tests.log
( timeout 600 python bench.py --config 3 --steps 6 --warmup 3 ) > gpurun_out/${T}_config3.json 2> gpurun_out/${T}_config3.err
tail -3 gpurun_out/${T}_smoke.log; grep -E "^FAILED|passed|failed" gpurun_out/${T}_tests.log | tail -20; head -c 300 gpurun_out/${T}_config3.json
