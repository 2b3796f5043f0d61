#!/bin/bash
# instruction count + time of the push kernel (1024^2 x 256 ppc) and the smoke / gapped tests
T=${1:-r2u}
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gapped" -p no:cacheprovider ) > gpurun_out/${T}_gapped_tests.log 2>&1
A="--nx 1024 --ny 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
( timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:cell_stream -s 3 -c 2 --csv --log-file gpurun_out/${T}_inst.csv python bench.py $A ) > gpurun_out/${T}_inst.log 2>&1
( timeout 300 python tools/one_push.py ) > gpurun_out/${T}_onepush.log 2>&1
tail -2 gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_gapped_tests.log; grep -E "inst_executed|time_duration|dram__bytes" gpurun_out/${T}_inst.csv | cut -d, -f5,13- | head -8; tail -1 gpurun_out/${T}_onepush.log
