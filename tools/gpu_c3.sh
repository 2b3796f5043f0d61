#!/bin/bash
# new unit tests + launch list of config 3 (Horowitz iterate)
T=${1:-r2c3}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "peer_send or kernel_variants" -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log
( timeout 600 python bench.py --config 3 --no-cpu-baseline ) > gpurun_out/${T}_config3.json 2> gpurun_out/${T}_config3.err
head -c 400 gpurun_out/${T}_config3.json; echo
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --config 3 --steps 2 --warmup 2 --no-cpu-baseline --no-parity ) > gpurun_out/${T}_launches.log 2>&1
tail -2 gpurun_out/${T}_launches.log | cut -c1-300
