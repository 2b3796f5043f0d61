#!/bin/bash
# compute-sanitizer over the kernels added at the end of round 2 (ring / pair deposit,
# peer send, counted insert through the single-sync gapped push)
T=${1:-r2san}
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "deposit or peer_send or gapped" -p no:cacheprovider ) > gpurun_out/${T}_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/${T}_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "test_deposit" -p no:cacheprovider ) > gpurun_out/${T}_racecheck.log 2>&1
echo "rc=$?" >> gpurun_out/${T}_racecheck.log
tail -4 gpurun_out/${T}_memcheck.log; tail -4 gpurun_out/${T}_racecheck.log
