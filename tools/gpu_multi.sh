#!/bin/bash
# N-GPU session: N ranks == 1 rank parity (tests/mgpu_check.py, dense + gapped) and bench
N=${1:-2}; T=${2:-r2m}
mkdir -p gpurun_out
for G in 0 1; do
  ( MGPU_GAPPED=$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$G tests/mgpu_check.py ) > gpurun_out/${T}_mgpu_check_n${N}_gapped${G}.log 2>&1
  echo "rc=$?" >> gpurun_out/${T}_mgpu_check_n${N}_gapped${G}.log
done
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --config 4 --steps 10 --warmup 3 ) > gpurun_out/${T}_bench_config4_n${N}.json 2> gpurun_out/${T}_bench_config4_n${N}.err
grep -E "OK|FAIL|rc=" gpurun_out/${T}_mgpu_check_n${N}_gapped*.log | tail -24; head -c 300 gpurun_out/${T}_bench_n${N}.json; echo; head -c 300 gpurun_out/${T}_bench_config4_n${N}.json
