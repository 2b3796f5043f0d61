#!/bin/bash
# N-GPU session: parity logs + bench with and without the overlapped migration + config 4
N=${1:-8}; T=${2:-r2x}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for G in 0 1; do
  ( MGPU_GAPPED=$G timeout 600 $R --master-port 2951$G tests/mgpu_check.py ) > gpurun_out/${T}_mgpu_check_n${N}_gapped${G}.log 2>&1
  echo "rc=$?" >> gpurun_out/${T}_mgpu_check_n${N}_gapped${G}.log
done
( timeout 600 $R --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err
( SKELETOR_B200_OVERLAP=0 timeout 600 $R --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 --no-parity --no-e2e ) > gpurun_out/${T}_bench_n${N}_nooverlap.json 2> gpurun_out/${T}_bench_n${N}_nooverlap.err
( timeout 600 $R --master-port 29534 bench.py --gpus $N --config 4 --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_config4_n${N}.json 2> gpurun_out/${T}_bench_config4_n${N}.err
grep -E "OK|FAIL|SKIP|rc=" gpurun_out/${T}_mgpu_check_n${N}_gapped*.log | tail -24
for f in gpurun_out/${T}_bench*n${N}*.json; do echo $f; head -c 260 $f; echo; done
