#!/bin/bash
# N-GPU session for the single-sync migration: parity (gapped) + bench with and without it
N=${1:-2}; T=${2:-r2ss}
mkdir -p gpurun_out
( MGPU_GAPPED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py ) > gpurun_out/${T}_mgpu_check_n${N}_gapped1.log 2>&1
echo "rc=$?" >> gpurun_out/${T}_mgpu_check_n${N}_gapped1.log
A="--gpus $N --steps 10 --warmup 3 --no-cpu-baseline"
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py $A ) > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err
( SKELETOR_B200_SINGLE_SYNC=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py $A ) > gpurun_out/${T}_bench_n${N}_3sync.json 2> gpurun_out/${T}_bench_n${N}_3sync.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py $A --config 4 ) > gpurun_out/${T}_bench_config4_n${N}.json 2> gpurun_out/${T}_bench_config4_n${N}.err
grep -E "OK|FAIL|SKIP|rc=" gpurun_out/${T}_mgpu_check_n${N}_gapped1.log | tail -14
python - <<PY
import json
for n in ("bench_n$N","bench_n${N}_3sync","bench_config4_n$N"):
    try:
        d=json.loads(open("gpurun_out/${T}_%s.json"%n).read().strip().splitlines()[-1])
        k=d.get("kernels",{}); r=d.get("roofline",{})
        print(n, "step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e", (d.get("e2e") or {}).get("value"), {a:(b.get("live_ms") or b.get("ms")) for a,b in k.items() if isinstance(b,dict)}, d.get("checks",{}).get("particles_bitexact"), d.get("checks",{}).get("sources_rel"))
    except Exception as e:
        print(n, "failed", e)
PY
