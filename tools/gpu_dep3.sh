#!/bin/bash
# pair deposit kernel: full GPU tests + benches (TSC, config 2, 16 ppc, CIC pair on/off)
T=${1:-r2dp}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${T}_tests.log | tail -12
A="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
run() { # name, env, args
  ( env $2 timeout 600 python bench.py $A $3 ) > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$1.json").read().strip().splitlines()[-1])
    k=d["kernels"]; c=d.get("checks",{})
    print("$1: step", round(d["ms_per_step"],3), {a:(b.get("live_ms") or b.get("ms")) for a,b in k.items() if isinstance(b,dict)}, "dep frac", [b.get("live_frac") or b.get("frac") for a,b in k.items() if "deposit" in a], c.get("particles_bitexact"), c.get("sources_rel"))
except Exception as e:
    print("$1 failed", e)
PY
}
run tsc_pair "SKB_DEP_PAIR=1" "--order 2"
run tsc_ring "SKB_DEP_PAIR=0" "--order 2"
run cic_pair "SKB_DEP_PAIR=1" ""
run cic_ring "SKB_DEP_PAIR=0" ""
run c2_pair "SKB_DEP_PAIR=1" "--config 2"
run c2_half "SKB_DEP_PAIR=0" "--config 2"
run p16_pair "SKB_DEP_PAIR=1" "--nx 1024 --ny 1024 --ppc 16"
run p16_half "SKB_DEP_PAIR=0" "--nx 1024 --ny 1024 --ppc 16"
