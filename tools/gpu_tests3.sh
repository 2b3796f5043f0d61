#!/bin/bash
# smoke + full GPU suite + default bench, with and without the single-sync migration
T=${1:-r2s}
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${T}_tests.log
( timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
( SKELETOR_B200_SINGLE_SYNC=0 timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/${T}_bench_3sync.json 2> gpurun_out/${T}_bench_3sync.err
( timeout 600 python bench.py --config 1 --no-cpu-baseline ) > gpurun_out/${T}_config1.json 2> gpurun_out/${T}_config1.err
( SKELETOR_B200_SINGLE_SYNC=0 timeout 600 python bench.py --config 1 --no-cpu-baseline ) > gpurun_out/${T}_config1_3sync.json 2> gpurun_out/${T}_config1_3sync.err
tail -3 gpurun_out/${T}_smoke.log; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${T}_tests.log | tail -20
python - <<PY
import json
for n in ("bench","bench_3sync","config1","config1_3sync"):
    try:
        d=json.loads(open("gpurun_out/${T}_%s.json"%n).read().strip().splitlines()[-1])
        k=d.get("kernels",{}); r=d.get("roofline",{})
        print(n, "step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e", (d.get("e2e") or {}).get("value"), "frac", r.get("frac"), {a:(b.get("live_ms") or b.get("ms")) for a,b in k.items() if isinstance(b,dict)}, d.get("checks",{}).get("particles_bitexact"), d.get("checks",{}).get("sources_rel"))
    except Exception as e:
        print(n, "failed", e)
PY
