#!/bin/bash
# final confirmation on one GPU, the way the driver runs it: smoke, the GPU suite, the
# default bench, the reference arm, config 4
T=${1:-r2fin}
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
( timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${T}_tests.log
( timeout 900 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
( timeout 900 python bench.py --impl reference ) > gpurun_out/${T}_reference.json 2> gpurun_out/${T}_reference.err
( timeout 600 python bench.py --config 4 --no-cpu-baseline ) > gpurun_out/${T}_config4.json 2> gpurun_out/${T}_config4.err
tail -3 gpurun_out/${T}_smoke.log; grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/${T}_tests.log | tail -8
python - <<PY
import json
for n in ("bench","reference","config4"):
    try:
        d=json.loads(open("gpurun_out/${T}_%s.json"%n).read().strip().splitlines()[-1])
        k=d.get("kernels",{}) or {}; r=d.get("roofline",{}) or {}
        print(n, "step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e", (d.get("e2e") or {}).get("value"), "frac", r.get("frac"), {a:(b.get("live_ms") or b.get("ms")) for a,b in k.items() if isinstance(b,dict)}, (d.get("checks") or {}).get("particles_bitexact"), (d.get("checks") or {}).get("sources_rel"), "launches", d.get("gpu_launches"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
