#!/bin/bash
# last look at the committed state: smoke, gapped + deposit tests, short default bench
T=${1:-r2last}
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gapped or deposit or peer" -p no:cacheprovider ) > gpurun_out/${T}_tests.log 2>&1
( timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_smoke.log; tail -1 gpurun_out/${T}_tests.log
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
k=d["kernels"]; print("step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "frac", d["roofline"]["frac"], {a:(b.get("live_ms") or b.get("ms")) for a,b in k.items() if isinstance(b,dict)}, d["checks"]["particles_bitexact"], d["checks"]["sources_rel"])
PY
