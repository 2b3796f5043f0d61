#!/bin/bash
# BASELINE configs 2-4 in driver format + TSC and 64-ppc variants of config 5 (1 GPU)
T=${1:-r2n}
mkdir -p gpurun_out
for C in 2 3 4; do
  ( timeout 600 python bench.py --config $C --steps 6 --warmup 3 ) > gpurun_out/${T}_config$C.json 2> gpurun_out/${T}_config$C.err
done
( timeout 600 python bench.py --order 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/${T}_tsc.json 2> gpurun_out/${T}_tsc.err
( timeout 600 python bench.py --nx 1024 --ny 1024 --ppc 64 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/${T}_ppc64.json 2> gpurun_out/${T}_ppc64.err
for f in gpurun_out/${T}_*.json; do echo $f; head -c 250 $f; echo; tail -3 ${f%.json}.err; done
