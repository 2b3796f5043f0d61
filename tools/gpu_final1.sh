#!/bin/bash
# final single-GPU numbers: default bench (config 5), reference arm, configs 1-4, TSC
T=${1:-r2z}
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
( timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${T}_reference.json 2> gpurun_out/${T}_reference.err
for C in 1 2 3 4; do
  ( timeout 600 python bench.py --config $C --steps 10 --warmup 3 ) > gpurun_out/${T}_config$C.json 2> gpurun_out/${T}_config$C.err
done
( SKB_DEP_HALF=0 timeout 600 python bench.py --config 2 --steps 10 --warmup 3 --no-parity ) > gpurun_out/${T}_config2_nohalf.json 2> gpurun_out/${T}_config2_nohalf.err
( timeout 600 python bench.py --order 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/${T}_tsc.json 2> gpurun_out/${T}_tsc.err
( timeout 600 python bench.py --nx 1024 --ny 1024 --ppc 16 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/${T}_ppc16.json 2> gpurun_out/${T}_ppc16.err
( SKB_DEP_HALF=0 timeout 600 python bench.py --nx 1024 --ny 1024 --ppc 16 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity ) > gpurun_out/${T}_ppc16_nohalf.json 2> gpurun_out/${T}_ppc16_nohalf.err
for f in gpurun_out/${T}_*.json; do echo $f; head -c 200 $f; echo; done
