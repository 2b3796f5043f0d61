#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient")
# usage: tools/gpurun_retry.sh [--gpus N] --timeout S -- 'command'
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
