#!/bin/bash
mkdir -p gpurun_out
( REPEAT=12 timeout 300 python tools/one_push.py ) > gpurun_out/steady_a.log 2>&1
( REPEAT=12 PAUSE=0.5 timeout 300 python tools/one_push.py ) > gpurun_out/steady_b.log 2>&1
tail -2 gpurun_out/steady_a.log; tail -2 gpurun_out/steady_b.log
