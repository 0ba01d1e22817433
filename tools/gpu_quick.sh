#!/bin/bash
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:_direct -s 6 -c 1 -o gpurun_out/q_prof_brk4 python bench.py --game breakout --steps 4 --warmup 3 $B > gpurun_out/q_ncu_brk4.log 2>&1
