#!/bin/bash
# quick single-GPU check: headline line with the driver's flags and the default, plus ncu captures of the SI / Amidar direct kernels (reps kept)
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/q_bench_k20.log 2>&1; tail -1 gpurun_out/q_bench_k20.log | cut -c1-300
timeout 600 python bench.py $B > gpurun_out/q_bench_k200.log 2>&1; tail -1 gpurun_out/q_bench_k200.log | cut -c1-300
for g in space_invaders amidar; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:_direct -s 6 -c 1 -o gpurun_out/q_prof_$g python bench.py --game $g --steps 4 --warmup 3 $B > gpurun_out/q_ncu_$g.log 2>&1
  ls -la gpurun_out/q_prof_$g.ncu-rep
done
