#!/bin/bash
# last check of a tree before a round ends: the GPU test tier, smoke(), the headline line with the driver's flags
set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-protocol --no-states --no-cpu-baseline > gpurun_out/q_final_k20.log 2>&1; tail -1 gpurun_out/q_final_k20.log | cut -c1-200
