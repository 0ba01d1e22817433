#!/bin/bash
# quick single-GPU check
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-protocol --no-states"
TBX_LIB_PATH=$PWD/toybox_b200/libtoybox_b200_stats.so timeout 300 python tools/si_stats.py 2000 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_area_kernels.py tests/test_gpu_parity.py -m gpu -q -x -k "amidar or Amidar or corridor or mixed" > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
timeout 300 python bench.py --game amidar --steps 50 --warmup 5 $B > gpurun_out/q_bench_amidar.log 2>&1; tail -1 gpurun_out/q_bench_amidar.log | cut -c1-900
