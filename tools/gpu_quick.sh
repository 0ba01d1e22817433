#!/bin/bash
# quick single-GPU check: the area-kernel and parity tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_area_kernels.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
