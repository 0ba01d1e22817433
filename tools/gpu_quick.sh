#!/bin/bash
# quick single-GPU check
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_area_kernels.py -m gpu -q -x -k "si_ or score" > gpurun_out/q_pytest.log 2>&1; tail -5 gpurun_out/q_pytest.log
