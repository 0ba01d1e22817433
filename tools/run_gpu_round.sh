set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wrappers.py -m gpu -x -q > gpurun_out/pytest_gpu21w.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu21w.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_wrappers.py > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu21.log
