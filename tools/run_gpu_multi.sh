#!/bin/bash
# N-GPU lines of round 2 (run under `gpurun --gpus N`): the headline bench line and the mixed-game sweep (BASELINE configs[4]) under torchrun
N=${1:-2}
TAG=${2:-r2}
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $R --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_${N}gpu.log 2>&1; grep '^{' gpurun_out/${TAG}_bench_${N}gpu.log | tail -1 | cut -c1-300
timeout 900 $R --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-protocol --no-states > gpurun_out/${TAG}_bench_${N}gpu_driver_flags.log 2>&1; grep '^{' gpurun_out/${TAG}_bench_${N}gpu_driver_flags.log | tail -1 | cut -c1-200
timeout 900 $R --master-port 29513 bench.py --gpus $N --mixed 1048576 --steps 512 --warmup 10 > gpurun_out/${TAG}_bench_cfg5_mixed_${N}gpu.log 2>&1; grep '^{' gpurun_out/${TAG}_bench_cfg5_mixed_${N}gpu.log | tail -1 | cut -c1-300
timeout 600 $R --master-port 29514 bench.py --impl reference --gpus $N --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_${N}gpu_reference_arm.log 2>&1; grep '^{' gpurun_out/${TAG}_bench_${N}gpu_reference_arm.log | tail -1 | cut -c1-200
