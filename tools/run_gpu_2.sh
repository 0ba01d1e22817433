set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench21_2gpu.log 2>&1; tail -1 gpurun_out/bench21_2gpu.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench21_2gpu_ref.log 2>&1; tail -1 gpurun_out/bench21_2gpu_ref.log | cut -c1-300
