#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; echo "2gpu rc $?"
tail -3 gpurun_out/bench_2gpu.log | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "2gpu ref rc $?"
tail -2 gpurun_out/bench_2gpu_ref.log | cut -c1-600
