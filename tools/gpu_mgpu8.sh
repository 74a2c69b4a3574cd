#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 1 --warmup 1 > gpurun_out/bench_8gpu.log 2>&1; echo "8gpu rc $?"
tail -1 gpurun_out/bench_8gpu.log | cut -c1-1500
