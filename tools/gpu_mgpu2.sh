#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_2gpu.log 2>&1; echo "2gpu rc $?"
tail -1 gpurun_out/bench_2gpu.log | cut -c1-900
