#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 1 --warmup 1 > gpurun_out/bench_4gpu.log 2>&1; echo "4gpu rc $?"
tail -1 gpurun_out/bench_4gpu.log | cut -c1-400
