#!/bin/bash
# bench line under torchrun on N GPUs of one box
tag=${1:-bm}; n=${2:-8}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err; echo "bench rc $?"
tail -c 1200 gpurun_out/${tag}_bench_${n}gpu.json
