#!/bin/bash
# (N GPUs) edge pass alone on every GPU at once vs one at a time; bench line under torchrun
tag=${1:-s3e}; n=${2:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/edge_bench_multi.py 64 > gpurun_out/${tag}_edge_multi.log 2>&1; echo "edge multi rc $?"
grep "^rank" gpurun_out/${tag}_edge_multi.log | sort
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err; echo "bench rc $?"
tail -c 1800 gpurun_out/${tag}_bench_${n}gpu.json
