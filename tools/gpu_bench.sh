#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_own.log 2>&1; echo "bench rc $?"
tail -1 gpurun_out/bench_own.log | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref rc $?"
tail -1 gpurun_out/bench_ref.log | cut -c1-1200
nproc
