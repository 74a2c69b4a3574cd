#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, default bench (own + reference arm)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/final_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc $?"
tail -1 gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc $?"
tail -1 gpurun_out/final_bench_ref.log | cut -c1-600
timeout 1500 python bench.py > gpurun_out/final_bench.log 2>&1; echo "bench rc $?"
tail -1 gpurun_out/final_bench.log | cut -c1-3500
