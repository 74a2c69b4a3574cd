#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tp_pytest.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/tp_pytest.log
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
timeout 1500 python bench.py --batch $1 --steps 2 --warmup 3 > gpurun_out/bench_own.log 2>&1; echo "bench rc $?"
tail -1 gpurun_out/bench_own.log | cut -c1-3500
