#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rc $?"
tail -4 gpurun_out/r1d_pytest.log
timeout 300 python tools_probe.py 256 512 4 > gpurun_out/r1d_probe_256.log 2>&1; echo "probe256 rc $?"
tail -3 gpurun_out/r1d_probe_256.log
timeout 600 python tools_probe.py 1024 2048 2 > gpurun_out/r1d_probe_full.log 2>&1; echo "probefull rc $?"
tail -3 gpurun_out/r1d_probe_full.log
timeout 1500 python bench.py --batch 64 --steps 1 --warmup 3 > gpurun_out/r1d_bench.log 2>&1; echo "bench rc $?"
tail -2 gpurun_out/r1d_bench.log
