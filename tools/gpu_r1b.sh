#!/bin/bash
# parity tests + quick probes of the scheduler (cycle buckets) at 256x512 and 1024x2048
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1b_pytest.log 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/r1b_pytest.log
timeout 300 python tools_probe.py 256 512 4 > gpurun_out/r1b_probe_256.log 2>&1; echo "probe256 rc $?"
tail -3 gpurun_out/r1b_probe_256.log
timeout 300 python tools_probe.py 256 512 4 oracle > gpurun_out/r1b_probe_256o.log 2>&1; echo "probe256o rc $?"
tail -3 gpurun_out/r1b_probe_256o.log
timeout 600 python tools_probe.py 1024 2048 2 > gpurun_out/r1b_probe_full.log 2>&1; echo "probefull rc $?"
tail -3 gpurun_out/r1b_probe_full.log
