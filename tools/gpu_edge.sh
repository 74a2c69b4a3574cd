#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tp_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/tp_pytest.log
timeout 600 python tools_probe.py 1024 2048 8 2>&1 | grep "iter" | cut -c1-220
