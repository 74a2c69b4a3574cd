#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/fix_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/fix_pytest.log
timeout 600 python tools_probe.py 1024 2048 2 2>&1 | grep -E "iter 1|ninst" | cut -c1-200
