#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/tp_pytest.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/tp_pytest.log
timeout 600 python tests/manual/cfg1_vs_reference.py 2>&1 | tail -4 | cut -c1-300
timeout 1400 python tools/gpu_robust.py 2>&1 | cut -c1-330
bash tools/gpu_probe.sh full 2>&1 | grep "iter 1\|pixel_pool" | cut -c1-200
