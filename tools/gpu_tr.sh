#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tp_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/tp_pytest.log
timeout 1400 python tools/gpu_robust.py
bash tools/gpu_probe.sh full
