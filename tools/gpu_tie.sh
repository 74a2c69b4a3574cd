#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/tie_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/tie_pytest.log
timeout 900 python tools/gpu_robust.py > gpurun_out/robust3.log 2>&1; echo rc $?; cat gpurun_out/robust3.log | cut -c1-330
