#!/bin/bash
# GPU suite + smoke (what the driver runs at round end)
tag=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/${tag}_smoke.log
