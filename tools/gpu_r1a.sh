#!/bin/bash
# round-1 first checkpoint: parity tests, first bench line, ncu launch list + edge-kernel capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r1a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1a_pytest.log 2>&1; echo "pytest rc $?"
timeout 900 python bench.py --batch 96 --steps 1 --warmup 3 > gpurun_out/r1a_bench.log 2>&1; echo "bench rc $?"
tail -2 gpurun_out/r1a_bench.log
CMD="python bench.py --batch 4 --height 256 --width 512 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r1a_plain_small.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1a_launches_small.csv $CMD > gpurun_out/r1a_ncu_small.log 2>&1
echo "ncu small rc $?"
CMD2="python bench.py --batch 2 --steps 1 --warmup 0 --no-cpu-baseline"
timeout 300 $CMD2 > gpurun_out/r1a_plain_full.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mn_edge -c 1 -o gpurun_out/r1a_edge $CMD2 > gpurun_out/r1a_ncu_full.log 2>&1
echo "ncu full rc $?"
ls -la gpurun_out
