#!/bin/bash
# launch list of the bench command at full resolution (batch 16: ~130 launches per step), only after
# the same command exited 0 without ncu
mkdir -p gpurun_out
CMD="python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ll_plain.log 2>&1; rc=$?; echo "plain rc $rc"
tail -1 gpurun_out/ll_plain.log | cut -c1-300
if [ $rc -eq 0 ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches_1024x2048_b16.csv $CMD > gpurun_out/ll_ncu.log 2>&1
echo "ncu rc $?"
fi
wc -l gpurun_out/r01_launches_1024x2048_b16.csv
