#!/bin/bash
# source-level profile of the merge kernel on a small image (ncu replays the kernel ~40x)
set -x
mkdir -p gpurun_out
CMD="python tools_probe.py 128 256 1"
timeout 300 $CMD > gpurun_out/ncu_merge_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mn_merge -c 1 -o gpurun_out/merge_prof -f $CMD > gpurun_out/ncu_merge.log 2>&1
echo "ncu rc $?"
tail -3 gpurun_out/ncu_merge_plain.log
ls -la gpurun_out
