#!/bin/bash
# source-level profile of the merge kernel on one 256x512 image (ncu replays the kernel ~40x)
mkdir -p gpurun_out
CMD="python tools_probe.py 256 512 1"
timeout 300 $CMD > gpurun_out/ncu_merge2_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:mn_merge -c 1 -o gpurun_out/r01_merge_256 -f $CMD > gpurun_out/ncu_merge2.log 2>&1
echo "ncu rc $?"
tail -2 gpurun_out/ncu_merge2.log
