#!/bin/bash
mkdir -p gpurun_out
CMD="python tools_probe.py 1024 2048 2"
timeout 300 $CMD > gpurun_out/ncu_edge_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mn_edge_pass -c 1 -o gpurun_out/edge_prof -f $CMD > gpurun_out/ncu_edge.log 2>&1
echo "ncu rc $?"
