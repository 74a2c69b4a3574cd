#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mn_edge_warp -s 3 -c 1 -f -o gpurun_out/r01_edge_warp python tools/edge_bench.py 1024 2048 9 10 16 4 > gpurun_out/ncu_e2.log 2>&1; echo "ncu rc $?"
tail -3 gpurun_out/ncu_e2.log
