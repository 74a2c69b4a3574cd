#!/bin/bash
mkdir -p gpurun_out
export MN_LIB_PATH=$PWD/tools/_prof/libmergenet_b200_prof.so
timeout 600 python tools_probe.py 1024 2048 2 > gpurun_out/prof_full.log 2>&1; echo "prof rc $?"
tail -2 gpurun_out/prof_full.log | head -1 | cut -c1-1800
