#!/bin/bash
# per-phase cycle buckets of the merge scheduler: a -DMN_PHASE_CYCLES build of the library (built here, on the
# box: nvcc is in the image), run on two 1024x2048 images through MN_LIB_PATH
mkdir -p gpurun_out /tmp/mn_prof
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -fmad=false -DMN_PHASE_CYCLES \
  -o /tmp/mn_prof/libmergenet_b200_prof.so mergenet_b200/csrc/mn_api.cu || exit 1
export MN_LIB_PATH=/tmp/mn_prof/libmergenet_b200_prof.so
timeout 600 python tools_probe.py 1024 2048 2 > gpurun_out/prof_full.log 2>&1; echo "prof rc $?"
tail -2 gpurun_out/prof_full.log | head -1 | cut -c1-1800
