#!/bin/bash
# round 2: per-phase cycle buckets of the merge scheduler (-DMN_PHASE_CYCLES build made here) on one
# 256x512 image and on 1024x2048 images; usage: tools/gpu_r2_prof.sh <tag> [B_full]
tag=${1:-r2}; bfull=${2:-2}
mkdir -p gpurun_out /tmp/mn_prof
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -fmad=false -DMN_PHASE_CYCLES $MN_NVCC_EXTRA \
  -o /tmp/mn_prof/libmergenet_b200_prof.so mergenet_b200/csrc/mn_api.cu || exit 1
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv > gpurun_out/${tag}_gpu.txt
free -g >> gpurun_out/${tag}_gpu.txt; nproc >> gpurun_out/${tag}_gpu.txt
timeout 300 python tools_probe.py 256 512 1 > gpurun_out/${tag}_plain_256.log 2>&1; echo "plain256 rc $?"
MN_LIB_PATH=/tmp/mn_prof/libmergenet_b200_prof.so timeout 300 python tools_probe.py 256 512 1 > gpurun_out/${tag}_prof_256.log 2>&1; echo "prof256 rc $?"
timeout 900 python tools_probe.py 1024 2048 $bfull > gpurun_out/${tag}_plain_full.log 2>&1; echo "plainfull rc $?"
MN_LIB_PATH=/tmp/mn_prof/libmergenet_b200_prof.so timeout 900 python tools_probe.py 1024 2048 $bfull > gpurun_out/${tag}_prof_full.log 2>&1; echo "proffull rc $?"
grep "iter 1" gpurun_out/${tag}_plain_256.log | cut -c1-300
grep "iter 1" gpurun_out/${tag}_plain_full.log | cut -c1-300
