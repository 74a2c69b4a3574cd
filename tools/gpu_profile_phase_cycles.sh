#!/bin/bash
# profiling recipe (round 2): fresh per-phase cycle buckets (profiling build made on the box) + ncu source-level capture of the
# merge kernel on one 256x512 image, exported to CSV on the box
tag=${1:-s2}
mkdir -p gpurun_out /tmp/mn_prof
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -fmad=false -DMN_PHASE_CYCLES $MN_NVCC_EXTRA \
  -o /tmp/mn_prof/libmergenet_b200_prof.so mergenet_b200/csrc/mn_api.cu || exit 1
timeout 300 python tools/probe.py 256 512 1 > gpurun_out/${tag}_plain_256.log 2>&1; echo "plain256 rc $?"
MN_LIB_PATH=/tmp/mn_prof/libmergenet_b200_prof.so timeout 300 python tools/probe.py 256 512 1 > gpurun_out/${tag}_prof_256.log 2>&1; echo "prof256 rc $?"
timeout 900 python tools/probe.py 1024 2048 2 > gpurun_out/${tag}_plain_full.log 2>&1; echo "plainfull rc $?"
MN_LIB_PATH=/tmp/mn_prof/libmergenet_b200_prof.so timeout 900 python tools/probe.py 1024 2048 2 > gpurun_out/${tag}_prof_full.log 2>&1; echo "proffull rc $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:mn_merge -c 1 -o gpurun_out/${tag}_merge_256 -f python tools/probe.py 256 512 1 > gpurun_out/${tag}_ncu_merge.log 2>&1
echo "ncu rc $?"
ncu -i gpurun_out/${tag}_merge_256.ncu-rep --page source --csv --print-source cuda > gpurun_out/${tag}_merge_src_cuda.csv 2>/dev/null
ncu -i gpurun_out/${tag}_merge_256.ncu-rep --page raw --csv > gpurun_out/${tag}_merge_raw.csv 2>/dev/null
grep "iter 1" gpurun_out/${tag}_plain_256.log | cut -c1-200
grep "iter 1" gpurun_out/${tag}_plain_full.log | cut -c1-200
