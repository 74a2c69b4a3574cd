#!/bin/bash
# round 2, session 3g: A/B -- partition pass at 3 vs 4 resident blocks per SM; L2 fetch granularity 32 / 64 / 128 with 148 images in flight
tag=${1:-s3g}
mkdir -p gpurun_out
timeout 300 python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_lp3.log 2>&1; echo "lp3:"; grep "iter 1" gpurun_out/${tag}_lp3.log | sed 's/.*"aggregate_ms": \([0-9.]*\).*/aggregate_ms \1/'
MN_LIB_PATH=$PWD/build_variants/lp4.so timeout 300 python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_lp4.log 2>&1; echo "lp4:"; grep "iter 1" gpurun_out/${tag}_lp4.log | sed 's/.*"aggregate_ms": \([0-9.]*\).*/aggregate_ms \1/'
for g in default 32 128; do
  if [ $g = default ]; then unset MN_L2_FETCH; else export MN_L2_FETCH=$g; fi
  timeout 600 python tools/probe.py 1024 2048 148 > gpurun_out/${tag}_l2_${g}.log 2>&1
  echo "L2 fetch $g:"; grep "L2 fetch" gpurun_out/${tag}_l2_${g}.log | head -1; grep "iter 1" gpurun_out/${tag}_l2_${g}.log | sed 's/.*"edge_ms": \([0-9.]*\).*"merge_ms": \([0-9.]*\).*/edge_ms \1 merge_ms \2/'
done
