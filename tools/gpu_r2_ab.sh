#!/bin/bash
# round 2 A/B: every build_variants/*.so (same library, different -D switches, built in the dev container) timed on
# one 256x512 image and on two 1024x2048 images; usage: tools/gpu_r2_ab.sh <tag> [variants...]
tag=${1:-ab}; shift
mkdir -p gpurun_out
vars=${@:-$(ls build_variants/*.so | xargs -n1 basename | sed 's/\.so$//')}
for v in $vars; do
  export MN_LIB_PATH=$PWD/build_variants/$v.so
  timeout 300 python tools/probe.py 256 512 1 > gpurun_out/${tag}_${v}_256.log 2>&1
  timeout 600 python tools/probe.py 1024 2048 2 > gpurun_out/${tag}_${v}_full.log 2>&1
  a=$(grep "iter 1" gpurun_out/${tag}_${v}_256.log | sed 's/.*"merge_ms": \([0-9.]*\).*/\1/')
  b=$(grep "iter 1" gpurun_out/${tag}_${v}_full.log | sed 's/.*"merge_ms": \([0-9.]*\).*/\1/')
  n=$(grep ninst gpurun_out/${tag}_${v}_full.log)
  echo "$v merge_ms 256x512: $a  1024x2048(B=2): $b  $n" | tee -a gpurun_out/${tag}_summary.txt
done
