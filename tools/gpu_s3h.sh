#!/bin/bash
# round 2, session 3h: partition pass (pipelined label / class loads) at 4 vs 3 resident blocks; GPU suite; ncu capture
tag=${1:-s3h}
mkdir -p gpurun_out
timeout 300 python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_lp4.log 2>&1; echo "lp4:"; grep "iter 1" gpurun_out/${tag}_lp4.log | sed 's/.*"aggregate_ms": \([0-9.]*\).*/aggregate_ms \1/'
MN_LIB_PATH=$PWD/build_variants/lp3.so timeout 300 python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_lp3.log 2>&1; echo "lp3:"; grep "iter 1" gpurun_out/${tag}_lp3.log | sed 's/.*"aggregate_ms": \([0-9.]*\).*/aggregate_ms \1/'
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mn_partition_logprob_kernel -c 1 -o gpurun_out/${tag}_logprob -f \
  python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_ncu_logprob.log 2>&1; echo "ncu logprob rc $?"
ncu -i gpurun_out/${tag}_logprob.ncu-rep --page raw --csv > gpurun_out/${tag}_logprob_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_logprob.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_logprob_src.csv 2>/dev/null
rm -f gpurun_out/${tag}_logprob.ncu-rep
