#!/bin/bash
# round 2, session 3f: GPU suite + ncu capture of the partition pass (K-specialised build)
tag=${1:-s3f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_probe16.log 2>&1; grep "iter 1" gpurun_out/${tag}_probe16.log | cut -c1-420
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mn_partition_logprob_kernel -c 1 -o gpurun_out/${tag}_logprob -f \
  python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_ncu_logprob.log 2>&1; echo "ncu logprob rc $?"
ncu -i gpurun_out/${tag}_logprob.ncu-rep --page raw --csv > gpurun_out/${tag}_logprob_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_logprob.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_logprob_src.csv 2>/dev/null
rm -f gpurun_out/${tag}_logprob.ncu-rep
