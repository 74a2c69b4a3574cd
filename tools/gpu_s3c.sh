#!/bin/bash
# round 2, session 3c: GPU suite (tensor-map tile kernel, bilinear resize, partition pass), edge-pass A/B on cfg4 / cfg2,
# partition-pass timing at batch 16 + its ncu capture
tag=${1:-s3c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/${tag}_pytest.log
{
echo "== cfg4 tile kernel, tensor-map loads"; python tools/edge_bench.py 512 512 81 16 32 20
echo "== cfg4 tile kernel, 1-D bulk copies"; MN_EDGE_NO_TMAP=1 python tools/edge_bench.py 512 512 81 16 32 20
echo "== cfg2 tile kernel, tensor-map loads"; MN_EDGE2_NCONS=0 python tools/edge_bench.py 1024 2048 9 10 16 20
echo "== cfg2 tile kernel, 1-D bulk copies"; MN_EDGE2_NCONS=0 MN_EDGE_NO_TMAP=1 python tools/edge_bench.py 1024 2048 9 10 16 20
echo "== cfg2 warp pipeline"; python tools/edge_bench.py 1024 2048 9 10 16 20
} > gpurun_out/${tag}_edge.log 2>&1
cat gpurun_out/${tag}_edge.log
timeout 600 python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_probe16.log 2>&1; grep "iter 1" gpurun_out/${tag}_probe16.log | cut -c1-420
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mn_partition_logprob_kernel -c 1 -o gpurun_out/${tag}_logprob -f \
  python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_ncu_logprob.log 2>&1; echo "ncu logprob rc $?"
ncu -i gpurun_out/${tag}_logprob.ncu-rep --page raw --csv > gpurun_out/${tag}_logprob_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_logprob.ncu-rep
