#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_libm_parity.py tests/test_gpu_parity.py -m gpu -q -x -k "warp_pipeline or edge or drop_in_c_abi or batch_api" > gpurun_out/e2b_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/e2b_pytest.log
for cfg in auto "7,2" "6,3" "5,4"; do
  if [ "$cfg" = auto ]; then unset MN_EDGE2_NCONS; else export MN_EDGE2_NCONS="$cfg"; fi
  timeout 300 python tools/edge_bench.py 1024 2048 9 10 16 20 2>&1 | grep edge_bench
done | tee gpurun_out/e2b_edge_bench.log
