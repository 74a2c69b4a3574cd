#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_libm_parity.py tests/test_gpu_parity.py -m gpu -q -x -k "warp_pipeline or edge or drop_in_c_abi or batch_api" > gpurun_out/e2c_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/e2c_pytest.log
run() { timeout 300 python tools/edge_bench.py 1024 2048 9 10 16 20 2>&1 | grep "clip=0" | sed "s/^/stages=${MN_EDGE2_STAGES:-2} /"; }
( unset MN_EDGE2_NCONS; run
  export MN_EDGE2_STAGES=3; MN_EDGE2_NCONS="7,2" run; MN_EDGE2_NCONS="6,2" run; MN_EDGE2_NCONS="5,3" run; MN_EDGE2_NCONS="4,4" run
  export MN_EDGE2_STAGES=4; MN_EDGE2_NCONS="6,2" run; MN_EDGE2_NCONS="4,3" run
) | tee gpurun_out/e2c_edge_bench.log
