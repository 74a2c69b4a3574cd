#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_post.py tests/test_gpu_parity.py -m gpu -q -x -k "post or rle or resize or logits or edge" > gpurun_out/post_pytest.log 2>&1; echo "pytest rc $?"
tail -15 gpurun_out/post_pytest.log
timeout 600 python tests/manual/post_bench.py > gpurun_out/post_bench.log 2>&1; echo "post_bench rc $?"
tail -6 gpurun_out/post_bench.log
