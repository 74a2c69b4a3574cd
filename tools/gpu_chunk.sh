#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch_api or chunked or drop_in_c_abi or same_different" > gpurun_out/chunk_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/chunk_pytest.log
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/chunk_bench.log 2>&1; echo "bench rc $?"
tail -1 gpurun_out/chunk_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('value',d['value'],'e2e',d['e2e'],'launches',d['gpu_launches'])"
