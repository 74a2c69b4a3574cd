#!/bin/bash
# round 2, session 3: 16-byte records -- GPU suite, probes, bench at the full batch
tag=${1:-s3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python tools_probe.py 256 512 1 > gpurun_out/${tag}_plain_256.log 2>&1; echo "plain256 rc $?"
timeout 600 python tools_probe.py 1024 2048 2 > gpurun_out/${tag}_plain_full.log 2>&1; echo "plainfull rc $?"
grep "iter 1" gpurun_out/${tag}_plain_256.log | cut -c1-300
grep "iter 1" gpurun_out/${tag}_plain_full.log | cut -c1-300
timeout 1200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench.log 2>&1; echo "bench rc $?"
tail -c 3000 gpurun_out/${tag}_bench.log
