#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "drop_in or golden or batch_api or invariants or coco" 2>&1 | tail -2
timeout 300 python tools_probe.py 256 512 2 2>&1 | grep "iter 1" | cut -c1-200
timeout 600 python tools_probe.py 1024 2048 2 2>&1 | grep "iter 1" | cut -c1-200
