#!/bin/bash
timeout 900 python -m pytest tests/test_libm_parity.py tests/test_gpu_parity.py -m gpu -q -x -k "libm or device or edge_pass or drop_in_c_abi" 2>&1 | tail -2
timeout 600 python tools_probe.py 1024 2048 8 2>&1 | grep "iter" | cut -c1-130
