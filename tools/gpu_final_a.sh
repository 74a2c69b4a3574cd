#!/bin/bash
# tests + ncu launch list (small bench) + full captures of the two HBM-bound kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/fa_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/fa_pytest.log
CMD="python bench.py --batch 4 --height 256 --width 512 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/fa_plain_small.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fa_launches_small.csv $CMD > gpurun_out/fa_ncu_small.log 2>&1
echo "ncu launches rc $?"
CMD2="python tools_probe.py 1024 2048 2"
timeout 300 $CMD2 > gpurun_out/fa_plain_full.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mn_edge_pass|mn_logprob" -c 2 -o gpurun_out/fa_hbm_kernels -f $CMD2 > gpurun_out/fa_ncu_full.log 2>&1
echo "ncu full rc $?"
