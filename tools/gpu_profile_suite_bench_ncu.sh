#!/bin/bash
# profiling recipe (round 2): GPU suite, bench line, ncu launch list of the bench command, ncu captures of the partition
# statistics pass (full) and of the merge kernel on the whole GPU (148 CTAs, 512x1024 images: the workspace of 148
# full-size images cannot be saved and restored between ncu's replay passes)
tag=${1:-s3b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
tail -c 2500 gpurun_out/${tag}_bench.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-single-image > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mn_partition_logprob_kernel -c 1 -o gpurun_out/${tag}_logprob -f \
  python tools/probe.py 1024 2048 16 > gpurun_out/${tag}_ncu_logprob.log 2>&1; echo "ncu logprob rc $?"
ncu -i gpurun_out/${tag}_logprob.ncu-rep --page raw --csv > gpurun_out/${tag}_logprob_raw.csv 2>/dev/null
timeout 1500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section SchedulerStats --section LaunchStats \
  --clock-control none -k regex:mn_merge_kernel -c 1 -o gpurun_out/${tag}_merge_b148 -f \
  python tools/probe.py 512 1024 148 > gpurun_out/${tag}_ncu_merge.log 2>&1; echo "ncu merge rc $?"
ncu -i gpurun_out/${tag}_merge_b148.ncu-rep --page raw --csv > gpurun_out/${tag}_merge_b148_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_merge_b148.ncu-rep --page details > gpurun_out/${tag}_merge_b148_details.txt 2>/dev/null
rm -f gpurun_out/${tag}_logprob.ncu-rep
ls -la gpurun_out | tail -12
