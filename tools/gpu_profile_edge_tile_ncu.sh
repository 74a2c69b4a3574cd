#!/bin/bash
# profiling recipe (round 2): GPU suite, bench line, Mode B wall time, ncu capture of the tile edge kernel (tensor-map loads) on cfg4
tag=${1:-s3i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"
tail -c 1500 gpurun_out/${tag}_bench.json
timeout 600 python tools/modeb_time.py > gpurun_out/${tag}_modeb.log 2>&1; cat gpurun_out/${tag}_modeb.log | tail -4
timeout 600 ncu --set full --clock-control none -k regex:mn_edge_pass_kernel -c 1 -o gpurun_out/${tag}_edge_tile -f \
  python tools/edge_bench.py 512 512 81 16 32 3 > gpurun_out/${tag}_ncu_edge.log 2>&1; echo "ncu edge rc $?"
ncu -i gpurun_out/${tag}_edge_tile.ncu-rep --page raw --csv > gpurun_out/${tag}_edge_tile_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_edge_tile.ncu-rep --page details > gpurun_out/${tag}_edge_tile_details.txt 2>/dev/null
rm -f gpurun_out/${tag}_edge_tile.ncu-rep
