#!/bin/bash
# bench code paths at a small shape: default host budget, and a pretend-small host (BH < B)
tag=${1:-bs}
mkdir -p gpurun_out
timeout 600 python bench.py --height 256 --width 512 --steps 2 --warmup 1 --no-cpu-baseline --no-single-image > gpurun_out/${tag}_a.json 2> gpurun_out/${tag}_a.err; echo "rc $?"
MN_BENCH_HOST_RAM_GB=1.2 timeout 600 python bench.py --height 256 --width 512 --steps 2 --warmup 1 --no-cpu-baseline --no-single-image > gpurun_out/${tag}_b.json 2> gpurun_out/${tag}_b.err; echo "rc $?"
python - <<PY
import json
for t in "ab":
    try:
        d=json.loads(open("gpurun_out/%s_%s.json" % ("${tag}", t)).read().strip().splitlines()[-1])
        print(t, "value", round(d["value"],1), "batch", d["config"]["batch_per_gpu"], "distinct", d["config"]["distinct_images_per_gpu"], "e2e", d["e2e"])
    except Exception as e:
        print(t, "ERR", e); print(open("gpurun_out/%s_%s.err" % ("${tag}", t)).read()[-1500:])
PY
