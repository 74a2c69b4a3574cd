#!/bin/bash
# dev: compare builds of the library (lib_*.so under $MN_VARIANT_DIR, default /tmp/mn_variants; build them with
# nvcc ... -DMN_REFILL_TARGET=... first) on the 1024x2048 probe
for f in ${MN_VARIANT_DIR:-/tmp/mn_variants}/lib_*.so; do
  export MN_LIB_PATH=$f
  echo "== $f"
  timeout 600 python tools_probe.py 1024 2048 2 2>&1 | grep -E "iter 1|ninst" | cut -c1-40,200-330 | tr '\n' ' '; echo
done
