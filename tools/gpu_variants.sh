#!/bin/bash
# dev: compare builds of the library (tools/_prof/lib_*.so) on the 1024x2048 probe
for f in tools/_prof/lib_*.so; do
  export MN_LIB_PATH=$PWD/$f
  echo "== $f"
  timeout 600 python tools_probe.py 1024 2048 2 2>&1 | grep -E "iter 1|ninst" | cut -c1-40,200-330 | tr '\n' ' '; echo
done
