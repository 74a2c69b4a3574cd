#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools_probe.py 256 512 2 > gpurun_out/probe_256.log 2>&1; echo "probe256 rc $?"
tail -3 gpurun_out/probe_256.log
if [ "$1" == "full" ]; then
timeout 600 python tools_probe.py 1024 2048 2 > gpurun_out/probe_full.log 2>&1; echo "probefull rc $?"
tail -3 gpurun_out/probe_full.log
fi
