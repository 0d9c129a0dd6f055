#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_trace' -s 7 -c 4 \
    -o gpurun_out/prof_r01_trace3 -f python scripts/perf_probe.py glossy 1000000 1920 1080 8 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
