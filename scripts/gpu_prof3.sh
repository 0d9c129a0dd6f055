#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01b.csv python scripts/perf_probe.py glossy 1000000 1920 1080 8 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_trace|k_surface|k_bounce|k_nee|k_accum|k_regen' -s 40 -c 8 \
    -o gpurun_out/prof_r01_staged -f python scripts/perf_probe.py glossy 1000000 1920 1080 8 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
