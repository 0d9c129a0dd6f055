#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_shade|k_trace_paths|k_trace_shadow' -c 9 \
    -o gpurun_out/prof_r01_first -f python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
