#!/bin/bash
# One gpurun call: GPU parity tests, stage timings, a short bench line, the ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/perf_probe.py glossy 1000000 1920 1080 32 count > gpurun_out/probe_glossy_count.log 2>&1
timeout 300 python scripts/perf_probe.py glossy 1000000 1920 1080 64 > gpurun_out/probe_glossy.log 2>&1
timeout 600 python bench.py --spp 64 --steps 3 --warmup 3 > gpurun_out/bench_spp64.json 2> gpurun_out/bench_spp64.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches_r01.csv python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/ncu_launch.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/probe_glossy_count.log gpurun_out/probe_glossy.log
cat gpurun_out/bench_spp64.json
tail -3 gpurun_out/bench_spp64.err
