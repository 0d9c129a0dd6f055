#!/bin/bash
# round 2, call a: statistical-parity probes at SURVEY 8(d)'s size + the parity suite
mkdir -p gpurun_out
(
for spec in "config3 480 270 64 0" "config3 480 270 64 1" "config2 480 270 64 0" "config4 320 180 64 0" "nested 96 96 64 0" "cornell 128 128 128 0"; do
  timeout 600 python scripts/stat_probe.py $spec 2>/dev/null
done
) > gpurun_out/r02a_stat_probe.jsonl
cat gpurun_out/r02a_stat_probe.jsonl | cut -c 1-900
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02a_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02a_pytest_gpu.log
