#!/bin/bash
# parity suite + the BASELINE configs at full size (stage timings)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
bash scripts/gpu_configs.sh
