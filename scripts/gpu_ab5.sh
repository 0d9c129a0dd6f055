#!/bin/bash
mkdir -p gpurun_out
V=raym0nade_b200/variants
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
(
timeout 300 python scripts/ab_probe.py main 32
RM_LIB_PATH=$V/tri4.so timeout 300 python scripts/ab_probe.py tri4 32
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab5.log
