#!/bin/bash
mkdir -p gpurun_out
V=raym0nade_b200/variants
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -16 gpurun_out/pytest_gpu.log
(
timeout 300 python scripts/ab_probe.py main 32
for v in leaf2 tb64 tb256; do
RM_LIB_PATH=$V/$v.so timeout 300 python scripts/ab_probe.py $v 32
done
timeout 300 python scripts/ab_probe.py main_sponza 32 scene=sponza
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab4.log
