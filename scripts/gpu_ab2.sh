#!/bin/bash
mkdir -p gpurun_out
V=raym0nade_b200/variants
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
(
RM_LIB_PATH=$V/bf.so   timeout 300 python scripts/ab_probe.py bf_stack18 32
timeout 300 python scripts/ab_probe.py new 32
timeout 300 python scripts/ab_probe.py new_refill28 32 trace_refill=28
timeout 300 python scripts/ab_probe.py new_refill30 32 trace_refill=30
timeout 300 python scripts/ab_probe.py new_refill32 32 trace_refill=32
timeout 300 python scripts/ab_probe.py new_wl2 32 trace_refill=28 trace_w_leaf=2 trace_w_inner=3
timeout 300 python scripts/ab_probe.py new_wi2 32 trace_refill=28 trace_w_leaf=3 trace_w_inner=2
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab2.log
