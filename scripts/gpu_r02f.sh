#!/bin/bash
# round 2, call f: device tree builder (PLOC + collapse) - parity, build time, A/B against the host SAH tree
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_trace.py -m gpu -x -q -k "secondary_ray_tree" ) 2>&1 | tail -15
(
timeout 300 python scripts/ab_probe.py device_tree 128
timeout 300 python scripts/ab_probe.py host_tree 128 tree_builder=0
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/r02f_ab.log
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
