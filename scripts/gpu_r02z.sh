#!/bin/bash
# round 2, call z: per-vertex records by dense vertex index: replay / statistical parity of the estimator and the A/B probe
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -x -q ) > gpurun_out/r02z_pytest_render.log 2>&1
tail -3 gpurun_out/r02z_pytest_render.log
( timeout 200 python scripts/ab_probe.py dense_vertex_records 128 ) 2>&1 | grep -v "Light object\|BVH has\|upload" | tee gpurun_out/r02z_ab_dense.log
