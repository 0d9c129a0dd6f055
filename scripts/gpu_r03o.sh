#!/bin/bash
# round 2, call 3o: union scan on a (tiles, 6 rows) grid - no integer division per item - and triangle boxes as 32-byte records read by one
# 256-bit load: tile scan checked against cub's on the test scenes and the bench scene, build phases
mkdir -p gpurun_out
( RM_SAH_SCAN=check timeout 600 python -m pytest tests/test_gpu_trace.py tests/test_gpu_tree.py -m gpu -x -q -k "not five_million and not one_million" ) 2>&1 | tail -3
RM_SAH_SCAN=check timeout 600 python scripts/ab_probe.py builder3check 1 lazy_tree=0 2>&1 | tail -2
RM_TIMING=2 timeout 600 python scripts/ab_probe.py builder3 8 lazy_tree=0 2>&1 | grep -v "validate\|textures" | tail -10 > gpurun_out/r03o_sweep_sah_build_phases.log
cat gpurun_out/r03o_sweep_sah_build_phases.log
