#!/bin/bash
# round 2, call 3h: the sweep-SAH builder with its own tile scan: checked bit for bit against cub's scan on the test scenes
# (RM_SAH_SCAN=check), memcheck, then build phases and kernel times on the bench scene
mkdir -p gpurun_out
( RM_SAH_SCAN=check timeout 600 python -m pytest tests/test_gpu_trace.py tests/test_gpu_tree.py -m gpu -x -q -k "(secondary_ray_tree and 2-3) or (refit and 3)" ) 2>&1 | tail -3
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q -m gpu \
    "tests/test_gpu_trace.py::test_secondary_ray_tree_finds_the_reference_hits[cornell-2-3]" \
    "tests/test_gpu_trace.py::test_secondary_ray_tree_finds_the_reference_hits[heightfield-2-3]" ) > gpurun_out/r03h_memcheck.log 2>&1
echo "memcheck rc $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned|Error" gpurun_out/r03h_memcheck.log | head -12
RM_SAH_SCAN=check RM_TIMING=2 timeout 600 python scripts/ab_probe.py builder3check 8 tree_builder=3 2>&1 | grep -v "validate\|textures" | tail -4
RM_TIMING=2 timeout 600 python scripts/ab_probe.py builder3 64 tree_builder=3 2>&1 | grep -v "validate\|textures" | tail -12 > gpurun_out/r03h_ab_builder3.log
cat gpurun_out/r03h_ab_builder3.log
