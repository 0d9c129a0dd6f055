#!/bin/bash
# round 2, call 3g: the sweep-SAH device builder ("tree_builder" 3): memcheck on small scenes, the seam and refit tests on its
# tree, then A/B of the three device builders on the bench scene (upload time, kernel times)
mkdir -p gpurun_out
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q -m gpu \
    "tests/test_gpu_trace.py::test_secondary_ray_tree_finds_the_reference_hits[cornell-2-3]" \
    "tests/test_gpu_trace.py::test_secondary_ray_tree_finds_the_reference_hits[heightfield-2-3]" ) > gpurun_out/r03g_memcheck.log 2>&1
echo "memcheck rc $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned|Error" gpurun_out/r03g_memcheck.log | head -12
( timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_tree.py -m gpu -x -q -k "secondary_ray_tree or refit" ) 2>&1 | tail -4
for b in 1 3 2; do
  RM_TIMING=1 timeout 600 python scripts/ab_probe.py builder$b 64 tree_builder=$b 2>&1 | grep -v "^rm_scene_upload: \(validate\|textures\)" | tail -12
done > gpurun_out/r03g_ab_builders.log 2>&1
cat gpurun_out/r03g_ab_builders.log
