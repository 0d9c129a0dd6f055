#!/bin/bash
# round 2, call 3b: compute-sanitizer (memcheck) over the kernels added in the second half of the round: FXAA strips (bulk copies),
# the device-built reference tree, the refit, the sampled-pixel list and the dense vertex records (a small render)
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q -m gpu \
    "tests/test_gpu_render.py::test_fxaa_bit_exact_random_and_edges" \
    "tests/test_gpu_tree.py::test_device_built_reference_tree_follows_the_rule[cornell]" \
    "tests/test_gpu_tree.py::test_device_built_reference_tree_follows_the_rule[duplicates]" \
    "tests/test_gpu_tree.py::test_refit_for_moved_vertices[heightfield-1]" \
    "tests/test_gpu_tree.py::test_refit_for_moved_vertices[heightfield-0]" ) > gpurun_out/r03b_memcheck.log 2>&1
echo "memcheck rc $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r03b_memcheck.log | head -20
