#!/bin/bash
# round 2, call 3c: compute-sanitizer (memcheck) over the wavefront estimator as it stands (sampled-pixel list, dense vertex records): the
# per-sample replay tests, the degenerate sample splits and smoke()
mkdir -p gpurun_out
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q -m gpu tests/test_gpu_render.py \
    -k "replay_indirect or replay_direct or degenerate or nine_nested" ) > gpurun_out/r03c_memcheck_render.log 2>&1
echo "memcheck rc $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r03c_memcheck_render.log | head -20
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | grep -E "ERROR SUMMARY|smoke|Invalid" | tee gpurun_out/r03c_memcheck_smoke.log
