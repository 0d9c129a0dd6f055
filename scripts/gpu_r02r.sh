#!/bin/bash
# round 2, call r: a-trous weight through exp2 approximations + contracted sums: parity (golden planes, 1080p bench frame) and timing
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_post.py -m gpu -x -q ) > gpurun_out/r02r_pytest_post.log 2>&1
tail -8 gpurun_out/r02r_pytest_post.log
timeout 300 python tests/tools/post_probe.py 8 2>/dev/null | tee gpurun_out/r02r_post_passes_1080p.json | cut -c 1-900
