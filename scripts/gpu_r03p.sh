#!/bin/bash
# round 2, call 3p: the byte decode of the 4-wide step with immediate PRMT selectors (2^23 from the kernel parameters): seam tests, A/B at 64 spp
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_trace.py -m gpu -x -q -k "secondary_ray_tree" ) 2>&1 | tail -2
timeout 300 python scripts/ab_probe.py prmt_imm_selectors 64 2>&1 | grep "wall" | tee gpurun_out/r03p_ab_prmt_immediate_selectors.log
