#!/bin/bash
# round 2, call w: k_direct_gen mapping A/B on the light-object scene
mkdir -p gpurun_out
(
timeout 200 python scripts/ab_probe.py thread_per_pixel 128
timeout 200 python scripts/ab_probe.py warp_per_pixel 128 direct_warp=1
) 2>&1 | grep -v "Light object\|BVH has\|upload" | tee gpurun_out/r02w_ab_direct.log
