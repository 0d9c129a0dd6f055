#!/bin/bash
# round 2, call u: where the first frame goes
mkdir -p gpurun_out
( timeout 200 python scripts/first_frame_probe.py device 64; timeout 200 python scripts/first_frame_probe.py host 64 ) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/r02u_first_frame.log
