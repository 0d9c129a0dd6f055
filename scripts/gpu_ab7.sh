#!/bin/bash
mkdir -p gpurun_out
(
timeout 300 python scripts/ab_probe.py wave4M 64
timeout 300 python scripts/ab_probe.py wave2M 64 wave_paths=2097152
timeout 300 python scripts/ab_probe.py wave8M 64 wave_paths=8388608
timeout 300 python scripts/ab_probe.py wave16M 64 wave_paths=16777216
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab7.log
