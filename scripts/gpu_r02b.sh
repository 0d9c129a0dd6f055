#!/bin/bash
# round 2, call b: 4-wide secondary-ray tree - parity + A/B against the binary tree; the new full-size statistical tests
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py -m gpu -q --durations=8 -k "secondary_ray_tree or full_size_scenes or texture_heavy or nested or replay or statistically" ) > gpurun_out/r02b_pytest.log 2>&1
tail -40 gpurun_out/r02b_pytest.log
(
timeout 300 python scripts/ab_probe.py binary 128 secondary_tree=1
timeout 300 python scripts/ab_probe.py wide 128 secondary_tree=2
timeout 300 python scripts/ab_probe.py wide_wl1 128 secondary_tree=2 trace_w_leaf=1
timeout 300 python scripts/ab_probe.py wide_wl3 128 secondary_tree=2 trace_w_leaf=3
timeout 300 python scripts/ab_probe.py wide_sm20 128 secondary_tree=2 smem_levels=20
timeout 300 python scripts/ab_probe.py wide_sm10 128 secondary_tree=2 smem_levels=10
timeout 300 python scripts/ab_probe.py wide_rf24 128 secondary_tree=2 trace_refill=24
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/r02b_ab_wide.log
timeout 300 python scripts/perf_probe.py glossy 1000000 1920 1080 16 count 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/r02b_counts_wide.log
