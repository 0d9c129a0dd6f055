#!/bin/bash
# round 2, call c: sorted-queue shading + 4-wide tree: full parity suite, A/B, ncu summary of the shading kernels
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r02c_pytest.log 2>&1
tail -12 gpurun_out/r02c_pytest.log
(
timeout 300 python scripts/ab_probe.py sorted_wide 128
timeout 300 python scripts/ab_probe.py sorted_binary 128 secondary_tree=1
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/r02c_ab.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_trace<rm::PathJob|k_trace<rm::ShadowJob|k_decide|k_continue|k_surface|k_nee|k_sort' -s 40 -c 22 \
    -f -o /tmp/prof_r02c python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/r02c_prof.log 2>&1
tail -3 gpurun_out/r02c_prof.log
python scripts/ncu_summary.py /tmp/prof_r02c.ncu-rep > gpurun_out/r02c_ncu_summary.txt 2>&1
cat gpurun_out/r02c_ncu_summary.txt | cut -c 1-330
