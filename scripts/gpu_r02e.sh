#!/bin/bash
# round 2, call e: sort by mode in window-contiguous blocks (A/B against call d), parity suite, reduce-scatter on 1 GPU is N/A
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
(
timeout 300 python scripts/ab_probe.py mode_sort 128
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/r02e_ab.log
timeout 300 python scripts/post_bench.py 3840 2160 | tee gpurun_out/r02e_fxaa_4k.json
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_decide|k_continue|k_nee|k_sort|k_fxaa' -s 12 -c 8 \
    -f -o /tmp/prof_r02e python scripts/ab_probe.py ncu 64 > gpurun_out/r02e_prof.log 2>&1
python scripts/ncu_summary.py /tmp/prof_r02e.ncu-rep > gpurun_out/r02e_ncu_summary.txt 2>&1
cut -c 1-330 gpurun_out/r02e_ncu_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_fxaa' -s 3 -c 1 \
    -f -o /tmp/prof_r02e_fxaa python scripts/post_bench.py 3840 2160 > gpurun_out/r02e_prof_fxaa.log 2>&1
python scripts/ncu_summary.py /tmp/prof_r02e_fxaa.ncu-rep | cut -c 1-400 | tee gpurun_out/r02e_ncu_fxaa.txt
python scripts/ncu_src.py /tmp/prof_r02e_fxaa.ncu-rep 'k_fxaa' 0 40 > gpurun_out/r02e_src_k_fxaa.txt 2>&1
