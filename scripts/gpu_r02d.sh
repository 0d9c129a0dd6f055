#!/bin/bash
# round 2, call d: batched ray reporting A/B, k_fxaa alone at 4K, steady-state ncu of the round's kernels with source pages
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py -m gpu -x -q -k "fxaa or closest or secondary or primary or replay_indirect" ) 2>&1 | tail -3
(
timeout 300 python scripts/ab_probe.py batched_report 128
timeout 300 python scripts/ab_probe.py batched_rf24 128 trace_refill=24
timeout 300 python scripts/ab_probe.py batched_rf30 128 trace_refill=30
timeout 300 python scripts/ab_probe.py batched_rf20 128 trace_refill=20
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/r02d_ab.log
timeout 300 python scripts/post_bench.py 3840 2160 | tee gpurun_out/r02d_fxaa_4k.json
timeout 300 python scripts/post_bench.py 1920 1080 | tee -a gpurun_out/r02d_fxaa_4k.json
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_trace<rm::PathJob|k_trace<rm::ShadowJob|k_decide|k_continue|k_surface|k_nee|k_regen' -s 24 -c 16 \
    -f -o /tmp/prof_r02d python scripts/ab_probe.py ncu 64 > gpurun_out/r02d_prof.log 2>&1
tail -2 gpurun_out/r02d_prof.log
python scripts/ncu_summary.py /tmp/prof_r02d.ncu-rep > gpurun_out/r02d_ncu_summary.txt 2>&1
cut -c 1-330 gpurun_out/r02d_ncu_summary.txt
python scripts/ncu_src.py /tmp/prof_r02d.ncu-rep 'k_trace<rm::PathJob' 0 70 > gpurun_out/r02d_src_k_trace_paths.txt 2>&1
python scripts/ncu_src.py /tmp/prof_r02d.ncu-rep 'k_nee' 0 70 > gpurun_out/r02d_src_k_nee.txt 2>&1
python scripts/ncu_src.py /tmp/prof_r02d.ncu-rep 'k_continue<0>' 0 50 > gpurun_out/r02d_src_k_continue.txt 2>&1
python scripts/ncu_src.py /tmp/prof_r02d.ncu-rep 'k_surface' 0 50 > gpurun_out/r02d_src_k_surface.txt 2>&1
