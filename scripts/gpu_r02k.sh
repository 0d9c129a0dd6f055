#!/bin/bash
# round 2, call k: two-pass FXAA (parity + timing), DoF device sort parity, ncu of the image-space passes, DRAM / L2 traffic of the
# traversal kernels on the 4-wide tree (-> profiles/traffic.json) and of the 5 M-triangle primary pass
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_post.py tests/test_gpu_render.py tests/test_gpu_trace.py -m gpu -x -q -k "fxaa or post or depth or five_million or console" ) 2>&1 | tail -4
timeout 300 python scripts/post_bench.py 3840 2160 | tee gpurun_out/r02k_fxaa_4k.json
timeout 300 python scripts/post_bench.py 1920 1080 | tee -a gpurun_out/r02k_fxaa_4k.json
timeout 600 python tests/tools/post_probe.py 8 2>/dev/null | tee gpurun_out/r02k_post_passes_1080p.json
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_atrous|k_spatial_clamp|k_filter_var|k_fxaa|k_dof' -c 14 \
    -f -o /tmp/prof_r02k_post python tests/tools/post_probe.py 8 > gpurun_out/r02k_prof_post.log 2>&1
python scripts/ncu_summary.py /tmp/prof_r02k_post.ncu-rep | cut -c 1-420 | tee gpurun_out/r02k_ncu_post_summary.txt
python scripts/ncu_src.py /tmp/prof_r02k_post.ncu-rep 'k_atrous' 0 45 > gpurun_out/r02k_src_k_atrous.txt 2>&1
bash scripts/gpu_traffic.sh r02k 2>&1 | cut -c 1-420 | tail -12
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:k_trace<rm::PrimaryJob' -s 3 -c 2 \
    -f -o /tmp/prof_r02k_c5 python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02k_prof_c5.log 2>&1
python scripts/ncu_summary.py /tmp/prof_r02k_c5.ncu-rep | cut -c 1-420 | tee gpurun_out/r02k_ncu_config5_primary.txt
