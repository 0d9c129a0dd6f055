#!/bin/bash
# round 2, call n: FXAA as one launch (k_fxaa_strip): parity, timing of the three forms, ncu of the strip kernel
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_trace.py -m gpu -x -q -k "fxaa or five_million" ) 2>&1 | tail -15
timeout 300 python scripts/post_bench.py 3840 2160 | tee gpurun_out/r02n_fxaa_4k.json
timeout 300 python scripts/post_bench.py 1920 1080 | tee gpurun_out/r02n_fxaa_1080p.json
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_fxaa_strip' -s 140 -c 2 \
    -f -o gpurun_out/r02n_prof_fxaa python scripts/post_bench.py 3840 2160 > gpurun_out/r02n_prof_fxaa.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02n_prof_fxaa.ncu-rep | cut -c 1-420 | tee gpurun_out/r02n_ncu_fxaa.txt
