#!/bin/bash
# round 2, call p: FXAA strip kernel with the prologue fixed; the device-built reference tree and the refit: parity + timing
mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "fxaa" ) > gpurun_out/r02p_pytest_fxaa.log 2>&1
FX=$?
tail -15 gpurun_out/r02p_pytest_fxaa.log
if [ $FX -eq 0 ]; then
  timeout 100 python scripts/post_bench.py 3840 2160 | tee gpurun_out/r02p_fxaa_4k.json
  timeout 100 python scripts/post_bench.py 1920 1080 | tee gpurun_out/r02p_fxaa_1080p.json
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_fxaa_strip' -s 90 -c 2 \
      -f -o gpurun_out/r02p_prof_fxaa python scripts/post_bench.py 3840 2160 > gpurun_out/r02p_prof_fxaa.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r02p_prof_fxaa.ncu-rep | cut -c 1-420 | tee gpurun_out/r02p_ncu_fxaa.txt
fi
( timeout 600 python -m pytest tests/test_gpu_tree.py -m gpu -x -q ) > gpurun_out/r02p_pytest_tree.log 2>&1
tail -25 gpurun_out/r02p_pytest_tree.log
timeout 200 python scripts/tree_bench.py | tee gpurun_out/r02p_tree_bench.json
