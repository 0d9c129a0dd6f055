#!/bin/bash
# source-level ncu capture of the shading stages (full-width launches a few rounds into the wavefront loop)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_bounce|k_surface|k_nee|k_regen|k_accum_shadow|k_direct_gen' -s 0 -c 26 \
    -f -o gpurun_out/prof_shade python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/prof_shade.log 2>&1
tail -3 gpurun_out/prof_shade.log
ls -la gpurun_out
