#!/bin/bash
# round 2, call v: steady-state ncu (--set full, source pages) of the shading stages
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_decide|k_continue|k_surface|k_nee|k_regen' -s 14 -c 8 \
    -f -o /tmp/prof_r02v python scripts/ab_probe.py ncu 64 > gpurun_out/r02v_prof.log 2>&1
tail -2 gpurun_out/r02v_prof.log
python scripts/ncu_summary.py /tmp/prof_r02v.ncu-rep > gpurun_out/r02v_ncu_summary.txt 2>&1
cut -c 1-330 gpurun_out/r02v_ncu_summary.txt
for k in k_surface k_regen k_decide 'k_continue<0>' k_nee; do
  python scripts/ncu_src.py /tmp/prof_r02v.ncu-rep "$k" 0 45 samples > "gpurun_out/r02v_src_$(echo $k | tr -d '<>').txt" 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_direct_gen' -s 2 -c 1 \
    -f -o /tmp/prof_r02v_direct python scripts/ab_probe.py ncu 64 > gpurun_out/r02v_prof_direct.log 2>&1
python scripts/ncu_summary.py /tmp/prof_r02v_direct.ncu-rep | cut -c 1-330 | tee -a gpurun_out/r02v_ncu_summary.txt
python scripts/ncu_src.py /tmp/prof_r02v_direct.ncu-rep k_direct_gen 0 45 samples > gpurun_out/r02v_src_k_direct_gen.txt 2>&1
