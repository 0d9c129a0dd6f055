#!/bin/bash
# A/B of trace-kernel variants on the bench workload + source-level ncu of the shading stages (summarised on the box)
mkdir -p gpurun_out
V=raym0nade_b200/variants
(
RM_LIB_PATH=$V/base.so timeout 300 python scripts/ab_probe.py base 32
RM_LIB_PATH=$V/bf.so   timeout 300 python scripts/ab_probe.py bf_stack24 32 stack_levels=24
RM_LIB_PATH=$V/bf.so   timeout 300 python scripts/ab_probe.py bf_stack18 32
RM_LIB_PATH=$V/pf.so   timeout 300 python scripts/ab_probe.py pf_stack18 32
RM_LIB_PATH=$V/bf.so   timeout 300 python scripts/ab_probe.py bf_refill16 32 trace_refill=16
RM_LIB_PATH=$V/bf.so   timeout 300 python scripts/ab_probe.py bf_refill28 32 trace_refill=28
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab1.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_bounce|k_surface|k_nee|k_regen|k_accum_shadow|k_direct_gen' -s 0 -c 13 \
    -f -o /tmp/prof_shade python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/prof_shade.log 2>&1
tail -2 gpurun_out/prof_shade.log
python scripts/ncu_summary.py /tmp/prof_shade.ncu-rep > gpurun_out/prof_shade_summary.txt 2>&1
for k in k_bounce k_surface k_nee k_direct_gen k_regen k_accum_shadow; do
  python scripts/ncu_src.py /tmp/prof_shade.ncu-rep $k 0 60 > gpurun_out/prof_shade_src_$k.txt 2>&1
done
SZ=$(stat -c %s /tmp/prof_shade.ncu-rep); echo "rep size $SZ"
if [ "$SZ" -lt 45000000 ]; then cp /tmp/prof_shade.ncu-rep gpurun_out/; fi
ls -la gpurun_out
