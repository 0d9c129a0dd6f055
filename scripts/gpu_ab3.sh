#!/bin/bash
mkdir -p gpurun_out
V=raym0nade_b200/variants
(
timeout 300 python scripts/ab_probe.py main_256x2 32 trace_refill=28
for v in s256x3 s128x4 s128x6 s64x8; do
RM_LIB_PATH=$V/$v.so timeout 300 python scripts/ab_probe.py $v 32 trace_refill=28
done
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_ab3.csv \
    python scripts/ab_probe.py ncu 32 > gpurun_out/launches_ab3.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_ab3.csv | tee gpurun_out/launches_ab3_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_bounce|k_surface' -s 4 -c 4 \
    -f -o /tmp/prof_shade2 python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/prof_shade2.log 2>&1
python scripts/ncu_summary.py /tmp/prof_shade2.ncu-rep > gpurun_out/prof_shade2_summary.txt 2>&1
for k in k_bounce k_surface; do
  python scripts/ncu_src.py /tmp/prof_shade2.ncu-rep $k 0 40 > gpurun_out/prof_shade2_src_$k.txt 2>&1
done
