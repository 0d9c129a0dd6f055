#!/bin/bash
# DRAM traffic of one steady-state (full-queue, 33.55 M rays) launch of the two traversal kernels: ncu --set full on the
# bench scene at 64 spp, third full round onwards; summarised on the box (the .ncu-rep stays there unless it is small)
mkdir -p gpurun_out
TAG=${1:-r01d}
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_trace<rm::PathJob|k_trace<rm::ShadowJob' -s 2 -c 8 \
    -f -o /tmp/traffic_${TAG} python scripts/ab_probe.py ncu 64 > gpurun_out/traffic_${TAG}.log 2>&1
tail -2 gpurun_out/traffic_${TAG}.log
python scripts/ncu_summary.py /tmp/traffic_${TAG}.ncu-rep | tee gpurun_out/traffic_${TAG}_summary.txt
python scripts/ncu_src.py /tmp/traffic_${TAG}.ncu-rep 'k_trace<rm::PathJob' 0 50 > gpurun_out/traffic_${TAG}_src_paths.txt 2>&1
SZ=$(stat -c %s /tmp/traffic_${TAG}.ncu-rep); echo "rep size $SZ"
if [ "$SZ" -lt 30000000 ]; then cp /tmp/traffic_${TAG}.ncu-rep gpurun_out/; fi
