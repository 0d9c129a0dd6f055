#!/bin/bash
# round 2, call 3n: ncu --set full of the sweep-SAH builder's kernels (levels 8-9 of the bench scene's build: thousands of nodes per level)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_union_fold|k_union_carries|k_union_scan|k_candidates|k_decide|k_mark|k_scatter|DeviceScanKernel' -s 72 -c 18 \
    -f -o /tmp/builder_r03n python scripts/ab_probe.py ncu 1 lazy_tree=0 > gpurun_out/r03n_ncu_builder.log 2>&1
tail -2 gpurun_out/r03n_ncu_builder.log
python scripts/ncu_summary.py /tmp/builder_r03n.ncu-rep | tee gpurun_out/r03n_ncu_sweep_sah_builder.txt
python scripts/ncu_src.py /tmp/builder_r03n.ncu-rep 'k_union_scan' 0 25 > gpurun_out/r03n_ncu_source_k_union_scan.txt 2>&1
