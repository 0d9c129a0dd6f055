#!/bin/bash
# round 2, call s: vote weights / refill threshold of the 4-wide trace (run-time options, no rebuild); ncu of k_atrous after the exp2 change
mkdir -p gpurun_out
(
timeout 200 python scripts/ab_probe.py main 64
for cfg in "2 24" "2 31" "3 28" "3 31" "4 28" "1 28"; do
  set -- $cfg
  timeout 120 python scripts/ab_probe.py wl$1_rf$2 64 trace_w_leaf=$1 trace_refill=$2
done
) 2>&1 | grep -v "Light object\|BVH has\|upload" | tee gpurun_out/r02s_ab_votes.log
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_atrous' -s 12 -c 2 \
    -f -o gpurun_out/r02s_prof_atrous python tests/tools/post_probe.py 8 > gpurun_out/r02s_prof_atrous.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02s_prof_atrous.ncu-rep | cut -c 1-420 | tee gpurun_out/r02s_ncu_atrous.txt
