#!/bin/bash
# A/B of library builds under raym0nade_b200/variants/*.so against the in-tree build (scripts/ab_probe.py); SPP=${1:-128}
mkdir -p gpurun_out
SPP=${1:-128}
(
timeout 300 python scripts/ab_probe.py main $SPP
for v in raym0nade_b200/variants/*.so; do
  [ -e "$v" ] && RM_LIB_PATH=$v timeout 300 python scripts/ab_probe.py $(basename $v .so) $SPP
done
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab.log
