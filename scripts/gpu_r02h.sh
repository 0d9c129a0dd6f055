#!/bin/bash
# round 2, call h: background tree refinement (tree_builder 2): parity suite, A/B, bench line
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
(
timeout 300 python scripts/ab_probe.py refine_default 128
timeout 300 python scripts/ab_probe.py device_only 128 tree_builder=1
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/r02h_ab.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
echo "bench rc $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02h_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_first_frame"], d["roofline"]["frac"], d["config"]["secondary_rays"])
print({k: (v.get("ms_per_step"), v.get("mrays_per_s_kernel_only")) for k, v in d["kernels"].items()})
PY
tail -3 gpurun_out/r02h_bench.err
