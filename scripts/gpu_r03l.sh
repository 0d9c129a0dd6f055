#!/bin/bash
# round 2, call 3l: nodes uploaded straight from the caller's array, finished positions skipped by the union scan: tree suites (tile scan
# checked against cub's), build phases, short-frame bench line
mkdir -p gpurun_out
( RM_SAH_SCAN=check timeout 600 python -m pytest tests/test_gpu_trace.py tests/test_gpu_tree.py -m gpu -x -q -k "not five_million and not one_million" ) 2>&1 | tail -3
RM_TIMING=2 timeout 600 python scripts/ab_probe.py builder3 8 lazy_tree=0 2>&1 | grep -v "validate\|textures" | tail -10 > gpurun_out/r03l_sweep_sah_build_phases.log
cat gpurun_out/r03l_sweep_sah_build_phases.log
timeout 600 python bench.py --spp 128 --steps 3 --warmup 3 --no-cpu --no-first-frame > gpurun_out/r03l_bench_spp128.json 2> gpurun_out/r03l_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r03l_bench_spp128.json"))
print("spp 128", {k: round(d[k], 1) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1), "upload ms", round(d["e2e"]["scene_upload_ms"], 1))
PY
