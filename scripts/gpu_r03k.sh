#!/bin/bash
# round 2, call 3k: the deferred build of the secondary-ray tree: tree / trace suites, the config-5 line (primary rays only: no build)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_trace.py -m gpu -x -q -k "not five_million and not one_million" ) 2>&1 | tail -4
timeout 300 python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r03k_bench_config5.json 2> gpurun_out/r03k_bench_config5.err; cut -c 1-700 gpurun_out/r03k_bench_config5.json
timeout 600 python bench.py --spp 128 --steps 3 --warmup 3 --no-cpu --no-first-frame > gpurun_out/r03k_bench_spp128.json 2> gpurun_out/r03k_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r03k_bench_spp128.json"))
print("spp 128", {k: round(d[k], 1) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1), "upload ms", round(d["e2e"]["scene_upload_ms"], 1))
PY
