#!/bin/bash
# round 2, call 3f: background refinement started lazily by the render (and cancellable): trace / tree suites, a short bench line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_tree.py -m gpu -x -q -k "not five_million and not one_million and not full_size" ) 2>&1 | tail -3
timeout 600 python bench.py --spp 256 --steps 3 --warmup 3 --no-cpu > gpurun_out/r03f_bench_spp256.json 2> gpurun_out/r03f_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r03f_bench_spp256.json"))
print({k: round(d[k], 1) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1), "upload ms", round(d["e2e"]["scene_upload_ms"], 1), "first", d["e2e_first_frame"]["total_s"], d["e2e_first_frame"]["secondary_tree"])
PY
timeout 600 python bench.py --spp 128 --steps 3 --warmup 3 --no-cpu --no-first-frame > gpurun_out/r03f_bench_spp128.json 2>> gpurun_out/r03f_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r03f_bench_spp128.json"))
print("spp128 (a frame as short as N=8's)", {k: round(d[k], 1) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1))
PY
