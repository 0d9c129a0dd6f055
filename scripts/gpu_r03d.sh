#!/bin/bash
# round 2, call 3d: the end-to-end loop with the prepared scene page-locked once (rm_prepared_pin) against pageable uploads
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -x -q -k "page_locked" ) 2>&1 | tail -2
for mode in "" "--pageable-scene"; do
  timeout 600 python bench.py --spp 256 --steps 3 --warmup 3 --no-cpu --no-first-frame $mode > gpurun_out/r03d_bench_spp256${mode}.json 2> gpurun_out/r03d_bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r03d_bench_spp256${mode}.json"))
print("${mode}" or "pinned", {k: round(d[k], 1) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1), "upload ms", round(d["e2e"]["scene_upload_ms"], 1))
PY
done
