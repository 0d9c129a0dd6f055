#!/bin/bash
# round 2, call 3i: the sweep-SAH builder as the default: whole GPU suite + smoke, bench lines at 128 / 256 spp (short frames, e2e)
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q; echo "pytest exit $?"; python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r03i_pytest_gpu_and_smoke.log 2>&1
grep -v "Light object\|BVH has" gpurun_out/r03i_pytest_gpu_and_smoke.log | tail -6
for spp in 128 256; do
timeout 600 python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu > gpurun_out/r03i_bench_spp$spp.json 2> gpurun_out/r03i_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r03i_bench_spp$spp.json"))
print("spp $spp", {k: round(d[k], 1) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1), "upload ms", round(d["e2e"]["scene_upload_ms"], 1), "first", d["e2e_first_frame"]["total_s"], d["e2e_first_frame"]["secondary_tree"])
PY
done
