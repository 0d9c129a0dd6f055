#!/bin/bash
# round 2, call y: the state of the code after the FXAA strip kernel, the device-built reference tree + refit, the a-trous exp2 weights, the
# sampled-pixel list and the lock step removed: whole GPU suite + smoke, the default bench line, the reference arm, the ncu launch list,
# the image-space timings, the tree timings
TAG=${1:-r02y}
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q; echo "pytest exit $?"; python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_pytest_gpu_and_smoke.log 2>&1
grep -v "Light object\|BVH has" gpurun_out/${TAG}_pytest_gpu_and_smoke.log | tail -6
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc $?"
kill $SMI
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "first", d["e2e_first_frame"]["total_s"], "frac", d["roofline"]["frac"], d["roofline"]["achieved"])
print({k: (round(v.get("ms_per_step", 0), 1), round(v.get("mrays_per_s_kernel_only", 0)), round(v.get("box_tests_per_ray", 0), 1)) for k, v in d["kernels"].items()})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c 1-400 gpurun_out/${TAG}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu --no-first-frame > gpurun_out/${TAG}_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt; head -24 gpurun_out/${TAG}_launches_summary.txt
timeout 200 python scripts/post_bench.py 3840 2160 | tee gpurun_out/${TAG}_fxaa_4k.json
timeout 300 python tests/tools/post_probe.py 8 2>/dev/null | tee gpurun_out/${TAG}_post_passes_1080p.json | cut -c 1-600
timeout 200 python scripts/tree_bench.py 2>/dev/null | tee gpurun_out/${TAG}_tree_bench.json
timeout 300 python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_config5.json 2> gpurun_out/${TAG}_bench_config5.err; cut -c 1-600 gpurun_out/${TAG}_bench_config5.json
