#!/bin/bash
# round 2, call t: per-pixel stages over the list of sampled pixels (k_regen / k_direct_gen / k_accum_direct), a-trous with row prefetch,
# first frame with the device-built reference tree: parity (render + post suites), A/B, timings
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_render.py tests/test_gpu_post.py -m gpu -x -q ) > gpurun_out/r02t_pytest.log 2>&1
tail -8 gpurun_out/r02t_pytest.log
timeout 300 python tests/tools/post_probe.py 8 2>/dev/null | tee gpurun_out/r02t_post_passes_1080p.json | cut -c 1-700
(
timeout 200 python scripts/ab_probe.py main 64
timeout 200 python scripts/ab_probe.py all_pixels 64 compact_pixels=0
) 2>&1 | grep -v "Light object\|BVH has\|upload" | tee gpurun_out/r02t_ab_compact.log
timeout 600 python bench.py --spp 256 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02t_bench_spp256.json 2> gpurun_out/r02t_bench_spp256.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02t_bench_spp256.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "first", d["e2e_first_frame"])
print({k: (round(v.get("ms_per_step", 0), 1), round(v.get("mrays_per_s_kernel_only", 0))) for k, v in d["kernels"].items()})
PY
