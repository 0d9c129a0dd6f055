#!/bin/bash
# round 2, call l: FXAA / a-trous after the second pass at them, the bench line with the reordered counting, time to 4096 spp at N=1,
# the reference arm, the ncu launch list of the bench command
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_post.py tests/test_gpu_render.py -m gpu -x -q -k "fxaa or post or depth" ) 2>&1 | tail -3
timeout 300 python scripts/post_bench.py 3840 2160 | tee gpurun_out/r02l_fxaa_4k.json
timeout 600 python tests/tools/post_probe.py 8 2>/dev/null | tee gpurun_out/r02l_post_passes_1080p.json | cut -c 1-700
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_fxaa' -s 6 -c 2 \
    -f -o /tmp/prof_r02l_fxaa python scripts/post_bench.py 3840 2160 > gpurun_out/r02l_prof_fxaa.log 2>&1
python scripts/ncu_summary.py /tmp/prof_r02l_fxaa.ncu-rep | cut -c 1-420 | tee gpurun_out/r02l_ncu_fxaa.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r02l_clocks.csv &
SMI=$!
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
echo "bench rc $?"
kill $SMI
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02l_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "first", d["e2e_first_frame"]["total_s"], "frac", d["roofline"]["frac"], d["roofline"]["achieved"])
print({k: (round(v.get("ms_per_step", 0), 1), round(v.get("mrays_per_s_kernel_only", 0)), round(v.get("box_tests_per_ray", 0), 1)) for k, v in d["kernels"].items()})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err; cut -c 1-600 gpurun_out/r02l_bench_ref.json
timeout 900 python bench.py --spp 4096 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02l_bench_spp4096_n1.json 2> gpurun_out/r02l_bench_spp4096_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02l_bench_spp4096_n1.json"))
print("spp4096 n1", {k: d[k] for k in ("value", "ms_per_step")}, d["time_to_spp_s"], "e2e", d["e2e"]["value"], d["e2e"]["time_to_spp_s"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02l_launches.csv \
    python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu --no-first-frame > gpurun_out/r02l_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02l_launches.csv | tee gpurun_out/r02l_launches_summary.txt | head -30
