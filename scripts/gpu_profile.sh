#!/bin/bash
# one GPU call: parity suite, default bench line, ncu launch list, ncu --set full of the hot kernels
mkdir -p gpurun_out
TAG=${1:-r01b}
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_${TAG}.csv &
SMI=$!
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
kill $SMI
tail -c 3000 gpurun_out/bench_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/launches_${TAG}.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_${TAG}.csv | tee gpurun_out/launches_${TAG}_summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_trace<rm::PathJob|k_trace<rm::ShadowJob|k_bounce|k_surface|k_nee|k_regen|k_accum_shadow' -s 28 -c 14 \
    -f -o gpurun_out/prof_${TAG} python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/prof_${TAG}.log 2>&1
tail -3 gpurun_out/prof_${TAG}.log
ls -la gpurun_out
