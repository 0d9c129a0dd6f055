#!/bin/bash
# one GPU call: parity suite, default bench line, ncu launch list, ncu --set full of the hot kernels (summarised on the box)
mkdir -p gpurun_out
TAG=${1:-r01g}
( timeout 1500 python -m pytest tests -m gpu -x -q; python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -5 gpurun_out/pytest_gpu_${TAG}.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_${TAG}.csv &
SMI=$!
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
kill $SMI
tail -c 3000 gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
cat gpurun_out/bench_ref_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/launches_${TAG}.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_${TAG}.csv | tee gpurun_out/launches_${TAG}_summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_trace<rm::PathJob|k_trace<rm::ShadowJob|k_bounce|k_surface|k_nee|k_regen|k_accum_shadow' -s 21 -c 14 \
    -f -o /tmp/prof_${TAG} python scripts/perf_probe.py glossy 1000000 1920 1080 4 > gpurun_out/prof_${TAG}.log 2>&1
tail -3 gpurun_out/prof_${TAG}.log
python scripts/ncu_summary.py /tmp/prof_${TAG}.ncu-rep > gpurun_out/prof_${TAG}_summary.txt 2>&1
python scripts/ncu_src.py /tmp/prof_${TAG}.ncu-rep 'k_trace<rm::PathJob' 0 60 > gpurun_out/prof_${TAG}_src_k_trace_paths.txt 2>&1
python scripts/ncu_src.py /tmp/prof_${TAG}.ncu-rep 'k_trace<rm::ShadowJob' 0 40 > gpurun_out/prof_${TAG}_src_k_trace_shadow.txt 2>&1
SZ=$(stat -c %s /tmp/prof_${TAG}.ncu-rep); echo "rep size $SZ"
if [ "$SZ" -lt 40000000 ]; then cp /tmp/prof_${TAG}.ncu-rep gpurun_out/; fi
ls -la gpurun_out
