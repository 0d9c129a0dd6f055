#!/bin/bash
# round 2, call g: the new bench line (N=1), the reference arm, the config5 line
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "bench rc $?"; tail -c 4500 gpurun_out/r02g_bench.json; tail -5 gpurun_out/r02g_bench.err
timeout 900 python bench.py --workload config5 --steps 5 --warmup 3 > gpurun_out/r02g_bench_config5.json 2> gpurun_out/r02g_bench_config5.err
echo "config5 rc $?"; tail -c 2500 gpurun_out/r02g_bench_config5.json; tail -5 gpurun_out/r02g_bench_config5.err
