#!/bin/bash
# quick GPU iteration: parity suite, then stage timings on the bench workload
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python scripts/perf_probe.py glossy 1000000 1920 1080 ${1:-32} > gpurun_out/probe_glossy.log 2>&1
cat gpurun_out/probe_glossy.log
timeout 600 python bench.py --spp ${1:-32} --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_small.json"))
    print("value %.1f Mrays/s  ms/step %.1f  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    for k, v in d["kernels"].items(): print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()})
    print(d["roofline"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_small.err").read()[-2000:])
PY
