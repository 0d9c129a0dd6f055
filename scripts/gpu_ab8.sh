#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
(
timeout 300 python scripts/ab_probe.py wave16M_256spp 256
timeout 300 python scripts/ab_probe.py wave32M_256spp 256 wave_paths=33554432
timeout 300 python scripts/ab_probe.py wave8M_256spp 256 wave_paths=8388608
) 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab8.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_w16.json 2> gpurun_out/bench_w16.err
grep '^{' gpurun_out/bench_w16.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3)); print({k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
