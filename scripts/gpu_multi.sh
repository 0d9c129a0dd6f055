#!/bin/bash
# N-GPU: the in-library reduce checked against the single-GPU frame, then the bench line exactly as the driver launches it
mkdir -p gpurun_out
N=${1:-2}
SPP=${2:-1024}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/reduce_check.py 2>&1 | grep -v "Light object\|BVH has\|OMP_NUM\|\*\*\*\*" | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --spp $SPP > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "exit $?"
grep '^{' gpurun_out/bench_${N}gpu.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3), d['config']['parallelism'])
"
tail -3 gpurun_out/bench_${N}gpu.err
