#!/bin/bash
# round 2, call i (2 GPUs): reduce-scatter parity against the single-GPU frame, bench line at N=2 (default spp)
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/reduce_check.py 2>&1 | grep -v "Light object\|BVH has" | tail -5
( timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "reduce_across or depth or post" ) 2>&1 | tail -3
( timeout 300 python -m pytest tests/test_gpu_post.py -m gpu -x -q ) 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02i_bench_2gpu.json 2> gpurun_out/r02i_bench_2gpu.err
echo "bench rc $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02i_bench_2gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"], d["config"]["parallelism"])
PY
tail -3 gpurun_out/r02i_bench_2gpu.err
