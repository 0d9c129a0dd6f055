#!/bin/bash
# round 2, call j (8 GPUs): time to 4096 spp on configs[2] at N = 8, 4, 2 and configs[3] (texture-heavy, 4K, 4096 spp) at N = 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # tag nproc args...
  tag=$1; np=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + np)) bench.py --gpus $np "$@" \
      > gpurun_out/r02j_${tag}.json 2> gpurun_out/r02j_${tag}.err
  echo "$tag rc $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02j_${tag}.json"))
    print("${tag}", {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "time_to_spp", d["time_to_spp_s"], "e2e", d["e2e"]["value"], d["e2e"]["time_to_spp_s"], d["clocks"])
except Exception as e:
    print("${tag} unreadable", e)
PY
  tail -2 gpurun_out/r02j_${tag}.err
}
run config3_spp4096_n8 8 --spp 4096 --steps 2 --warmup 3
run config4_spp4096_n8 8 --workload config4 --steps 2 --warmup 3
run config3_spp4096_n4 4 --spp 4096 --steps 2 --warmup 3
run config3_spp4096_n2 2 --spp 4096 --steps 2 --warmup 3
