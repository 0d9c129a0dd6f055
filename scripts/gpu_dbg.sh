#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_trace.py -m gpu -q ) 2>&1 | tail -15
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model
from oracle import refbind
scene, args = scenes.glossy_dielectric(60_000, 128, 72, 32)
R = refbind.RefScene(scene); ctx = Context(0).upload(Model(scene))
tot = lambda o: {k: float(o[k]["radiance"].astype(np.float64).sum()) for k in ("Dd", "Ds", "Id", "Is")}
for seed in (5, 6, 7, 8):
    print("gpu", seed, tot(ctx.render(args, seed=seed)))
for sb in (100, 200, 300):
    print("cpu", sb, tot(R.render(args, threads=8, seed_base=sb)))
PY
