"""The device's sweep-SAH builder on configs[4]'s 5 M-triangle scene: build time, tree shape, and a low-resolution render through it.
    python scripts/big_tree_check.py [n_tris]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
scene, args = scenes.five_million(n, 480, 270)
model = Model(scene)
ctx = Context(0)
ctx.upload(model); ctx.synchronize()
t0 = time.perf_counter(); info = ctx.tree_info(); ctx.synchronize(); t_first = time.perf_counter() - t0      # the deferred build runs here
ctx.set_option("lazy_tree", 0)
ts = []
for _ in range(2):
    t0 = time.perf_counter(); ctx.upload(model); ctx.synchronize(); ts.append(time.perf_counter() - t0)
ctx.set_option("tree_builder", 1)
t0 = time.perf_counter(); ctx.upload(model); ctx.synchronize(); t_ploc = time.perf_counter() - t0
ploc = ctx.tree_info()
ctx.set_option("tree_builder", 3)
ctx.upload(model)
out = ctx.render(args.replace(spp=8, P_Direct=0.5), seed=1)
ok = all(bool(np.isfinite(out[k]["radiance"]).all()) for k in ("Dd", "Ds", "Id", "Is"))
print(json.dumps({"faces": model.n_faces, "s_first_build_incl_allocations": t_first, "s_upload_incl_sweep_sah_build_pageable": min(ts), "s_upload_incl_ploc_build_pageable": t_ploc,
                  "sweep_sah_tree": info, "ploc_tree": ploc, "render_8spp_finite": ok, "energy": float(out["Id"]["radiance"].astype(np.float64).sum())}))
