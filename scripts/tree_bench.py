"""BVH::build on the host (rm_prepare_scene) against the device (rm_prepare_scene_device), and rm_scene_refit, on configs[2]'s
scene: wall-clock seconds per call, one JSON line.   python scripts/tree_bench.py [n_tris]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
scene, args = scenes.glossy_dielectric(n, 1920, 1080, 0)
ctx = Context(0)
raw = np.ascontiguousarray(scene.positions, np.float32).reshape(-1, 3, 3)


def best(f, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts), r


ctx.tree_build(raw)                                   # warm-up: allocations, cub temp
t_tree_dev, _ = best(lambda: ctx.tree_build(raw))
t_prep_host, host = best(lambda: Model(scene), 2)
t_prep_dev, dev = best(lambda: Model(scene, ctx), 2)
ctx.set_option("lazy_tree", 0)                       # the secondary-ray tree is built inside the upload that is timed below
ctx.set_option("tree_builder", 1)
ctx.upload(dev)
ctx.synchronize()
t_upload_ploc, _ = best(lambda: (ctx.upload(dev), ctx.synchronize()))
ctx.set_option("tree_builder", 3)
ctx.upload(dev)
ctx.synchronize()
pos = np.frombuffer((__import__("ctypes").c_char * (36 * dev.n_faces)).from_address(dev.desc.positions), np.float32).copy()
t_refit, _ = best(lambda: (ctx.refit(pos), ctx.synchronize()))
t_upload, _ = best(lambda: (ctx.upload(dev), ctx.synchronize()))
print(json.dumps({"faces": dev.n_faces, "nodes": int(dev.desc.n_nodes), "s_tree_build_device_incl_h2d_d2h": t_tree_dev, "s_prepare_scene_host": t_prep_host,
                  "s_prepare_scene_device_tree": t_prep_dev, "s_scene_refit_incl_h2d": t_refit, "s_scene_upload_incl_device_secondary_tree": t_upload,
                  "s_scene_upload_incl_ploc_tree": t_upload_ploc, "secondary_tree": ctx.tree_info()}))
