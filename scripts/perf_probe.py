"""Perf probe: time the stages of one render on the GPU (wall clock around stream syncs)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model

which = sys.argv[1] if len(sys.argv) > 1 else "glossy"
n_tris, w, h, spp = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
count = len(sys.argv) > 6 and sys.argv[6] == "count"
t = time.time()
if which == "glossy": scene, args = scenes.glossy_dielectric(n_tris, w, h, spp)
elif which == "sponza": scene, args = scenes.sponza_scale(n_tris, w, h, spp, tex_size=1024)
elif which == "five": scene, args = scenes.five_million(n_tris, w, h); args = args.replace(spp=spp)
elif which == "cornell": scene, args = scenes.cornell_box(w, h, spp)
elif which == "texture": scene, args = scenes.texture_heavy(n_tris, w, h, spp)
print("scene gen %.2fs faces %d" % (time.time() - t, scene.n_faces))
t = time.time(); model = Model(scene); print("host prepare %.2fs" % (time.time() - t))
ctx = Context(0)
t = time.time(); ctx.upload(model); ctx.synchronize(); print("upload %.3fs  %.1f MB" % (time.time() - t, ctx.scene_bytes() / 1e6))
if count: ctx.set_option("count_tests", 1)
def timed(name, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        ctx.stats_reset(); ctx.synchronize(); t0 = time.time(); fn(); ctx.synchronize(); best = min(best, time.time() - t0)
    s = ctx.stats()
    extra = ""
    if count and s["rays"]: extra = " B/ray %.1f T/ray %.1f bytes/ray %.0f" % (s["box"] / s["rays"], s["tri"] / s["rays"], (32 * s["box"] + 36 * s["tri"]) / s["rays"])
    print("%-16s %8.3f ms  rays %11d  %8.1f Mrays/s launches %d%s" % (name, best * 1e3, s["rays"], s["rays"] / best / 1e6, s["launches"], extra))
    return best
timed("primary", lambda: ctx.trace_primary(args, download=False))
timed("gbuffer", lambda: ctx.gbuffer(args, download=False))
if spp > 0:
    tr = timed("render_samples", lambda: ctx.render_samples(args, seed=1), reps=2)
    print("pixel-samples/s %.1f M" % (w * h * spp / tr / 1e6))
    t0 = time.time(); ctx.resolve(args); print("resolve+download %.3fs" % (time.time() - t0))
