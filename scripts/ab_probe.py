"""A/B perf probe: per-kernel-kind device times of one render (CUDA events inside the library).
usage: [RM_LIB_PATH=variant.so] python scripts/ab_probe.py label spp [opt=value ...]
Scene: the bench workload (glossy/dielectric 1M tris, 1080p)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model

label, spp = sys.argv[1], int(sys.argv[2])
opts = dict(a.split("=") for a in sys.argv[3:])
which = opts.pop("scene", "glossy")
import pickle
cache = "/tmp/ab_scene_%s_%d.pkl" % (which, spp)
if os.path.exists(cache):
    scene, args = pickle.load(open(cache, "rb"))
else:
    if which == "glossy": scene, args = scenes.glossy_dielectric(1000000, 1920, 1080, spp)
    elif which == "sponza": scene, args = scenes.sponza_scale(260000, 1920, 1080, spp, tex_size=1024)
    try: pickle.dump((scene, args), open(cache, "wb"), protocol=4)
    except Exception as e: print("scene cache not written:", e)
model = Model(scene)
ctx = Context(0)
for k, v in opts.items(): ctx.set_option(k, int(v))         # before the upload: some options shape what is staged
t0 = time.time(); ctx.upload(model); ctx.synchronize(); up1 = time.time() - t0
t0 = time.time(); ctx.upload(model); ctx.synchronize(); up2 = time.time() - t0
print("%-22s upload %.1f ms, again %.1f ms (host tree cache cleared: " % (label, up1 * 1e3, up2 * 1e3), end="")
ctx.set_option("tree_builder", int(opts.get("tree_builder", 3)))      # clears the host-tree cache (3 = the library default)
t0 = time.time(); ctx.upload(model); ctx.synchronize(); print("%.1f ms)  tree %s" % ((time.time() - t0) * 1e3, ctx.tree_info()), flush=True)
ctx.trace_primary(args, download=False); ctx.gbuffer(args, download=False)
ctx.render_samples(args, seed=1); ctx.synchronize()          # warm-up
ctx.set_option("tree_wait", 1)                               # the background-refined tree (if any) is in place before the clock starts
best = None
for rep in range(2):
    ctx.set_option("time_kernels", 1); ctx.stats_reset(); ctx.synchronize()
    t0 = time.time(); ctx.render_samples(args, seed=2 + rep); ctx.synchronize(); wall = (time.time() - t0) * 1e3
    k = ctx.stats_kernels(); s = ctx.stats(); ctx.set_option("time_kernels", 0)
    if best is None or wall < best[0]: best = (wall, k, s)
wall, k, s = best
print("%-22s wall %8.1f ms  %7.1f Mrays/s | paths %7.1f ms (%6.0f Mr/s)  shadow %7.1f ms (%6.0f Mr/s)  shade %7.1f ms | other %6.1f ms  launches %d" % (
    label, wall, s["rays"] / wall / 1e3, k["paths"]["ms"], k["paths"]["rays"] / max(k["paths"]["ms"], 1e-9) / 1e3,
    k["shadow"]["ms"], k["shadow"]["rays"] / max(k["shadow"]["ms"], 1e-9) / 1e3, k["shade"]["ms"],
    wall - k["paths"]["ms"] - k["shadow"]["ms"] - k["shade"]["ms"], s["launches"]), flush=True)
