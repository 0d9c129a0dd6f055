"""Where the first frame of a fresh process goes (host clock): context creation, scene preparation (host tree | device tree),
rm_scene_upload (RM_TIMING=1 prints its phases on stderr), first render.   python scripts/first_frame_probe.py host|device [spp]"""
import os, sys, time
os.environ.setdefault("RM_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model
from raym0nade_b200.ctypes_defs import HITINFO_DTYPE, RADIANCE_DTYPE

mode = sys.argv[1] if len(sys.argv) > 1 else "device"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
scene, args = scenes.glossy_dielectric(1_000_000, 1920, 1080, spp)
npix = args.width * args.height
g, pl = np.zeros(npix, HITINFO_DTYPE), [np.zeros(npix, RADIANCE_DTYPE) for _ in range(4)]
t = [time.perf_counter()]
def lap(): t.append(time.perf_counter()); return (t[-1] - t[-2]) * 1e3
if mode == "host":
    m = Model(scene); t_prep = lap()
    ctx = Context(0); t_ctx = lap()
else:
    ctx = Context(0); t_ctx = lap()
    m = Model(scene, ctx); t_prep = lap()
ctx.upload(m); ctx.synchronize(); t_up = lap()
ctx.render_into(args, 1, g, pl); t_render = lap()
ctx.upload(m); ctx.synchronize(); t_up2 = lap()
ctx.render_into(args, 2, g, pl); t_render2 = lap()
print("%s tree: context %.1f ms, prepare %.1f ms, upload %.1f ms, render %d spp %.1f ms | again: upload %.1f ms, render %.1f ms | tree %s" % (
    mode, t_ctx, t_prep, t_up, spp, t_render, t_up2, t_render2, ctx.tree_info()))
