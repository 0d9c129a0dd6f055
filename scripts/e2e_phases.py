"""Host-clock phases of one end-to-end frame (what bench.py's e2e loop runs per step), each followed by a synchronize:
    python scripts/e2e_phases.py [spp]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model
from raym0nade_b200.ctypes_defs import HITINFO_DTYPE, RADIANCE_DTYPE

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
scene, args = scenes.glossy_dielectric(1_000_000, 1920, 1080, spp)
npix = args.width * args.height
torch.cuda.set_device(0)
ctx = Context(0, stream=torch.cuda.current_stream().cuda_stream)
model = Model(scene)
model.pin()
g = torch.empty(npix * HITINFO_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(HITINFO_DTYPE)
pl = [torch.empty(npix * RADIANCE_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(RADIANCE_DTYPE) for _ in range(4)]
ctx.set_option("tree_cache", 0)
for rep in range(3):
    t = [time.perf_counter()]
    def lap():
        ctx.synchronize(); t.append(time.perf_counter()); return (t[-1] - t[-2]) * 1e3
    ctx.upload(model); a = lap()
    ctx.trace_primary(args, download=False); b = lap()
    ctx.gbuffer(args, download=False); c = lap()
    ctx.render_samples(args, seed=rep + 1); d = lap()
    ctx.resolve(args, download=False); e = lap()
    ctx.download_resolved(g, pl); f = lap()
    print("spp %d: upload %.1f  primary %.1f  gbuffer %.1f  render_samples %.1f  resolve %.1f  download %.1f  | total %.1f ms" % (spp, a, b, c, d, e, f, (t[-1] - t[0]) * 1e3))
