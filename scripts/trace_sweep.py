"""Perf experiment: trace-kernel throughput for several engine settings (runtime options)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene, args = scenes.glossy_dielectric(1_000_000, 1920, 1080, spp)
ctx = Context(0).upload(Model(scene))
ctx.trace_primary(args, download=False); ctx.gbuffer(args, download=False)

def run(label, max_depth=16, **opts):
    for k, v in opts.items(): ctx.set_option(k, v)
    ctx.set_option("max_depth", max_depth)
    ctx.render_samples(args, seed=1); ctx.synchronize()
    ctx.set_option("time_kernels", 1); ctx.stats_reset()
    t0 = time.time(); ctx.render_samples(args, seed=1); ctx.synchronize(); wall = time.time() - t0
    k = ctx.stats_kernels(); ctx.set_option("time_kernels", 0)
    out = "%-34s wall %7.1f ms |" % (label, wall * 1e3)
    for kind in ("paths", "shadow"):
        out += " %s %7.2f ms %6.1f Mr/s (%d launches) |" % (kind, k[kind]["ms"], k[kind]["rays"] / max(k[kind]["ms"], 1e-9) / 1e3, k[kind]["launches"])
    out += " shade %7.2f ms" % k["shade"]["ms"]
    print(out, flush=True)

variants = [("refill22 1:1", dict(trace_refill=22, trace_w_inner=1, trace_w_leaf=1))]
if len(sys.argv) > 2:
    variants = []
    for spec in sys.argv[2:]:
        if spec.startswith("s"):          # shade launch shape: s<block>:<ctas per sm>:<lockstep>
            b, c, l = (int(x) for x in spec[1:].split(":"))
            variants.append(("shade block%d ctas%d lockstep%d" % (b, c, l), dict(shade_block=b, shade_ctas=c, shade_lockstep=l)))
            continue
        r, wi, wl = (int(x) for x in spec.split(":"))
        variants.append(("refill%d %d:%d" % (r, wi, wl), dict(trace_refill=r, trace_w_inner=wi, trace_w_leaf=wl)))
for label, o in variants:
    run(label, 16, **o)
