"""Image-space passes at 1080p on the bench scene: device time of rm_spatial_clamp / rm_filter / rm_postprocess beside
the reference's Photo::spatialClamp / Photo::filter / Photo::postProcessing on the host cores (oracle/_ref), one JSON line.
Algorithmic bytes per pixel: clamp 4 planes x (16 B read + 16 B write) = 128; filter = pack (88 + 36)
+ variance pass 128 + 5 a-trous passes x (64 read + 36 G + 64 write) = 1072."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model

w, h, spp = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene, args = scenes.glossy_dielectric(1_000_000, w, h, spp)
ctx = Context(0).upload(Model(scene))
o = ctx.render(args, seed=1)
planes = [o[k] for k in ("Dd", "Ds", "Id", "Is")]

def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        ctx.upload_resolved(args, o["gbuffer"], planes); ctx.synchronize()
        t0 = time.perf_counter(); fn(); ctx.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3

rgb = np.zeros((h, w, 3), np.float32)
out = {"frame": "%dx%d" % (w, h), "gpu_ms": {}, "cpu_ms": {}, "cpu_threads": {"spatialClamp": 1, "filter": 4, "postProcessing": 1}}
out["gpu_ms"]["spatial_clamp"] = timed(lambda: ctx.spatial_clamp(args))
out["gpu_ms"]["filter"] = timed(lambda: ctx.filter(args))
out["gpu_ms"]["postprocess_full_bloom_fxaa_incl_d2h"] = timed(lambda: ctx.postprocess(args, 63 | 256 | 512))
out["gpu_ms"]["postprocess_full_incl_d2h"] = timed(lambda: ctx.postprocess(args, 63))
npix = w * h
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists("MEASURED_PEAKS.json") else 6650.0
out["roofline"] = {"peak_gbs": peak,
                   "spatial_clamp": {"bytes": 128 * npix, "gbs": 128 * npix / out["gpu_ms"]["spatial_clamp"] / 1e6},
                   "filter": {"bytes": 1072 * npix, "gbs": 1072 * npix / out["gpu_ms"]["filter"] / 1e6}}
for k in ("spatial_clamp", "filter"): out["roofline"][k]["frac"] = out["roofline"][k]["gbs"] / peak
try:
    from oracle import refbind
    if refbind.available("plain"):
        saved = os.dup(1); os.dup2(2, 1)
        t0 = time.perf_counter(); refbind.denoise(o["gbuffer"], *planes, w, h, 1); out["cpu_ms"]["spatialClamp"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); refbind.denoise(o["gbuffer"], *planes, w, h, 2); out["cpu_ms"]["filter"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); refbind.postprocess(o["gbuffer"], *planes, w, h, args.exposure, 63 | 256 | 512); out["cpu_ms"]["postProcessing_full_bloom_fxaa"] = (time.perf_counter() - t0) * 1e3
        os.dup2(saved, 1); os.close(saved)
except Exception as e:
    out["cpu_ms"]["error"] = str(e)
print(json.dumps(out))
