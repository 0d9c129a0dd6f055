"""Exploratory: render a scene on the GPU and with the reference oracle, dump PNGs + stats."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model
from oracle import refbind
from PIL import Image

os.makedirs("gpurun_out", exist_ok=True)
which = sys.argv[1] if len(sys.argv) > 1 else "cornell"
w, h, spp = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
if which == "cornell": scene, args = scenes.cornell_box(w, h, spp)
elif which == "sky": scene, args = scenes.heightfield_scene(20000, w, h, spp, with_sky=True)
elif which == "hf": scene, args = scenes.heightfield_scene(20000, w, h, spp)
elif which == "glass": scene, args = scenes.glossy_dielectric(60000, w, h, spp)
elif which == "tex": scene, args = scenes.texture_heavy(40000, w, h, spp, tex_size=128, n_materials=8)
model = Model(scene)
ctx = Context(0).upload(model)
t = time.time(); out = ctx.render(args, seed=1); tg = time.time() - t
st = ctx.stats()
print("gpu render %.3fs rays %d launches %d" % (tg, st["rays"], st["launches"]))
img = ctx.postprocess(args, refbind.SHADE["Full"])
Image.fromarray((np.clip(img, 0, 1) * 255).astype(np.uint8)).save("gpurun_out/%s_gpu.png" % which)
R = refbind.RefScene(scene)
ro = R.render(args, threads=os.cpu_count())
print("ref render %.3fs rays %d  (%.2f Mrays/s on %d threads)" % (ro["seconds"], ro["rays"], ro["rays"] / ro["seconds"] / 1e6, os.cpu_count()))
rimg = refbind.postprocess(ro["gbuffer"], ro["Dd"], ro["Ds"], ro["Id"], ro["Is"], args.width, args.height, args.exposure, refbind.SHADE["Full"])
Image.fromarray((np.clip(rimg, 0, 1) * 255).astype(np.uint8)).save("gpurun_out/%s_ref.png" % which)
for k in ["Dd", "Ds", "Id", "Is"]:
    a, b = out[k]["radiance"].astype(np.float64), ro[k]["radiance"].astype(np.float64)
    print(k, "mean gpu %.5f ref %.5f | nan gpu %d ref %d" % (np.nanmean(a), np.nanmean(b), np.isnan(a).sum(), np.isnan(b).sum()))
print("rays/pixel-sample gpu %.3f ref %.3f" % (st["rays"] / (w * h * max(spp, 1)), ro["rays"] / (w * h * max(spp, 1))))
