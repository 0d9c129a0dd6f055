"""Golden vectors for the image-space passes (spatialClamp, filter, bloom, depthFeildBlur), produced by the compiled reference
(oracle/_ref, i.e. the unmodified src/image.cpp).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_post.py

Inputs are real reference renders (G-buffer + four radiance planes) of two tiny scenes - a closed box and a
height field under a sky, whose background pixels carry NaN positions - with a few planted outliers so the
clamp fires; outputs are what Photo::spatialClamp / Photo::filter / Photo::postProcessing make of them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind          # noqa: E402
from raym0nade_b200 import scenes   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    g = {}
    rs = np.random.default_rng(2024)
    for name, (scene, args) in dict(box=scenes.cornell_box(40, 40, 16), hf=scenes.heightfield_scene(3000, 56, 32, 8, with_sky=True)).items():
        R = refbind.RefScene(scene)
        o = R.render(args, threads=1)
        R.close()
        w, h = args.width, args.height
        planes = [o[k].copy() for k in ("Dd", "Ds", "Id", "Is")]
        for p in planes:                                    # fireflies
            for i in rs.integers(0, w * h, 6):
                p["radiance"][i] *= np.float32(400.0)
        g[name + "_wh"] = np.array([w, h], np.int32)
        g[name + "_exposure"] = np.array([args.exposure], np.float32)
        g[name + "_gbuffer"] = o["gbuffer"]
        for k, p in zip(("Dd", "Ds", "Id", "Is"), planes):
            g["%s_in_%s" % (name, k)] = p
        for stages in (1, 2, 3):
            d = refbind.denoise(o["gbuffer"], *planes, w, h, stages)
            for k in ("Dd", "Ds", "Id", "Is"):
                g["%s_s%d_%s" % (name, stages, k)] = d[k]
        for opts in (refbind.SHADE["Full"] | 256, refbind.SHADE["Full"] | 256 | 512):
            g["%s_post_%d" % (name, opts)] = refbind.postprocess(o["gbuffer"], *planes, w, h, args.exposure, opts)
        # Photo::depthFeildBlur on the shaded frame (before gamma, as postProcessing orders it), two lens settings
        shaded = refbind.postprocess(o["gbuffer"], *planes, w, h, args.exposure, refbind.SHADE["BaseColor"])
        g[name + "_dof_in"] = shaded
        g[name + "_dof_cam"] = np.asarray(args.position, np.float32)
        for j, (focus, coc) in enumerate(((3.0, 4.0), (1.5, 24.0))):
            g["%s_dof_%d_params" % (name, j)] = np.array([focus, coc], np.float32)
            g["%s_dof_%d" % (name, j)] = refbind.depth_field_blur(o["gbuffer"], shaded, args.position, focus, coc)
    np.savez_compressed(os.path.join(OUT, "post_vectors.npz"), **g)
    print("wrote post_vectors.npz with", len(g), "arrays")


if __name__ == "__main__":
    main()
