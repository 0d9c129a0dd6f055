"""Statistical parity probe: GPU render vs the compiled reference (oracle/_ref) on the same scene and camera at equal spp.

    python scripts/stat_probe.py <scene> <width> <height> <spp> [exact_secondary]

scene: config2 (260 K tris + HDR sky), config3 (1 M tris glossy / dielectric), config4 (texture-heavy, reduced: 120 K tris,
12 materials x 256^2 with mips, normal maps and alpha cut-outs), cornell.  Prints one JSON line with the figures SURVEY.md
section 8(d) states its parity bounds on: per-plane z-scores from the Var planes (mean |z|, P(|z| > 4)), the relMSE against the
CPU-vs-CPU noise floor and the mean-image energy ratio - for GPU vs CPU and, as the yardstick, for CPU seed A vs CPU seed B.
tests/test_gpu_render.py holds the same computation as assertions; this script is how the bounds were looked at first.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from parity_stats import compare_renders
from oracle import refbind


def make(which, w, h, spp):
    if which == "config2":
        return scenes.sponza_scale(260_000, w, h, spp, tex_size=1024)
    if which == "config3":
        return scenes.glossy_dielectric(1_000_000, w, h, spp)
    if which == "config4":
        return scenes.texture_heavy(120_000, w, h, spp, tex_size=256, n_materials=12)
    if which == "cornell":
        return scenes.cornell_box(w, h, spp)
    if which == "nested":
        return scenes.nested_glass(w, h, spp)
    raise SystemExit("unknown scene " + which)


def main():
    which, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    exact = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    scene, args = make(which, w, h, spp)
    ctx = Context(0).upload(Model(scene))
    ctx.set_option("exact_secondary", exact)
    t = time.time()
    ga = ctx.render(args, seed=5)
    tg = time.time() - t
    gb = ctx.render(args, seed=6)
    threads = os.cpu_count() or 8
    saved = os.dup(1)
    os.dup2(2, 1)
    R = refbind.RefScene(scene)
    ca = R.render(args, threads=threads, seed_base=100)
    cb = R.render(args, threads=threads, seed_base=200)
    os.dup2(saved, 1)
    out = {"scene": which, "tris": scene.n_faces, "width": w, "height": h, "spp": spp, "exact_secondary": exact,
           "gpu_s": tg, "cpu_s": ca["seconds"], "threads": threads,
           "gpu_vs_cpu": compare_renders(ga, ca, args), "cpu_vs_cpu": compare_renders(ca, cb, args),
           "gpu_vs_gpu": compare_renders(ga, gb, args)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
