"""GPU test of the C++ host side: the console binary (raym0nade_b200/host) renders a scene through the C ABI and exports
the reference's set of PNGs (src/render.cpp:635-674); the images equal what the Python host gets from the same library
calls - byte for byte where only the G-buffer is involved, up to fp32 summation order where sampled radiance is."""
import os
import subprocess

import numpy as np
import pytest

from raym0nade_b200 import build, scenes
from raym0nade_b200.api import Context, Model
from test_cpu_console import read_png

pytestmark = pytest.mark.gpu

EXPORTS = ["DiffuseColor", "DiffuseColor_FXAA", "shapeNormal", "surfaceNormal",
           "Direct_Diffuse", "Direct_Specular", "Indirect_Diffuse", "Indirect_Specular", "Raw", "Raw_Bloom", "Raw_FXAA", "Raw_Bloom_FXAA",
           "Direct_Diffuse_Filter", "Direct_Specular_Filter", "Indirect_Diffuse_Filter", "Indirect_Specular_Filter",
           "Filter", "Filter_Bloom", "Filter_FXAA", "Filter_Bloom_FXAA"]
EXPORTS_DOF = ["BaseColor_DepthFieldBlur", "Filter_DepthFieldBlur", "Filter_DepthFieldBlur_Bloom", "Filter_DepthFieldBlur_FXAA",
               "Filter_DepthFieldBlur_Bloom_FXAA"]


def to_bytes(img):
    """Photo::save's conversion (src/image.cpp:516-518): byte(pixel * 255), truncating"""
    with np.errstate(invalid="ignore"):
        return (np.nan_to_num(img, nan=0.0) * np.float32(255)).astype(np.uint8)


def test_console_render_exports_match_the_python_host(tmp_path):
    console = build.HOST_OUT if os.path.exists(build.HOST_OUT) else build.build_host()
    scene, a0 = scenes.cornell_box(96, 96, 8)
    scene.save(str(tmp_path / "cornell.rmscene"))
    g = lambda v: " ".join("%.9g" % np.float32(x) for x in v)
    d, r, u, p = (np.float32(a0.direction), np.float32(a0.right), np.float32(a0.up), np.float32(a0.position))
    text = "\n".join([g(d), g(r), g(u), g([np.dot(p, d), np.dot(p, r), np.dot(p, u)]),
                      g([a0.accuracy, 3.0, 4.0, a0.exposure]), "96 96", "8 1 %s" % g([a0.P_Direct]), str(tmp_path / "out")])
    args = scenes.RenderArgs.from_console(text)                 # the Python mirror parses the very same text
    assert np.allclose(args.position, a0.position, atol=1e-6) and args.CoC == 4.0 and args.spp == 8
    script = "create model box\n%s/\ncornell.rmscene\nnull\ncreate args a\n%s\nrender box a\nexit\n" % (tmp_path, text)
    run = subprocess.run([console], input=script, capture_output=True, text=True, timeout=300, env=dict(os.environ, RM_SEED="0", RM_RANK="0", RM_WORLD="1", RM_DEVICE="0"))       # one process, one GPU, whatever the launcher of the test run set
    assert run.returncode == 0, run.stderr
    assert "Rendering completed in" in run.stdout and "Post processing finished." in run.stdout, run.stdout + run.stderr
    png = {}
    for tag in EXPORTS + EXPORTS_DOF:
        path = str(tmp_path / ("out(%s).png" % tag))
        assert os.path.exists(path), (tag, run.stderr)
        png[tag] = read_png(path)
        assert png[tag].shape == (96, 96, 3), tag
    assert png["Raw"].mean() > 10 and png["shapeNormal"].std() > 10            # not black, not flat

    # the same sequence of library calls from the Python host
    ctx = Context(0).upload(Model(scene))
    ctx.render(args, seed=0)
    exact = {"DiffuseColor": ctx.postprocess(args, 1), "DiffuseColor_FXAA": ctx.postprocess(args, 1 | 512),
             "shapeNormal": ctx.postprocess(args, 64), "surfaceNormal": ctx.postprocess(args, 128)}
    ctx.spatial_clamp(args)
    full = 4 | 8 | 16 | 32 | 1 | 2
    sampled = {"Raw": ctx.postprocess(args, full), "Direct_Diffuse": ctx.postprocess(args, 4 | 16)}
    ctx.filter(args)
    sampled["Filter"] = ctx.postprocess(args, full)
    exact["BaseColor_DepthFieldBlur"] = ctx.postprocess(args, 1 | 1024)
    ctx.close()
    for tag, img in exact.items():                               # G-buffer only: deterministic, so byte-identical
        finite = np.isfinite(img).all(-1)
        assert finite.mean() > 0.9, tag
        assert np.array_equal(png[tag][finite], to_bytes(img)[finite]), tag
    for tag, img in sampled.items():                             # fp32 atomic adds land in a different order run to run
        diff = np.abs(png[tag].astype(np.int32) - to_bytes(img).astype(np.int32))
        assert (diff > 1).mean() < 0.005, (tag, diff.max(), (diff > 1).mean())
