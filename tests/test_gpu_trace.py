"""GPU parity, stage K1 and the per-ray seam: CUDA traversal vs the reference oracle.

Bar (BASELINE.json north_star): the primary-ray triangle index must match the reference
exactly and the hit distance t within 1e-5 relative (we expect and assert bit equality).
"""
import numpy as np
import pytest

from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model

pytestmark = pytest.mark.gpu

T_REL_TOL = 1e-5   # the north star's stated tolerance; asserted in addition to bit equality


def _compare_primary(ref, scene, args, threads=8):
    model = Model(scene)
    R = ref.RefScene(scene)
    ctx = Context(0).upload(model)
    tri, t = ctx.trace_primary(args)
    rtri, rt = R.trace_primary(args, threads=threads)
    assert np.array_equal(tri, rtri), "tri_idx mismatches: %d of %d" % ((tri != rtri).sum(), tri.size)
    hit = rtri >= 0
    assert np.all(np.isinf(t[~hit])) and np.all(np.isinf(rt[~hit]))
    rel = np.abs(t[hit] - rt[hit]) / rt[hit]
    assert rel.max(initial=0.0) <= T_REL_TOL
    assert np.array_equal(t.view(np.uint32), rt.view(np.uint32)), "t not bit-equal"
    ctx.close()
    return hit.mean()


def test_primary_cornell(ref):
    scene, args = scenes.cornell_box(512, 512, 0)
    assert _compare_primary(ref, scene, args) > 0.9


def test_primary_heightfield_20k(ref):
    scene, args = scenes.heightfield_scene(20_000, 480, 270)
    assert _compare_primary(ref, scene, args) > 0.3


def test_primary_sponza_scale_260k(ref):
    scene, args = scenes.sponza_scale(260_000, 960, 540, tex_size=64, sky_size=(256, 128))
    assert _compare_primary(ref, scene, args) > 0.3


def test_primary_cutout_materials(ref):
    """alpha cut-outs exercise the <=8 re-trace loop of Model::rayHit (src/model.cpp:332-341)"""
    scene, args = scenes.texture_heavy(40_000, 320, 180, tex_size=64, n_materials=8)
    assert _compare_primary(ref, scene, args) > 0.3


@pytest.mark.parametrize("size", [(1, 1), (7, 3), (37, 53), (130, 65)])
def test_primary_ragged_frame_sizes(ref, size):
    """frames that do not fill the 8x4 pixel tiles a warp fetches (K1): the partial tiles at the right and bottom edges"""
    w, h = size
    scene, args = scenes.heightfield_scene(3000, w, h, 0, with_sky=True)
    _compare_primary(ref, scene, args)
    scene, args = scenes.cornell_box(w, h, 0)
    _compare_primary(ref, scene, args)


def test_per_ray_seam_edge_cases(ref):
    """empty batch, a single ray, a ray that starts far outside the scene and points away, axis-parallel rays"""
    scene, _ = scenes.cornell_box()
    ctx = Context(0).upload(Model(scene))
    R = ref.RefScene(scene)
    tri, t = ctx.trace_closest(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert tri.size == 0 and t.size == 0
    org = np.array([[0, 1, 0], [50, 50, 50], [0, 1, 0], [0.3, 1.2, 0.1], [0, 1, 0]], np.float32)
    d = np.array([[0, 0, -1], [0.57735, 0.57735, 0.57735], [1, 0, 0], [0, 1, 0], [0, -1, 0]], np.float32)
    tri, t = ctx.trace_closest(org, d)
    rtri, rt = R.trace_closest(org, d)
    assert np.array_equal(tri, rtri)
    assert np.array_equal(np.isnan(t), np.isnan(rt)) and np.array_equal(t[~np.isnan(t)].view(np.uint32), rt[~np.isnan(rt)].view(np.uint32))
    tri1, t1 = ctx.trace_closest(org[:1], d[:1])
    assert tri1[0] == tri[0] and t1[0] == t[0]
    ctx.close()


def test_primary_glossy_1m_1080p_full_size(ref):
    """BASELINE config 3's scene and frame at full size: every one of the 2 073 600 primary hits against the reference"""
    import os
    scene, args = scenes.glossy_dielectric(1_000_000, 1920, 1080, 0)
    assert _compare_primary(ref, scene, args, threads=os.cpu_count() or 8) > 0.3


def test_primary_five_million_4k_full_size(ref):
    """BASELINE config 5 at full size: 5 M triangles, 3840x2160 - tri_idx identical and t bit-equal for all
    8 294 400 pixels, then the FXAA pass over the shaded 4K frame bit-equal to Photo::FXAA"""
    import os
    scene, args = scenes.five_million(5_000_000, 3840, 2160)
    model = Model(scene)
    R = ref.RefScene(scene)
    ctx = Context(0).upload(model)
    tri, t = ctx.trace_primary(args)
    rtri, rt = R.trace_primary(args, threads=os.cpu_count() or 8)
    assert np.array_equal(tri, rtri), "tri_idx mismatches: %d of %d" % ((tri != rtri).sum(), tri.size)
    assert np.array_equal(t.view(np.uint32), rt.view(np.uint32)), "t not bit-equal"
    assert 0.3 < (rtri >= 0).mean() < 1.0
    # FXAA on the 4K frame shaded from the G-buffer (base colour x facing ratio stands in for the lit image: the
    # pass only sees an rgb frame); compared with the reference's own Photo::FXAA on the same input
    g = ctx.gbuffer(args.replace(spp=0))
    hit = ~np.isnan(g["position"][:, 0])
    view = np.asarray(args.direction, np.float32)
    facing = np.abs(g["shapeNormal"] @ view).astype(np.float32)
    img = np.where(hit[:, None], g["baseColor"] * facing[:, None], np.float32(0.1)).astype(np.float32).reshape(args.height, args.width, 3)
    got = ctx.fxaa(img)
    want = ref.fxaa(img)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert not np.array_equal(got, img)            # the pass did smooth edges
    ctx.close()


def _random_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene.positions.reshape(-1, 3).min(0), scene.positions.reshape(-1, 3).max(0)
    org = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    org[:, 1] = hi[1] * (0.3 + rng.random(n)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    # a share of axis-parallel directions: exercises the |d| < 1e-4 "parallel" branch of rayInBox
    d[: n // 16, 0] = 0.0
    d[n // 16: n // 8, 2] = 5e-5
    return org, d


@pytest.mark.parametrize("which", ["cornell", "heightfield", "cutout"])
def test_closest_and_occluded_random_rays(ref, which):
    if which == "cornell":
        scene, _ = scenes.cornell_box()
    elif which == "heightfield":
        scene, _ = scenes.heightfield_scene(20_000)
    else:
        scene, _ = scenes.texture_heavy(40_000, tex_size=64, n_materials=8)
    model = Model(scene)
    R = ref.RefScene(scene)
    ctx = Context(0).upload(model)
    org, d = _random_rays(scene, 20_000, 11)
    tri, t = ctx.trace_closest(org, d)
    rtri, rt = R.trace_closest(org, d)
    assert np.array_equal(tri, rtri)
    assert np.array_equal(t.view(np.uint32), rt.view(np.uint32))
    # occlusion: aim at, just before and just behind the closest hit, and at infinity
    hit = rtri >= 0
    aim = np.full(org.shape[0], np.inf, np.float32)
    aim[hit] = rt[hit] * np.where(np.arange(hit.sum()) % 3 == 0, 0.5, np.where(np.arange(hit.sum()) % 3 == 1, 1.0, 1.5)).astype(np.float32)
    occ = ctx.trace_occluded(org, d, aim)
    rocc = R.trace_occluded(org, d, aim)
    assert np.array_equal(occ, rocc)
    assert 0 < occ.sum() < occ.size
    ctx.close()


@pytest.mark.parametrize("tree,builder", [(1, 0), (2, 0), (2, 1), (2, 3)])
@pytest.mark.parametrize("which", ["cornell", "heightfield", "cutout", "glossy"])
def test_secondary_ray_tree_finds_the_reference_hits(ref, which, tree, builder):
    """The estimator's bounce and shadow rays traverse a second tree over the same triangles (fast_bvh.cpp: binned SAH,
    leaves <= 3; tree = 1) - by default in its 4-wide form with 8-bit quantised child boxes and a conservative slab test
    (wide_bvh.cpp; tree = 2), built on the host (builder = 0) or, the default, on the device by Morton sort + PLOC clustering
    + collapse (gpu_bvh.cu; builder = 1) - with the reference's triangle test.  It must find the closest accepted triangle the reference finds:
    checked here ray by ray against the compiled reference through a test hook that sends the per-ray seam through that
    tree.  Equal-t ties between triangles (shared edges) are the only freedom a different visit order has: t is
    bit-equal on every ray (with alpha cut-outs: on all but <= 0.02 %), the triangle index and the occlusion booleans on all
    but a sliver."""
    if which == "cornell":
        scene, _ = scenes.cornell_box()
    elif which == "heightfield":
        scene, _ = scenes.heightfield_scene(20_000)
    elif which == "cutout":
        scene, _ = scenes.texture_heavy(40_000, tex_size=64, n_materials=8)
    else:
        scene, _ = scenes.glossy_dielectric(200_000, 64, 36, 0)
    model = Model(scene)
    R = ref.RefScene(scene)
    ctx = Context(0)
    ctx.set_option("seam_secondary_tree", tree)          # before the upload: the binary form is only built when asked for
    ctx.set_option("tree_builder", builder)
    ctx.upload(model)
    org, d = _random_rays(scene, 50_000, 23)
    tri, t = ctx.trace_closest(org, d)
    rtri, rt = R.trace_closest(org, d)
    t_differs = t.view(np.uint32) != rt.view(np.uint32)
    if which == "cutout":
        # a tie at a shared edge can put the alpha test (and so the <= 8 re-traces) on the neighbouring triangle: 2 of 50 000 rays
        assert t_differs.mean() <= 2e-4, t_differs.sum()
    else:
        assert not t_differs.any(), "closest t differs on %d rays" % t_differs.sum()
    assert (tri != rtri).mean() <= 1e-3, (tri != rtri).mean()
    hit = rtri >= 0
    aim = np.full(org.shape[0], np.inf, np.float32)
    aim[hit] = rt[hit] * np.where(np.arange(hit.sum()) % 3 == 0, 0.5, np.where(np.arange(hit.sum()) % 3 == 1, 1.0, 1.5)).astype(np.float32)
    occ = ctx.trace_occluded(org, d, aim)
    rocc = R.trace_occluded(org, d, aim)
    assert (occ != rocc).mean() <= 1e-3, (occ != rocc).mean()
    ctx.close()


def test_upload_from_a_page_locked_prepared_scene():
    """rm_prepared_pin: the same scene uploaded from page-locked arrays (a DMA out of them) gives the same hits; pinning twice and
    releasing are harmless"""
    scene, args = scenes.texture_heavy(40_000, 160, 90, tex_size=64, n_materials=8)
    model = Model(scene)
    ctx = Context(0).upload(model)
    tri, t = ctx.trace_primary(args)
    model.pin()
    model.pin()
    ctx.upload(model)
    tri2, t2 = ctx.trace_primary(args)
    assert np.array_equal(tri, tri2) and np.array_equal(t.view(np.uint32), t2.view(np.uint32))
    out = ctx.render(args.replace(spp=2), seed=1)
    assert np.isfinite(out["Id"]["radiance"]).all()
    model.pin(False)
    ctx.upload(model)
    tri3, _ = ctx.trace_primary(args)
    assert np.array_equal(tri, tri3)
    ctx.close()
    model.close()


def test_work_counters_match_reference_traversal(ref):
    """box / triangle test counts come out identical because the traversal order is identical
    (these are the B and T of the bytes-per-ray roofline figure)"""
    from oracle import refbind
    if not refbind.available("count"):
        pytest.skip("counting oracle not built")
    scene, args = scenes.heightfield_scene(20_000, 160, 90)
    model = Model(scene)
    R = refbind.RefScene(scene, flavour="count")
    ctx = Context(0).upload(model)
    ctx.set_option("count_tests", 1)
    ctx.stats_reset()
    ctx.trace_primary(args)
    s = ctx.stats()
    _, _, cnt = R.trace_primary(args, counters=True)
    assert (s["rays"], s["box"], s["tri"]) == (int(cnt[0]), int(cnt[1]), int(cnt[2]))
    ctx.close()
