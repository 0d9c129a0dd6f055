"""GPU parity, SURVEY.md section 8f row 3: the reference's tree built on the device (rm_tree_build: BVH::build, src/bvh.cpp:18-54)
and the refit of an uploaded scene's trees for moved vertices (rm_scene_refit).

What is held to the reference bit for bit: the tree's shape (node count, heap layout, every leaf range - they follow from the
face count alone), every box given the faces beneath it (Face::aabb + Box::operator+ are min / max), and the hits rays find in
it.  What the reference itself does not define - which of several equal keys std::nth_element leaves left of a median, the order
inside a range, the axis where two fp32 running sums tie - is checked as a property instead: every split separates its two
halves along the axis the rule names for the faces actually in the node."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model, RmSceneDesc, lib, _check, _p

pytestmark = pytest.mark.gpu


def _centres(pos):
    """Face::center(): (v0 + v1 + v2) * (1 / 3) in fp32, glm's operation order"""
    return ((pos[:, 0] + pos[:, 1]) + pos[:, 2]) * np.float32(np.float32(1.0) / np.float32(3.0))


def _check_tree(raw_pos, nodes, perm, n_ref_nodes=None):
    """every property of BVH::dfs_build the reference defines, for the tree (nodes, perm) over raw_pos [n][3][3]"""
    n = raw_pos.shape[0]
    assert sorted(perm.tolist()) == list(range(n)), "perm is not a permutation"
    pos = raw_pos[perm]
    cen = _centres(pos).astype(np.float64)
    fmin = pos.min(axis=1)
    fmax = pos.max(axis=1)
    if n_ref_nodes is not None:
        assert nodes.shape[0] == n_ref_nodes
    stack = [(1, 0, n)]
    leaves = inner = near_ties = 0
    box = {}
    order = []
    while stack:                      # ranges as dfs_build hands them down
        u, L, R = stack.pop()
        order.append((u, L, R))
        if R - L <= 10:
            assert (nodes["faceL"][u], nodes["faceR"][u]) == (L, R), (u, L, R)
            leaves += 1
            continue
        assert nodes["faceR"][u] == 0 and nodes["faceL"][u] == 0, u
        inner += 1
        c = cen[L:R]
        D = (c * c).sum(0) - c.sum(0) ** 2 / (R - L)
        axis = 0
        if D[1] > D[0]:
            axis = 1
        if D[2] > D[0] and D[2] > D[1]:
            axis = 2
        M = (L + R) // 2
        sep = [cen[L:M, a].max() <= cen[M:R, a].min() for a in range(3)]
        if not sep[axis]:
            # the rule's axis is decided by sums in another order on the device (prefix sums): only a near-tie may flip it
            alt = [a for a in range(3) if sep[a]]
            assert alt, "node %d [%d, %d) is not split along any axis" % (u, L, R)
            assert abs(D[alt[0]] - D[axis]) <= 1e-9 * max(abs(D).max(), 1e-30), (u, D, alt)
            near_ties += 1
        stack.append((u << 1, L, M))
        stack.append((u << 1 | 1, M, R))
    for u, L, R in reversed(order):   # children before parents
        if R - L <= 10:
            lo, hi = fmin[L:R].min(0), fmax[L:R].max(0)
        else:
            lo = np.minimum(box[u << 1][0], box[u << 1 | 1][0])
            hi = np.maximum(box[u << 1][1], box[u << 1 | 1][1])
        box[u] = (lo, hi)
        assert np.array_equal(nodes["v0"][u], lo) and np.array_equal(nodes["v1"][u], hi), u
    used = np.zeros(nodes.shape[0], bool)
    used[[u for u, _, _ in order]] = True
    rest = nodes[~used]
    assert not rest["faceL"].any() and not rest["faceR"].any() and not rest["v0"].any() and not rest["v1"].any(), "never-written slots must be zero"
    return leaves, inner, near_ties


@pytest.mark.parametrize("which", ["tiny", "cornell", "heightfield", "cutout", "duplicates"])
def test_device_built_reference_tree_follows_the_rule(which):
    """rm_tree_build against BVH::dfs_build's definition, node by node (numpy, double precision)"""
    if which == "tiny":
        scene, _ = scenes.cornell_box()
        scene.positions = scene.positions[:7].copy()
    elif which == "cornell":
        scene, _ = scenes.cornell_box()
    elif which == "heightfield":
        scene, _ = scenes.heightfield_scene(20_000)
    elif which == "cutout":
        scene, _ = scenes.texture_heavy(40_000, tex_size=64, n_materials=8)
    else:
        scene, _ = scenes.heightfield_scene(3_000)
        scene.positions = np.concatenate([scene.positions] * 3)       # every face three times: ties across every median
    raw = np.ascontiguousarray(scene.positions, np.float32).reshape(-1, 3, 3)
    ctx = Context(0)
    nodes, perm = ctx.tree_build(raw)
    leaves, inner, ties = _check_tree(raw, nodes, perm, lib().rm_tree_node_count(raw.shape[0]))
    assert leaves == inner + 1
    if which in ("heightfield", "cutout"):
        assert ties <= inner // 100
    ctx.close()


@pytest.mark.parametrize("which", ["cornell", "heightfield", "cutout", "glossy"])
def test_device_built_reference_tree_against_the_reference(ref, which):
    """Model(raw, ctx): the scene prepared with the device-built tree.  Same shape as the reference's own tree (every leaf range),
    the reference's rule at every node, and primary rays find the
    reference's hits: t bit-equal, the raw face equal (ties at shared edges aside)."""
    if which == "cornell":
        scene, args = scenes.cornell_box(256, 256, 0)
    elif which == "heightfield":
        scene, args = scenes.heightfield_scene(20_000, 480, 270)
    elif which == "cutout":
        scene, args = scenes.texture_heavy(40_000, 320, 180, tex_size=64, n_materials=8)
    else:
        scene, args = scenes.glossy_dielectric(200_000, 480, 270, 0)
    ctx = Context(0)
    host = Model(scene)
    dev = Model(scene, ctx)
    dev.validate()
    hn, dn = host.nodes(), dev.nodes()
    assert hn.shape == dn.shape
    assert np.array_equal(hn["faceL"], dn["faceL"]) and np.array_equal(hn["faceR"], dn["faceR"])
    hp, dp = host.permutation(), dev.permutation()
    # (which faces a leaf holds is the reference's only up to ties: a regular grid has a tie across every median, and there
    # std::nth_element and a sort choose differently - the rule itself is checked node by node in the test above)
    _check_tree(np.ascontiguousarray(scene.positions, np.float32).reshape(-1, 3, 3), dn, dp)
    assert np.array_equal(hn["v0"][1], dn["v0"][1]) and np.array_equal(hn["v1"][1], dn["v1"][1])       # the scene bounds
    ctx.upload(dev)
    tri, t = ctx.trace_primary(args)
    R = ref.RefScene(scene)
    rtri, rt = R.trace_primary(args, threads=8)
    t_differs = t.view(np.uint32) != rt.view(np.uint32)
    assert t_differs.mean() <= (2e-4 if which == "cutout" else 1e-5), t_differs.sum()
    hit = (tri >= 0) & (rtri >= 0)
    assert np.array_equal(tri >= 0, rtri >= 0) or t_differs.any()
    raw_dev, raw_ref = dp[tri[hit]], hp[rtri[hit]]
    assert (raw_dev != raw_ref).mean() <= 1e-3, (raw_dev != raw_ref).mean()
    # the per-ray seam and the estimator run on the device-prepared scene like on any other
    out = ctx.render(args.replace(spp=2), seed=1)
    assert np.isfinite(out["Id"]["radiance"]).all()
    ctx.close()


def test_device_built_tree_one_million(ref):
    """configs[2]'s scene: the build on the device, the rule checked per node, and the primary hits of the full 1080p frame"""
    import os
    scene, args = scenes.glossy_dielectric(1_000_000, 1920, 1080, 0)
    raw = np.ascontiguousarray(scene.positions, np.float32).reshape(-1, 3, 3)
    ctx = Context(0)
    dev = Model(scene, ctx)
    _check_tree(raw, dev.nodes(), dev.permutation(), lib().rm_tree_node_count(raw.shape[0]))
    ctx.upload(dev)
    tri, t = ctx.trace_primary(args)
    host = Model(scene)
    rtri, rt = ref.RefScene(scene).trace_primary(args, threads=os.cpu_count() or 8)
    assert (t.view(np.uint32) != rt.view(np.uint32)).mean() <= 1e-5
    hit = (tri >= 0) & (rtri >= 0)
    assert (dev.permutation()[tri[hit]] != host.permutation()[rtri[hit]]).mean() <= 1e-3
    ctx.close()


def _upload_desc(ctx, model, nodes, positions):
    """rm_scene_upload of `model`'s prepared scene with another node array and other (post-build order) positions"""
    d = RmSceneDesc()
    C.memmove(C.byref(d), C.byref(model.desc), C.sizeof(RmSceneDesc))
    d.nodes = nodes.ctypes.data
    d.positions = positions.ctypes.data
    _check(lib().rm_scene_upload(ctx.h, C.byref(d)))
    ctx.model = model


def _host_refit(nodes, pos):
    """dfs_build's box arithmetic over a fixed topology (numpy): leaf boxes from their faces, inner boxes from their children"""
    out = nodes.copy()
    fmin, fmax = pos.min(axis=1), pos.max(axis=1)
    order, stack = [], [1]
    while stack:
        u = stack.pop()
        order.append(u)
        if not nodes["faceR"][u]:
            stack += [u << 1, u << 1 | 1]
    for u in reversed(order):
        if nodes["faceR"][u]:
            L, R = nodes["faceL"][u], nodes["faceR"][u]
            out["v0"][u], out["v1"][u] = fmin[L:R].min(0), fmax[L:R].max(0)
        else:
            out["v0"][u] = np.minimum(out["v0"][u << 1], out["v0"][u << 1 | 1])
            out["v1"][u] = np.maximum(out["v1"][u << 1], out["v1"][u << 1 | 1])
    return out


@pytest.mark.parametrize("builder", [3, 1, 0])
@pytest.mark.parametrize("which", ["heightfield", "glossy"])
def test_refit_for_moved_vertices(ref, which, builder):
    """rm_scene_refit: the vertices of an uploaded scene move (a travelling wave, amplitude ~ a few triangle sizes).  Afterwards
      * primary rays - the reference's tree, refitted in place - return bit for bit what a fresh upload of the same topology
        with host-recomputed boxes returns, and the reference's own answers for the moved scene (its own new tree) in t;
      * the refitted secondary-ray tree (device-built, builder = 1, or host-built, builder = 0) finds the same closest hits and
        occlusion answers as the reference does on the moved scene."""
    if which == "heightfield":
        scene, args = scenes.heightfield_scene(20_000, 320, 180)
    else:
        scene, args = scenes.glossy_dielectric(200_000, 320, 180, 0)
    model = Model(scene)
    ctx = Context(0)
    ctx.set_option("tree_builder", builder)
    ctx.set_option("lazy_tree", 0)           # the tree is built by the upload, so that it is there to be refitted
    ctx.upload(model)
    n = model.n_faces
    old = np.frombuffer((C.c_char * (36 * n)).from_address(model.desc.positions), np.float32).reshape(n, 3, 3).copy()
    ext = old.reshape(-1, 3).max(0) - old.reshape(-1, 3).min(0)
    new = old.copy()
    new[..., 1] += (0.02 * ext[1] * np.sin(old[..., 0] * (12.0 / ext[0])) * np.cos(old[..., 2] * (9.0 / max(ext[2], 1e-6)))).astype(np.float32)
    new[..., 0] += (0.01 * ext[0] * np.sin(old[..., 2] * (7.0 / max(ext[2], 1e-6)))).astype(np.float32)
    new = np.ascontiguousarray(new, np.float32)
    with pytest.raises(Exception):
        ctx.refit(new[:-1])                  # a face count other than the uploaded scene's is refused
    ctx.refit(new)
    tri, t = ctx.trace_primary(args)
    # (a) the same topology with boxes recomputed on the host, uploaded afresh
    ctx2 = Context(0)
    ctx2.set_option("tree_builder", builder)
    _upload_desc(ctx2, model, _host_refit(model.nodes(), new), new)
    tri2, t2 = ctx2.trace_primary(args)
    assert np.array_equal(tri, tri2) and np.array_equal(t.view(np.uint32), t2.view(np.uint32))
    # (b) the reference on the moved scene (raw order restored; it builds its own tree)
    raw_new = np.empty_like(new)
    raw_new[model.permutation()] = new
    moved = dataclasses.replace(scene, positions=raw_new.reshape(np.asarray(scene.positions).shape))
    R = ref.RefScene(moved)
    rtri, rt = R.trace_primary(args, threads=8)
    assert (t.view(np.uint32) != rt.view(np.uint32)).mean() <= 1e-5
    assert 0.2 < (rtri >= 0).mean()
    # (c) the refitted secondary-ray tree through the per-ray seam
    from test_gpu_trace import _random_rays
    org, d = _random_rays(moved, 50_000, 5)
    rtri, rt = R.trace_closest(org, d)
    differs = []
    for c in (ctx, ctx2):
        c.set_option("seam_secondary_tree", 2)
        stri, st = c.trace_closest(org, d)
        differs.append(int((st.view(np.uint32) != rt.view(np.uint32)).sum()))
        hit = rtri >= 0
        aim = np.full(org.shape[0], np.inf, np.float32)
        aim[hit] = rt[hit] * np.where(np.arange(hit.sum()) % 3 == 0, 0.5, np.where(np.arange(hit.sum()) % 3 == 1, 1.0, 1.5)).astype(np.float32)
        assert (c.trace_occluded(org, d, aim) != R.trace_occluded(org, d, aim)).mean() <= 1e-3
    # t is the reference's on every ray but a grazing one or two (rayInBox has no slack on its near plane: a triangle lying in a
    # face of its leaf box can be missed by the reference's own tree and found by the conservative test of this one) - and the
    # refitted tree is no different in that from one built afresh for the moved vertices
    assert max(differs) <= 5 and abs(differs[0] - differs[1]) <= 3, differs
    info = ctx.tree_info()
    assert info["in_use"] and info["device_built"] == bool(builder)
    # and the estimator runs on the refitted scene
    out = ctx.render(args.replace(spp=2), seed=3)
    out2 = ctx2.render(args.replace(spp=2), seed=3)
    a, b = out["Id"]["radiance"].astype(np.float64), out2["Id"]["radiance"].astype(np.float64)
    assert np.isfinite(a).all() and abs(a.sum() - b.sum()) <= 0.05 * abs(b.sum()) + 1e-6
    ctx.close()
    ctx2.close()


def test_secondary_tree_is_built_when_first_needed():
    """The sweep-SAH build of the secondary-ray tree is deferred to the first call that needs the tree (rm_ensure_secondary_tree):
    an upload followed by primary rays and a G-buffer launches no builder kernel, the first render builds the tree, and the frame
    is the frame of a context whose upload built it ("lazy_tree" 0); vertices moved in between are what it is built from."""
    scene, args = scenes.glossy_dielectric(60_000, 160, 90, 4)
    model = Model(scene)
    lazy, eager = Context(0), Context(0)
    eager.set_option("lazy_tree", 0)
    for c in (lazy, eager):
        c.stats_reset()
        c.upload(model)
    l0, e0 = lazy.stats()["launches"], eager.stats()["launches"]
    assert l0 + 50 < e0, (l0, e0)                      # 20-odd levels of the builder at a handful of launches each
    lazy.trace_primary(args)
    lazy.gbuffer(args)
    assert lazy.stats()["launches"] <= l0 + 4
    a, b = lazy.render(args, seed=9), eager.render(args, seed=9)
    for plane in ("Dd", "Ds", "Id", "Is"):              # (same tree, same draws; float atomics sum in their own order)
        x, y = a[plane]["radiance"].astype(np.float64), b[plane]["radiance"].astype(np.float64)
        assert np.abs(x - y).max() <= 2e-4 * (1.0 + np.abs(x).max()), plane
        assert abs(x.sum() - y.sum()) <= 1e-5 * abs(x.sum()) + 1e-6, plane
    ia, ib = lazy.tree_info(), eager.tree_info()
    assert ia == ib and ia["builder"] == "sweep_sah" and ia["in_use"]
    # moved vertices before the first render: the deferred build sees the new positions (the refit has nothing to refit yet)
    n = model.n_faces
    old = np.frombuffer((C.c_char * (36 * n)).from_address(model.desc.positions), np.float32).reshape(n, 3, 3).copy()
    new = old.copy()
    new[..., 1] += (0.05 * np.sin(old[..., 0] * 3.0)).astype(np.float32)
    new = np.ascontiguousarray(new, np.float32)
    c1, c2 = Context(0), Context(0)
    c2.set_option("lazy_tree", 0)
    for c in (c1, c2):
        c.upload(model)
        c.refit(new)
    r1, r2 = c1.render(args, seed=11), c2.render(args, seed=11)
    s1, s2 = float(r1["Id"]["radiance"].astype(np.float64).sum()), float(r2["Id"]["radiance"].astype(np.float64).sum())
    assert np.isfinite(r1["Id"]["radiance"]).all() and abs(s1 - s2) <= 0.05 * abs(s2) + 1e-6
    for c in (lazy, eager, c1, c2):
        c.close()


@pytest.mark.parametrize("k", [1, 2, 3, 5])
def test_sweep_sah_tree_of_a_handful_of_triangles(ref, k):
    """the device's sweep-SAH builder on scenes of 1, 2, 3 and 5 triangles (no level loop at all; one split; one leaf; the first
    real collapse): the seam through its 4-wide tree finds the reference's hits, and a render comes back finite"""
    from test_gpu_trace import _random_rays
    b = scenes.SceneBuilder("handful")
    light, grey = b.emissive(255), b.diffuse(180, 180, 180)
    rng = np.random.default_rng(k)
    for i in range(k):
        c = np.array([1.5 * i, 0.0, 0.3 * i])
        tri = c + rng.uniform(-1.0, 1.0, (3, 3))
        b.add(tri[None], None, None, light if i == 0 else grey)
    scene = b.build()
    args = scenes.camera((0.5 * k, 0.5, 6.0), (0.5 * k, 0.0, 0.0), 64, 48, hfov_tan=0.8, exposure=1.0, P_Direct=0.5, spp=4)
    model = Model(scene)
    ctx = Context(0)
    ctx.set_option("seam_secondary_tree", 2)
    ctx.upload(model)
    info = ctx.tree_info()
    assert info["builder"] == "sweep_sah" and info["in_use"] and info["levels"] >= 1
    org, d = _random_rays(scene, 5_000, 3)
    tri, t = ctx.trace_closest(org, d)
    rtri, rt = ref.RefScene(scene).trace_closest(org, d)
    assert np.array_equal(t.view(np.uint32), rt.view(np.uint32)) and np.array_equal(tri, rtri)
    out = ctx.render(args, seed=2)
    for plane in ("Dd", "Ds", "Id", "Is"):
        assert np.isfinite(out[plane]["radiance"]).all()
    ctx.close()
