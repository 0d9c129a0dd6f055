"""CPU tests of the oracles: the restatement (oracle/port) against the golden vectors the
reference itself produced (tests/golden/reference_vectors.npz, made by make_golden.py from
oracle/_ref), and - where oracle/_ref is present - the golden file against a fresh run of the
reference, so a stale fixture cannot go unnoticed."""
import hashlib
import os

import numpy as np
import pytest

from oracle import portbind
from raym0nade_b200 import scenes

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same(a, b):
    """bit-equal, NaNs in the same places"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(bits(np.nan_to_num(a, nan=0.0)), bits(np.nan_to_num(b, nan=0.0)))


def _scene(name):
    return {"cornell": lambda: scenes.cornell_box(64, 64, 0), "hf": lambda: scenes.heightfield_scene(3000, 96, 54, 0, with_sky=True),
            "tex": lambda: scenes.texture_heavy(6000, 96, 54, 0, tex_size=32, n_materials=8)}[name]()


def test_port_geometry_kernels_match_reference_vectors():
    assert same(portbind.ray_in_box(G["box_rays"], G["box_boxes"], G["box_tlr_in"]), G["box_tlr_out"])
    assert same(portbind.ray_triangle(G["tri_rays"], G["tri_tris"]), G["tri_t"])
    assert same(portbind.barycentric(G["bary_tris"], G["bary_p"]), G["bary_out"])


def test_port_bsdf_matches_reference_vectors():
    for which, name in [(0, "bsdf_out"), (1, "brdf_out"), (2, "btdf_out")]:
        got = portbind.bsdf_eval(which, G["bsdf_surf"], G["bsdf_V"], G["bsdf_L"])
        assert same(got, G[name]), name
    assert (G["brdf_out"] > 0).any() and (G["btdf_out"] > 0).any()


def test_port_radiance_split_matches_reference_vectors():
    assert same(portbind.accumulate(G["acc_base"], G["acc_s7"]), G["acc_out"])


def test_port_rng_mapping_matches_reference_vectors():
    out = portbind.uniform_from_u32(G["rng_u32"])
    assert same(out, G["rng_out"])
    assert out.min() >= np.float32(1e-6) and out.max() < 1.0          # the open range the estimator relies on


def test_port_fxaa_matches_reference_vectors():
    assert same(portbind.fxaa(G["fxaa_in"]), G["fxaa_out"])


@pytest.mark.parametrize("name", ["cornell", "hf", "tex"])
def test_port_scene_stages_match_reference_vectors(name):
    scene, args = _scene(name)
    h = hashlib.sha256(scene.positions.tobytes() + scene.uvs.tobytes() + scene.normals.tobytes()).digest()
    assert np.array_equal(np.frombuffer(h, np.uint8), G[name + "_scene_hash"]), "scene generator is not reproducing the golden scene"
    P = portbind.PortScene(scene)
    nodes, perm = P.bvh()
    assert np.array_equal(perm, G[name + "_perm"])
    assert nodes.tobytes() == G[name + "_nodes"].tobytes()
    tri, t = P.trace_primary(args, threads=4)
    assert np.array_equal(tri, G[name + "_tri"]) and same(t, G[name + "_t"])
    ctri, ct = P.trace_closest(G[name + "_rays_o"], G[name + "_rays_d"])
    assert np.array_equal(ctri, G[name + "_rays_tri"]) and same(ct, G[name + "_rays_t"])
    assert np.array_equal(P.trace_occluded(G[name + "_rays_o"], G[name + "_rays_d"], G[name + "_rays_aim"]), G[name + "_rays_occ"])
    g, rg = P.gbuffer(args, threads=4), G[name + "_gbuffer"]
    for k in ["shapeNormal", "surfaceNormal", "emission", "baseColor", "position", "specular", "roughness", "metallic", "opacity", "eta"]:
        assert same(g[k], rg[k]), (name, k)
    assert np.array_equal(g["id"], rg["id"]) and np.array_equal(g["entering"], rg["entering"])
    if name == "tex":
        for which in range(4):
            assert same(P.material_fetch(1, which, G["tex_uvd"]), G["tex_fetch%d" % which]), which
    if name == "hf":
        assert same(P.sky_get(G["sky_dirs"]), G["sky_out"])
    P.close()


def test_port_counts_box_and_triangle_tests():
    scene, args = _scene("hf")
    P = portbind.PortScene(scene)
    _, _, cnt = P.trace_primary(args, threads=2, counters=True)
    rays, box, tri = [int(c) for c in cnt]
    assert rays >= args.width * args.height and box > rays and tri > 0
    P.close()


def test_golden_vectors_are_current(ref):
    """regenerate a few vectors with the compiled reference and compare with the committed file"""
    assert same(ref.ray_in_box(G["box_rays"], G["box_boxes"], G["box_tlr_in"]), G["box_tlr_out"])
    assert same(ref.ray_triangle(G["tri_rays"], G["tri_tris"]), G["tri_t"])
    assert same(ref.bsdf_eval(1, G["bsdf_surf"], G["bsdf_V"], G["bsdf_L"]), G["brdf_out"])
    assert same(ref.fxaa(G["fxaa_in"]), G["fxaa_out"])
    scene, args = _scene("cornell")
    R = ref.RefScene(scene)
    tri, t = R.trace_primary(args, threads=2)
    assert np.array_equal(tri, G["cornell_tri"]) and same(t, G["cornell_t"])
    R.close()


def test_port_postprocess_matches_reference(ref):
    scene, args = scenes.cornell_box(48, 48, 4)
    R = ref.RefScene(scene)
    o = R.render(args, threads=2)
    for opts in [ref.SHADE["Full"], ref.SHADE["Full"] | ref.SHADE["DoFXAA"], ref.SHADE["shapeNormal"]]:
        want = ref.postprocess(o["gbuffer"], o["Dd"], o["Ds"], o["Id"], o["Is"], 48, 48, args.exposure, opts)
        got = portbind.postprocess(o["gbuffer"], o["Dd"], o["Ds"], o["Id"], o["Is"], 48, 48, args.exposure, opts)
        assert same(got, want), opts
    R.close()


# --------------------------------------------------------------------------- image-space passes
GP = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_vectors.npz"))
PLANES = ("Dd", "Ds", "Id", "Is")


def same_planes(a, b):
    return all(same(a[k]["radiance"], b[k]["radiance"]) and same(a[k]["Var"], b[k]["Var"]) for k in PLANES)


@pytest.mark.parametrize("name", ["box", "hf"])
@pytest.mark.parametrize("stages", [1, 2, 3])
def test_port_clamp_and_filter_match_reference_vectors(name, stages):
    """spatialClamp (1), filter (2) and both in the reference's order (3): the restatement is bit-equal"""
    w, h = (int(v) for v in GP[name + "_wh"])
    inp = [GP["%s_in_%s" % (name, k)] for k in PLANES]
    got = portbind.denoise(GP[name + "_gbuffer"], *inp, w, h, stages)
    want = {k: GP["%s_s%d_%s" % (name, stages, k)] for k in PLANES}
    assert same_planes(got, want)
    assert any(not same(got[k]["radiance"], i["radiance"]) for k, i in zip(PLANES, inp))       # the pass did something


@pytest.mark.parametrize("name", ["box", "hf"])
def test_port_bloom_matches_reference_vectors(name):
    w, h = (int(v) for v in GP[name + "_wh"])
    inp = [GP["%s_in_%s" % (name, k)] for k in PLANES]
    for opts in (63 | 256, 63 | 256 | 512):
        got = portbind.postprocess(GP[name + "_gbuffer"], *inp, w, h, float(GP[name + "_exposure"][0]), opts)
        assert same(got, GP["%s_post_%d" % (name, opts)]), opts


def test_post_golden_vectors_are_current(ref):
    for name in ("box", "hf"):
        w, h = (int(v) for v in GP[name + "_wh"])
        inp = [GP["%s_in_%s" % (name, k)] for k in PLANES]
        got = ref.denoise(GP[name + "_gbuffer"], *inp, w, h, 3)
        assert same_planes(got, {k: GP["%s_s3_%s" % (name, k)] for k in PLANES})


@pytest.mark.parametrize("name", ["box", "hf"])
def test_port_depth_of_field_matches_reference_vectors(name):
    """Photo::depthFeildBlur: the depth-ordered disc scatter with its 0.99 gain cap, restated bit for bit (the pass itself
    stays on the host - its result depends on the visiting order)"""
    for j in (0, 1):
        focus, coc = (float(v) for v in GP["%s_dof_%d_params" % (name, j)])
        got = portbind.depth_field_blur(GP[name + "_gbuffer"], GP[name + "_dof_in"], GP[name + "_dof_cam"], focus, coc)
        assert same(got, GP["%s_dof_%d" % (name, j)]), (name, j)
        assert not same(got, GP[name + "_dof_in"])
