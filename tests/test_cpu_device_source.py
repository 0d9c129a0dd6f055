"""The kernels' own source on the CPU: tests/tools/device_on_host.cpp compiles the device headers of raym0nade_b200/csrc
(dev_math / dev_trace / dev_surface / dev_bsdf / dev_texture .cuh) with g++ through a small shim and runs their deterministic
functions against the golden vectors the compiled reference produced (tests/golden/reference_vectors.npz).  The GPU suite
checks the same functions where they really run; this file keeps an arithmetic regression from slipping through a CPU-only
run.  The image-space kernels (FXAA, spatial clamp, filterVar + a-trous, shade / bloom / gamma) are run whole: launched on the
host thread by thread with the launch shapes of rm_render.cu - kernels with a shared-memory tile as real threads meeting at a
barrier - against tests/golden/post_vectors.npz; and the traversal engine runs as one emulated warp (32 real threads meeting at
every vote and shuffle) against the golden primary hits, closest hits and occlusion answers.  Bars: bit-equal for the arithmetic that is + - * / sqrt only; a few ulp where libm (powf, atan2f, acosf, sinf) is
involved - on the host that is glibc, the reference's own libm, and in this container every one of these comparisons comes
out bit-equal (0 ulp)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from raym0nade_b200 import scenes
from raym0nade_b200.api import Model, RmSceneDesc
from raym0nade_b200.ctypes_defs import HITINFO_DTYPE

pytestmark = pytest.mark.timeout(600)          # the emulated warps and blocks are real threads meeting at barriers: never hang the suite

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


@pytest.fixture(scope="module")
def doh(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    cxx = os.environ.get("CXX") or shutil.which("g++")
    out = str(tmp_path_factory.mktemp("doh") / "libdoh.so")
    cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-pthread", "-I", CUDA_INC, "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "raym0nade_b200", "csrc"), "-I", os.path.join(ROOT, "tests", "tools"),
           os.path.join(ROOT, "tests", "tools", "device_on_host.cpp"),
           os.path.join(ROOT, "raym0nade_b200", "csrc", "fast_bvh.cpp"), os.path.join(ROOT, "raym0nade_b200", "csrc", "rm_error.cpp"),      # the secondary-ray tree builder, as it is
           "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(out)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.doh_ray_in_box.argtypes = [i64, vp, vp, vp, i32]
    L.doh_ray_triangle.argtypes = [i64, vp, vp, vp]
    L.doh_barycentric.argtypes = [i64, vp, vp, vp]
    L.doh_bsdf.argtypes = [i32, i64, vp, vp, vp, vp]
    L.doh_uniform_from_u32.argtypes = [vp, i32, vp]
    L.doh_material_fetch.argtypes = [C.POINTER(RmSceneDesc), i32, i32, i64, vp, vp]
    L.doh_sky_get.argtypes = [C.POINTER(RmSceneDesc), i64, vp, vp]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def same(a, b):
    """bit-equal, NaNs in the same places"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a, nan=0.0).view(np.uint32), np.nan_to_num(b, nan=0.0).view(np.uint32))


def ulps(a, b):
    """largest distance in units of the last place between two finite float32 arrays"""
    a, b = np.asarray(a, np.float32).ravel(), np.asarray(b, np.float32).ravel()
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia, ib = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia), np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return int(np.abs(ia - ib).max()) if a.size else 0


def test_slab_test_both_paths(doh):
    """rayInBox (src/geometry.cpp:40-61) as slab_axis emulates it - early return, partial tL and all - and the fast path for
    rays without a parallel axis; the golden rays include 256 exactly axis-parallel ones and infinite tR"""
    rays, boxes = _f32(G["box_rays"]), _f32(G["box_boxes"])
    for variant in (0, 1):
        tlr = _f32(G["box_tlr_in"]).copy()
        doh.doh_ray_in_box(len(rays), _p(rays), _p(boxes), _p(tlr), variant)
        # the reference's early return leaves tR untouched where the emulation parks -1 (slab_axis `kill`): both mean "missed",
        # and every caller only asks tL < tR; compare tL bit for bit and tR wherever the box was entered
        want = G["box_tlr_out"]
        hit_ref, hit_dev = want[:, 0] < want[:, 1], tlr[:, 0] < tlr[:, 1]
        assert np.array_equal(hit_ref, hit_dev), variant
        assert same(tlr[hit_ref], want[hit_ref]), variant
        assert same(tlr[:, 0], want[:, 0]), variant          # the partial tL a missed box leaves decides the child order (bvh.cpp:75-87)
        assert hit_ref.sum() > 100 and (~hit_ref).sum() > 100


def test_triangle_test_and_barycentrics(doh):
    rays, tris = _f32(G["tri_rays"]), _f32(G["tri_tris"])
    t = np.zeros(len(rays), np.float32)
    doh.doh_ray_triangle(len(rays), _p(rays), _p(tris), _p(t))
    assert same(t, G["tri_t"])                                 # includes the 128 degenerate-edge triangles and the misses (+inf)
    assert np.isfinite(G["tri_t"]).mean() > 0.02
    tri, p = _f32(G["bary_tris"]), _f32(G["bary_p"])
    out = np.zeros((len(p), 3), np.float32)
    doh.doh_barycentric(len(p), _p(tri), _p(p), _p(out))
    assert same(out, G["bary_out"])


def test_bsdf_evaluation(doh):
    surf = np.ascontiguousarray(G["bsdf_surf"], HITINFO_DTYPE)
    V, Ldir = _f32(G["bsdf_V"]), _f32(G["bsdf_L"])
    for which, name in [(0, "bsdf_out"), (1, "brdf_out"), (2, "btdf_out")]:
        out = np.zeros((len(surf), 3), np.float32)
        doh.doh_bsdf(which, len(surf), _p(surf), _p(V), _p(Ldir), _p(out))
        want = G[name]
        assert np.array_equal(np.isnan(out), np.isnan(want)) and np.array_equal(out == 0, want == 0), name
        ok = np.isfinite(want) & (want != 0)
        assert np.allclose(out[ok], want[ok], rtol=2e-6, atol=0), (name, ulps(out[ok], want[ok]))
        assert (out.view(np.uint32)[ok] == want.view(np.uint32)[ok]).mean() > 0.9, name     # almost everywhere to the bit


def test_rng_float_mapping(doh):
    u32 = np.ascontiguousarray(G["rng_u32"], np.uint32)
    out = np.zeros(len(u32), np.float32)
    doh.doh_uniform_from_u32(_p(u32), len(u32), _p(out))
    assert same(out, G["rng_out"])


def test_material_fetches_and_sky_lookup(doh):
    scene, _ = scenes.texture_heavy(6000, 96, 54, 0, tex_size=32, n_materials=8)
    m = Model(scene)
    uvd = _f32(G["tex_uvd"])                                   # includes NaN footprints (no mip selection) and far out-of-range uv
    for which in range(4):
        out = np.zeros((len(uvd), 4), np.float32)
        doh.doh_material_fetch(C.byref(m.desc), 1, which, len(uvd), _p(uvd), _p(out))
        want = G["tex_fetch%d" % which]
        assert np.array_equal(np.isnan(out), np.isnan(want)), which
        ok = np.isfinite(want)
        assert ulps(out[ok], want[ok]) <= 2, (which, ulps(out[ok], want[ok]))
    scene, _ = scenes.heightfield_scene(3000, 96, 54, 0, with_sky=True)
    m = Model(scene)
    dirs = _f32(G["sky_dirs"])
    out = np.zeros((len(dirs), 3), np.float32)
    doh.doh_sky_get(C.byref(m.desc), len(dirs), _p(dirs), _p(out))
    assert ulps(out, G["sky_out"]) <= 2, ulps(out, G["sky_out"])


# --------------------------------------------------------------------------- whole kernels, run on the host thread by thread
GP = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_vectors.npz"))
PLANES = ("Dd", "Ds", "Id", "Is")


def _kernels(doh):
    vp, i32 = C.c_void_p, C.c_int32
    doh.doh_fxaa.argtypes = [vp, vp, i32, i32]
    doh.doh_denoise.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32]
    doh.doh_postprocess.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_float, i32, vp]
    return doh


def test_fxaa_kernel(doh):
    """k_fxaa - shared-memory luma tile, one barrier - run as 256 real threads per block: bit-equal to Photo::FXAA"""
    L = _kernels(doh)
    src = _f32(G["fxaa_in"])
    out = np.zeros_like(src)
    L.doh_fxaa(_p(src), _p(out), src.shape[1], src.shape[0])
    assert same(out, G["fxaa_out"])
    assert not same(out, src)


@pytest.mark.parametrize("name", ["box", "hf"])
@pytest.mark.parametrize("stages", [1, 2, 3])
def test_clamp_and_filter_kernels(doh, name, stages):
    """k_spatial_clamp (1), k_filter_pack + k_filter_var + five k_atrous passes (2), both in the reference's order (3), with the
    launch shapes of rm_spatial_clamp / rm_filter"""
    L = _kernels(doh)
    w, h = (int(v) for v in GP[name + "_wh"])
    planes = [np.ascontiguousarray(GP["%s_in_%s" % (name, k)]).copy() for k in PLANES]
    g = np.ascontiguousarray(GP[name + "_gbuffer"])
    L.doh_denoise(_p(g), *[_p(p) for p in planes], w, h, stages)
    for k, got in zip(PLANES, planes):
        want = GP["%s_s%d_%s" % (name, stages, k)]
        if stages == 1:
            assert same(got["radiance"], want["radiance"]) and same(got["Var"], want["Var"]), k       # + - * / only: to the bit
        else:                                                  # the weights go through exp2 approximations (k_atrous): the bar of tests/test_gpu_post.py
            for f in ("radiance", "Var"):
                a, b = got[f].astype(np.float64), want[f].astype(np.float64)
                assert np.array_equal(np.isnan(a), np.isnan(b)), (k, f)
                ok = np.isfinite(b)
                bad = np.abs(a[ok] - b[ok]) > 2e-4 * (1.0 + np.abs(b[ok]))
                assert bad.mean() <= 2e-3, (k, f, bad.mean(), np.abs(a[ok] - b[ok]).max())
                assert np.median(np.abs(a[ok] - b[ok]) / (1.0 + np.abs(b[ok]))) <= 2e-6, (k, f)
    assert any(not same(p["radiance"], GP["%s_in_%s" % (name, k)]["radiance"]) for k, p in zip(PLANES, planes))


@pytest.mark.parametrize("name", ["box", "hf"])
def test_shade_bloom_gamma_fxaa_kernels(doh, name):
    L = _kernels(doh)
    w, h = (int(v) for v in GP[name + "_wh"])
    planes = [np.ascontiguousarray(GP["%s_in_%s" % (name, k)]) for k in PLANES]
    g = np.ascontiguousarray(GP[name + "_gbuffer"])
    for opts in (63 | 256, 63 | 256 | 512):
        out = np.zeros((h, w, 3), np.float32)
        L.doh_postprocess(_p(g), *[_p(p) for p in planes], w, h, float(GP[name + "_exposure"][0]), opts, _p(out))
        want = GP["%s_post_%d" % (name, opts)]
        assert np.array_equal(np.isnan(out), np.isnan(want)), opts
        ok = np.isfinite(want)
        assert np.abs(out[ok] - want[ok]).max() <= 2e-6, (opts, np.abs(out[ok] - want[ok]).max())


# --------------------------------------------------------------------------- the traversal engine itself, as one emulated warp
def _scene(name):
    return {"cornell": lambda: scenes.cornell_box(64, 64, 0), "hf": lambda: scenes.heightfield_scene(3000, 96, 54, 0, with_sky=True),
            "tex": lambda: scenes.texture_heavy(6000, 96, 54, 0, tex_size=32, n_materials=8)}[name]()


@pytest.mark.parametrize("name", ["cornell", "hf", "tex"])
def test_trace_engine_as_an_emulated_warp(doh, name):
    """trace_engine (dev_trace.cuh) - persistent warp, ballot / popc lane refill, vote between the inner and the leaf step,
    per-lane stack, alpha cut-out re-traces - run as 32 real threads that meet at every __ballot_sync / __shfl_sync, on the
    reference's own tree: primary hits, closest hits and occlusion answers equal the compiled reference's golden vectors,
    triangle index for triangle index and t to the bit (the north star's first correctness bar, here without a GPU)"""
    from raym0nade_b200.ctypes_defs import RmRenderArgs
    L = doh
    vp, i32 = C.c_void_p, C.c_int32
    L.doh_trace_primary.argtypes = [C.POINTER(RmSceneDesc), C.POINTER(RmRenderArgs), vp, vp, vp]
    L.doh_trace_closest.argtypes = [C.POINTER(RmSceneDesc), i32, vp, vp, vp, vp]
    L.doh_trace_occluded.argtypes = [C.POINTER(RmSceneDesc), i32, vp, vp, vp, vp]
    scene, args = _scene(name)
    m = Model(scene)
    a = args.to_c()
    n = args.width * args.height
    tri, t, counts = np.full(n, -7, np.int32), np.zeros(n, np.float32), np.zeros(3, np.uint64)
    L.doh_trace_primary(C.byref(m.desc), C.byref(a), _p(tri), _p(t), _p(counts))
    assert np.array_equal(tri, G[name + "_tri"])
    assert same(t, G[name + "_t"])
    rays, box, tests = (int(c) for c in counts)
    assert rays >= n and box > rays and tests > 0              # cut-out re-traces count as rays again
    if name == "tex":
        assert rays > n                                        # the cut-out materials of this scene were re-traced through
    o, d = _f32(G[name + "_rays_o"]), _f32(G[name + "_rays_d"])
    ctri, ct = np.zeros(len(o), np.int32), np.zeros(len(o), np.float32)
    L.doh_trace_closest(C.byref(m.desc), len(o), _p(o), _p(d), _p(ctri), _p(ct))
    assert np.array_equal(ctri, G[name + "_rays_tri"]) and same(ct, G[name + "_rays_t"])
    aim = _f32(G[name + "_rays_aim"])
    occ = np.zeros(len(o), np.uint8)
    L.doh_trace_occluded(C.byref(m.desc), len(o), _p(o), _p(d), _p(aim), _p(occ))
    assert np.array_equal(occ, G[name + "_rays_occ"])


@pytest.mark.parametrize("name", ["cornell", "hf", "tex"])
@pytest.mark.parametrize("smem_levels", [14, 3])
def test_secondary_ray_tree_finds_the_reference_hits(doh, name, smem_levels):
    """The estimator's bounce and shadow rays traverse the library's second tree (fast_bvh.cpp: binned SAH, leaves <= 3, explicit
    child blocks, its own triangle order).  Same box test, same triangle test, same engine: the closest accepted hit is the
    reference's - t to the bit, the same triangle unless two triangles tie at that t - and so is every occlusion answer.
    smem_levels 14 is the context's setting; 3 forces deferred children into the engine's local spill array."""
    L = doh
    vp, i32 = C.c_void_p, C.c_int32
    L.doh_trace_closest_secondary.argtypes = [C.POINTER(RmSceneDesc), i32, vp, vp, vp, vp, i32, vp]
    L.doh_trace_occluded_secondary.argtypes = [C.POINTER(RmSceneDesc), i32, vp, vp, vp, vp, i32]
    scene, _ = _scene(name)
    m = Model(scene)
    o, d = _f32(G[name + "_rays_o"]), _f32(G[name + "_rays_d"])
    tri, t, shape = np.zeros(len(o), np.int32), np.zeros(len(o), np.float32), np.zeros(4, np.int32)
    assert L.doh_trace_closest_secondary(C.byref(m.desc), len(o), _p(o), _p(d), _p(tri), _p(t), smem_levels, _p(shape)) == 0
    assert same(t, G[name + "_rays_t"])
    differ = tri != G[name + "_rays_tri"]
    assert differ.mean() < 0.01                                # only exact ties may name another triangle of the same surface point
    if differ.any():
        pos = np.float32(scene.positions)[m.permutation()]
        for i in np.nonzero(differ)[0]:                        # both triangles contain the hit point: re-test the other one on the host
            assert tri[i] >= 0 and G[name + "_rays_tri"][i] >= 0
            hit = o[i].astype(np.float64) + d[i].astype(np.float64) * float(t[i])
            for f in (tri[i], G[name + "_rays_tri"][i]):
                v = pos[f].astype(np.float64)
                nrm = np.cross(v[1] - v[0], v[2] - v[0])
                assert abs(np.dot(hit - v[0], nrm)) <= 1e-3 * np.linalg.norm(nrm) * (1.0 + np.abs(hit).max())
    if name != "cornell":
        assert shape[1] > 3                                    # deeper than the small stack: the spill array was really in play
    aim = _f32(G[name + "_rays_aim"])
    occ = np.zeros(len(o), np.uint8)
    assert L.doh_trace_occluded_secondary(C.byref(m.desc), len(o), _p(o), _p(d), _p(aim), _p(occ), smem_levels) == 0
    assert np.array_equal(occ, G[name + "_rays_occ"])


@pytest.mark.parametrize("name", ["cornell", "hf", "tex"])
def test_gbuffer_kernel(doh, name):
    """k_gbuffer over the golden primary hits: getHitInfo with ray differentials, smooth and mapped normals, trilinear
    material fetches with mip selection, sky emission on a miss, the +0.04 red nudge of near-grey base colours - every field
    of every pixel equals the compiled reference's HitInfo to the bit (on the device, texture-driven fields sit within 2e-5:
    CUDA's powf; here libm is the reference's)"""
    from raym0nade_b200.ctypes_defs import RmRenderArgs
    L = doh
    vp = C.c_void_p
    L.doh_gbuffer.argtypes = [C.POINTER(RmSceneDesc), C.POINTER(RmRenderArgs), vp, vp, vp, vp, vp]
    scene, args = _scene(name)
    m = Model(scene)
    a = args.to_c()
    n = args.width * args.height
    g, sav, n_ind = np.zeros(n, HITINFO_DTYPE), np.zeros((n, 3), np.float32), np.full(n, -1, np.int32)
    tri, t = np.ascontiguousarray(G[name + "_tri"], np.int32), np.ascontiguousarray(G[name + "_t"], np.float32)
    L.doh_gbuffer(C.byref(m.desc), C.byref(a), _p(tri), _p(t), _p(g), _p(sav), _p(n_ind))
    want = G[name + "_gbuffer"]
    for k in ["shapeNormal", "surfaceNormal", "emission", "baseColor", "position", "specular", "roughness", "metallic", "opacity", "eta"]:
        assert same(g[k], want[k]), (name, k)
    assert np.array_equal(g["id"], want["id"]) and np.array_equal(g["entering"], want["entering"])
    hit = tri >= 0
    assert hit.any() and np.isnan(g["position"][~hit]).all()
    # the un-nudged base colour kept for the resolve differs from the stored one by exactly +0.04 in red, nowhere else
    d = g["baseColor"].astype(np.float64) - sav.astype(np.float64)
    assert np.abs(d[:, 1:]).max() == 0.0 and ((np.abs(d[:, 0] - 4e-2) < 1e-6) | (d[:, 0] == 0.0)).all()
    if name == "cornell":
        assert (d[:, 0] != 0).any()                            # grey walls: the nudge is in play
    assert (n_ind == 0).all()                                  # spp 0: nothing to sample


def test_radiance_split(doh):
    """accum_split, the device's accumulateInwardRadiance: the projection of a sample's bsdf onto {white, base colour} that
    separates demodulated diffuse from specular, with its three special cases (black light, black base, near-white base)"""
    doh.doh_accumulate.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    base, s7 = _f32(G["acc_base"]), _f32(G["acc_s7"])
    out = np.zeros((len(s7), 8), np.float32)
    doh.doh_accumulate(len(s7), _p(base), _p(s7), _p(out))
    assert same(out, G["acc_out"])


# --------------------------------------------------------------------------- next-event estimation, sample by sample
@pytest.mark.parametrize("secondary_tree", [0, 1])
@pytest.mark.parametrize("which", ["cornell", "sky", "glossy"])
def test_direct_light_samples_replay_the_reference(doh, ref, which, secondary_tree):
    """One direct-light sample per pixel drawn by the device code (light-object pick by lum(bsdf)*power/d^2, face pick from the
    area*luminance CDF, uniform point + cosine rejection - or an environment texel through the guided CDF search -, BSDF
    evaluation and clamp, then rayHit_test on the shadow ray through the emulated engine) against the reference's own
    sampleDirectLight fed the same 32-bit draws through its pre-loaded mt19937 (oracle/ref_harness.cpp): every pixel agrees on
    whether a sample arrives, and every arriving sample equals the reference's LightSample to the bit."""
    from raym0nade_b200 import rng
    from raym0nade_b200.ctypes_defs import RmRenderArgs
    doh.doh_replay_direct.argtypes = [C.POINTER(RmSceneDesc), C.POINTER(RmRenderArgs), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32]
    if which == "cornell":
        scene, args = scenes.cornell_box(40, 40, 1)
    elif which == "sky":
        scene, args = scenes.heightfield_scene(8_000, 48, 27, with_sky=True)
    else:
        scene, args = scenes.glossy_dielectric(30_000, 48, 27)
    R = ref.RefScene(scene)
    m = Model(scene)
    g = np.ascontiguousarray(R.gbuffer(args, threads=4))
    n = args.width * args.height
    out7, status = np.zeros((n, 7), np.float32), np.zeros(n, np.uint8)
    a = args.to_c()
    seed = 77
    assert doh.doh_replay_direct(C.byref(m.desc), C.byref(a), _p(g), seed, _p(out7), _p(status), secondary_tree) == 0
    arrived = 0
    for p in range(n):
        if not np.isfinite(g[p]["position"]).any() or np.linalg.norm(g[p]["emission"]) > 0:
            assert status[p] == 0                              # a miss or an emissive surface draws nothing (src/render.cpp:425-437)
            continue
        s, used = R.replay_direct(args, g[p], rng.stream_u32(seed, p, 0, rng.STREAM_DIRECT, 624))
        assert used < 624
        assert (s is not None) == (status[p] == 2), p
        if s is not None:
            assert np.array_equal(s.view(np.uint32), out7[p].view(np.uint32)), (p, s, out7[p])
            arrived += 1
    assert arrived > n // 4 and (status == 1).sum() > 20       # lit and shadowed pixels both occur
    R.close()


# --------------------------------------------------------------------------- the direct-light kernels, launched as on the device
def _calc_var(rad, var, exposure):
    """calcVar of renderPixel (src/render.cpp:510-516), fp32"""
    f32 = np.float32
    rad = (rad * f32(exposure)).astype(f32)
    var = (var * f32(exposure * exposure)).astype(f32)
    var = var - ((rad[0] * rad[0] + rad[1] * rad[1]) + rad[2] * rad[2])
    return rad, max(f32(var), f32(0))


@pytest.mark.parametrize("which,spp,warp_per_pixel", [("cornell", 1, 0), ("glossy", 1, 0), ("sky", 3, 1), ("sky", 3, 0), ("cornell", 3, 0)])
def test_direct_light_kernels_resolve_to_the_reference_planes(doh, ref, which, spp, warp_per_pixel):
    """k_direct_gen (both mappings: a thread or a warp per pixel) -> ShadowJob through the engine -> k_accum_direct ->
    k_finalise, launched block by block as real threads with a CTA barrier and a warp context per 32 lanes: the resolved
    direct-diffuse and direct-specular planes equal what the reference's own sampleDirectLight / accumulateInwardRadiance /
    calcVar give for the same draws - to the bit where the sample weights are formed identically (one sample per pixel, or an
    environment-lit scene), to 2e-6 where the reference harness's 1/(fails+1) * 1/spp is the kernel's 1/(spp*(fails+1))"""
    from raym0nade_b200 import rng
    from raym0nade_b200.ctypes_defs import RADIANCE_DTYPE, RmRenderArgs
    f32 = np.float32
    doh.doh_direct_planes.argtypes = [C.POINTER(RmSceneDesc), C.POINTER(RmRenderArgs), C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_void_p]
    if which == "cornell":
        scene, args = scenes.cornell_box(32, 32, 1)
    elif which == "sky":
        scene, args = scenes.heightfield_scene(8_000, 40, 24, with_sky=True)
    else:
        scene, args = scenes.glossy_dielectric(30_000, 40, 24)
    R = ref.RefScene(scene)
    m = Model(scene)
    g = np.ascontiguousarray(R.gbuffer(args, threads=4))
    n = args.width * args.height
    Dd, Ds = np.zeros(n, RADIANCE_DTYPE), np.zeros(n, RADIANCE_DTYPE)
    a = args.to_c()
    seed = 91
    items = doh.doh_direct_planes(C.byref(m.desc), C.byref(a), _p(g), seed, spp, warp_per_pixel, 0, _p(Dd), _p(Ds))
    assert items > 0 and items % spp == 0                      # one contiguous block of spp items per sampled pixel
    exact = spp == 1 or which == "sky"
    lit = 0
    for p in range(n):
        acc = np.zeros(8, f32)
        if np.isfinite(g[p]["position"]).any() and not np.linalg.norm(g[p]["emission"]) > 0:
            for si in range(spp):
                s, _ = R.replay_direct(args, g[p], rng.stream_u32(seed, p, si, rng.STREAM_DIRECT, 624))
                if s is not None:
                    s = s.copy()
                    s[6] = f32(s[6]) * (f32(1.0) / f32(spp))           # mulWeight(samples, 1 / spp_direct), src/render.cpp:505
                    acc = acc + ref.accumulate(g[p]["baseColor"][None], s[None])[0]
        lit += bool(acc[:3].any())
        for j, plane in enumerate((Dd, Ds)):
            rad, var = _calc_var(acc[4 * j:4 * j + 3], acc[4 * j + 3], args.exposure)
            if exact:
                assert np.array_equal(rad.view(np.uint32), plane["radiance"][p].view(np.uint32)), (p, j, rad, plane["radiance"][p])
                assert f32(var).view(np.uint32) == plane["Var"][p].view(np.uint32), (p, j)
            else:
                assert np.allclose(plane["radiance"][p], rad, rtol=2e-6, atol=1e-9), (p, j, rad, plane["radiance"][p])
                assert abs(float(plane["Var"][p]) - float(var)) <= 1e-5 * (1.0 + abs(float(var))), (p, j)
    assert lit > n // 5
    R.close()


# --------------------------------------------------------------------------- the whole indirect estimator, kernels and round loop
@pytest.mark.parametrize("secondary_tree", [0, 1])
@pytest.mark.parametrize("which", ["cornell", "sky", "glossy"])
def test_indirect_wavefront_replays_the_reference(doh, ref, which, secondary_tree):
    """The wavefront estimator itself on the CPU: primary hits -> k_gbuffer -> rounds of [k_plan, k_regen, PathJob through the
    engine, k_surface, k_bounce, k_nee, k_shadow_gate, ShadowJob through the engine, k_accum_shadow] with the round loop of
    rm_render_samples restated around the kernels -> k_publish_max, k_commit_hold, k_finalise; every block a set of real threads
    with a CTA barrier and a warp context per 32 lanes.  One indirect sample per opaque pixel, sixteen on glass.  Each pixel is
    then replayed through the reference's own sampleIndirectLightFromFirstIntersection fed the pixel's Philox draws
    (recursion, Russian roulette into NEE, nested dielectrics, absorption, rejection sampling - src/render.cpp:121-423): every
    pixel's indirect-diffuse and indirect-specular radiance agrees to 2e-3 of the pixel's scale (the bar of the GPU replay test,
    which asks it of 97 % of the pixels) and nearly all to the bit - the rest differ in the order several samples of one pixel
    were added (atomics)."""
    from raym0nade_b200 import rng
    from raym0nade_b200.ctypes_defs import RADIANCE_DTYPE, RmRenderArgs
    f32 = np.float32
    doh.doh_indirect_planes.argtypes = [C.POINTER(RmSceneDesc), C.POINTER(RmRenderArgs), C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    if which == "cornell":
        scene, args = scenes.cornell_box(24, 24, 1)
    elif which == "sky":
        scene, args = scenes.heightfield_scene(8_000, 32, 18, with_sky=True)
    else:
        scene, args = scenes.glossy_dielectric(30_000, 32, 18)
    a1 = args.replace(spp=1, P_Direct=0.0)
    R = ref.RefScene(scene)
    m = Model(scene)
    n = a1.width * a1.height
    Id, Is, g_out, rounds = np.zeros(n, RADIANCE_DTYPE), np.zeros(n, RADIANCE_DTYPE), np.zeros(n, HITINFO_DTYPE), np.zeros(1, np.int32)
    a = a1.to_c()
    seed = 1234
    n_glass = doh.doh_indirect_planes(C.byref(m.desc), C.byref(a), seed, 1, secondary_tree, _p(g_out), _p(Id), _p(Is), _p(rounds))
    assert n_glass >= 0 and 2 <= rounds[0] <= 40
    rg = R.gbuffer(a1.replace(spp=0), threads=4)
    if which == "glossy":
        assert n_glass > 0 and n_glass == int((rg["opacity"][np.isfinite(rg["position"][:, 0])] <= 1 - 1e-4).sum())
    pixels = exact = 0
    for p in range(n):
        g = rg[p]
        if np.isnan(g["position"][0]):
            assert not Id["radiance"][p].any() and not Is["radiance"][p].any()
            continue
        n_s = 1 if g["opacity"] > 1 - 1e-4 else 16             # src/render.cpp:498-501
        acc = np.zeros(8, f32)
        base = np.zeros(3, f32) if g["opacity"] < 1e-4 else g["baseColor"]
        for si in range(n_s):
            samples, used = R.replay_indirect(a1, p % a1.width, p // a1.width, g, rng.stream_u32(seed, p, si, rng.STREAM_INDIRECT, 624))
            assert used < 624
            for s in samples:
                s = s.copy()
                s[6] = f32(s[6]) * (f32(1.0) / f32(n_s))       # mulWeight(samples, 1 / spp_indirect)
                acc = acc + ref.accumulate(base[None], s[None])[0]
        same_bits = True
        for j, plane in enumerate((Id, Is)):
            rad = (acc[4 * j:4 * j + 3] * f32(a1.exposure)).astype(f32)
            err = np.abs(plane["radiance"][p] - rad).max() / (np.abs(rad).max() + 1e-6)
            assert err < 2e-3, (p, j, rad, plane["radiance"][p])
            same_bits = same_bits and np.array_equal(rad.view(np.uint32), plane["radiance"][p].view(np.uint32))
        pixels += 1
        exact += same_bits
    assert pixels > n // 2 and exact >= 0.9 * pixels, (pixels, exact)
    assert Id["radiance"].sum() > 0
    R.close()


@pytest.mark.parametrize("name", ["box", "hf"])
def test_depth_of_field_kernels(doh, name):
    """k_dof_prepare -> the reference's stable sort by camera distance -> k_dof_tile_lists -> k_dof_gather, as dof_device chains
    them: every destination replays, in the reference's visiting order, the sources whose disc reaches it.  + - * / sqrt only:
    bit-equal to Photo::depthFeildBlur wherever every pixel has a finite depth (the closed box); a pixel without a hit carries
    a NaN depth, for which the reference's own loop bounds are int(NaN) - there the frames agree on the pixels no such source
    can reach (the same bar as tests/test_gpu_post.py)."""
    doh.doh_depth_field_blur.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p]
    w, h = (int(v) for v in GP[name + "_wh"])
    g = np.ascontiguousarray(GP[name + "_gbuffer"])
    src, cam = _f32(GP[name + "_dof_in"]), _f32(GP[name + "_dof_cam"])
    finite = np.isfinite(g["position"][:, 0]).all()
    for j in (0, 1):
        focus, coc = (float(v) for v in GP["%s_dof_%d_params" % (name, j)])
        out = np.zeros_like(src)
        reach = doh.doh_depth_field_blur(_p(g), _p(src), _p(cam), focus, coc, w, h, _p(out))
        want = GP["%s_dof_%d" % (name, j)]
        assert reach >= 1
        if finite:
            assert same(out, want), (name, j)
        else:
            agree = np.isclose(out, want, rtol=0, atol=1e-6).all(-1)
            assert agree.mean() > 0.5, (name, j, agree.mean())
        assert not same(out, src)
