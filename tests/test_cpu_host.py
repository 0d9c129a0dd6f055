"""CPU tests of the host side: the C-ABI library loads and exports what the header declares,
host-side scene preparation reproduces the reference's load-time derivations bit for bit,
argument parsing follows the reference console, errors are loud, and the multi-GPU reduction
protocol is exact (world_size 2 over gloo)."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

from raym0nade_b200 import api, multi_gpu, rng, scenes
from raym0nade_b200.api import Model, RmError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "raym0nade_b200.h")).read()
    declared = set(re.findall(r"\b(rm_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    L = api.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert b"sm_100a" in L.rm_version()


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RmError) as e:
        api.Context(0)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_tree_node_count_is_the_reference_node_array_length():
    """rm_tree_node_count = nodeCount(1, n) + 1 (src/bvh.cpp:44-49): SURVEY.md's figures (8 @ 40 faces, 65 536 @ 260 K, 262 144 @ 1 M,
    1 048 576 @ 5 M), the golden trees' lengths, and the host builder's own array for a spread of sizes"""
    L = api.lib()
    for n, want in [(40, 8), (260_000, 65_536), (1_000_000, 262_144), (5_000_000, 1_048_576), (10, 2), (11, 4), (1, 2), (0, 0), (-3, 0)]:
        assert L.rm_tree_node_count(n) == want, (n, L.rm_tree_node_count(n))
    for name in ("cornell", "hf", "tex"):
        assert L.rm_tree_node_count(G[name + "_perm"].shape[0]) == G[name + "_nodes"].shape[0]
    for n in (12, 21, 22, 23, 100, 1000, 2999):
        scene, _ = scenes.heightfield_scene(n)
        m = Model(scene)
        assert L.rm_tree_node_count(m.n_faces) == m.desc.n_nodes, (n, m.n_faces)


def test_device_tree_entry_points_fail_loudly_without_a_gpu():
    """rm_tree_build / rm_prepare_scene_device / rm_scene_refit need a context, and there is none without a device: no host fallback"""
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = api.lib()
    pos = np.zeros((12, 9), np.float32)
    nodes, perm = np.zeros(4, api.BVHNODE_DTYPE), np.zeros(12, np.int32)
    assert L.rm_tree_build(None, api._p(pos), 12, api._p(nodes), 4, api._p(perm)) != 0
    assert b"null" in L.rm_last_error()
    assert L.rm_scene_refit(None, api._p(pos), 12) != 0
    scene, _ = scenes.cornell_box()
    raw = scene.to_c()
    h = C.c_void_p()
    assert L.rm_prepare_scene_device(None, C.byref(raw), C.byref(h)) != 0 and not h.value
    with pytest.raises(RmError):
        Model(scene, api.Context(0))


@pytest.mark.parametrize("name", ["cornell", "hf", "tex"])
def test_host_prepare_matches_reference_vectors(name):
    scene, _ = {"cornell": lambda: scenes.cornell_box(64, 64, 0), "hf": lambda: scenes.heightfield_scene(3000, 96, 54, 0, with_sky=True),
                "tex": lambda: scenes.texture_heavy(6000, 96, 54, 0, tex_size=32, n_materials=8)}[name]()
    m = Model(scene)
    assert np.array_equal(m.permutation(), G[name + "_perm"])
    assert m.nodes().tobytes() == G[name + "_nodes"].tobytes()
    if name == "hf":
        _, cdf = m.sky()
        assert np.array_equal(cdf[-16:], G["sky_cdf_tail"])


def test_host_prepare_matches_reference_directly(ref):
    for scene in [scenes.cornell_box()[0], scenes.texture_heavy(20_000, tex_size=64, n_materials=8)[0],
                  scenes.heightfield_scene(20_000, with_sky=True)[0], scenes.glossy_dielectric(120_000)[0]]:
        m, R = Model(scene), ref.RefScene(scene)
        rn, rp = R.bvh()
        assert np.array_equal(m.permutation(), rp) and m.nodes().tobytes() == rn.tobytes(), scene.name
        for a, b in zip(m.lights(), R.lights()):
            assert np.array_equal(a["center"], b["center"]) and np.array_equal(a["color"], b["color"])
            assert np.float32(a["power"]) == np.float32(b["power"]) and np.array_equal(a["cdf"], b["cdf"])
            assert np.array_equal(a["faces"], b["faces"])
        assert len(m.lights()) == len(R.lights())
        sd, sc = m.sky()
        rd, rc = R.sky()
        assert np.array_equal(sd, rd) and np.array_equal(sc, rc)
        # mip chains of every texture slot
        slots = [ref.SLOT_DIFFUSE, ref.SLOT_SPECULAR, ref.SLOT_EMISSIVE, ref.SLOT_NORMALS]
        for mi in range(m.desc.n_materials):
            md = m.material(mi)
            for k in range(4):
                if md.tex[k] < 0:
                    continue
                mine, depth = m.texture_levels(md.tex[k])
                theirs, rdepth = R.texture_levels(mi, slots[k])
                assert depth == rdepth and len(mine) == len(theirs)
                assert all(np.array_equal(x, y) for x, y in zip(mine, theirs))
            assert bool(md.has_fully_transparent_part) == any(
                (scene.textures[scene.materials[mi].tex_diffuse][..., 3] < 255).any() for _ in [0] if scene.materials[mi].tex_diffuse >= 0)
        R.close()


def test_prepare_rejects_bad_input():
    scene, _ = scenes.cornell_box()
    bad = scenes.RawScene(positions=scene.positions, uvs=scene.uvs, normals=scene.normals,
                          meshes=[(0, scene.n_faces + 5, 0)], materials=scene.materials, textures=scene.textures)
    with pytest.raises(RmError):
        Model(bad)
    rgb_as_diffuse = scenes.RawScene(positions=scene.positions, uvs=scene.uvs, normals=scene.normals, meshes=scene.meshes,
                                     materials=scene.materials, textures=[t[..., :3].copy() for t in scene.textures])
    with pytest.raises(RmError):
        Model(rgb_as_diffuse)      # the reference would stride an RGB8 diffuse texture by 4 (src/material.cpp:58)


def test_render_args_follow_the_reference_console():
    txt = """0.987117 -0.16 0
             0 0 1
             -0.16 -0.987117 0
             6.9 -0.2 -3.5
             0.00048 0.0 0.0 512.0
             2048 1152
             320 12 0.7
             output/Bistro"""                      # docs/renderArguments.txt, arg_interior_table
    a = scenes.RenderArgs.from_console(txt)
    assert (a.width, a.height, a.spp, a.threads) == (2048, 1152, 320, 12)
    d, r, u = np.float32([0.987117, -0.16, 0]), np.float32([0, 0, 1]), np.float32([-0.16, -0.987117, 0])
    want = np.float32(6.9) * d + np.float32(-0.2) * r + np.float32(-3.5) * u       # position = D*direction + R*right + U*up
    assert np.allclose(a.position, want, rtol=0, atol=1e-6)
    assert abs(a.P_Direct - 0.7) < 1e-6 and a.savePath == "output/Bistro"
    c = a.to_c()
    assert c.width == 2048 and abs(c.exposure - 512.0) < 1e-6
    with pytest.raises(ValueError):
        scenes.RenderArgs.from_console("1 2 3")


def test_rng_stream_statement_matches_reference_mapping():
    got = rng.uniform_from_u32(G["rng_u32"])
    assert np.array_equal(got.view(np.uint32), G["rng_out"].view(np.uint32))
    # Philox4x32-10 known answers (Random123 kat_vectors): zero key/counter and all-ones
    z = rng.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    o = rng.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in o] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    a = rng.stream_u32(7, 11, 3, rng.STREAM_INDIRECT, 10)
    b = rng.stream_u32(7, 11, 3, rng.STREAM_DIRECT, 10)
    assert len(a) == 10 and not np.array_equal(a, b)


def test_sample_sharding_partitions_every_index_once():
    for total in [0, 1, 7, 44, 1024]:
        for world in [1, 2, 4, 8]:
            counts = [multi_gpu.local_sample_count(total, r, world) for r in range(world)]
            assert sum(counts) == total
            seen = sorted(r + k * world for r in range(world) for k in range(counts[r]))
            assert seen == list(range(total))


# --------------------------------------------------------------------------- world_size 2 over gloo
def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    npix, n_samples = 64, 40
    rs = np.random.default_rng(5)
    clum = rs.random((n_samples, npix)).astype(np.float32)
    clum[3, ::4] = 500.0          # fireflies owned by rank 1 (sample 3): must be dropped in a quarter of the pixels
    clum[8, 1::4] = 3.0           # large but below 16/17 of the total: must survive

    class Acc:                     # numpy stand-in for the CUDA accumulators, same protocol
        def __init__(self):
            mine = clum[rank::world]
            self.sum = np.zeros((npix, 2), np.float32)
            self.sum[:, 0], self.sum[:, 1] = mine.sum(0), mine.shape[0]
            self.hold = mine.max(0)
            self.max = self.hold.copy()
            self.rad = np.zeros((npix, 16), np.float32)
            self.rad[:, 8] = mine.sum(0) - self.hold          # everything but the held-back sample

        def accum_view(self):
            return self.sum, self.max

        def accum_after_reduce(self, rank, world):
            total, gmax = self.sum[:, 0], self.max
            drop = (self.hold == gmax) & (self.hold / (total - self.hold + np.float32(1e-4)) > 16.0)
            self.rad[:, 8] += np.where(drop, 0.0, self.hold).astype(np.float32)

        def accum_radiance(self):
            return self.rad

    acc = Acc()
    multi_gpu.reduce_frame(acc, dist, rank, world, torch.from_numpy)
    # the scattered form on fresh accumulators: every rank ends up with the frame's sums in its own slice of the pixels
    acc2 = Acc()
    first, count = multi_gpu.reduce_frame_scatter(acc2, dist, rank, world, torch.from_numpy, npix)
    mine = [None] * world
    dist.all_gather_object(mine, (first, count, acc2.rad[first:first + count, 8].copy()))
    if rank == 0:
        whole = np.full(npix, np.nan, np.float32)
        for f, c, v in mine:
            assert np.isnan(whole[f:f + c]).all()              # slices do not overlap
            whole[f:f + c] = v
        q.put((acc.rad[:, 8].copy(), whole))
    dist.destroy_process_group()


def test_two_rank_reduction_equals_single_process_clamp():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, got_scattered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process statement of src/render.cpp:534-547 on the same samples
    rs = np.random.default_rng(5)
    clum = rs.random((40, 64)).astype(np.float32)
    clum[3, ::4] = 500.0
    clum[8, 1::4] = 3.0
    total = clum.sum(0)
    keep = ~(clum / (total - clum + np.float32(1e-4)) > 16.0)
    want = (clum * keep).sum(0)
    assert keep[3, ::4].sum() == 0 and keep[8].all()
    assert np.allclose(got, want, rtol=1e-5)
    assert np.allclose(got_scattered, want, rtol=1e-5)          # the slices of the scattered exchange, put together, are the same frame


def test_frame_slices_cover_every_pixel_once():
    """rm_frame_slice's partition: ceil(npix / world) pixels per rank, the tail clipped - also when world does not divide npix"""
    for npix in [1, 7, 64, 1920 * 1080, 3840 * 2160 + 5]:
        for world in [1, 2, 3, 8, 64]:
            cover = np.zeros(npix, np.int8) if npix < 10000 else None
            total, end = 0, 0
            for r in range(world):
                first, count = multi_gpu.frame_slice(npix, r, world)
                assert first == min(end, npix) or count == 0
                end = first + count
                total += count
                if cover is not None:
                    cover[first:first + count] += 1
            assert total == npix and (cover is None or (cover == 1).all())


@pytest.mark.parametrize("which", ["one", "cornell", "heightfield", "glossy"])
def test_secondary_ray_tree_invariants(which):
    """the host builder of the secondary-ray tree (csrc/fast_bvh.cpp): every triangle in exactly one leaf, boxes enclose,
    links in range, leaf size and depth within their bounds (checked inside rm_secondary_tree_stats)"""
    from raym0nade_b200 import api
    if which == "one":
        pos = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    elif which == "cornell":
        pos = scenes.cornell_box(8, 8, 0)[0].positions
    elif which == "heightfield":
        pos = scenes.heightfield_scene(20_000)[0].positions
    else:
        pos = scenes.glossy_dielectric(120_000, 8, 8, 0)[0].positions
    n = np.asarray(pos).reshape(-1, 9).shape[0]
    if which == "heightfield":          # coincident triangles (zero-extent centroid boxes) must still terminate and cover everything
        pos = np.concatenate([np.asarray(pos).reshape(-1, 9), np.repeat(np.asarray(pos).reshape(-1, 9)[:1], 40, axis=0)])
        n = pos.shape[0]
    for leaf_max, cap in [(3, 22), (4, 21), (1, 26), (15, 8)]:
        st = api.secondary_tree_stats(pos, cap, leaf_max)
        assert st["largest_leaf"] <= leaf_max and st["leaves"] >= (n + leaf_max - 1) // leaf_max
        min_depth = int(np.ceil(np.log2(max(1.0, n / leaf_max)))) + 1
        assert st["depth"] <= max(cap, min_depth) + 1, st


@pytest.mark.parametrize("which", ["one", "cornell", "heightfield", "glossy", "flat"])
def test_wide_secondary_tree_invariants(which):
    """csrc/wide_bvh.cpp: the 4-wide collapse with 8-bit quantised child boxes keeps every triangle in exactly one leaf and
    every decoded child box around the vertices beneath it (checked inside rm_wide_tree_stats, in double precision)"""
    from raym0nade_b200 import api
    if which == "one":
        pos = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    elif which == "cornell":
        pos = scenes.cornell_box(8, 8, 0)[0].positions
    elif which == "heightfield":
        pos = scenes.heightfield_scene(20_000)[0].positions
    elif which == "flat":               # axis-aligned, zero-thickness geometry far from the origin: flat quantisation grids
        g = np.stack(np.meshgrid(np.arange(40, dtype=np.float32), np.arange(40, dtype=np.float32)), -1).reshape(-1, 2)
        z = np.full((g.shape[0], 1), 1000.25, np.float32)
        a, b, c = np.concatenate([g, z], 1), np.concatenate([g + [1, 0], z], 1), np.concatenate([g + [0, 1], z], 1)
        pos = np.concatenate([a, b, c], 1).astype(np.float32)
    else:
        pos = scenes.glossy_dielectric(120_000, 8, 8, 0)[0].positions
    n = np.asarray(pos).reshape(-1, 9).shape[0]
    st = api.wide_tree_stats(pos)
    assert st["leaves"] >= (n + 2) // 3 and st["nodes"] >= 1
    if n > 1000:
        assert st["children_per_node"] > 2.5, st            # the collapse fills its nodes (2 = nothing gained over the binary tree)
        assert 3 * st["levels"] <= 14 + 82, st              # deferred children fit the traversal stack (dev_trace.cuh)


# --------------------------------------------------------------------------- static check of the compiled kernels
def test_sass_has_the_fetch_widths_and_warp_primitives_the_design_states():
    """DESIGN.md section 4 in the machine code (cuobjdump -sass, no GPU): every traversal kernel fetches nodes and triangles
    with 128/256-bit global loads and refills idle lanes with vote + shuffle; no kernel uses a tensor-core instruction
    (nothing on the path is a dense contraction); the image-space stencils stay free of local memory."""
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sass_mix
    kernels = sass_mix.mix(api.LIB_PATH)
    names = sass_mix.demangle(list(kernels))
    by_name = {sass_mix.short(names[k]): c for k, c in kernels.items()}
    assert len(by_name) >= 35
    trace = {n: c for n, c in by_name.items() if "k_trace<" in n}
    assert len(trace) == 20                                   # 5 jobs x {counting, plain} x {binary trees, 4-wide tree}
    for n, c in trace.items():
        assert c["LDG.128"] + c["LDG.256"] >= 8, (n, dict(c))     # node pairs and triangle records, 16 B or 32 B per load
        assert c["VOTE"] >= 3 and c["SHFL"] >= 8, (n, dict(c))    # ballot / shfl compaction of the ray queue
        assert c["BAR"] == 0, n                                   # warps run independently: no CTA-wide barrier in traversal
        assert c["inst"] < 2000, (n, c["inst"])                   # the loop stays inside the instruction cache
    assert all(c["TENSOR"] == 0 for c in by_name.values())
    for n in ("k_fxaa", "k_atrous", "k_shade_gamma", "k_bloom_pass", "k_dof_gather", "k_finalise"):
        assert by_name[n]["LDL"] == 0 and by_name[n]["STL"] == 0, n


# --------------------------------------------------------------------------- rm_scene_validate (host only)
def _scene_for(which):
    if which == "cornell":
        return scenes.cornell_box(64, 64, 0)[0]
    if which == "sky":
        return scenes.heightfield_scene(3000, 96, 54, 0, with_sky=True)[0]
    if which == "textured":
        return scenes.texture_heavy(4000, 96, 54, 4, tex_size=64, n_materials=4)[0]
    return scenes.glossy_dielectric(20000, 96, 54, 0)[0]


@pytest.mark.parametrize("which", ["cornell", "sky", "textured", "glossy"])
def test_prepared_scenes_pass_validation(which):
    Model(_scene_for(which)).validate()


def test_validation_names_the_inconsistency():
    """rm_scene_upload refuses, on the host and with a message, every scene whose indices would send a kernel out of bounds"""
    import ctypes as C
    model = Model(_scene_for("textured"))
    L = api.lib()

    def refused(mutate, needle):
        d = api.RmSceneDesc.from_buffer_copy(model.desc)           # shallow copy: same arrays, private header
        keep = mutate(d)                                            # noqa: F841  (keeps replacement arrays alive)
        rc = L.rm_scene_validate(C.byref(d))
        msg = L.rm_last_error().decode()
        assert rc == -1 and needle in msg, (rc, msg)

    def with_nodes(edit):
        def f(d):
            nodes = model.nodes().copy()
            edit(nodes, d)
            d.nodes = nodes.ctypes.data
            return nodes
        return f

    first_leaf = int(np.nonzero(model.nodes()["faceR"] != 0)[0][0])
    refused(with_nodes(lambda n, d: n["faceR"].__setitem__(first_leaf, d.n_faces + 7)), "leaf range")
    refused(with_nodes(lambda n, d: n["faceL"].__setitem__(first_leaf, -3)), "leaf range")
    refused(with_nodes(lambda n, d: n["faceR"].__setitem__(first_leaf, 0)), "children lie beyond")     # a leaf turned inner at the bottom level
    refused(with_nodes(lambda n, d: n["faceL"].__setitem__(first_leaf, n["faceR"][first_leaf] - 1)), "leaves own")

    # a leaf reference holds its face count in 4 bits: a root leaf of 16 faces must not reach the device
    def long_leaf(n, d):
        n["faceL"][1], n["faceR"][1] = 0, d.n_faces
    refused(with_nodes(long_leaf), "at most 15")

    def bad_face_material(d):
        fm = np.ctypeslib.as_array(C.cast(d.face_material, C.POINTER(C.c_int32)), (d.n_faces,)).copy()
        fm[5] = d.n_materials
        d.face_material = fm.ctypes.data
        return fm
    refused(bad_face_material, "face 5: material index")

    def bad_material_texture(d):
        mats = (api.RmMaterialDesc * d.n_materials)(*[d.materials[i] for i in range(d.n_materials)])
        mats[0].tex[0] = d.n_textures
        d.materials = C.cast(mats, C.POINTER(api.RmMaterialDesc))
        return mats
    refused(bad_material_texture, "texture index out of range")

    def swapped_slots(d):                                           # an RGB8 normal map in the diffuse slot
        mats = (api.RmMaterialDesc * d.n_materials)(*[d.materials[i] for i in range(d.n_materials)])
        i = next(i for i in range(d.n_materials) if mats[i].tex[3] >= 0)
        mats[i].tex[0] = mats[i].tex[3]
        d.materials = C.cast(mats, C.POINTER(api.RmMaterialDesc))
        return mats
    refused(swapped_slots, "channels in slot 0")

    def null_level(d):
        texs = (api.RmTextureDesc * d.n_textures)(*[d.textures[i] for i in range(d.n_textures)])
        texs[0].levels[1] = None
        d.textures = C.cast(texs, C.POINTER(api.RmTextureDesc))
        return texs
    refused(null_level, "level 1 is NULL")

    def empty_level(d):                                             # more levels than the image has (a level of zero width)
        texs = (api.RmTextureDesc * d.n_textures)(*[d.textures[i] for i in range(d.n_textures)])
        texs[0].width, texs[0].height, texs[0].map_depth = 4, 64, 4
        d.textures = C.cast(texs, C.POINTER(api.RmTextureDesc))
        return texs
    refused(empty_level, "is empty")

    refused(lambda d: setattr(d, "n_materials", 0), "at least one material")
    refused(lambda d: setattr(d, "sky_width", 8), "bad sky size")
    refused(lambda d: setattr(d, "positions", None), "missing geometry")

    lit = Model(_scene_for("cornell"))

    def empty_light(d):
        lights = (api.RmLightDesc * d.n_lights)(*[d.lights[i] for i in range(d.n_lights)])
        lights[0].n_faces = 0
        d.lights = C.cast(lights, C.POINTER(api.RmLightDesc))
        return lights
    d = api.RmSceneDesc.from_buffer_copy(lit.desc)
    keep = empty_light(d)                                           # noqa: F841
    assert L.rm_scene_validate(C.byref(d)) == -1 and "light 0: no faces" in L.rm_last_error().decode()
    assert L.rm_scene_validate(None) == -1


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """the drop-in boundary is a C ABI: include/raym0nade_b200.h compiles as pedantic C99 and a C program links against the
    library and calls a host-only entry point (what a cgo / JNI / ctypes binding relies on)"""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "raym0nade_b200.h"\n'
                   "int main(void) {\n"
                   "    RmSceneDesc d; memset(&d, 0, sizeof d);\n"
                   "    int rc = rm_scene_validate(&d);\n"
                   '    printf("%d|%s|%s|%d %d %d %d\\n", rc, rm_last_error(), rm_version(), (int)sizeof(RmHitInfo), (int)sizeof(RmRadiance), (int)sizeof(RmBvhNode), (int)sizeof(RmRenderArgs));\n'
                   "    return 0;\n}\n")
    exe = str(tmp_path / "abi")
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                        "-L", os.path.dirname(api.LIB_PATH), "-lraym0nade_b200", "-Wl,-rpath," + os.path.dirname(api.LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True).stdout.strip().split("|")
    assert out[0] == "-1" and "missing geometry" in out[1] and "sm_100a" in out[2]
    assert out[3] == "88 16 32 80"                          # HitInfo, RadianceData, BVH_Node of the reference; RenderArgs minus threads / savePath


# --------------------------------------------------------------------------- the binding a maintainer adds to the reference tree
@pytest.mark.parametrize("which", ["cornell", "sky", "textured", "glossy"])
def test_reference_tree_binding_fills_the_same_scene(which):
    """raym0nade_b200/host/reference_tree/gpu_bridge.cpp, compiled against the unmodified reference: the RmSceneDesc it fills
    from a Model the reference built equals the one rm_prepare_scene derives from the same raw scene - nodes, triangle order,
    materials, mip chains, light objects, sky - bit for bit, and passes rm_scene_validate"""
    import ctypes as C
    from oracle import refbind
    if not refbind.available("bridge"):
        pytest.skip("oracle/_ref/libraym_bridge.so not built (make -C oracle bridge, needs /root/reference)")
    scene = _scene_for(which)
    R = refbind.RefScene(scene, flavour="bridge")
    b = R.L.ref_bridge_create(R.h)
    d = api.RmSceneDesc.from_address(R.L.ref_bridge_desc(b))
    m = Model(scene)
    p = m.desc
    assert api.lib().rm_scene_validate(C.byref(d)) == 0, api.lib().rm_last_error()

    def arr(addr, dtype, count):
        return np.frombuffer((C.c_char * (np.dtype(dtype).itemsize * count)).from_address(addr), dtype) if count else np.zeros(0, dtype)

    def pv(ptr):
        return ptr if isinstance(ptr, int) or ptr is None else C.cast(ptr, C.c_void_p).value

    for f in ("n_faces", "n_nodes", "n_materials", "n_textures", "n_lights", "sky_width", "sky_height"):
        assert getattr(d, f) == getattr(p, f), f
    nf = d.n_faces
    for f, per in (("positions", 9), ("uvs", 6), ("normals", 9)):
        assert np.array_equal(arr(pv(getattr(d, f)), np.uint32, nf * per), arr(pv(getattr(p, f)), np.uint32, nf * per)), f
    assert np.array_equal(arr(pv(d.face_material), np.int32, nf), arr(pv(p.face_material), np.int32, nf))
    # every node the traversal can reach (slots the build never wrote are garbage in the reference, zero in ours)
    dn, pn = arr(pv(d.nodes), np.uint8, d.n_nodes * 32).reshape(-1, 32), arr(pv(p.nodes), np.uint8, p.n_nodes * 32).reshape(-1, 32)
    face_r = pn.view(np.int32).reshape(-1, 8)[:, 7]
    stack, seen = [1], 0
    while stack:
        u = stack.pop()
        assert np.array_equal(dn[u], pn[u]), u
        seen += 1
        if face_r[u] == 0:
            stack += [2 * u, 2 * u + 1]
    assert seen > 0
    for i in range(d.n_materials):
        a, c = d.materials[i], p.materials[i]
        assert list(a.tex) == list(c.tex) and a.has_fully_transparent_part == c.has_fully_transparent_part
        assert (a.opacity, a.ior, a.roughness, list(a.transmitting_color)) == (c.opacity, c.ior, c.roughness, list(c.transmitting_color))
    for i in range(d.n_textures):
        a, c = d.textures[i], p.textures[i]
        assert (a.width, a.height, a.channels, a.map_depth) == (c.width, c.height, c.channels, c.map_depth), i
        for l in range(a.map_depth):
            n = (a.width >> l) * (a.height >> l) * a.channels
            assert np.array_equal(arr(a.levels[l], np.uint8, n), arr(c.levels[l], np.uint8, n)), (i, l)
    for i in range(d.n_lights):
        a, c = d.lights[i], p.lights[i]
        assert (list(a.center), list(a.color), a.power, a.n_faces) == (list(c.center), list(c.color), c.power, c.n_faces)
        for f, per in (("face_positions", 9), ("face_normals", 9), ("face_cdf", 1)):
            assert np.array_equal(arr(getattr(a, f), np.uint32, a.n_faces * per), arr(getattr(c, f), np.uint32, a.n_faces * per)), (i, f)
    ns = d.sky_width * d.sky_height
    assert np.array_equal(arr(pv(d.sky_data), np.uint32, ns * 3), arr(pv(p.sky_data), np.uint32, ns * 3))
    assert np.array_equal(arr(pv(d.sky_cdf), np.uint32, ns), arr(pv(p.sky_cdf), np.uint32, ns))
    R.L.ref_bridge_destroy(b)
    R.close()


def test_reference_tree_binding_without_a_gpu_leaves_the_photo_alone():
    """render_multiThread_b200 on a box without a B200: the library's error is printed, the reference's Photo is not touched,
    and nothing falls back to the CPU renderer"""
    import ctypes as C
    import torch
    from oracle import refbind
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not refbind.available("bridge"):
        pytest.skip("oracle/_ref/libraym_bridge.so not built (make -C oracle bridge, needs /root/reference)")
    scene, args = scenes.cornell_box(16, 16, 2)
    R = refbind.RefScene(scene, flavour="bridge")
    rgb = np.zeros((16, 16, 3), np.float32)
    a = args.to_c()
    assert R.L.ref_bridge_render(R.h, C.byref(a), 63, rgb.ctypes.data_as(C.c_void_p), None) == -1
    assert not rgb.any()
    R.close()


@pytest.fixture(scope="module")
def sweep_mirror(tmp_path_factory):
    """tests/tools/sah_sweep_host.cpp: the level loop of the device's sweep-SAH builder on the CPU, around csrc/sah_sweep.h"""
    import subprocess
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "vector_types.h")):
        pytest.skip("no CUDA headers")
    so = str(tmp_path_factory.mktemp("sweep") / "sah_sweep_host.so")
    r = subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-shared", "-fPIC", "-I", cuda_inc, "-I", os.path.join(ROOT, "raym0nade_b200", "csrc"),
                        os.path.join(ROOT, "tests", "tools", "sah_sweep_host.cpp"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(so)
    L.sah_sweep_host.restype = C.c_int
    L.sah_sweep_host.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    return L


def _half_area(lo, hi):
    d = hi - lo
    return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]


@pytest.mark.parametrize("which", ["one", "two", "three", "cornell", "heightfield", "duplicates", "glossy"])
def test_sweep_sah_builder_steps(sweep_mirror, which):
    """csrc/sah_sweep.h - the per-element code of the device's sweep-SAH builder (gpu_sah_bvh.cu) - driven level by level on the
    CPU: a binary tree over every triangle exactly once, counts and boxes consistent, depth within the cap's reach, and a SAH
    cost no worse than the object-median tree's over the same lists (the depth-cap fallback alone)"""
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    if which in ("one", "two", "three"):
        k = {"one": 1, "two": 2, "three": 3}[which]
        pos = np.concatenate([tri + 2.0 * i for i in range(k)])
    elif which == "cornell":
        pos = np.asarray(scenes.cornell_box(8, 8, 0)[0].positions, np.float32).reshape(-1, 9)
    elif which == "heightfield":
        pos = np.asarray(scenes.heightfield_scene(20_000)[0].positions, np.float32).reshape(-1, 9)
    elif which == "duplicates":       # coincident triangles: every split costs the same; the tie-break and the depth rule must still end the build
        pos = np.concatenate([np.repeat(tri, 500, axis=0), np.asarray(scenes.cornell_box(8, 8, 0)[0].positions, np.float32).reshape(-1, 9)])
    else:
        pos = np.asarray(scenes.glossy_dielectric(60_000, 8, 8, 0)[0].positions, np.float32).reshape(-1, 9)
    pos = np.ascontiguousarray(pos)
    n = pos.shape[0]

    def build(cap):
        lo, hi = np.zeros((2 * n, 4), np.float32), np.zeros((2 * n, 4), np.float32)
        left, right, count = (np.zeros(2 * n, np.int32) for _ in range(3))
        root = C.c_int(-1)
        levels = sweep_mirror.sah_sweep_host(pos.ctypes.data, n, cap, lo.ctypes.data, hi.ctypes.data, left.ctypes.data, right.ctypes.data, count.ctypes.data, C.byref(root))
        assert levels >= 0, levels
        return lo[:, :3], hi[:, :3], left, right, count, root.value, levels

    def walk(tree):
        lo, hi, left, right, count, root, _ = tree
        seen = np.zeros(n, np.int32)
        cost, depth_wide, stack = 0.0, 0, [(root, 0)]
        root_area = max(float(_half_area(lo[root], hi[root])), 1e-30)
        while stack:
            b, d = stack.pop()
            if right[b] < 0:
                t = ~left[b]
                assert 0 <= t < n and b == t and count[b] == 1
                seen[t] += 1
                v = pos[t].reshape(3, 3)
                assert (v.min(0) >= lo[b]).all() and (v.max(0) <= hi[b]).all()
                continue
            l, r = left[b], right[b]
            assert n <= b < 2 * n - 1 and count[b] == count[l] + count[r]
            assert (np.minimum(lo[l], lo[r]) == lo[b]).all() and (np.maximum(hi[l], hi[r]) == hi[b]).all()
            if count[b] > 3:
                depth_wide = max(depth_wide, d + 1)
                cost += float(_half_area(lo[b], hi[b])) / root_area
            stack += [(l, d + 1), (r, d + 1)]
        assert (seen == 1).all()
        return cost, depth_wide

    sah = build(22)
    assert sah[4][sah[5]] == n
    cost, depth = walk(sah)
    need = int(np.ceil(np.log2(max(1.0, n / 3)))) + 1
    assert depth <= max(22, need) + 1, depth
    if n > 3:
        median_cost, _ = walk(build(0))          # cap 0: every node takes the median fallback
        assert cost <= median_cost * 1.0001, (cost, median_cost)
        if which in ("heightfield", "glossy"):
            assert cost <= 0.9 * median_cost, (cost, median_cost)
