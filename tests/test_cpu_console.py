"""CPU tests of the C++ host side (raym0nade_b200/host): the console follows the reference's prompts and messages
(src/myconsole.cpp), Model reads what the Python side writes and prepares the same scene, the PNG writer produces files
a decoder accepts, and a render without a B200 fails loudly instead of falling back to anything."""
import os
import re
import struct
import subprocess
import zlib

import numpy as np
import pytest

from raym0nade_b200 import build, scenes
from raym0nade_b200.api import Model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ARGS_INTERIOR_TABLE = """0.987117 -0.16 0
0 0 1
-0.16 -0.987117 0
6.9 -0.2 -3.5
0.00048 0.0 0.0 512.0
2048 1152
320 12 0.7
"""                                                 # docs/renderArguments.txt, arg_interior_table (savePath follows)


@pytest.fixture(scope="module")
def console():
    return build.build_host()


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    """the test-only driver (tests/tools/host_check.cpp) linked against the host sources without their main()"""
    out = str(tmp_path_factory.mktemp("host_check") / "host_check")
    cmd = [os.environ.get("CXX", "g++")] + build.HOST_FLAGS + [os.path.join(ROOT, "tests", "tools", "host_check.cpp")] + \
        build.host_sources(with_main=False) + ["-o", out] + build.host_link_flags() + ["-Wl,-rpath," + os.path.dirname(build.OUT)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def run_console(binary, script, timeout=120):
    r = subprocess.run([binary], input=script, capture_output=True, text=True, timeout=timeout)
    return r.returncode, r.stdout, r.stderr


def read_png(path):
    """minimal decoder for what Photo::save writes: 8-bit RGB, one IDAT, filter 0 rows"""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(b):
        n, typ = struct.unpack(">I4s", b[pos:pos + 8])
        data = b[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + data) & 0xFFFFFFFF
        chunks.append((typ, data))
        pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, ctype, comp, filt, lace = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, filt, lace) == (8, 2, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, 1 + 3 * w)
    assert not raw[:, 0].any()
    return raw[:, 1:].reshape(h, w, 3)


def test_console_commands_follow_the_reference(console, tmp_path):
    scene, _ = scenes.cornell_box(64, 64, 0)
    scene.save(str(tmp_path / "cornell.rmscene"))
    script = ("create model box\n%s/\ncornell.rmscene\nnull\n"
              "create model box\n"
              "view model box\n"
              "create args a1\n%soutput/Bistro\n"
              "create args a1\n"
              "view args a1\n"
              "view args nope\nrender nope a1\nrender box nope\nfrobnicate\n"
              "delete args a1\ndelete args a1\ndelete model box\ndelete model box\nexit\n") % (tmp_path, ARGS_INTERIOR_TABLE)
    rc, out, err = run_console(console, script)
    assert rc == 0, err
    for line in ["Model (box) created.", "Model (box) is already exists.", "Faces: %d" % scene.n_faces,
                 "Model Path: %s/cornell.rmscene" % tmp_path, "BVH has built with size 8",
                 "RenderArgs (a1) created.", "Args (a1) is already exists.", "Args (nope) does not exists.",
                 "Model (nope) does not exists.", "Unknown command.", "RenderArgs (a1) deleted.", "Args (a1) does not exists.",
                 "Model (box) deleted.", "Model (box) does not exists.",
                 "width, height: 2048 1152", "spp: 320", "threads: 12", "P_Direct: 0.7", "savePath: output/Bistro", "exposure: 512"]:
        assert line in out, line
    # position = D*direction + R*right + U*up in fp32 (src/myconsole.cpp:51), same as the Python mirror computes it
    a = scenes.RenderArgs.from_console(ARGS_INTERIOR_TABLE + "output/Bistro")
    got = [float(x) for x in re.search(r"position : (\S+) (\S+) (\S+)", out).groups()]
    assert np.allclose(got, a.position, rtol=1e-5, atol=1e-6)


def test_console_survives_bad_input(console, tmp_path):
    script = ("create model m\n%s/\nmissing.rmscene\nnull\nview model m\n"
              "create model f\n%s/\nmodel.fbx\nnull\n"
              "create args broken\n1 2 3\nnot numbers\nview args broken\n") % (tmp_path, tmp_path)          # input ends without `exit`
    rc, out, err = run_console(console, script)
    assert rc == 0
    assert "cannot open" in err and "Faces: 0" in out                      # a failed load leaves an empty model (src/model.cpp:185-188)
    assert "reads .rmscene and .obj" in err
    assert "was not created" in out and "Args (broken) does not exists." in out


def test_damaged_scene_files_are_refused(console, tmp_path):
    scene, _ = scenes.cornell_box(64, 64, 0)
    scene.save(str(tmp_path / "ok.rmscene"))
    good = open(tmp_path / "ok.rmscene", "rb").read()
    open(tmp_path / "short.rmscene", "wb").write(good[: len(good) // 2])
    open(tmp_path / "magic.rmscene", "wb").write(b"NOTSCENE" + good[8:])
    open(tmp_path / "huge.rmscene", "wb").write(good[:8] + struct.pack("<6i", 2**30, 1, 1, 0, 0, 0) + good[32:])
    open(tmp_path / "neg.rmscene", "wb").write(good[:8] + struct.pack("<6i", -5, 1, 1, 0, 0, 0) + good[32:])
    script = "".join("create model %s\n%s/\n%s.rmscene\nnull\nview model %s\n" % (n, tmp_path, n, n) for n in ("short", "magic", "huge", "neg"))
    rc, out, err = run_console(console, script + "exit\n")
    assert rc == 0, err
    assert out.count("Faces: 0") == 4
    assert err.count("Error loading model") == 4


def test_render_without_a_gpu_is_a_loud_no(console, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    scene, _ = scenes.cornell_box(64, 64, 0)
    scene.save(str(tmp_path / "cornell.rmscene"))
    script = "create model box\n%s/\ncornell.rmscene\nnull\ncreate args a\n%s%s/out\nrender box a\nexit\n" % (tmp_path, ARGS_INTERIOR_TABLE, tmp_path)
    rc, out, err = run_console(console, script)
    assert rc == 0
    assert "No CUDA context" in err and "no CPU path" in err
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".png")]       # nothing was rendered, nothing was written


@pytest.mark.parametrize("which", ["cornell", "sky_textures"])
def test_model_reads_rmscene_and_prepares_the_same_scene(host_check, tmp_path, which):
    if which == "cornell":
        scene, _ = scenes.cornell_box(64, 64, 0)
    else:
        scene, _ = scenes.heightfield_scene(3000, 96, 54, 0, with_sky=True)
    scene.save(str(tmp_path / "s.rmscene"))
    out = str(tmp_path / "dump.bin")
    r = subprocess.run([host_check, "dump", str(tmp_path) + "/", "s.rmscene", "embedded", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    b = open(out, "rb").read()
    nf, nmesh, nmat, ntex, sw, sh, nnodes, nlights = struct.unpack("<8i", b[:32])
    pos = 32

    def take(dtype, count):
        nonlocal pos
        a = np.frombuffer(b, dtype, count, pos)
        pos += a.nbytes
        return a

    assert nf == scene.n_faces and nmesh == len(scene.meshes) and nmat == len(scene.materials) and ntex == len(scene.textures)
    assert np.array_equal(take("<f4", nf * 9), np.float32(scene.positions).ravel())
    assert np.array_equal(take("<f4", nf * 6), np.float32(scene.uvs).ravel())
    assert np.array_equal(take("<f4", nf * 9), np.float32(scene.normals).ravel())
    assert np.array_equal(take("<i4", nmesh * 3).reshape(-1, 3), np.array(scene.meshes, np.int32).reshape(-1, 3))
    mats = take("<f4", nmat * 10).reshape(nmat, 10)
    for i, m in enumerate(scene.materials):
        assert list(mats[i, :4].view(np.int32)) == [m.tex_diffuse, m.tex_specular, m.tex_emissive, m.tex_normals]
        assert np.allclose(mats[i, 4:], [m.opacity, m.ior, m.roughness, *m.transmitting_color], rtol=1e-7)
    if scene.sky is None:
        assert sw == 0 and sh == 0
    else:
        assert (sh, sw) == scene.sky.shape[:2]
        assert np.array_equal(take("<f4", sw * sh * 3), np.float32(scene.sky).ravel())
    # the prepared scene (BVH nodes, permuted triangles) is the one the Python host prepares from the same arrays
    model = Model(scene)
    assert nnodes == model.desc.n_nodes and nlights == model.desc.n_lights
    assert take(np.uint8, nnodes * 32).tobytes() == model.nodes().tobytes()
    assert np.array_equal(take("<f4", nf * 9).reshape(nf, 3, 3), np.float32(scene.positions)[model.permutation()])
    for t in scene.textures:
        w, h, c = take("<i4", 3)
        assert (h, w, c) == t.shape
        assert np.array_equal(take(np.uint8, w * h * c), np.asarray(t, np.uint8).ravel())
    assert pos == len(b)


def test_sky_name_null_drops_an_embedded_sky(console, tmp_path):
    scene, _ = scenes.heightfield_scene(3000, 96, 54, 0, with_sky=True)
    scene.save(str(tmp_path / "s.rmscene"))
    rc, out, err = run_console(console, "create model a\n%s/\ns.rmscene\nnull\nexit\n" % tmp_path)
    assert rc == 0 and "Model (a) created." in out


def test_model_reads_obj_mtl_and_sky_files(host_check, tmp_path):
    (tmp_path / "quad.mtl").write_text(
        "newmtl wall\nKd 0.5 0.25 1.0\n"
        "newmtl lamp\nKd 1 1 1\nKe 1 1 1\n"
        "newmtl Beer\nKd 1 1 1\nd 0.3\n")
    (tmp_path / "quad.obj").write_text(
        "mtllib quad.mtl\n"
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\nv 1 1 1\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
        "usemtl wall\nf 1/1/1 2/2/1 3/3/1 4/4/1\n"          # a quad: fanned into two triangles
        "usemtl lamp\nf 5 6 7\n"                             # no vt / vn: zero uv, geometric normal
        "usemtl Beer\nf -3//1 -2//1 -1//1\n"                 # negative (relative) indices
        "usemtl wall\nf 1/1/1 3/3/1 4/4/1\n")                # back to the first material: joins its mesh
    # sky: a 4x2 little-endian PFM (bottom row first) and the same image as a flat Radiance .hdr
    sky = np.arange(24, dtype=np.float32).reshape(2, 4, 3) * 0.25
    with open(tmp_path / "sky.pfm", "wb") as f:
        f.write(b"PF\n4 2\n-1.0\n" + sky[::-1].astype("<f4").tobytes())
    out = str(tmp_path / "dump.bin")
    r = subprocess.run([host_check, "dump", str(tmp_path) + "/", "quad.obj", "sky.pfm", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    b = open(out, "rb").read()
    nf, nmesh, nmat, ntex, sw, sh, nnodes, nlights = struct.unpack("<8i", b[:32])
    assert (nf, nmesh, nmat, sw, sh) == (5, 3, 3, 4, 2)
    assert ntex == 4                                          # three Kd texels + one Ke texel
    assert nlights == 1                                       # the lamp mesh (emissive texture) forms a light object even with a sky
    pos = np.frombuffer(b, "<f4", nf * 9, 32).reshape(nf, 3, 3)
    assert np.array_equal(pos[0], [[0, 0, 0], [1, 0, 0], [1, 1, 0]]) and np.array_equal(pos[1], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    assert np.array_equal(pos[2], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])       # wall's second run, same mesh
    assert np.array_equal(pos[3], [[0, 0, 1], [1, 0, 1], [1, 1, 1]])       # lamp
    assert np.array_equal(pos[4], pos[3])                                   # Beer through negative indices
    off = 32 + nf * 9 * 4
    uvs = np.frombuffer(b, "<f4", nf * 6, off).reshape(nf, 3, 2)
    assert np.array_equal(uvs[0], [[0, 0], [1, 0], [1, 1]]) and not uvs[3].any()
    off += nf * 6 * 4
    nrm = np.frombuffer(b, "<f4", nf * 9, off).reshape(nf, 3, 3)
    assert np.array_equal(nrm[3], [[0, 0, 1]] * 3)
    off += nf * 9 * 4
    meshes = np.frombuffer(b, "<i4", nmesh * 3, off).reshape(nmesh, 3)
    assert meshes.tolist() == [[0, 3, 0], [3, 4, 1], [4, 5, 2]]
    off += nmesh * 12
    mats = np.frombuffer(b, "<f4", nmat * 10, off).reshape(nmat, 10)
    assert np.allclose(mats[0, 4:7], [1.0, 1.0, 0.8])                      # opaque defaults (src/material.cpp:109-111)
    assert np.allclose(mats[2, 4:], [0.0, 1.25, 5e-3, 0.8, 0.7, 0.55])     # "Beer" with d < 0.99 (src/material.cpp:306-322)
    assert mats[1, :4].view(np.int32)[2] >= 0                              # lamp has an emissive texture
    off += nmat * 40
    got_sky = np.frombuffer(b, "<f4", sw * sh * 3, off).reshape(sh, sw, 3)
    assert np.array_equal(got_sky, sky)                                    # rows top to bottom, as hdr_to_array gives them

    # the same sky as a flat RGBE file: exactly representable values only
    rgbe = np.zeros((2, 4, 4), np.uint8)
    img = np.zeros((2, 4, 3), np.float32)
    for y in range(2):
        for x in range(4):
            m = np.array([128 + 8 * x, 64 + y, 255 - x], np.uint8)
            e = 128 + x - y
            rgbe[y, x] = [*m, e]
            img[y, x] = m.astype(np.float32) * np.float32(2.0) ** (e - 136)
    with open(tmp_path / "sky.hdr", "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 4\n" + rgbe.tobytes())
    r = subprocess.run([host_check, "dump", str(tmp_path) + "/", "quad.obj", "sky.hdr", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    b = open(out, "rb").read()
    assert np.array_equal(np.frombuffer(b, "<f4", 24, off).reshape(2, 4, 3), img)


def test_rle_hdr_scanlines(host_check, tmp_path):
    """new-style run-length encoded Radiance scanlines (width >= 8): runs and literal spans per channel"""
    w, h = 16, 2
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[..., 0] = 200                                   # a pure run
    rgbe[..., 1] = np.arange(w) * 3 + 1                  # pure literals
    rgbe[..., 2] = np.where(np.arange(w) < 9, 7, np.arange(w))      # run then literals
    rgbe[..., 3] = 130
    body = b""
    for y in range(h):
        body += bytes([2, 2, w >> 8, w & 255])
        body += bytes([128 + w, 200])
        body += bytes([w]) + rgbe[y, :, 1].tobytes()
        body += bytes([128 + 9, 7, w - 9]) + rgbe[y, 9:, 2].tobytes()
        body += bytes([128 + w, 130])
    with open(tmp_path / "rle.hdr", "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w) + body)
    (tmp_path / "t.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    out = str(tmp_path / "dump.bin")
    r = subprocess.run([host_check, "dump", str(tmp_path) + "/", "t.obj", "rle.hdr", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    b = open(out, "rb").read()
    nf, nmesh, nmat, ntex, sw, sh, _, _ = struct.unpack("<8i", b[:32])
    assert (nf, sw, sh) == (1, w, h)
    off = 32 + nf * 96 + nmesh * 12 + nmat * 40
    want = rgbe[..., :3].astype(np.float32) * np.float32(2.0) ** (130 - 136)
    assert np.array_equal(np.frombuffer(b, "<f4", w * h * 3, off).reshape(h, w, 3), want)


def test_png_writer_round_trips(host_check, tmp_path):
    rng = np.random.default_rng(3)
    for (h, w) in [(1, 1), (7, 13), (64, 96)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        img[: h // 2] = 17                                # compressible part
        raw, png = str(tmp_path / "a.raw"), str(tmp_path / "a.png")
        img.tofile(raw)
        r = subprocess.run([host_check, "png", raw, str(w), str(h), png], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert np.array_equal(read_png(png), img)
