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


# --------------------------------------------------------------------------- texture file decoders (host/image_io.cpp)
def _png_bytes(rows, width, height, depth, ctype, filters=None, palette=None, trns=None, idat_split=1, header_height=None, interlace=0):
    """encode packed scanlines `rows` (h x stride uint8) as a PNG, applying the given filter type per row"""
    def chunk(typ, data):
        return struct.pack(">I", len(data)) + typ + data + struct.pack(">I", zlib.crc32(typ + data) & 0xFFFFFFFF)
    samples = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bpp = max(1, samples * depth // 8)
    rows = np.asarray(rows, np.uint8)
    out = bytearray()
    prev = np.zeros(rows.shape[1], np.int32)
    for y in range(height):
        cur = rows[y].astype(np.int32)
        ft = 0 if filters is None else filters[y % len(filters)]
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
        b = prev
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = b
        elif ft == 3:
            pred = (a + b) // 2
        else:
            p = a + b - c
            pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
        out.append(ft)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur
    z = zlib.compress(bytes(out), 6)
    parts = [z[i * len(z) // idat_split:(i + 1) * len(z) // idat_split] for i in range(idat_split)]
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, header_height or height, depth, ctype, 0, 0, interlace))
    if palette is not None:
        png += chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None:
        png += chunk(b"tRNS", bytes(trns))
    png += chunk(b"tEXt", b"Comment\x00made by the test")          # an ancillary chunk the decoder must skip
    for part in parts:
        png += chunk(b"IDAT", part)
    return png + chunk(b"IEND", b"")


def _decode(host_check, tmp_path, name, data):
    path, out = str(tmp_path / name), str(tmp_path / (name + ".bin"))
    open(path, "wb").write(data)
    r = subprocess.run([host_check, "image", path, out], capture_output=True, text=True)
    if r.returncode != 0:
        return None, r.stderr
    b = open(out, "rb").read()
    w, h = struct.unpack("<2i", b[:8])
    return np.frombuffer(b, np.uint8, w * h * 4, 8).reshape(h, w, 4), ""


def test_png_decoder_colour_types_depths_and_filters(host_check, tmp_path):
    rng = np.random.default_rng(11)
    w, h = 13, 9
    # RGBA 8-bit, every filter type in turn, IDAT split in three
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    rgba[2:5] = rgba[2]                                         # rows that repeat: Up / Paeth do real work
    got, err = _decode(host_check, tmp_path, "rgba.png", _png_bytes(rgba.reshape(h, -1), w, h, 8, 6, filters=[0, 1, 2, 3, 4], idat_split=3))
    assert got is not None, err
    assert np.array_equal(got, rgba)
    # RGB 8-bit with a colour key (tRNS): the keyed colour turns transparent, everything else opaque
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    rgb[4, 5] = [10, 20, 30]
    got, err = _decode(host_check, tmp_path, "rgb.png", _png_bytes(rgb.reshape(h, -1), w, h, 8, 2, filters=[4, 3], trns=[0, 10, 0, 20, 0, 30]))
    assert got is not None, err
    assert np.array_equal(got[..., :3], rgb)
    assert got[4, 5, 3] == 0 and (np.delete(got[..., 3].ravel(), 4 * w + 5) == 255).all()
    # RGB 16-bit: the high byte of every sample survives (png_set_strip_16, src/material.cpp:243-245)
    rgb16 = rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)
    got, err = _decode(host_check, tmp_path, "rgb16.png", _png_bytes(rgb16.astype(">u2").view(np.uint8).reshape(h, -1), w, h, 16, 2, filters=[1, 4]))
    assert got is not None, err
    assert np.array_equal(got[..., :3], (rgb16 >> 8).astype(np.uint8)) and (got[..., 3] == 255).all()
    # grey + alpha 8-bit
    ga = rng.integers(0, 256, (h, w, 2), dtype=np.uint8)
    got, err = _decode(host_check, tmp_path, "ga.png", _png_bytes(ga.reshape(h, -1), w, h, 8, 4, filters=[2, 3]))
    assert got is not None, err
    assert np.array_equal(got[..., 0], ga[..., 0]) and np.array_equal(got[..., 2], ga[..., 0]) and np.array_equal(got[..., 3], ga[..., 1])
    # grey 4-bit and 1-bit: packed MSB first, expanded to the full 0..255 range
    for depth in (4, 1):
        g = rng.integers(0, 1 << depth, (h, w), dtype=np.uint8)
        per = 8 // depth
        padded = np.zeros((h, (w + per - 1) // per * per), np.uint8)
        padded[:, :w] = g
        packed = np.zeros((h, padded.shape[1] // per), np.uint8)
        for k in range(per):
            packed |= padded[:, k::per] << (8 - depth * (k + 1))
        got, err = _decode(host_check, tmp_path, "g%d.png" % depth, _png_bytes(packed, w, h, depth, 0, filters=[0, 2]))
        assert got is not None, err
        assert np.array_equal(got[..., 1], (g.astype(np.int32) * 255 // ((1 << depth) - 1)).astype(np.uint8))
    # palette 8-bit with per-entry alpha
    pal = rng.integers(0, 256, (5, 3), dtype=np.uint8)
    idx = rng.integers(0, 5, (h, w), dtype=np.uint8)
    got, err = _decode(host_check, tmp_path, "pal.png", _png_bytes(idx, w, h, 8, 3, filters=[0, 1], palette=pal, trns=[255, 128, 0]))
    assert got is not None, err
    assert np.array_equal(got[..., :3], pal[idx])
    assert np.array_equal(got[..., 3], np.array([255, 128, 0, 255, 255], np.uint8)[idx])


def test_png_decoder_refuses_damaged_files(host_check, tmp_path):
    img = np.arange(4 * 4 * 3, dtype=np.uint8).reshape(4, 12)
    good = _png_bytes(img, 4, 4, 8, 2)
    flipped = bytearray(good)
    flipped[-20] ^= 0x55                                        # inside the IDAT payload: its CRC no longer matches
    for name, data, needle in [("crc.png", bytes(flipped), "CRC"), ("short.png", good[:40], ""),
                               ("size.png", _png_bytes(img[:3], 4, 3, 8, 2, header_height=4), "zlib stream"),
                               ("lace.png", _png_bytes(img, 4, 4, 8, 2, interlace=1), "interlaced")]:
        got, err = _decode(host_check, tmp_path, name, data)
        assert got is None and needle in err, (name, err)


def _dxt_colours(c0, c1, punch):
    def e(c):
        r, g, b = (c >> 11) & 31, (c >> 5) & 63, c & 31
        return np.array([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], np.int32)
    p0, p1 = e(c0), e(c1)
    if c0 > c1 or not punch:
        return [(*p0, 255), (*p1, 255), (*((2 * p0 + p1) // 3), 255), (*((p0 + 2 * p1) // 3), 255)]
    return [(*p0, 255), (*p1, 255), (*((p0 + p1) // 2), 255), (0, 0, 0, 0)]


def _dds_header(w, h, fourcc=None, bits=0, masks=(0, 0, 0, 0)):
    pf_flags = 0x4 if fourcc else (0x40 | (0x1 if masks[3] else 0))
    pf = struct.pack("<II4sIIIII", 32, pf_flags, fourcc or b"\0\0\0\0", bits, *masks)
    return b"DDS " + struct.pack("<IIIIIII", 124, 0x1007, h, w, 0, 0, 1) + b"\0" * 44 + pf + struct.pack("<IIIII", 0x1000, 0, 0, 0, 0)


def test_dds_decoder_block_compression_and_masks(host_check, tmp_path):
    rng = np.random.default_rng(5)
    w, h = 10, 6                                                # not multiples of 4: edge blocks are cropped
    bw, bh = (w + 3) // 4, (h + 3) // 4
    assert len(_dds_header(w, h, b"DXT1")) == 128

    def expected(blocks, alpha=None):
        img = np.zeros((bh * 4, bw * 4, 4), np.uint8)
        for by in range(bh):
            for bx in range(bw):
                c0, c1, idx, punch = blocks[by][bx]
                pal = _dxt_colours(c0, c1, punch)
                for i in range(16):
                    img[by * 4 + i // 4, bx * 4 + i % 4] = pal[(idx >> (2 * i)) & 3]
                if alpha is not None:
                    img[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4, 3] = alpha[by][bx]
        return img[:h, :w]

    # DXT1, both endpoint orders (four-colour and three-colour + transparent modes)
    blocks = [[(int(rng.integers(0, 65536)), int(rng.integers(0, 65536)), int(rng.integers(0, 2**32)), True) for _ in range(bw)] for _ in range(bh)]
    blocks[0][0] = (0x1234, 0xF00F, blocks[0][0][2], True)      # c0 < c1: index 3 is transparent black
    blocks[0][1] = (0xF00F, 0x1234, blocks[0][1][2], True)
    body = b"".join(struct.pack("<HHI", c0, c1, idx) for row in blocks for (c0, c1, idx, _) in row)
    got, err = _decode(host_check, tmp_path, "a.dds", _dds_header(w, h, b"DXT1") + body)
    assert got is not None, err
    assert np.array_equal(got, expected(blocks))

    # DXT5: interpolated alpha in both modes, colour always four-colour
    alpha_blocks, body = [], b""
    for by in range(bh):
        arow = []
        for bx in range(bw):
            a0, a1 = (200, 40) if (bx + by) % 2 == 0 else (40, 200)
            bits = int(rng.integers(0, 2**48))
            if a0 > a1:
                pal = [a0, a1] + [((7 - k) * a0 + k * a1) // 7 for k in range(1, 7)]
            else:
                pal = [a0, a1] + [((5 - k) * a0 + k * a1) // 5 for k in range(1, 5)] + [0, 255]
            arow.append(np.array([pal[(bits >> (3 * i)) & 7] for i in range(16)], np.uint8).reshape(4, 4))
            c0, c1, idx, _ = blocks[by][bx]
            body += bytes([a0, a1]) + bits.to_bytes(6, "little") + struct.pack("<HHI", c0, c1, idx)
        alpha_blocks.append(arow)
    got, err = _decode(host_check, tmp_path, "b.dds", _dds_header(w, h, b"DXT5") + body)
    assert got is not None, err
    assert np.array_equal(got, expected([[(c0, c1, idx, False) for (c0, c1, idx, _) in row] for row in blocks], alpha_blocks))

    # DXT3: explicit 4-bit alpha
    alpha_blocks, body = [], b""
    for by in range(bh):
        arow = []
        for bx in range(bw):
            nib = rng.integers(0, 16, 16, dtype=np.uint8)
            arow.append((nib * 17).reshape(4, 4))
            packed = bytes(int(nib[2 * k]) | (int(nib[2 * k + 1]) << 4) for k in range(8))
            c0, c1, idx, _ = blocks[by][bx]
            body += packed + struct.pack("<HHI", c0, c1, idx)
        alpha_blocks.append(arow)
    got, err = _decode(host_check, tmp_path, "c.dds", _dds_header(w, h, b"DXT3") + body)
    assert got is not None, err
    assert np.array_equal(got, expected([[(c0, c1, idx, False) for (c0, c1, idx, _) in row] for row in blocks], alpha_blocks))

    # uncompressed: 32-bit BGRA (the usual A8R8G8B8 masks) and 24-bit RGB without alpha
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    got, err = _decode(host_check, tmp_path, "d.dds", _dds_header(w, h, None, 32, (0x00FF0000, 0x0000FF00, 0x000000FF, 0xFF000000)) + rgba[..., [2, 1, 0, 3]].tobytes())
    assert got is not None, err
    assert np.array_equal(got, rgba)
    got, err = _decode(host_check, tmp_path, "e.dds", _dds_header(w, h, None, 24, (0x0000FF, 0x00FF00, 0xFF0000, 0)) + rgba[..., :3].tobytes())
    assert got is not None, err
    assert np.array_equal(got[..., :3], rgba[..., :3]) and (got[..., 3] == 255).all()

    # refused: a format outside the list, truncated data
    got, err = _decode(host_check, tmp_path, "f.dds", _dds_header(w, h, b"DX10") + b"\0" * 64)
    assert got is None and "DX10" in err
    got, err = _decode(host_check, tmp_path, "g.dds", _dds_header(w, h, b"DXT1") + b"\0" * 8)
    assert got is None and "truncated" in err


def test_obj_materials_with_texture_maps(host_check, tmp_path):
    """map_Kd / map_Bump / map_Ks in a .mtl: RGBA8 for colour slots, RGB8 for the normal map, constants replaced by maps"""
    rng = np.random.default_rng(2)
    albedo = rng.integers(0, 256, (4, 8, 4), dtype=np.uint8)
    albedo[..., 3] = 255
    albedo[0, 0, 3] = 0                                         # a cut-out texel -> hasFullyTransparentPart in the prepared scene
    normal = rng.integers(0, 256, (2, 2, 3), dtype=np.uint8)
    open(tmp_path / "albedo.png", "wb").write(_png_bytes(albedo.reshape(4, -1), 8, 4, 8, 6, filters=[4]))
    open(tmp_path / "normal.dds", "wb").write(_dds_header(2, 2, None, 24, (0x0000FF, 0x00FF00, 0xFF0000, 0)) + normal.tobytes())
    (tmp_path / "t.mtl").write_text("newmtl m\nKd 1 0 0\nmap_Kd albedo.png\nmap_Bump -bm 1.0 normal.dds\nmap_Ks missing.png\n")
    (tmp_path / "t.obj").write_text("mtllib t.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nusemtl m\nf 1/1 2/2 3/3\n")
    out = str(tmp_path / "dump.bin")
    r = subprocess.run([host_check, "dump", str(tmp_path) + "/", "t.obj", "null", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "missing.png" in r.stderr                            # said so, carried on without the map
    b = open(out, "rb").read()
    nf, nmesh, nmat, ntex, sw, sh, nnodes, nlights = struct.unpack("<8i", b[:32])
    assert (nf, nmat, ntex) == (1, 1, 3)                        # Kd texel, albedo map, normal map
    off = 32 + nf * 96 + nmesh * 12
    slots = np.frombuffer(b, "<i4", 4, off)
    assert slots[0] == 1 and slots[3] == 2 and slots[1] == -1   # the map replaced the Kd constant; no specular map
    off += nmat * 40 + (nnodes * 32) + nf * 36
    texs = []
    for _ in range(ntex):
        w, h, c = struct.unpack("<3i", b[off:off + 12])
        texs.append(np.frombuffer(b, np.uint8, w * h * c, off + 12).reshape(h, w, c))
        off += 12 + w * h * c
    assert np.array_equal(texs[1], albedo) and np.array_equal(texs[2], normal)


# --------------------------------------------------------------------------- render_multiThread against a mock device
@pytest.fixture(scope="module")
def mock_console(tmp_path_factory):
    """the console linked with tests/tools/mock_device.cpp: the device entry points log their calls instead of rendering"""
    out = str(tmp_path_factory.mktemp("mock") / "raym0nade_mock")
    cmd = [os.environ.get("CXX", "g++")] + build.HOST_FLAGS + [os.path.join(ROOT, "tests", "tools", "mock_device.cpp")] + \
        build.host_sources() + ["-o", out] + build.host_link_flags() + ["-Wl,-rpath," + os.path.dirname(build.OUT)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.mark.parametrize("coc", [0.0, 4.0])
def test_render_multithread_call_order_and_exports(mock_console, tmp_path, coc):
    """the host's render_multiThread (src/render.cpp:593-676 in the reference): one upload, one render, then the exports
    in the reference's order with the clamp and the denoiser between the groups; the depth-of-field group only with a
    blur circle, and only it carries focus / CoC / camera position (src/render.cpp:664-668)"""
    scene, _ = scenes.cornell_box(64, 64, 0)
    scene.save(str(tmp_path / "cornell.rmscene"))
    args = "0 0 -1\n1 0 0\n0 -1 0\n-3.5 0.25 0\n0.01 2.5 %g 8\n24 16\n5 3 0.7\n%s/frame\n" % (coc, tmp_path)
    log = str(tmp_path / "calls.log")
    r = subprocess.run([mock_console], input="create model box\n%s/\ncornell.rmscene\nnull\ncreate args a\n%srender box a\nexit\n" % (tmp_path, args),
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RM_MOCK_LOG=log, RM_SEED="77", RM_DEVICE="0"))
    assert r.returncode == 0, r.stderr
    assert "Rendering completed in" in r.stdout and "Rays traced: 1000" in r.stdout and "Post processing finished." in r.stdout
    calls = open(log).read().split("\n")[:-1]
    kinds = [c.split()[0] for c in calls]
    groups = ["postprocess"] * 4 + ["spatial_clamp", "download_resolved"] + ["postprocess"] * 8 + ["filter", "download_resolved"] + ["postprocess"] * 8
    assert kinds == ["context_create", "stats_reset", "scene_upload", "render"] + groups + (["postprocess"] * 5 if coc > 0 else [])
    assert calls[2] == "scene_upload faces=%d nodes=8 lights=1" % scene.n_faces
    assert calls[3] == "render 24x16 spp=5 seed=77"
    opts = [int(re.search(r"options=(\d+)", c).group(1)) for c in calls if c.startswith("postprocess")]
    full = 63
    want = [1, 1 | 512, 64, 128, 20, 36, 24, 40, full, full | 256, full | 512, full | 768, 20, 36, 24, 40, full, full | 256, full | 512, full | 768]
    if coc > 0:
        want += [1 | 1024, full | 1024, full | 1024 | 256, full | 1024 | 512, full | 1024 | 768]
    assert opts == want
    tags = ["DiffuseColor", "DiffuseColor_FXAA", "shapeNormal", "surfaceNormal", "Direct_Diffuse", "Direct_Specular", "Indirect_Diffuse",
            "Indirect_Specular", "Raw", "Raw_Bloom", "Raw_FXAA", "Raw_Bloom_FXAA", "Direct_Diffuse_Filter", "Direct_Specular_Filter",
            "Indirect_Diffuse_Filter", "Indirect_Specular_Filter", "Filter", "Filter_Bloom", "Filter_FXAA", "Filter_Bloom_FXAA"]
    if coc > 0:
        tags += ["BaseColor_DepthFieldBlur", "Filter_DepthFieldBlur", "Filter_DepthFieldBlur_Bloom", "Filter_DepthFieldBlur_FXAA",
                 "Filter_DepthFieldBlur_Bloom_FXAA"]
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".png")) == sorted("frame(%s).png" % t for t in tags)
    for tag, o in zip(tags, want):                         # each file holds the frame its own postprocess call returned, as byte(pixel * 255)
        img = read_png(str(tmp_path / ("frame(%s).png" % tag)))
        assert img.shape == (16, 24, 3)
        assert (img[..., 0] == int(np.float32(o) / np.float32(2048) * np.float32(255))).all(), tag
        assert np.array_equal(img[0, :, 1], (np.arange(24, dtype=np.float32) / np.float32(24) * np.float32(255)).astype(np.uint8))
        assert (img[..., 2] == 255).all()
    # the lens parameters reach the library only with the depth-of-field exports; exposure always
    post = [c for c in calls if c.startswith("postprocess")]
    assert all("focus=0 CoC=0 pos=0,0,0 exposure=8" in c for c in post[:20])
    if coc > 0:
        assert all("focus=2.5 CoC=4 pos=0.25,0,3.5 exposure=8" in c for c in post[20:])


def test_two_process_launch_shards_samples_and_exports_on_rank_0(mock_console, tmp_path):
    """one process per GPU (RM_RANK / RM_WORLD, or torchrun's RANK / WORLD_SIZE / LOCAL_RANK): the unique id travels through
    a file, every rank renders its interleaved sample shard and joins rm_reduce, rank 0 alone resolves and writes the
    exports.  The device calls are the mock's; their sequence is the one scripts/reduce_check.py runs on real GPUs."""
    scene, _ = scenes.cornell_box(64, 64, 0)
    scene.save(str(tmp_path / "cornell.rmscene"))
    script = ("create model box\n%s/\ncornell.rmscene\nnull\ncreate args a\n0 0 -1\n1 0 0\n0 -1 0\n-3.5 0 0\n0.01 0 0 8\n24 16\n6 1 0.7\n%s/frame\n"
              "render box a\nrender box a\nexit\n") % (tmp_path, tmp_path)               # two frames: the communicator is set up once
    comm = str(tmp_path / "comm.id")
    (tmp_path / "script.txt").write_text(script)
    barrier = tmp_path / "barrier"
    barrier.mkdir()
    procs, logs, outs = [], [], []
    for rank, env in [(1, dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")), (0, dict(RM_RANK="0", RM_WORLD="2", RM_DEVICE="0"))]:
        logs.append(str(tmp_path / ("calls%d.log" % rank)))          # rank 1 starts first and has to wait for the id file
        outs.append(str(tmp_path / ("stdout%d.txt" % rank)))
        procs.append(subprocess.Popen([mock_console], stdin=open(tmp_path / "script.txt"), stdout=open(outs[-1], "w"), stderr=subprocess.STDOUT,
                                      env=dict(os.environ, RM_MOCK_LOG=logs[-1], RM_MOCK_BARRIER=str(barrier), RM_COMM_FILE=comm, RM_SEED="5", **env)))
    for p in procs:
        p.wait(timeout=120)
    outs = [(open(f).read(),) for f in outs]
    assert all(p.returncode == 0 for p in procs), outs
    calls1, calls0 = [open(f).read().split("\n")[:-1] for f in logs]
    frame = ["scene_upload faces=%d nodes=8 lights=1" % scene.n_faces, "trace_primary 24x16 host=0", "gbuffer host=0"]
    shard = lambda r: "render_samples begin=%d stride=2 spp=6 seed=5 reset=1" % r
    idsum = sum((i * 7 + 3) & 255 for i in range(128))
    assert calls1 == ["context_create 1", "comm_init rank=1 world=2 idsum=%d" % idsum] + \
        2 * (["stats_reset"] + frame + [shard(1), "reduce root=0", "synchronize"])
    head = ["context_create 0", "comm_unique_id", "comm_init rank=0 world=2 idsum=%d" % idsum]
    assert calls0[:3] == head
    per_frame = ["stats_reset"] + frame + [shard(0), "reduce root=0", "resolve host=1", "download_resolved gbuffer_only"]
    rest = calls0[3:]
    assert len(rest) % 2 == 0 and rest[:len(rest) // 2] == rest[len(rest) // 2:]          # both frames took the same path
    assert rest[:len(per_frame)] == per_frame
    assert [c.split()[0] for c in rest[len(per_frame):len(rest) // 2]] == \
        ["postprocess"] * 4 + ["spatial_clamp", "download_resolved"] + ["postprocess"] * 8 + ["filter", "download_resolved"] + ["postprocess"] * 8
    assert "Rank 1: sample shard rendered and reduced" in outs[0][0] and "rank 0 of 2" in outs[1][0]
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".png")]) == 20            # written once, by rank 0
    assert not os.path.exists(comm)                                                      # the id file is gone once everybody has joined


def test_fxaa_program_sequence_against_the_mock_device(tmp_path):
    """raym0nade_fxaa (the reference's fxaa.cpp): PNG -> linear light -> rm_fxaa -> gammaCorrection through rm_postprocess ->
    PNG.  The device calls are the mock's: what is checked is the decoding table, the staging of the frame as the only
    radiance plane, the shade options that pass it through, and the file that comes out."""
    tool = str(tmp_path / "fxaa_mock")
    cmd = [os.environ.get("CXX", "g++")] + build.HOST_FLAGS + [os.path.join(ROOT, "tests", "tools", "mock_device.cpp")] + \
        build.host_sources(main="fxaa_tool.cpp") + ["-o", tool] + build.host_link_flags() + ["-Wl,-rpath," + os.path.dirname(build.OUT)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (6, 10, 3), dtype=np.uint8)
    open(tmp_path / "in.png", "wb").write(_png_bytes(img.reshape(6, -1), 10, 6, 8, 2, filters=[1, 4]))
    log = str(tmp_path / "calls.log")
    r = subprocess.run([tool, str(tmp_path / "in.png"), str(tmp_path / "out.png")], capture_output=True, text=True, env=dict(os.environ, RM_MOCK_LOG=log))
    assert r.returncode == 0, r.stderr
    assert "width: 10 height: 6" in r.stdout
    calls = open(log).read().split("\n")[:-1]
    assert [c.split()[0] for c in calls] == ["context_create", "fxaa", "upload_resolved", "postprocess", "context_destroy"]
    lin = lambda k: float(np.float32(np.float32(k) / np.float32(255)) ** np.float32(2.2))
    first = [float(x) for x in re.search(r"first=(\S+),(\S+),(\S+)", calls[1]).groups()]
    assert calls[1].startswith("fxaa 10x6") and np.allclose(first, [lin(k) for k in img[0, 0]], rtol=2e-6)
    staged = [float(x) for x in re.search(r"Dd0=(\S+),(\S+),(\S+)", calls[2]).groups()]
    assert np.allclose(staged, first, rtol=1e-7) and "Ds0=0 base0=0" in calls[2]          # the frame is the only thing in the Photo
    assert "options=20 " in calls[3] and "exposure=1" in calls[3]                        # Direct_Diffuse without BaseColor / Emission
    out = read_png(str(tmp_path / "out.png"))
    assert out.shape == (6, 10, 3) and (out[..., 0] == int(np.float32(20) / np.float32(2048) * np.float32(255))).all() and (out[..., 2] == 255).all()
    # a missing input is reported, nothing is written
    r = subprocess.run([tool, str(tmp_path / "nope.png"), str(tmp_path / "o2.png")], capture_output=True, text=True, env=dict(os.environ, RM_MOCK_LOG=log))
    assert r.returncode == 1 and "Could not open file for reading" in r.stderr and not os.path.exists(tmp_path / "o2.png")


def test_fxaa_program_without_a_gpu_is_a_loud_no(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    build.build_host()
    open(tmp_path / "in.png", "wb").write(_png_bytes(np.zeros((2, 6), np.uint8), 2, 2, 8, 2))
    r = subprocess.run([build.HOST_FXAA_OUT, str(tmp_path / "in.png"), str(tmp_path / "out.png")], capture_output=True, text=True)
    assert r.returncode == 1 and "No CUDA context" in r.stderr and not os.path.exists(tmp_path / "out.png")
