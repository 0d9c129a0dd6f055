"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference -> `make -C oracle ref`):
    python tests/golden/make_golden.py
The reference ships no tests or golden vectors of its own, so these
fixtures - outputs of the reference's own compiled code on seeded inputs - are what pins the
restatement (oracle/port) and the CUDA path on machines where the reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402
from raym0nade_b200 import scenes  # noqa: E402
from raym0nade_b200.ctypes_defs import HITINFO_DTYPE  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def unit(v):
    return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(f32)


def random_surfaces(rs, n):
    s = np.zeros(n, HITINFO_DTYPE)
    nrm = unit(rs.normal(size=(n, 3)))
    s["shapeNormal"] = nrm
    s["surfaceNormal"] = unit(nrm + 0.2 * rs.normal(size=(n, 3)))
    s["baseColor"] = rs.random((n, 3)).astype(f32)
    s["baseColor"][: n // 16] = 0.0
    s["position"] = rs.normal(size=(n, 3)).astype(f32)
    s["specular"] = 0.04
    s["roughness"] = (1e-3 + rs.random(n) ** 2).astype(f32)
    s["metallic"] = np.where(rs.random(n) < 0.3, 0.99, 0.0).astype(f32) * rs.random(n).astype(f32)
    glass = rs.random(n) < 0.4
    s["opacity"] = np.where(glass, 0.0, 1.0).astype(f32)
    s["eta"] = np.where(glass, np.where(rs.random(n) < 0.5, 1.25, 0.8), 1.0).astype(f32)
    s["entering"] = np.where(glass, rs.random(n) < 0.5, True)
    return s


def main():
    rs = np.random.default_rng(20261017)
    g = {}
    # ---- geometry kernels
    n = 4096
    rays = np.concatenate([rs.normal(size=(n, 3)) * 2, unit(rs.normal(size=(n, 3)))], 1).astype(f32)
    rays[:256, 3] = 0.0                      # exactly parallel
    rays[256:512, 4] = 5e-5                  # inside the |d| < 1e-4 band
    rays[512:768, 5] = -9.9e-5
    lo = rs.normal(size=(n, 3)).astype(f32)
    boxes = np.concatenate([lo, lo + rs.random((n, 3)).astype(f32) * 3], 1).astype(f32)
    tlr = np.stack([np.full(n, 1e-4, f32), np.where(rs.random(n) < 0.5, np.inf, rs.random(n) * 10).astype(f32)], 1)
    g["box_rays"], g["box_boxes"], g["box_tlr_in"] = rays, boxes, tlr
    g["box_tlr_out"] = refbind.ray_in_box(rays, boxes, tlr)
    tris = (rs.normal(size=(n, 9)) * 1.5).astype(f32)
    tris[:128, 3:6] = tris[:128, 0:3]        # degenerate edge1
    g["tri_rays"], g["tri_tris"] = rays, tris
    g["tri_t"] = refbind.ray_triangle(rays, tris)
    pts = rs.normal(size=(1024, 3)).astype(f32)
    g["bary_tris"], g["bary_p"] = tris[:1024], pts
    g["bary_out"] = refbind.barycentric(tris[:1024], pts)
    # ---- BSDF evaluation
    surf = random_surfaces(rs, 2048)
    V = unit(surf["surfaceNormal"] + 0.8 * rs.normal(size=(2048, 3)))
    L = unit(rs.normal(size=(2048, 3)))
    g["bsdf_surf"], g["bsdf_V"], g["bsdf_L"] = surf, V, L
    for which, name in [(0, "bsdf_out"), (1, "brdf_out"), (2, "btdf_out")]:
        g[name] = refbind.bsdf_eval(which, surf, V, L)
    # ---- radiance split
    base = rs.random((2048, 3)).astype(f32)
    base[:128] = base[:128, :1]              # grey -> "white" branch
    base[128:192] = 0.0                      # black -> specular-only branch
    s7 = np.concatenate([rs.random((2048, 3)) * 2, rs.random((2048, 3)) * 5, rs.random((2048, 1))], 1).astype(f32)
    s7[192:256, 3:6] = 0.0                   # no light -> nothing accumulated
    g["acc_base"], g["acc_s7"] = base, s7
    g["acc_out"] = refbind.accumulate(base, s7)
    # ---- RNG float mapping
    u32 = np.concatenate([np.array([0, 1, 2, 127, 128, 129, 2 ** 31, 2 ** 32 - 1, 2 ** 32 - 128, 2 ** 32 - 129], np.uint64),
                          rs.integers(0, 2 ** 32, 2000, dtype=np.uint64)]).astype(np.uint32)
    g["rng_u32"] = u32
    g["rng_out"] = refbind.uniform_from_u32(u32)
    # ---- FXAA
    img = rs.random((40, 48, 3)).astype(f32)
    img[10:20] *= 0.05
    img[:, 20:30] *= 0.3
    g["fxaa_in"] = img
    g["fxaa_out"] = refbind.fxaa(img)
    # ---- scenes: BVH, primary hits, G-buffer
    for name, (scene, args) in {"cornell": scenes.cornell_box(64, 64, 0),
                                "hf": scenes.heightfield_scene(3000, 96, 54, 0, with_sky=True),
                                "tex": scenes.texture_heavy(6000, 96, 54, 0, tex_size=32, n_materials=8)}.items():
        R = refbind.RefScene(scene)
        nodes, perm = R.bvh()
        tri, t = R.trace_primary(args, threads=4)
        g[name + "_scene_hash"] = np.frombuffer(__import__("hashlib").sha256(scene.positions.tobytes() + scene.uvs.tobytes() + scene.normals.tobytes()).digest(), np.uint8)
        g[name + "_nodes"], g[name + "_perm"], g[name + "_tri"], g[name + "_t"] = nodes, perm, tri, t
        g[name + "_gbuffer"] = R.gbuffer(args, threads=4)
        org = (rs.normal(size=(2000, 3)) * 2 + np.array([0, 2.5, 0])).astype(f32)
        d = unit(rs.normal(size=(2000, 3)))
        ctri, ct = R.trace_closest(org, d)
        aim = np.where(ctri >= 0, ct * rs.choice([0.5, 1.0, 1.5], 2000), np.inf).astype(f32)
        g[name + "_rays_o"], g[name + "_rays_d"], g[name + "_rays_tri"], g[name + "_rays_t"] = org, d, ctri, ct
        g[name + "_rays_aim"], g[name + "_rays_occ"] = aim, R.trace_occluded(org, d, aim)
        if name == "tex":
            uvd = np.concatenate([rs.random((512, 2)) * 3 - 1, (rs.random((512, 1)) ** 3) * 0.2], 1).astype(f32)
            uvd[:64, 2] = np.nan
            g["tex_uvd"] = uvd
            for which in range(4):
                g["tex_fetch%d" % which] = R.material_fetch(1, which, uvd)
        if name == "hf":
            dirs = unit(rs.normal(size=(512, 3)))
            g["sky_dirs"], g["sky_out"] = dirs, R.sky_get(dirs)
            sd, sc = R.sky()
            g["sky_cdf_tail"] = sc[-16:]
        R.close()
    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **g)
    print("wrote", os.path.join(OUT, "reference_vectors.npz"), sum(v.nbytes for v in g.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
