"""ctypes binding for oracle/_ref/libraym_ref*.so - the UNMODIFIED reference compiled as a
CPU oracle (see oracle/ref_harness.cpp, oracle/Makefile).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.  Nothing under raym0nade_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from raym0nade_b200.ctypes_defs import (BVHNODE_DTYPE, HITINFO_DTYPE, RADIANCE_DTYPE, RmRawScene, RmRenderArgs)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

# aiTextureType values the hot path reads
SLOT_DIFFUSE, SLOT_SPECULAR, SLOT_EMISSIVE, SLOT_NORMALS = 1, 2, 4, 6


def lib_path(flavour="plain"):
    """'plain' / 'count': the reference + harness; 'bridge': the same plus the reference-tree binding
    (raym0nade_b200/host/reference_tree), linked against the product library (make -C oracle bridge)"""
    return os.path.join(_HERE, "_ref", {"plain": "libraym_ref.so", "count": "libraym_ref_count.so", "bridge": "libraym_bridge.so"}[flavour])


def available(flavour="plain") -> bool:
    return os.path.exists(lib_path(flavour))


def load(flavour="plain"):
    """flavour 'plain' = stock path (+ ray counter); 'count' = also counts box/triangle tests."""
    if flavour in _LIBS:
        return _LIBS[flavour]
    path = lib_path(flavour)
    if not os.path.exists(path):
        raise RuntimeError("%s missing: run `make -C oracle ref` where /root/reference is mounted" % path)
    L = C.CDLL(path)
    vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_float)
    L.ref_scene_create.restype = vp
    L.ref_scene_create.argtypes = [C.POINTER(RmRawScene)]
    L.ref_scene_destroy.argtypes = [vp]
    L.ref_node_count.argtypes = [vp]
    L.ref_light_count.argtypes = [vp]
    L.ref_bvh_export.argtypes = [vp, vp, vp]
    L.ref_light_export.argtypes = [vp, i32, vp, vp]
    L.ref_light_faces.argtypes = [vp, i32, vp, vp]
    L.ref_sky_export.argtypes = [vp, vp, vp]
    L.ref_texture_level.restype = i64
    L.ref_texture_level.argtypes = [vp, i32, i32, i32, vp, vp]
    L.ref_trace_primary.argtypes = [vp, C.POINTER(RmRenderArgs), i32, vp, vp, vp]
    L.ref_trace_closest.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.ref_trace_occluded.argtypes = [vp, i64, vp, vp, vp, vp]
    L.ref_render.argtypes = [vp, C.POINTER(RmRenderArgs), i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.ref_replay_indirect.argtypes = [vp, C.POINTER(RmRenderArgs), i32, i32, vp, vp, i32, vp, i32, vp]
    L.ref_replay_direct.argtypes = [vp, C.POINTER(RmRenderArgs), vp, vp, i32, vp, vp]
    L.ref_uniform_from_u32.argtypes = [vp, i32, vp]
    L.ref_gbuffer.argtypes = [vp, C.POINTER(RmRenderArgs), i32, vp]
    L.ref_fxaa.argtypes = [vp, vp, i32, i32]
    L.ref_postprocess.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_float, i32, vp]
    L.ref_denoise.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32]
    L.ref_postprocess_full.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_float, i32, vp, C.c_float, C.c_float, vp]
    L.ref_depth_field_blur.argtypes = [vp, vp, vp, i32, i32, vp, C.c_float, C.c_float]
    L.ref_kat_ray_in_box.argtypes = [i64, vp, vp, vp]
    L.ref_kat_ray_triangle.argtypes = [i64, vp, vp, vp]
    L.ref_kat_barycentric.argtypes = [i64, vp, vp, vp]
    L.ref_kat_material_fetch.argtypes = [vp, i32, i32, i64, vp, vp]
    L.ref_kat_sky_get.argtypes = [vp, i64, vp, vp]
    L.ref_kat_bsdf.argtypes = [i32, i64, vp, vp, vp, vp]
    L.ref_kat_precise_refraction.argtypes = [i64, vp, vp, vp]
    L.ref_kat_accumulate.argtypes = [i64, vp, vp, vp]
    L.ref_kat_absorb.argtypes = [i64, vp, vp, vp]
    L.ref_build_flavour.restype = C.c_char_p
    if flavour == "bridge":
        L.ref_bridge_create.restype = vp
        L.ref_bridge_create.argtypes = [vp]
        L.ref_bridge_desc.restype = vp
        L.ref_bridge_desc.argtypes = [vp]
        L.ref_bridge_destroy.argtypes = [vp]
        L.ref_bridge_render.argtypes = [vp, C.POINTER(RmRenderArgs), i32, vp, vp]
    _LIBS[flavour] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefScene:
    """A reference `Model` built from a RawScene through Model's public members."""

    def __init__(self, raw_scene, flavour="plain"):
        self.L = load(flavour)
        self.raw = raw_scene
        self._c = raw_scene.to_c()
        self.h = self.L.ref_scene_create(C.byref(self._c))
        self.n_faces = raw_scene.n_faces

    def close(self):
        if self.h:
            self.L.ref_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- structure export
    def bvh(self):
        n = self.L.ref_node_count(self.h)
        nodes = np.zeros(n, BVHNODE_DTYPE)
        perm = np.zeros(self.n_faces, np.int32)
        self.L.ref_bvh_export(self.h, _p(nodes), _p(perm))
        return nodes, perm

    def lights(self):
        out = []
        for i in range(self.L.ref_light_count(self.h)):
            ccp = np.zeros(7, np.float32)
            nf = C.c_int32(0)
            self.L.ref_light_export(self.h, i, _p(ccp), C.byref(nf))
            faces = np.zeros((nf.value, 3, 3), np.float32)
            cdf = np.zeros(nf.value, np.float32)
            self.L.ref_light_faces(self.h, i, _p(faces), _p(cdf))
            out.append(dict(center=ccp[0:3].copy(), color=ccp[3:6].copy(), power=float(ccp[6]), faces=faces, cdf=cdf))
        return out

    def sky(self):
        w, h = self._c.sky_width, self._c.sky_height
        data = np.zeros((h, w, 3), np.float32)
        cdf = np.zeros(h * w, np.float32)
        if w:
            self.L.ref_sky_export(self.h, _p(data), _p(cdf))
        return data, cdf

    def texture_levels(self, material, slot):
        depth = C.c_int32(0)
        levels = []
        for lv in range(8):
            nbytes = self.L.ref_texture_level(self.h, material, slot, lv, None, C.byref(depth))
            if nbytes == 0:
                break
            buf = np.zeros(nbytes, np.uint8)
            self.L.ref_texture_level(self.h, material, slot, lv, _p(buf), C.byref(depth))
            levels.append(buf)
        return levels, depth.value

    # ---- tracing
    def trace_primary(self, args, threads=8, counters=False):
        a = args.to_c()
        n = args.width * args.height
        tri = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        cnt = np.zeros(3, np.uint64)
        self.L.ref_trace_primary(self.h, C.byref(a), threads, _p(tri), _p(t), _p(cnt))
        return (tri, t, cnt) if counters else (tri, t)

    def trace_closest(self, org, dirs, counters=False):
        org, dirs = _f32(org), _f32(dirs)
        n = org.shape[0]
        tri = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        cnt = np.zeros(3, np.uint64)
        self.L.ref_trace_closest(self.h, n, _p(org), _p(dirs), _p(tri), _p(t), _p(cnt))
        return (tri, t, cnt) if counters else (tri, t)

    def trace_occluded(self, org, dirs, aim):
        org, dirs, aim = _f32(org), _f32(dirs), _f32(aim)
        out = np.zeros(org.shape[0], np.uint8)
        self.L.ref_trace_occluded(self.h, org.shape[0], _p(org), _p(dirs), _p(aim), _p(out))
        return out

    # ---- rendering
    def render(self, args, threads=8, seed_base=0):
        """Returns dict(gbuffer, Dd, Ds, Id, Is (structured arrays), seconds, rays, box, tri)."""
        a = args.to_c()
        n = args.width * args.height
        g = np.zeros(n, HITINFO_DTYPE)
        planes = [np.zeros(n, RADIANCE_DTYPE) for _ in range(4)]
        sec = C.c_double(0)
        cnt = np.zeros(3, np.uint64)
        self.L.ref_render(self.h, C.byref(a), threads, seed_base, _p(g), *[_p(p) for p in planes], C.byref(sec), _p(cnt))
        return dict(gbuffer=g, Dd=planes[0], Ds=planes[1], Id=planes[2], Is=planes[3], seconds=sec.value,
                    rays=int(cnt[0]), box=int(cnt[1]), tri=int(cnt[2]))

    def gbuffer(self, args, threads=8):
        a = args.to_c()
        g = np.zeros(args.width * args.height, HITINFO_DTYPE)
        self.L.ref_gbuffer(self.h, C.byref(a), threads, _p(g))
        return g

    def replay_indirect(self, args, x, y, g_entry, u32, max_samples=8):
        a = args.to_c()
        u32 = np.ascontiguousarray(u32, np.uint32)
        g = np.ascontiguousarray(g_entry, HITINFO_DTYPE).reshape(1)
        out = np.zeros((max_samples, 7), np.float32)
        used = C.c_int32(0)
        n = self.L.ref_replay_indirect(self.h, C.byref(a), x, y, _p(g), _p(u32), len(u32), _p(out), max_samples, C.byref(used))
        return out[:min(n, max_samples)], used.value

    def replay_direct(self, args, g_entry, u32):
        a = args.to_c()
        u32 = np.ascontiguousarray(u32, np.uint32)
        g = np.ascontiguousarray(g_entry, HITINFO_DTYPE).reshape(1)
        out = np.zeros(7, np.float32)
        used = C.c_int32(0)
        n = self.L.ref_replay_direct(self.h, C.byref(a), _p(g), _p(u32), len(u32), _p(out), C.byref(used))
        return (out if n else None), used.value

    # ---- per-function known answers that need a scene
    def material_fetch(self, material, which, uvd):
        uvd = _f32(uvd)
        out = np.zeros((uvd.shape[0], 4), np.float32)
        self.L.ref_kat_material_fetch(self.h, material, which, uvd.shape[0], _p(uvd), _p(out))
        return out

    def sky_get(self, dirs):
        dirs = _f32(dirs)
        out = np.zeros((dirs.shape[0], 3), np.float32)
        self.L.ref_kat_sky_get(self.h, dirs.shape[0], _p(dirs), _p(out))
        return out


# ---- scene-free known answers
def ray_in_box(rays, boxes, tlr, flavour="plain"):
    L = load(flavour)
    rays, boxes = _f32(rays), _f32(boxes)
    tlr = _f32(tlr).copy()
    L.ref_kat_ray_in_box(rays.shape[0], _p(rays), _p(boxes), _p(tlr))
    return tlr


def ray_triangle(rays, tris):
    L = load()
    rays, tris = _f32(rays), _f32(tris)
    t = np.zeros(rays.shape[0], np.float32)
    L.ref_kat_ray_triangle(rays.shape[0], _p(rays), _p(tris), _p(t))
    return t


def barycentric(tris, p):
    L = load()
    tris, p = _f32(tris), _f32(p)
    out = np.zeros((p.shape[0], 3), np.float32)
    L.ref_kat_barycentric(p.shape[0], _p(tris), _p(p), _p(out))
    return out


def bsdf_eval(which, surf, in_dirs, out_dirs):
    L = load()
    surf = np.ascontiguousarray(surf, HITINFO_DTYPE)
    in_dirs, out_dirs = _f32(in_dirs), _f32(out_dirs)
    out = np.zeros((surf.shape[0], 3), np.float32)
    L.ref_kat_bsdf(which, surf.shape[0], _p(surf), _p(in_dirs), _p(out_dirs), _p(out))
    return out


def precise_refraction(surf, in_dirs):
    L = load()
    surf = np.ascontiguousarray(surf, HITINFO_DTYPE)
    in_dirs = _f32(in_dirs)
    out = np.zeros((surf.shape[0], 4), np.float32)
    L.ref_kat_precise_refraction(surf.shape[0], _p(surf), _p(in_dirs), _p(out))
    return out


def accumulate(base_colors, samples7):
    L = load()
    base_colors, samples7 = _f32(base_colors), _f32(samples7)
    out = np.zeros((samples7.shape[0], 8), np.float32)
    L.ref_kat_accumulate(samples7.shape[0], _p(base_colors), _p(samples7), _p(out))
    return out


def absorb(absorb_rgb, dist):
    L = load()
    absorb_rgb, dist = _f32(absorb_rgb), _f32(dist)
    out = np.zeros((dist.shape[0], 3), np.float32)
    L.ref_kat_absorb(dist.shape[0], _p(absorb_rgb), _p(dist), _p(out))
    return out


def uniform_from_u32(u32):
    L = load()
    u32 = np.ascontiguousarray(u32, np.uint32)
    out = np.zeros(u32.shape[0], np.float32)
    L.ref_uniform_from_u32(_p(u32), u32.shape[0], _p(out))
    return out


def fxaa(rgb):
    L = load()
    rgb = _f32(rgb)
    h, w = rgb.shape[:2]
    out = np.zeros_like(rgb)
    L.ref_fxaa(_p(rgb), _p(out), w, h)
    return out


def postprocess_full(gbuffer, Dd, Ds, Id, Is, width, height, exposure, shade_options, camera_position, focus, coc):
    """Photo::postProcessing with depth of field available"""
    L = load()
    out = np.zeros((height, width, 3), np.float32)
    cam = _f32(np.asarray(camera_position, np.float32))
    L.ref_postprocess_full(_p(gbuffer), _p(Dd), _p(Ds), _p(Id), _p(Is), width, height, C.c_float(exposure), shade_options, _p(cam),
                           C.c_float(focus), C.c_float(coc), _p(out))
    return out


def depth_field_blur(gbuffer, rgb, camera_position, focus, coc):
    """Photo::depthFeildBlur on an rgb frame [h][w][3] with the frame's G-buffer"""
    L = load()
    rgb = _f32(rgb)
    h, w = rgb.shape[:2]
    out = np.zeros_like(rgb)
    cam = _f32(np.asarray(camera_position, np.float32))
    L.ref_depth_field_blur(_p(gbuffer), _p(rgb), _p(out), w, h, _p(cam), C.c_float(focus), C.c_float(coc))
    return out


def denoise(gbuffer, Dd, Ds, Id, Is, width, height, stages):
    """Photo::spatialClamp (stages & 1) then Photo::filter (stages & 2) on copies of the four planes"""
    L = load()
    planes = [np.ascontiguousarray(p).copy() for p in (Dd, Ds, Id, Is)]
    L.ref_denoise(_p(gbuffer), *[_p(p) for p in planes], width, height, stages)
    return dict(Dd=planes[0], Ds=planes[1], Id=planes[2], Is=planes[3])


def postprocess(gbuffer, Dd, Ds, Id, Is, width, height, exposure, shade_options):
    L = load()
    out = np.zeros((height, width, 3), np.float32)
    L.ref_postprocess(_p(gbuffer), _p(Dd), _p(Ds), _p(Id), _p(Is), width, height, C.c_float(exposure), shade_options, _p(out))
    return out


# Photo::ShadeOption (include/image.h:17-36)
SHADE = dict(BaseColor=1, Emission=2, DirectLight=4, IndirectLight=8, Diffuse=16, Specular=32, shapeNormal=64,
             surfaceNormal=128, DoBloom=256, DoFXAA=512, DoDepthFieldBlur=1024)
SHADE["Full"] = 4 | 8 | 16 | 32 | 1 | 2
