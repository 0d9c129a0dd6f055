"""ctypes binding for oracle/_build/libraym_port.so - the glm-free CPU restatement (oracle/port).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from raym0nade_b200.ctypes_defs import BVHNODE_DTYPE, HITINFO_DTYPE, RmRawScene, RmRenderArgs

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libraym_port.so")
_L = None


def load():
    global _L
    if _L is not None:
        return _L
    src = [os.path.join(_HERE, "port", f) for f in os.listdir(os.path.join(_HERE, "port"))]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)
    L = C.CDLL(LIB)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.port_scene_create.restype = vp
    L.port_scene_create.argtypes = [C.POINTER(RmRawScene)]
    L.port_scene_destroy.argtypes = [vp]
    L.port_node_count.argtypes = [vp]
    L.port_bvh_export.argtypes = [vp, vp, vp]
    L.port_trace_primary.argtypes = [vp, C.POINTER(RmRenderArgs), i32, vp, vp, vp]
    L.port_trace_closest.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.port_trace_occluded.argtypes = [vp, i64, vp, vp, vp, vp]
    L.port_gbuffer.argtypes = [vp, C.POINTER(RmRenderArgs), i32, vp]
    L.port_fxaa.argtypes = [vp, vp, i32, i32]
    L.port_postprocess.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_float, i32, vp]
    L.port_denoise.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32]
    L.port_bloom.argtypes = [vp, i32, i32]
    L.port_depth_field_blur.argtypes = [vp, vp, i32, i32, vp, C.c_float, C.c_float]
    L.port_kat_ray_in_box.argtypes = [i64, vp, vp, vp]
    L.port_kat_ray_triangle.argtypes = [i64, vp, vp, vp]
    L.port_kat_barycentric.argtypes = [i64, vp, vp, vp]
    L.port_kat_bsdf.argtypes = [i32, i64, vp, vp, vp, vp]
    L.port_kat_accumulate.argtypes = [i64, vp, vp, vp]
    L.port_kat_material_fetch.argtypes = [vp, i32, i32, i64, vp, vp]
    L.port_kat_sky_get.argtypes = [vp, i64, vp, vp]
    L.port_uniform_from_u32.argtypes = [vp, i32, vp]
    _L = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class PortScene:
    def __init__(self, raw_scene):
        self.L = load()
        self._c = raw_scene.to_c()
        self.h = self.L.port_scene_create(C.byref(self._c))
        self.n_faces = raw_scene.n_faces

    def close(self):
        if self.h:
            self.L.port_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bvh(self):
        n = self.L.port_node_count(self.h)
        nodes, perm = np.zeros(n, BVHNODE_DTYPE), np.zeros(self.n_faces, np.int32)
        self.L.port_bvh_export(self.h, _p(nodes), _p(perm))
        return nodes, perm

    def trace_primary(self, args, threads=8, counters=False):
        a = args.to_c()
        n = args.width * args.height
        tri, t, cnt = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(3, np.uint64)
        self.L.port_trace_primary(self.h, C.byref(a), threads, _p(tri), _p(t), _p(cnt))
        return (tri, t, cnt) if counters else (tri, t)

    def trace_closest(self, org, dirs, counters=False):
        org, dirs = _f32(org), _f32(dirs)
        n = org.shape[0]
        tri, t, cnt = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(3, np.uint64)
        self.L.port_trace_closest(self.h, n, _p(org), _p(dirs), _p(tri), _p(t), _p(cnt))
        return (tri, t, cnt) if counters else (tri, t)

    def trace_occluded(self, org, dirs, aim):
        org, dirs, aim = _f32(org), _f32(dirs), _f32(aim)
        out = np.zeros(org.shape[0], np.uint8)
        self.L.port_trace_occluded(self.h, org.shape[0], _p(org), _p(dirs), _p(aim), _p(out))
        return out

    def gbuffer(self, args, threads=8):
        a = args.to_c()
        g = np.zeros(args.width * args.height, HITINFO_DTYPE)
        self.L.port_gbuffer(self.h, C.byref(a), threads, _p(g))
        return g

    def material_fetch(self, material, which, uvd):
        uvd = _f32(uvd)
        out = np.zeros((uvd.shape[0], 4), np.float32)
        self.L.port_kat_material_fetch(self.h, material, which, uvd.shape[0], _p(uvd), _p(out))
        return out

    def sky_get(self, dirs):
        dirs = _f32(dirs)
        out = np.zeros((dirs.shape[0], 3), np.float32)
        self.L.port_kat_sky_get(self.h, dirs.shape[0], _p(dirs), _p(out))
        return out


def ray_in_box(rays, boxes, tlr):
    rays, boxes, tlr = _f32(rays), _f32(boxes), _f32(tlr).copy()
    load().port_kat_ray_in_box(rays.shape[0], _p(rays), _p(boxes), _p(tlr))
    return tlr


def ray_triangle(rays, tris):
    rays, tris = _f32(rays), _f32(tris)
    t = np.zeros(rays.shape[0], np.float32)
    load().port_kat_ray_triangle(rays.shape[0], _p(rays), _p(tris), _p(t))
    return t


def barycentric(tris, p):
    tris, p = _f32(tris), _f32(p)
    out = np.zeros((p.shape[0], 3), np.float32)
    load().port_kat_barycentric(p.shape[0], _p(tris), _p(p), _p(out))
    return out


def bsdf_eval(which, surf, in_dirs, out_dirs):
    surf = np.ascontiguousarray(surf, HITINFO_DTYPE)
    in_dirs, out_dirs = _f32(in_dirs), _f32(out_dirs)
    out = np.zeros((surf.shape[0], 3), np.float32)
    load().port_kat_bsdf(which, surf.shape[0], _p(surf), _p(in_dirs), _p(out_dirs), _p(out))
    return out


def accumulate(base_colors, samples7):
    base_colors, samples7 = _f32(base_colors), _f32(samples7)
    out = np.zeros((samples7.shape[0], 8), np.float32)
    load().port_kat_accumulate(samples7.shape[0], _p(base_colors), _p(samples7), _p(out))
    return out


def uniform_from_u32(u32):
    u32 = np.ascontiguousarray(u32, np.uint32)
    out = np.zeros(u32.shape[0], np.float32)
    load().port_uniform_from_u32(_p(u32), u32.shape[0], _p(out))
    return out


def fxaa(rgb):
    rgb = _f32(rgb)
    h, w = rgb.shape[:2]
    out = np.zeros_like(rgb)
    load().port_fxaa(_p(rgb), _p(out), w, h)
    return out


def depth_field_blur(gbuffer, rgb, camera_position, focus, coc):
    """Photo::depthFeildBlur on an rgb frame [h][w][3] with the frame's G-buffer"""
    out = _f32(rgb).copy()
    h, w = out.shape[:2]
    cam = _f32(np.asarray(camera_position, np.float32))
    load().port_depth_field_blur(_p(gbuffer), _p(out), w, h, _p(cam), C.c_float(focus), C.c_float(coc))
    return out


def denoise(gbuffer, Dd, Ds, Id, Is, width, height, stages):
    """Photo::spatialClamp (stages & 1) then Photo::filter (stages & 2) on copies of the four planes"""
    planes = [np.ascontiguousarray(p).copy() for p in (Dd, Ds, Id, Is)]
    load().port_denoise(_p(gbuffer), *[_p(p) for p in planes], width, height, stages)
    return dict(Dd=planes[0], Ds=planes[1], Id=planes[2], Is=planes[3])


def postprocess(gbuffer, Dd, Ds, Id, Is, width, height, exposure, shade_options):
    out = np.zeros((height, width, 3), np.float32)
    load().port_postprocess(_p(gbuffer), _p(Dd), _p(Ds), _p(Id), _p(Is), width, height, C.c_float(exposure), shade_options, _p(out))
    return out
