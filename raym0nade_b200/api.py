"""Python host side over the C ABI (include/raym0nade_b200.h).

Mirrors the reference's host interface for the hot path: a `Model` that is prepared once
(BVH build, mip chains, light objects, sky CDF - what `Model::Model` does, src/model.cpp:172-215)
and a `render_multiThread(model, args)` that runs the pixel loop (src/render.cpp:593-626) -
here on one B200 through libraym0nade_b200.so.  There is no CPU fallback: if the CUDA
library is missing or no device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .ctypes_defs import (BVHNODE_DTYPE, HITINFO_DTYPE, RADIANCE_DTYPE, RmRawScene, RmRenderArgs)
from .scenes import RawScene, RenderArgs

_HERE = os.path.dirname(os.path.abspath(__file__))
# RM_LIB_PATH: A/B perf experiments load another build of the same CUDA library (scripts/ab_probe.py)
LIB_PATH = os.environ.get("RM_LIB_PATH") or os.path.join(_HERE, "libraym0nade_b200.so")
_LIB = None


class RmError(RuntimeError):
    pass


class RmTextureDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32), ("map_depth", C.c_int32),
                ("levels", C.c_void_p * 8)]


class RmMaterialDesc(C.Structure):
    _fields_ = [("tex", C.c_int32 * 4), ("opacity", C.c_float), ("ior", C.c_float), ("roughness", C.c_float),
                ("transmitting_color", C.c_float * 3), ("has_fully_transparent_part", C.c_int32), ("_pad", C.c_int32)]


class RmLightDesc(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("color", C.c_float * 3), ("power", C.c_float), ("n_faces", C.c_int32),
                ("face_positions", C.c_void_p), ("face_normals", C.c_void_p), ("face_cdf", C.c_void_p)]


class RmSceneDesc(C.Structure):
    _fields_ = [("n_faces", C.c_int32), ("n_nodes", C.c_int32), ("n_materials", C.c_int32), ("n_textures", C.c_int32),
                ("n_lights", C.c_int32), ("sky_width", C.c_int32), ("sky_height", C.c_int32), ("_pad", C.c_int32),
                ("nodes", C.c_void_p), ("positions", C.c_void_p), ("uvs", C.c_void_p), ("normals", C.c_void_p),
                ("face_material", C.c_void_p), ("materials", C.POINTER(RmMaterialDesc)),
                ("textures", C.POINTER(RmTextureDesc)), ("lights", C.POINTER(RmLightDesc)),
                ("sky_data", C.c_void_p), ("sky_cdf", C.c_void_p)]


# every symbol include/raym0nade_b200.h declares
EXPORTS = ["rm_prepare_scene", "rm_prepared_desc", "rm_prepared_permutation", "rm_prepared_free", "rm_prepared_pin", "rm_last_error",
           "rm_version", "rm_context_create", "rm_context_destroy", "rm_context_synchronize", "rm_scene_validate", "rm_scene_upload",
           "rm_scene_device_bytes", "rm_scene_h2d_bytes", "rm_trace_closest", "rm_trace_occluded", "rm_trace_primary", "rm_gbuffer",
           "rm_render_samples", "rm_accum_view", "rm_accum_after_reduce", "rm_accum_radiance", "rm_resolve", "rm_download_resolved", "rm_render", "rm_fxaa",
           "rm_fxaa_device", "rm_postprocess", "rm_spatial_clamp", "rm_filter", "rm_upload_resolved", "rm_depth_field_blur",
           "rm_checkpoint_bytes", "rm_checkpoint_save", "rm_checkpoint_load",
           "rm_comm_unique_id", "rm_comm_init", "rm_reduce", "rm_comm_destroy", "rm_reduce_scatter", "rm_frame_slice", "rm_resolve_slice", "rm_secondary_tree_stats", "rm_wide_tree_stats", "rm_tree_info",
           "rm_tree_node_count", "rm_tree_build", "rm_prepare_scene_device", "rm_scene_refit", "rm_stats_reset", "rm_stats_read", "rm_stats_kernels", "rm_set_option"]


def lib():
    """Load libraym0nade_b200.so (built in-tree by raym0nade_b200.build).  Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RmError("%s is missing - run `python -m raym0nade_b200.build` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    ARGS = C.POINTER(RmRenderArgs)
    L.rm_last_error.restype = C.c_char_p
    L.rm_version.restype = C.c_char_p
    L.rm_prepare_scene.argtypes = [C.POINTER(RmRawScene), C.POINTER(vp)]
    L.rm_prepared_desc.restype = C.POINTER(RmSceneDesc)
    L.rm_prepared_desc.argtypes = [vp]
    L.rm_prepared_permutation.restype = C.POINTER(C.c_int32)
    L.rm_prepared_permutation.argtypes = [vp]
    L.rm_prepared_free.argtypes = [vp]
    L.rm_prepared_pin.argtypes = [vp, i32]
    L.rm_context_create.argtypes = [i32, vp, C.POINTER(vp)]
    L.rm_context_destroy.argtypes = [vp]
    L.rm_context_synchronize.argtypes = [vp]
    L.rm_scene_upload.argtypes = [vp, C.POINTER(RmSceneDesc)]
    L.rm_scene_validate.argtypes = [C.POINTER(RmSceneDesc)]
    L.rm_scene_device_bytes.restype = i64
    L.rm_scene_device_bytes.argtypes = [vp]
    L.rm_scene_h2d_bytes.restype = i64
    L.rm_scene_h2d_bytes.argtypes = [vp]
    L.rm_trace_closest.argtypes = [vp, i64, vp, vp, vp, vp]
    L.rm_trace_occluded.argtypes = [vp, i64, vp, vp, vp, vp]
    L.rm_trace_primary.argtypes = [vp, ARGS, vp, vp]
    L.rm_gbuffer.argtypes = [vp, ARGS, vp]
    L.rm_render_samples.argtypes = [vp, ARGS, i32, i32, u64, i32]
    L.rm_accum_view.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64)]
    L.rm_accum_after_reduce.argtypes = [vp, i32, i32]
    L.rm_accum_radiance.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.rm_resolve.argtypes = [vp, ARGS, vp, vp, vp, vp]
    L.rm_render.argtypes = [vp, ARGS, u64, vp, vp, vp, vp, vp]
    L.rm_download_resolved.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rm_fxaa.argtypes = [vp, vp, vp, i32, i32]
    L.rm_fxaa_device.argtypes = [vp, vp, vp, i32, i32]
    L.rm_postprocess.argtypes = [vp, ARGS, i32, vp]
    L.rm_secondary_tree_stats.argtypes = [vp, i32, i32, i32, vp]
    L.rm_wide_tree_stats.argtypes = [vp, i32, i32, vp]
    L.rm_tree_info.argtypes = [vp, vp]
    L.rm_tree_node_count.argtypes = [i32]
    L.rm_tree_node_count.restype = i32
    L.rm_tree_build.argtypes = [vp, vp, i32, vp, i32, vp]
    L.rm_prepare_scene_device.argtypes = [vp, C.POINTER(RmRawScene), C.POINTER(vp)]
    L.rm_scene_refit.argtypes = [vp, vp, i32]
    L.rm_comm_unique_id.argtypes = [vp]
    L.rm_comm_init.argtypes = [vp, vp, i32, i32]
    L.rm_reduce.argtypes = [vp, i32]
    L.rm_comm_destroy.argtypes = [vp]
    L.rm_reduce_scatter.argtypes = [vp]
    L.rm_frame_slice.argtypes = [vp, vp, vp]
    L.rm_resolve_slice.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.rm_checkpoint_bytes.restype = i64
    L.rm_checkpoint_bytes.argtypes = [ARGS]
    L.rm_checkpoint_save.argtypes = [vp, vp, i64]
    L.rm_checkpoint_load.argtypes = [vp, ARGS, vp, i64]
    L.rm_upload_resolved.argtypes = [vp, ARGS, vp, vp, vp, vp, vp]
    L.rm_depth_field_blur.argtypes = [vp, ARGS, vp, vp]
    L.rm_spatial_clamp.argtypes = [vp, ARGS]
    L.rm_filter.argtypes = [vp, ARGS]
    L.rm_stats_reset.argtypes = [vp]
    L.rm_stats_read.argtypes = [vp, vp]
    L.rm_set_option.argtypes = [vp, C.c_char_p, i64]
    L.rm_stats_kernels.argtypes = [vp, vp, vp, vp]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise RmError("raym0nade_b200 error %d: %s" % (rc, lib().rm_last_error().decode()))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Model:
    """Host-side prepared scene = the reference's loaded `Model` (include/model.h:25-43)."""

    def __init__(self, raw: RawScene, ctx: "Optional[Context]" = None):
        """ctx given: the reference's tree is built on that context's device (rm_prepare_scene_device) instead of by the host recursion"""
        self.raw = raw
        self._c = raw.to_c()
        h = C.c_void_p()
        if ctx is None:
            _check(lib().rm_prepare_scene(C.byref(self._c), C.byref(h)))
        else:
            _check(lib().rm_prepare_scene_device(ctx.h, C.byref(self._c), C.byref(h)))
        self.h = h
        self.desc = lib().rm_prepared_desc(h).contents

    @property
    def n_faces(self):
        return self.desc.n_faces

    def permutation(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().rm_prepared_permutation(self.h), (self.desc.n_faces,)).copy()

    def nodes(self) -> np.ndarray:
        buf = (C.c_char * (32 * self.desc.n_nodes)).from_address(self.desc.nodes)
        return np.frombuffer(buf, BVHNODE_DTYPE).copy()

    def _f32_view(self, addr, n):
        if not addr or n == 0:
            return np.zeros(0, np.float32)
        return np.frombuffer((C.c_char * (4 * n)).from_address(addr), np.float32).copy()

    def lights(self):
        out = []
        for i in range(self.desc.n_lights):
            l = self.desc.lights[i]
            out.append(dict(center=np.array(l.center[:], np.float32), color=np.array(l.color[:], np.float32),
                            power=float(l.power), faces=self._f32_view(l.face_positions, l.n_faces * 9).reshape(-1, 3, 3),
                            cdf=self._f32_view(l.face_cdf, l.n_faces)))
        return out

    def sky(self):
        n = self.desc.sky_width * self.desc.sky_height
        return (self._f32_view(self.desc.sky_data, n * 3).reshape(self.desc.sky_height, self.desc.sky_width, 3),
                self._f32_view(self.desc.sky_cdf, n))

    def texture_levels(self, tex_index):
        t = self.desc.textures[tex_index]
        out = []
        for l in range(t.map_depth):
            nbytes = (t.width >> l) * (t.height >> l) * t.channels
            out.append(np.frombuffer((C.c_char * nbytes).from_address(t.levels[l]), np.uint8).copy())
        return out, t.map_depth

    def material(self, i):
        return self.desc.materials[i]

    def validate(self):
        """rm_scene_validate on the prepared scene (host only); raises RmError naming the first inconsistency"""
        _check(lib().rm_scene_validate(C.byref(self.desc)))

    def pin(self, on=True):
        """page-lock (or release) the prepared arrays: later uploads of this model are DMAs straight out of them"""
        _check(lib().rm_prepared_pin(self.h, 1 if on else 0))
        self._pinned = bool(on)

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "_pinned", False):
                lib().rm_prepared_pin(self.h, 0)
            lib().rm_prepared_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One CUDA device + one staged scene.  `stream` = a raw cudaStream_t (int) or None."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        h = C.c_void_p()
        _check(lib().rm_context_create(device, C.c_void_p(stream or 0), C.byref(h)))
        self.h = h
        self.model = None

    def upload(self, model: Model):
        _check(lib().rm_scene_upload(self.h, C.byref(model.desc)))
        self.model = model
        return self

    def tree_build(self, positions):
        """BVH::build on the device: (nodes, perm) for raw positions [n][3][3]"""
        pos = _f32(positions).reshape(-1, 9)
        n = pos.shape[0]
        nn = lib().rm_tree_node_count(n)
        nodes, perm = np.zeros(nn, BVHNODE_DTYPE), np.zeros(n, np.int32)
        _check(lib().rm_tree_build(self.h, _p(pos), n, _p(nodes), nn, _p(perm)))
        return nodes, perm

    def refit(self, positions):
        """moved vertices (post-build order, [n][3][3]): both trees of the uploaded scene are refitted on the device"""
        pos = _f32(positions).reshape(-1, 9)
        _check(lib().rm_scene_refit(self.h, _p(pos), pos.shape[0]))

    def tree_info(self):
        """the 4-wide secondary-ray tree bounce / shadow rays currently traverse: dict(device_built, builder ("host" | "ploc" | "ploc+host refinement" |
        "sweep_sah"), refined (the host builder's tree has been swapped in), nodes, levels, in_use)"""
        out = np.zeros(4, np.int32)
        _check(lib().rm_tree_info(self.h, _p(out)))
        return dict(device_built=bool(out[0]), builder=("host", "ploc", "ploc+host refinement", "sweep_sah")[int(out[0])], refined=int(out[0]) == 2, nodes=int(out[1]), levels=int(out[2]),
                    in_use=bool(out[3]))

    def scene_bytes(self):
        return lib().rm_scene_device_bytes(self.h)

    def scene_h2d_bytes(self):
        return lib().rm_scene_h2d_bytes(self.h)

    def set_option(self, name, value):
        _check(lib().rm_set_option(self.h, name.encode(), int(value)))

    def synchronize(self):
        _check(lib().rm_context_synchronize(self.h))

    def stats_reset(self):
        _check(lib().rm_stats_reset(self.h))

    def stats(self):
        out = np.zeros(4, np.uint64)
        _check(lib().rm_stats_read(self.h, _p(out)))
        return dict(rays=int(out[0]), box=int(out[1]), tri=int(out[2]), launches=int(out[3]))

    def stats_kernels(self):
        """per kernel kind (primary, paths, shadow, shade): rays/box/tri counters, device ms, launches"""
        cnt, ms, n = np.zeros(9, np.uint64), np.zeros(4, np.float64), np.zeros(4, np.uint64)
        _check(lib().rm_stats_kernels(self.h, _p(cnt), _p(ms), _p(n)))
        kinds = ["primary", "paths", "shadow", "shade"]
        return {k: dict(rays=int(cnt[3 * i]) if i < 3 else 0, box=int(cnt[3 * i + 1]) if i < 3 else 0,
                        tri=int(cnt[3 * i + 2]) if i < 3 else 0, ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(kinds)}

    # ---- per-ray seam
    def trace_closest(self, org, dirs):
        org, dirs = _f32(org), _f32(dirs)
        n = org.shape[0]
        tri, t = np.zeros(n, np.int32), np.zeros(n, np.float32)
        _check(lib().rm_trace_closest(self.h, n, _p(org), _p(dirs), _p(tri), _p(t)))
        return tri, t

    def trace_occluded(self, org, dirs, aim):
        org, dirs, aim = _f32(org), _f32(dirs), _f32(aim)
        out = np.zeros(org.shape[0], np.uint8)
        _check(lib().rm_trace_occluded(self.h, org.shape[0], _p(org), _p(dirs), _p(aim), _p(out)))
        return out

    # ---- per-pixel stages
    def trace_primary(self, args: RenderArgs, download=True):
        a = args.to_c()
        n = args.width * args.height
        tri = np.zeros(n, np.int32) if download else None
        t = np.zeros(n, np.float32) if download else None
        _check(lib().rm_trace_primary(self.h, C.byref(a), _p(tri), _p(t)))
        return tri, t

    def trace_primary_into(self, args: RenderArgs, tri_idx, t):
        """rm_trace_primary with caller-owned (e.g. pinned) host arrays"""
        a = args.to_c()
        _check(lib().rm_trace_primary(self.h, C.byref(a), _p(tri_idx), _p(t)))

    def gbuffer(self, args: RenderArgs, download=True):
        a = args.to_c()
        g = np.zeros(args.width * args.height, HITINFO_DTYPE) if download else None
        _check(lib().rm_gbuffer(self.h, C.byref(a), _p(g)))
        return g

    def render_samples(self, args: RenderArgs, sample_begin=0, sample_stride=1, seed=0, reset=True):
        a = args.to_c()
        _check(lib().rm_render_samples(self.h, C.byref(a), sample_begin, sample_stride, seed, int(reset)))

    def accum_view(self):
        ps, ns, pm, nm = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64()
        _check(lib().rm_accum_view(self.h, C.byref(ps), C.byref(ns), C.byref(pm), C.byref(nm)))
        return ps.value, ns.value, pm.value, nm.value

    def accum_after_reduce(self, rank, world):
        _check(lib().rm_accum_after_reduce(self.h, rank, world))

    def accum_radiance(self):
        p, n = C.c_void_p(), C.c_int64()
        _check(lib().rm_accum_radiance(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def resolve(self, args: RenderArgs, download=True):
        a = args.to_c()
        n = args.width * args.height
        planes = [np.zeros(n, RADIANCE_DTYPE) if download else None for _ in range(4)]
        _check(lib().rm_resolve(self.h, C.byref(a), *[_p(p) for p in planes]))
        return dict(Dd=planes[0], Ds=planes[1], Id=planes[2], Is=planes[3])

    def download_resolved(self, gbuffer, planes):
        """copy the resolved G-buffer + [Dd, Ds, Id, Is] into caller-owned (e.g. pinned) arrays"""
        _check(lib().rm_download_resolved(self.h, _p(gbuffer), *[_p(p) for p in planes]))

    def render_into(self, args: RenderArgs, seed, gbuffer, planes):
        """rm_render with caller-owned host output arrays"""
        a = args.to_c()
        _check(lib().rm_render(self.h, C.byref(a), seed, _p(gbuffer), *[_p(p) for p in planes]))

    def render(self, args: RenderArgs, seed=0, download=True):
        """The whole render_multiThread pixel loop on this GPU; returns gbuffer + 4 planes."""
        a = args.to_c()
        n = args.width * args.height
        g = np.zeros(n, HITINFO_DTYPE) if download else None
        planes = [np.zeros(n, RADIANCE_DTYPE) if download else None for _ in range(4)]
        _check(lib().rm_render(self.h, C.byref(a), seed, _p(g), *[_p(p) for p in planes]))
        return dict(gbuffer=g, Dd=planes[0], Ds=planes[1], Id=planes[2], Is=planes[3])

    # ---- post pass
    def fxaa(self, rgb):
        rgb = _f32(rgb)
        h, w = rgb.shape[:2]
        out = np.zeros_like(rgb)
        _check(lib().rm_fxaa(self.h, _p(rgb), _p(out), w, h))
        return out

    def fxaa_device(self, d_in: int, d_out: int, width: int, height: int):
        _check(lib().rm_fxaa_device(self.h, C.c_void_p(d_in), C.c_void_p(d_out), width, height))

    # ---- multi-GPU exchange inside the library (NCCL)
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (call on rank 0, hand to every rank)"""
        uid = np.zeros(128, np.uint8)
        _check(lib().rm_comm_unique_id(_p(uid)))
        return uid

    def comm_init(self, uid, rank: int, world: int):
        uid = np.ascontiguousarray(uid, np.uint8)
        assert uid.size == 128
        _check(lib().rm_comm_init(self.h, _p(uid), rank, world))

    def reduce(self, root: int = 0):
        """the three-step frame reduction over the context's communicator, on the context's stream"""
        _check(lib().rm_reduce(self.h, root))

    def reduce_scatter(self):
        """the frame reduction with its last step scattered: this rank ends up holding its slice of the summed frame"""
        _check(lib().rm_reduce_scatter(self.h))

    def frame_slice(self):
        """(first pixel, pixel count) of the frame this rank holds after reduce_scatter (the whole frame without one)"""
        a, b = C.c_int64(0), C.c_int64(0)
        _check(lib().rm_frame_slice(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def resolve_slice(self, args: RenderArgs, gbuffer=None, planes=(None, None, None, None)):
        """finalise this rank's slice and write it into the given WHOLE-frame host arrays (any may be None)"""
        a = args.to_c()
        _check(lib().rm_resolve_slice(self.h, C.byref(a), _p(gbuffer), *[_p(p) for p in planes]))

    def checkpoint_save(self, args: RenderArgs):
        """the un-finalised accumulators of the current frame as a byte blob"""
        a = args.to_c()
        n = lib().rm_checkpoint_bytes(C.byref(a))
        blob = np.zeros(n, np.uint8)
        _check(lib().rm_checkpoint_save(self.h, _p(blob), n))
        return blob

    def checkpoint_load(self, args: RenderArgs, blob):
        a = args.to_c()
        blob = np.ascontiguousarray(blob, np.uint8)
        _check(lib().rm_checkpoint_load(self.h, C.byref(a), _p(blob), blob.size))

    def upload_resolved(self, args: RenderArgs, gbuffer, planes):
        """stage host Photo buffers (G-buffer + [Dd, Ds, Id, Is]) as the resolved frame"""
        a = args.to_c()
        g = np.ascontiguousarray(gbuffer)
        pl = [np.ascontiguousarray(p) for p in planes]
        _check(lib().rm_upload_resolved(self.h, C.byref(a), _p(g), *[_p(p) for p in pl]))

    def spatial_clamp(self, args: RenderArgs):
        """Photo::spatialClamp on the resolved planes (device, in place)"""
        a = args.to_c()
        _check(lib().rm_spatial_clamp(self.h, C.byref(a)))

    def filter(self, args: RenderArgs):
        """Photo::filter on the resolved planes (device, in place)"""
        a = args.to_c()
        _check(lib().rm_filter(self.h, C.byref(a)))

    def resolved(self, args: RenderArgs):
        """the resolved G-buffer + four planes as they currently stand on the device"""
        n = args.width * args.height
        g = np.zeros(n, HITINFO_DTYPE)
        planes = [np.zeros(n, RADIANCE_DTYPE) for _ in range(4)]
        self.download_resolved(g, planes)
        return dict(gbuffer=g, Dd=planes[0], Ds=planes[1], Id=planes[2], Is=planes[3])

    def depth_field_blur(self, args: RenderArgs, rgb):
        """Photo::depthFeildBlur on an rgb frame [h][w][3] with the resolved G-buffer; focus / CoC / position from args"""
        a = args.to_c()
        rgb = _f32(rgb)
        out = np.zeros_like(rgb)
        _check(lib().rm_depth_field_blur(self.h, C.byref(a), _p(rgb), _p(out)))
        return out

    def postprocess(self, args: RenderArgs, shade_options: int):
        a = args.to_c()
        out = np.zeros((args.height, args.width, 3), np.float32)
        _check(lib().rm_postprocess(self.h, C.byref(a), shade_options, _p(out)))
        return out

    def close(self):
        if getattr(self, "h", None):
            lib().rm_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def secondary_tree_stats(positions, depth_cap=22, leaf_max=3):
    """build + verify the secondary-ray tree on the host; returns dict(blocks, depth, leaves, largest_leaf)"""
    pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 9)
    out = np.zeros(4, np.int32)
    _check(lib().rm_secondary_tree_stats(_p(pos), pos.shape[0], depth_cap, leaf_max, _p(out)))
    return dict(blocks=int(out[0]), depth=int(out[1]), leaves=int(out[2]), largest_leaf=int(out[3]))


def wide_tree_stats(positions, depth_cap=22):
    """build + verify the 4-wide secondary-ray tree on the host; returns dict(nodes, levels, leaves, children_per_node)"""
    pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 9)
    out = np.zeros(4, np.int32)
    _check(lib().rm_wide_tree_stats(_p(pos), pos.shape[0], depth_cap, _p(out)))
    return dict(nodes=int(out[0]), levels=int(out[1]), leaves=int(out[2]), children_per_node=out[3] / 100.0)


def render_multiThread(model: Model, args: RenderArgs, device: int = 0, seed: int = 0):
    """Drop-in for the reference's `render_multiThread(Model&, const RenderArgs&)`
    (include/render.h:43): runs the pixel loop and returns the Photo buffers
    (G-buffer + the four RadianceData planes) instead of writing PNGs."""
    ctx = Context(device).upload(model)
    try:
        return ctx.render(args, seed=seed)
    finally:
        ctx.close()
