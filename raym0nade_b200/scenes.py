"""Synthetic, seeded, trig-free scene generators for the five BASELINE.json configs.

The reference ships no assets (its `model/` directory is git-ignored) and loads scenes
through assimp, so every scene here is procedural and is produced directly in the raw
form `Model::processMesh` / `processMaterial` would hand on (src/model.cpp:88-168):
per-face corner positions / uvs / normals, one material per mesh, RGBA8 / RGB8 level-0
textures, and an optional equirectangular fp32 sky.

All arithmetic is +, -, *, /, sqrt on float64 followed by one rounding to float32 and
integer hashing for noise - no sin/cos - so a given (generator, size, seed) yields
bit-identical arrays on any machine.  That is what lets `tests/golden/` pin hit maps.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .ctypes_defs import (RmRawMaterial, RmRawMesh, RmRawScene, RmRawTexture,
                          RmRenderArgs)

F32 = np.float32


# --------------------------------------------------------------------------- containers
@dataclass
class RawMaterial:
    tex_diffuse: int = -1
    tex_specular: int = -1
    tex_emissive: int = -1
    tex_normals: int = -1
    opacity: float = 1.0      # Material() defaults, src/material.cpp:109-111
    ior: float = 1.0
    roughness: float = 0.8
    transmitting_color: Tuple[float, float, float] = (0.0, 0.0, 0.0)

    @staticmethod
    def glass(color=(1.0, 1.0, 1.0), roughness=5e-3):
        """What loadMaterialProperties sets for an aiMaterial with opacity < 0.99
        (src/material.cpp:306-328)."""
        return RawMaterial(opacity=0.0, ior=1.25, roughness=roughness, transmitting_color=tuple(color))


@dataclass
class RawScene:
    positions: np.ndarray                      # (n,3,3) f32
    uvs: np.ndarray                            # (n,3,2) f32
    normals: np.ndarray                        # (n,3,3) f32
    meshes: List[Tuple[int, int, int]]         # (face_begin, face_end, material)
    materials: List[RawMaterial]
    textures: List[np.ndarray]                 # (h,w,c) uint8, c in {3,4}
    sky: Optional[np.ndarray] = None           # (h,w,3) f32
    name: str = "scene"
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_faces(self) -> int:
        return int(self.positions.shape[0])

    def save(self, path: str) -> None:
        """Write the `.rmscene` container the C++ host's Model reads (raym0nade_b200/host/model.cpp, loadRmScene):
        the arrays of an RmRawScene, little-endian, in struct order."""
        with open(path, "wb") as f:
            sky = None if self.sky is None else np.ascontiguousarray(self.sky, dtype="<f4")
            f.write(b"RMSCENE1")
            f.write(np.array([self.n_faces, len(self.meshes), len(self.materials), len(self.textures),
                              0 if sky is None else sky.shape[1], 0 if sky is None else sky.shape[0]], "<i4").tobytes())
            for a in (self.positions, self.uvs, self.normals):
                f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())
            f.write(np.array(self.meshes, "<i4").reshape(-1, 3).tobytes())
            for m in self.materials:
                f.write(np.array([m.tex_diffuse, m.tex_specular, m.tex_emissive, m.tex_normals], "<i4").tobytes())
                f.write(np.array([m.opacity, m.ior, m.roughness, *m.transmitting_color], "<f4").tobytes())
            for t in self.textures:
                t = np.ascontiguousarray(t, dtype=np.uint8)
                f.write(np.array([t.shape[1], t.shape[0], t.shape[2]], "<i4").tobytes())
                f.write(t.tobytes())
            if sky is not None:
                f.write(sky.tobytes())

    def to_c(self) -> RmRawScene:
        """Build the C view (include/rm_types.h).  Arrays stay owned by this object."""
        pos = np.ascontiguousarray(self.positions, dtype=F32)
        uvs = np.ascontiguousarray(self.uvs, dtype=F32)
        nrm = np.ascontiguousarray(self.normals, dtype=F32)
        n_tex = len(self.textures)
        texs = (RmRawTexture * max(n_tex, 1))()
        keep = [pos, uvs, nrm, texs]
        for i, t in enumerate(self.textures):
            t = np.ascontiguousarray(t, dtype=np.uint8)
            keep.append(t)
            texs[i].width, texs[i].height, texs[i].channels = t.shape[1], t.shape[0], t.shape[2]
            texs[i].pixels = t.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        mats = (RmRawMaterial * max(len(self.materials), 1))()
        for i, m in enumerate(self.materials):
            mats[i].tex_diffuse, mats[i].tex_specular = m.tex_diffuse, m.tex_specular
            mats[i].tex_emissive, mats[i].tex_normals = m.tex_emissive, m.tex_normals
            mats[i].opacity, mats[i].ior, mats[i].roughness = m.opacity, m.ior, m.roughness
            for k in range(3):
                mats[i].transmitting_color[k] = m.transmitting_color[k]
        meshes = (RmRawMesh * max(len(self.meshes), 1))()
        for i, (b, e, m) in enumerate(self.meshes):
            meshes[i].face_begin, meshes[i].face_end, meshes[i].material = b, e, m
        c = RmRawScene()
        c.n_faces, c.n_meshes = self.n_faces, len(self.meshes)
        c.n_materials, c.n_textures = len(self.materials), n_tex
        fp = ctypes.POINTER(ctypes.c_float)
        c.positions, c.uvs, c.normals = pos.ctypes.data_as(fp), uvs.ctypes.data_as(fp), nrm.ctypes.data_as(fp)
        c.meshes, c.materials, c.textures = meshes, mats, texs
        if self.sky is not None:
            sky = np.ascontiguousarray(self.sky, dtype=F32)
            keep.append(sky)
            c.sky_width, c.sky_height = sky.shape[1], sky.shape[0]
            c.sky_rgb = sky.ctypes.data_as(fp)
        else:
            c.sky_width = c.sky_height = 0
            c.sky_rgb = None
        keep += [mats, meshes]
        self._keep = keep
        return c


@dataclass
class RenderArgs:
    """Mirror of the reference's RenderArgs (include/render.h:8-15), same field names."""
    position: Tuple[float, float, float]
    direction: Tuple[float, float, float]
    up: Tuple[float, float, float]
    right: Tuple[float, float, float]
    accuracy: float
    focus: float = 0.0
    CoC: float = 0.0
    exposure: float = 1.0
    P_Direct: float = 0.7
    width: int = 512
    height: int = 512
    spp: int = 64
    threads: int = 8
    savePath: str = "output/render"

    @staticmethod
    def from_console(text: str) -> "RenderArgs":
        """Parse the reference console's `create args` prompt order
        (src/myconsole.cpp:27-67; docs/renderArguments.txt): direction / right / up /
        position given as (D,R,U) coefficients / accuracy focus CoC exposure /
        width height / spp threads P_Direct / savePath."""
        tok = text.split()
        if len(tok) < 22:
            raise ValueError("RenderArgs needs 22 whitespace-separated fields, got %d" % len(tok))
        f = [F32(x) for x in tok[:16]]
        d, r, u = np.array(f[0:3], F32), np.array(f[3:6], F32), np.array(f[6:9], F32)
        D, R, U = f[9], f[10], f[11]
        pos = (D * d + R * r) + U * u          # float32, left to right (src/myconsole.cpp:51)
        return RenderArgs(position=tuple(map(float, pos)), direction=tuple(map(float, d)),
                          up=tuple(map(float, u)), right=tuple(map(float, r)),
                          accuracy=float(f[12]), focus=float(f[13]), CoC=float(f[14]), exposure=float(f[15]),
                          width=int(tok[16]), height=int(tok[17]), spp=int(tok[18]), threads=int(tok[19]),
                          P_Direct=float(F32(tok[20])), savePath=tok[21])

    def to_c(self) -> RmRenderArgs:
        c = RmRenderArgs()
        for k in range(3):
            c.position[k], c.direction[k] = self.position[k], self.direction[k]
            c.up[k], c.right[k] = self.up[k], self.right[k]
        c.accuracy, c.focus, c.CoC = self.accuracy, self.focus, self.CoC
        c.exposure, c.P_Direct = self.exposure, self.P_Direct
        c.width, c.height, c.spp = self.width, self.height, self.spp
        return c

    def replace(self, **kw) -> "RenderArgs":
        d = dict(self.__dict__)
        d.update(kw)
        return RenderArgs(**d)


def look_at(position, target, world_up=(0.0, 1.0, 0.0)):
    """direction/right/up in the reference's convention: right = d x world_up,
    up = d x right, i.e. `up` points DOWN the image (src/render.cpp:466-470)."""
    p, t, w = np.array(position, np.float64), np.array(target, np.float64), np.array(world_up, np.float64)
    d = t - p
    d /= np.sqrt((d * d).sum())
    r = np.cross(d, w)
    r /= np.sqrt((r * r).sum())
    u = np.cross(d, r)
    f = lambda v: tuple(float(F32(x)) for x in v)
    return f(p), f(d), f(u), f(r)


def camera(position, target, width, height, hfov_tan, **kw) -> RenderArgs:
    """hfov_tan = tan(half horizontal field of view); accuracy = pixel pitch at film distance 1."""
    p, d, u, r = look_at(position, target)
    return RenderArgs(position=p, direction=d, up=u, right=r,
                      accuracy=float(F32(2.0 * hfov_tan / width)), width=width, height=height, **kw)


# --------------------------------------------------------------------------- noise
def _hash2(ix, iy, seed):
    """Integer lattice hash -> float64 in [0,1).  Pure uint32 arithmetic."""
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B1) + iy.astype(np.uint64) * np.uint64(0x85EBCA77)
         + np.uint64(seed) * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    return h.astype(np.float64) / 4294967296.0


def value_noise(x, y, seed, period=None):
    """Smooth value noise in [0,1); x,y float64 arrays.  period wraps the lattice (tileable)."""
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = x - x0, y - y0
    ix, iy = x0.astype(np.int64), y0.astype(np.int64)
    ix1, iy1 = ix + 1, iy + 1
    if period is not None:
        ix, iy, ix1, iy1 = ix % period, iy % period, ix1 % period, iy1 % period
    off = 1 << 20
    a = _hash2(ix + off, iy + off, seed)
    b = _hash2(ix1 + off, iy + off, seed)
    c = _hash2(ix + off, iy1 + off, seed)
    d = _hash2(ix1 + off, iy1 + off, seed)
    sx, sy = fx * fx * (3.0 - 2.0 * fx), fy * fy * (3.0 - 2.0 * fy)
    return (a * (1 - sx) + b * sx) * (1 - sy) + (c * (1 - sx) + d * sx) * sy


def fbm(x, y, seed, octaves=4, period=None):
    amp, tot, s = 0.5, 0.0, 0.0
    for o in range(octaves):
        p = None if period is None else period * (1 << o)
        s = s + amp * value_noise(x * (1 << o), y * (1 << o), seed + 101 * o, p)
        tot += amp
        amp *= 0.5
    return s / tot


# --------------------------------------------------------------------------- textures
def solid_rgba(r, g, b, a=255):
    return np.array([[[r, g, b, a]]], np.uint8)


def noise_albedo(size, seed, base=(180, 150, 120), checker=8, cutout=False):
    """RGBA8 albedo: tileable fbm tint over a checker; optional alpha cut-out discs."""
    v, u = np.meshgrid(np.arange(size, dtype=np.float64), np.arange(size, dtype=np.float64), indexing="ij")
    n = fbm(u * 8.0 / size, v * 8.0 / size, seed, 4, period=8)
    chk = (((u * checker // size).astype(np.int64) + (v * checker // size).astype(np.int64)) & 1).astype(np.float64)
    shade = 0.55 + 0.45 * n
    shade = shade * (0.75 + 0.25 * chk)
    img = np.empty((size, size, 4), np.uint8)
    for k in range(3):
        img[..., k] = np.clip(np.floor(base[k] * shade), 0, 255).astype(np.uint8)
    img[..., 3] = 255
    if cutout:
        cu = (u * 4.0 / size) % 1.0 - 0.5
        cv = (v * 4.0 / size) % 1.0 - 0.5
        img[..., 3] = np.where(cu * cu + cv * cv < 0.09, 0, 255).astype(np.uint8)
    return img


def noise_normal_map(size, seed, strength=2.0):
    """RGB8 tangent-space normal map from a tileable height field (central differences)."""
    v, u = np.meshgrid(np.arange(size, dtype=np.float64), np.arange(size, dtype=np.float64), indexing="ij")
    hgt = lambda uu, vv: fbm(uu * 16.0 / size, vv * 16.0 / size, seed, 3, period=16)
    dx = (hgt(u + 1, v) - hgt(u - 1, v)) * strength * size / 64.0
    dy = (hgt(u, v + 1) - hgt(u, v - 1)) * strength * size / 64.0
    nz = np.ones_like(dx)
    inv = 1.0 / np.sqrt(dx * dx + dy * dy + nz * nz)
    img = np.empty((size, size, 3), np.uint8)
    img[..., 0] = np.clip(np.floor((-dx * inv * 0.5 + 0.5) * 255.0 + 0.5), 0, 255).astype(np.uint8)
    img[..., 1] = np.clip(np.floor((-dy * inv * 0.5 + 0.5) * 255.0 + 0.5), 0, 255).astype(np.uint8)
    img[..., 2] = np.clip(np.floor((nz * inv * 0.5 + 0.5) * 255.0 + 0.5), 0, 255).astype(np.uint8)
    return img


def gradient_sky(width, height, sun_dir=(0.35, 0.8, 0.45), sun_radiance=5e3, sun_cos=0.9985):
    """fp32 equirect sky: vertical gradient + sun disc.  Texel (u,v) -> direction by the
    reference's own mapping (src/sampling.cpp:455-457) but with polynomial sin/cos
    replaced by an exact parametrisation: we only need *a* smooth radiance field, the
    direction->texel lookup is the reference's business."""
    v = (np.arange(height, dtype=np.float64) + 0.5) / height          # 0 = zenith
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    # rational approximation of (cos phi) on [0,1] -> [1,-1], monotone, exact arithmetic
    cy = 1.0 - 2.0 * v
    sy = np.sqrt(np.maximum(0.0, 1.0 - cy * cy))
    # azimuth on the unit circle via the rational parametrisation of the circle
    t = 2.0 * u - 1.0                                                  # [-1,1]
    q = 4.0 * t * (1.0 - np.abs(t))                                    # parabola "sine", C1
    q = 0.225 * (q * np.abs(q) - q) + q                                # refined; |q| <= 1
    t2 = ((u + 0.25) % 1.0) * 2.0 - 1.0
    c = 4.0 * t2 * (1.0 - np.abs(t2))
    c = 0.225 * (c * np.abs(c) - c) + c
    dirx = -sy[:, None] * q[None, :]
    diry = np.broadcast_to(cy[:, None], (height, width))
    dirz = sy[:, None] * c[None, :]
    s = np.array(sun_dir, np.float64)
    s /= np.sqrt((s * s).sum())
    cosang = dirx * s[0] + diry * s[1] + dirz * s[2]
    up = np.clip(diry * 0.5 + 0.5, 0.0, 1.0)
    sky = np.empty((height, width, 3), np.float64)
    sky[..., 0] = 0.25 + 0.35 * (1 - up)
    sky[..., 1] = 0.35 + 0.35 * (1 - up)
    sky[..., 2] = 0.55 + 0.25 * (1 - up)
    sky *= np.where(diry < 0.0, 0.35, 1.0)[..., None]
    sun = np.where(cosang > sun_cos, sun_radiance, 0.0)
    sky[..., 0] += sun
    sky[..., 1] += sun * 0.95
    sky[..., 2] += sun * 0.85
    return sky.astype(F32)


# --------------------------------------------------------------------------- mesh builder
class SceneBuilder:
    def __init__(self, name):
        self.name = name
        self.pos, self.uv, self.nrm = [], [], []
        self.meshes: List[Tuple[int, int, int]] = []
        self.materials: List[RawMaterial] = []
        self.textures: List[np.ndarray] = []
        self.n = 0

    def texture(self, img) -> int:
        self.textures.append(np.ascontiguousarray(img, np.uint8))
        return len(self.textures) - 1

    def material(self, m: RawMaterial) -> int:
        self.materials.append(m)
        return len(self.materials) - 1

    def diffuse(self, r, g, b, **kw) -> int:
        return self.material(RawMaterial(tex_diffuse=self.texture(solid_rgba(r, g, b)), **kw))

    def emissive(self, value=255) -> int:
        """Emission exists only through an emissive texture (src/material.cpp:365-372)."""
        return self.material(RawMaterial(tex_diffuse=self.texture(solid_rgba(255, 255, 255)),
                                         tex_emissive=self.texture(solid_rgba(value, value, value))))

    def glossy(self, r, g, b, roughness, metallic) -> int:
        """SPECULAR texture: G = roughness, B = metallic (src/material.cpp:374-383)."""
        spec = solid_rgba(0, int(round(roughness * 255)), int(round(metallic * 255)))
        return self.material(RawMaterial(tex_diffuse=self.texture(solid_rgba(r, g, b)), tex_specular=self.texture(spec)))

    def add(self, pos, uv, nrm, material):
        pos = np.asarray(pos, np.float64).reshape(-1, 3, 3)
        n = pos.shape[0]
        if nrm is None:
            e1, e2 = pos[:, 1] - pos[:, 0], pos[:, 2] - pos[:, 0]
            fn = np.cross(e1, e2)
            ln = np.sqrt((fn * fn).sum(-1, keepdims=True))
            fn = fn / np.where(ln > 0, ln, 1.0)
            nrm = np.repeat(fn[:, None, :], 3, axis=1)
        if uv is None:
            uv = np.broadcast_to(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), (n, 3, 2))
        self.pos.append(pos.astype(F32))
        self.uv.append(np.asarray(uv, np.float64).reshape(-1, 3, 2).astype(F32))
        self.nrm.append(np.asarray(nrm, np.float64).reshape(-1, 3, 3).astype(F32))
        self.meshes.append((self.n, self.n + n, material))
        self.n += n

    def quad(self, p00, p10, p11, p01, material, uv_scale=1.0):
        p = np.array([p00, p10, p11, p01], np.float64)
        pos = np.array([[p[0], p[1], p[2]], [p[0], p[2], p[3]]])
        t = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64) * uv_scale
        uv = np.array([[t[0], t[1], t[2]], [t[0], t[2], t[3]]])
        self.add(pos, uv, None, material)

    def box(self, lo, hi, material, skip_bottom=False, rot=None):
        lo, hi = np.array(lo, np.float64), np.array(hi, np.float64)
        c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                      [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]])
        if rot is not None:   # rotation about the vertical axis through the box centre; rot = (cos, sin) exact pair
            ctr = (lo + hi) / 2
            d = c - ctr
            cs, sn = rot
            c = np.stack([d[:, 0] * cs + d[:, 2] * sn, d[:, 1], -d[:, 0] * sn + d[:, 2] * cs], 1) + ctr
        faces = [(4, 5, 6, 7), (1, 0, 3, 2), (5, 1, 2, 6), (0, 4, 7, 3), (7, 6, 2, 3)]
        if not skip_bottom:
            faces.append((0, 1, 5, 4))
        pos, uv = [], []
        t = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64)
        for a, b, cc, d in faces:
            pos += [[c[a], c[b], c[cc]], [c[a], c[cc], c[d]]]
            uv += [[t[0], t[1], t[2]], [t[0], t[2], t[3]]]
        self.add(np.array(pos), np.array(uv), None, material)

    def heightfield(self, nx, nz, x0, x1, z0, z1, height_fn, material, uv_tiles=8.0):
        """(nx x nz) quads -> 2*nx*nz triangles with smooth vertex normals."""
        gx = x0 + (x1 - x0) * np.arange(nx + 1, dtype=np.float64) / nx
        gz = z0 + (z1 - z0) * np.arange(nz + 1, dtype=np.float64) / nz
        X, Z = np.meshgrid(gx, gz, indexing="ij")
        Y = height_fn(X, Z)
        P = np.stack([X, Y, Z], -1)
        dx, dz = (x1 - x0) / nx, (z1 - z0) / nz
        Yx = np.empty_like(Y)
        Yx[1:-1] = (Y[2:] - Y[:-2]) / (2 * dx)
        Yx[0], Yx[-1] = (Y[1] - Y[0]) / dx, (Y[-1] - Y[-2]) / dx
        Yz = np.empty_like(Y)
        Yz[:, 1:-1] = (Y[:, 2:] - Y[:, :-2]) / (2 * dz)
        Yz[:, 0], Yz[:, -1] = (Y[:, 1] - Y[:, 0]) / dz, (Y[:, -1] - Y[:, -2]) / dz
        N = np.stack([-Yx, np.ones_like(Y), -Yz], -1)
        N /= np.sqrt((N * N).sum(-1, keepdims=True))
        UV = np.stack([np.broadcast_to((np.arange(nx + 1) * (uv_tiles / nx))[:, None], Y.shape),
                       np.broadcast_to((np.arange(nz + 1) * (uv_tiles / nz))[None, :], Y.shape)], -1)
        self._grid(P, N, UV, material, flip=True)

    def _grid(self, P, N, UV, material, flip=False):
        a, b, c, d = (P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:])
        na, nb, nc, nd = (N[:-1, :-1], N[1:, :-1], N[1:, 1:], N[:-1, 1:])
        ta, tb, tc, td = (UV[:-1, :-1], UV[1:, :-1], UV[1:, 1:], UV[:-1, 1:])
        if flip:
            b, d, nb, nd, tb, td = d, b, nd, nb, td, tb
        pos = np.stack([np.stack([a, b, c], -2), np.stack([a, c, d], -2)], 2).reshape(-1, 3, 3)
        nrm = np.stack([np.stack([na, nb, nc], -2), np.stack([na, nc, nd], -2)], 2).reshape(-1, 3, 3)
        uv = np.stack([np.stack([ta, tb, tc], -2), np.stack([ta, tc, td], -2)], 2).reshape(-1, 3, 2)
        self.add(pos, uv, nrm, material)

    def cube_sphere(self, center, radius, res, material, displace=None, uv_tiles=2.0, stretch=(1.0, 1.0, 1.0)):
        """Normalised-cube sphere: 6 faces x res x res quads -> 12*res^2 triangles.
        displace(dir (..,3)) -> radial scale.  Normals by finite differences of the
        displaced surface (exact arithmetic only)."""
        center = np.array(center, np.float64)
        st = np.array(stretch, np.float64)
        g = -1.0 + 2.0 * np.arange(res + 1, dtype=np.float64) / res
        A, B = np.meshgrid(g, g, indexing="ij")
        # the six cube faces as maps (a,b) -> point on the cube, and whether to flip winding
        faces = [(lambda a, b: (np.ones_like(a), a, b), False), (lambda a, b: (-np.ones_like(a), a, b), True),
                 (lambda a, b: (a, np.ones_like(a), b), True), (lambda a, b: (a, -np.ones_like(a), b), False),
                 (lambda a, b: (a, b, np.ones_like(a)), False), (lambda a, b: (a, b, -np.ones_like(a)), True)]

        def surf(a, b, fmap):
            D = np.stack(fmap(a, b), -1)
            D = D / np.sqrt((D * D).sum(-1, keepdims=True))
            r = radius * (displace(D) if displace is not None else 1.0)
            return D, D * np.asarray(r)[..., None] * st

        eps = 1e-3
        for fmap, flip in faces:
            Dn, P = surf(A, B, fmap)
            _, Pu = surf(A + eps, B, fmap)
            _, Pv = surf(A, B + eps, fmap)
            Nn = np.cross(Pu - P, Pv - P)
            Nn /= np.sqrt((Nn * Nn).sum(-1, keepdims=True))
            sgn = np.where((Nn * Dn).sum(-1, keepdims=True) < 0, -1.0, 1.0)
            Nn = Nn * sgn
            UV = np.stack([(A * 0.5 + 0.5) * uv_tiles, (B * 0.5 + 0.5) * uv_tiles], -1)
            self._grid(P + center, Nn, UV, material, flip=flip)

    def build(self, sky=None) -> RawScene:
        return RawScene(positions=np.concatenate(self.pos), uvs=np.concatenate(self.uv),
                        normals=np.concatenate(self.nrm), meshes=self.meshes, materials=self.materials,
                        textures=self.textures, sky=sky, name=self.name)


# --------------------------------------------------------------------------- config 1
def cornell_box(width=512, height=512, spp=64):
    """Config 1: Cornell box, 5 walls + 2 boxes + 1 emissive quad = 38 triangles,
    diffuse materials as 1x1 RGBA8 textures, no sky.  512x512, spp 64, P_Direct 0.7, exposure 8."""
    b = SceneBuilder("cornell")
    white, red, green = b.diffuse(200, 200, 200), b.diffuse(200, 30, 30), b.diffuse(30, 200, 30)
    light = b.emissive(255)
    b.quad((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1), white)          # floor
    b.quad((-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, 1, -1), white)              # ceiling
    b.quad((-1, -1, -1), (-1, 1, -1), (1, 1, -1), (1, -1, -1), white)          # back
    b.quad((-1, -1, -1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1), red)            # left
    b.quad((1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1), green)              # right
    b.box((-0.65, -1.0, -0.55), (-0.05, 0.2, 0.05), white, skip_bottom=True, rot=(0.96, 0.28))
    b.box((0.1, -1.0, 0.05), (0.7, -0.4, 0.65), white, skip_bottom=True, rot=(0.96, -0.28))
    b.quad((-0.25, 0.995, -0.25), (0.25, 0.995, -0.25), (0.25, 0.995, 0.25), (-0.25, 0.995, 0.25), light)
    scene = b.build()
    args = camera((0.0, 0.0, 3.4), (0.0, 0.0, 0.0), width, height, hfov_tan=0.41,
                  exposure=8.0, P_Direct=0.7, spp=spp)
    return scene, args


def nested_glass(width=96, height=96, spp=16, shells=9):
    """Not a BASELINE config: the Cornell room with `shells` concentric glass spheres (each its own material, so each its
    own entry of the reference's Medium multimap, src/render.cpp:13-42) around a small diffuse core - a ray towards the
    core is inside `shells` nested dielectrics at once."""
    b = SceneBuilder("nested_glass")
    white, red, green = b.diffuse(200, 200, 200), b.diffuse(200, 30, 30), b.diffuse(30, 200, 30)
    light = b.emissive(255)
    b.quad((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1), white)
    b.quad((-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, 1, -1), white)
    b.quad((-1, -1, -1), (-1, 1, -1), (1, 1, -1), (1, -1, -1), white)
    b.quad((-1, -1, -1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1), red)
    b.quad((1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1), green)
    b.quad((-0.25, 0.995, -0.25), (0.25, 0.995, -0.25), (0.25, 0.995, 0.25), (-0.25, 0.995, 0.25), light)
    for k in range(shells):
        tint = (1.0 - 0.01 * (k % 3), 1.0 - 0.012 * ((k + 1) % 3), 1.0 - 0.008 * ((k + 2) % 3))
        m = b.material(RawMaterial(opacity=0.0, ior=1.05 + 0.03 * k, roughness=5e-3, transmitting_color=tint))
        b.cube_sphere((0.0, -0.2, 0.0), 0.75 - 0.06 * k, 10, m)
    b.cube_sphere((0.0, -0.2, 0.0), 0.75 - 0.06 * shells - 0.05, 6, b.diffuse(220, 160, 60))
    scene = b.build()
    args = camera((0.0, 0.0, 3.4), (0.0, 0.0, 0.0), width, height, hfov_tan=0.41, exposure=8.0, P_Direct=0.7, spp=spp)
    return scene, args


# --------------------------------------------------------------------------- config 2
def _terrain(seed, amp=1.0, freq=0.35):
    return lambda X, Z: amp * (fbm(X * freq + 37.0, Z * freq + 11.0, seed, 5) - 0.5) * 2.0


def sponza_scale(n_tris=260_000, width=1920, height=1080, spp=256, tex_size=1024, sky_size=(2048, 1024), seed=1):
    """Config 2: heightfield courtyard + colonnade of displaced columns and spheres,
    ~n_tris triangles, 8 diffuse materials with tex_size^2 RGBA8 albedo, HDR gradient
    sky with a sun disc.  No emissive meshes (see SURVEY.md section 7 on sky + lights)."""
    b = SceneBuilder("sponza_scale")
    bases = [(190, 170, 140), (150, 120, 100), (120, 140, 110), (200, 200, 190),
             (170, 90, 70), (90, 110, 150), (210, 180, 120), (130, 130, 130)]
    mats = [b.material(RawMaterial(tex_diffuse=b.texture(noise_albedo(tex_size, seed * 100 + i, bases[i]))))
            for i in range(8)]
    # budget: 55 % terrain, 45 % columns/spheres
    g = int(np.sqrt(n_tris * 0.55 / 2))
    b.heightfield(g, g, -12, 12, -12, 12, _terrain(seed, 0.6), mats[0], uv_tiles=12.0)
    n_obj = 24
    res = max(2, int(np.sqrt(n_tris * 0.45 / (12 * n_obj))))
    k = 0
    for i in range(6):
        for j in range(4):
            cx, cz = -7.5 + 3.0 * i, -6.0 + 4.0 * j
            m = mats[1 + (k % 7)]
            if (i + j) & 1:   # column: stretched sphere with fluting
                disp = lambda D, s=seed + k: 1.0 + 0.06 * (fbm(D[..., 0] * 6 + 3, D[..., 2] * 6 + 5, s, 2) - 0.5)
                b.cube_sphere((cx, 1.6, cz), 0.55, res, m, disp, stretch=(1.0, 4.0, 1.0))
            else:
                disp = lambda D, s=seed + k: 1.0 + 0.25 * (fbm(D[..., 0] * 3 + D[..., 1] * 2 + 9, D[..., 2] * 3 + 1, s, 3) - 0.5)
                b.cube_sphere((cx, 1.0, cz), 0.9, res, m, disp)
            k += 1
    sky = gradient_sky(sky_size[0], sky_size[1])
    scene = b.build(sky)
    args = camera((-10.5, 3.2, 9.5), (0.0, 0.8, 0.0), width, height, hfov_tan=0.55,
                  exposure=1.0, P_Direct=0.7, spp=spp)
    return scene, args


# --------------------------------------------------------------------------- config 3
def glossy_dielectric(n_tris=1_000_000, width=1920, height=1080, spp=1024, seed=3):
    """Config 3: displaced glossy / glass blobs on a ground heightfield, ~n_tris triangles
    (all far below the 1e-2 area threshold so smooth normals engage, src/model.cpp:249-252),
    lit by emissive quads.  Materials: glossy via 1x1 SPECULAR textures, glass via
    opacity 0 / ior 1.25 / roughness 5e-3 (src/material.cpp:306-328)."""
    b = SceneBuilder("glossy_dielectric")
    ground = b.glossy(170, 170, 175, 0.4, 0.0)
    gl = [b.glossy(220, 180, 60, 0.05, 0.99), b.glossy(200, 60, 50, 0.15, 0.0), b.glossy(60, 90, 200, 0.25, 0.0),
          b.glossy(230, 230, 230, 0.08, 0.99), b.glossy(70, 180, 90, 0.4, 0.0)]
    glass = [b.material(RawMaterial.glass((1.0, 1.0, 1.0))), b.material(RawMaterial.glass((0.8, 0.7, 0.55)))]
    light = b.emissive(255)
    g = int(np.sqrt(n_tris * 0.30 / 2))
    b.heightfield(g, g, -8, 8, -8, 8, _terrain(seed, 0.15, 0.5), ground, uv_tiles=8.0)
    n_obj = 16
    res = max(2, int(np.sqrt(n_tris * 0.70 / (12 * n_obj))))
    k = 0
    for i in range(4):
        for j in range(4):
            cx, cz = -4.5 + 3.0 * i, -4.5 + 3.0 * j
            if (i * 4 + j) % 4 == 1:
                m = glass[k % 2]
                disp = None if k % 3 == 0 else (lambda D, s=seed + k: 1.0 + 0.08 * (fbm(D[..., 0] * 2 + 1, D[..., 1] * 2 + D[..., 2] * 2, s, 2) - 0.5))
            else:
                m = gl[k % 5]
                disp = lambda D, s=seed + k: 1.0 + 0.3 * (fbm(D[..., 0] * 3 + D[..., 1] + 7, D[..., 2] * 3 - D[..., 1] + 2, s, 4) - 0.5)
            b.cube_sphere((cx, 1.15, cz), 0.95, res, m, disp)
            k += 1
    for (lx, lz) in [(-3.0, -3.0), (3.0, 3.0), (0.0, 0.0)]:
        b.quad((lx - 0.8, 5.0, lz - 0.8), (lx + 0.8, 5.0, lz - 0.8), (lx + 0.8, 5.0, lz + 0.8), (lx - 0.8, 5.0, lz + 0.8), light)
    scene = b.build()
    args = camera((-9.0, 4.5, 9.0), (0.0, 0.8, 0.0), width, height, hfov_tan=0.5,
                  exposure=2.0, P_Direct=0.7, spp=spp)
    return scene, args


# --------------------------------------------------------------------------- config 4
def texture_heavy(n_tris=500_000, width=3840, height=2160, spp=4096, tex_size=2048, n_materials=32, seed=4):
    """Config 4: ~n_tris triangles, n_materials materials each with tex_size^2 RGBA8 albedo
    (a quarter of them with alpha cut-outs -> hasFullyTransparentPart) + tex_size^2 RGB8
    normal map; emissive quads for light."""
    b = SceneBuilder("texture_heavy")
    mats = []

    def make_pair(i):
        base = (120 + (i * 37) % 120, 110 + (i * 53) % 130, 100 + (i * 71) % 140)
        return (noise_albedo(tex_size, seed * 1000 + i, base, checker=4 + i % 5, cutout=(i % 4 == 3)),
                noise_normal_map(tex_size, seed * 2000 + i))

    # the textures are independent: large ones are generated on a thread pool (numpy releases the GIL), same arrays either way
    if tex_size >= 512:
        import concurrent.futures
        import os
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            pairs = list(pool.map(make_pair, range(n_materials)))
    else:
        pairs = [make_pair(i) for i in range(n_materials)]
    for alb, nrm in pairs:
        d = b.texture(alb)
        n = b.texture(nrm)
        mats.append(b.material(RawMaterial(tex_diffuse=d, tex_normals=n)))
    light = b.emissive(255)
    g = int(np.sqrt(n_tris * 0.4 / 2))
    b.heightfield(g, g, -10, 10, -10, 10, _terrain(seed, 0.3, 0.4), mats[0], uv_tiles=10.0)
    n_obj = n_materials - 1
    res = max(2, int(np.sqrt(n_tris * 0.6 / (12 * n_obj))))
    for k in range(n_obj):
        i, j = k % 6, k // 6
        cx, cz = -7.5 + 3.0 * i, -7.5 + 3.0 * j
        m = mats[1 + k]
        disp = lambda D, s=seed + k: 1.0 + 0.2 * (fbm(D[..., 0] * 2 + 5, D[..., 2] * 2 + D[..., 1], s, 3) - 0.5)
        b.cube_sphere((cx, 1.3, cz), 1.0, res, m, disp, uv_tiles=2.0)
    for (lx, lz) in [(-5.0, -5.0), (5.0, 5.0), (-5.0, 5.0), (5.0, -5.0)]:
        b.quad((lx - 1, 6.0, lz - 1), (lx + 1, 6.0, lz - 1), (lx + 1, 6.0, lz + 1), (lx - 1, 6.0, lz + 1), light)
    scene = b.build()
    args = camera((-11.0, 5.0, 11.0), (0.0, 0.8, 0.0), width, height, hfov_tan=0.5,
                  exposure=2.0, P_Direct=0.7, spp=spp)
    return scene, args


# --------------------------------------------------------------------------- config 5
def five_million(n_tris=5_000_000, width=3840, height=2160, seed=5):
    """Config 5: fine heightfield + blobs, ~n_tris triangles, for the primary-ray
    {tri_idx, t} exact check and the FXAA pass at 4K."""
    b = SceneBuilder("five_million")
    mats = [b.diffuse(180, 160, 130), b.diffuse(120, 150, 180), b.diffuse(190, 100, 90), b.diffuse(140, 180, 120)]
    light = b.emissive(255)
    g = int(np.sqrt(n_tris * 0.6 / 2))
    b.heightfield(g, g, -12, 12, -12, 12, _terrain(seed, 0.8, 0.3), mats[0], uv_tiles=12.0)
    n_obj = 12
    res = max(2, int(np.sqrt(n_tris * 0.4 / (12 * n_obj))))
    for k in range(n_obj):
        i, j = k % 4, k // 4
        cx, cz = -6.0 + 4.0 * i, -4.0 + 4.0 * j
        disp = lambda D, s=seed + k: 1.0 + 0.35 * (fbm(D[..., 0] * 4 + 3, D[..., 2] * 4 + D[..., 1] * 2, s, 4) - 0.5)
        b.cube_sphere((cx, 1.6, cz), 1.2, res, mats[1 + k % 3], disp)
    b.quad((-2, 8.0, -2), (2, 8.0, -2), (2, 8.0, 2), (-2, 8.0, 2), light)
    scene = b.build()
    args = camera((-12.0, 6.0, 12.0), (0.0, 0.5, 0.0), width, height, hfov_tan=0.5,
                  exposure=2.0, P_Direct=0.7, spp=0)
    return scene, args


def heightfield_scene(n_tris=20_000, width=128, height=72, spp=4, seed=7, with_sky=False):
    """Small generic test scene (not a BASELINE config): terrain + 2 blobs + a light."""
    b = SceneBuilder("heightfield")
    mats = [b.diffuse(180, 160, 130), b.glossy(200, 200, 210, 0.2, 0.5)]
    g = max(2, int(np.sqrt(n_tris * 0.6 / 2)))
    b.heightfield(g, g, -6, 6, -6, 6, _terrain(seed, 0.5, 0.4), mats[0], uv_tiles=6.0)
    res = max(2, int(np.sqrt(n_tris * 0.4 / 24)))
    for k, (cx, cz) in enumerate([(-1.5, 0.0), (1.8, -1.0)]):
        disp = lambda D, s=seed + k: 1.0 + 0.3 * (fbm(D[..., 0] * 3 + 1, D[..., 2] * 3 + D[..., 1], s, 3) - 0.5)
        b.cube_sphere((cx, 1.2, cz), 0.9, res, mats[1], disp)
    sky = None
    if with_sky:
        sky = gradient_sky(256, 128)
    else:
        light = b.emissive(255)
        b.quad((-1, 5.0, -1), (1, 5.0, -1), (1, 5.0, 1), (-1, 5.0, 1), light)
    scene = b.build(sky)
    args = camera((-6.0, 3.5, 6.0), (0.0, 0.6, 0.0), width, height, hfov_tan=0.5, exposure=2.0, spp=spp)
    return scene, args
