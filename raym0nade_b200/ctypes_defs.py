"""ctypes images of the plain-C structs in include/rm_types.h (data layout only)."""
import ctypes as C


class RmRawTexture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32), ("_pad", C.c_int32),
                ("pixels", C.POINTER(C.c_uint8))]


class RmRawMaterial(C.Structure):
    _fields_ = [("tex_diffuse", C.c_int32), ("tex_specular", C.c_int32), ("tex_emissive", C.c_int32),
                ("tex_normals", C.c_int32), ("opacity", C.c_float), ("ior", C.c_float), ("roughness", C.c_float),
                ("transmitting_color", C.c_float * 3)]


class RmRawMesh(C.Structure):
    _fields_ = [("face_begin", C.c_int32), ("face_end", C.c_int32), ("material", C.c_int32)]


class RmRawScene(C.Structure):
    _fields_ = [("n_faces", C.c_int32), ("n_meshes", C.c_int32), ("n_materials", C.c_int32), ("n_textures", C.c_int32),
                ("positions", C.POINTER(C.c_float)), ("uvs", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
                ("meshes", C.POINTER(RmRawMesh)), ("materials", C.POINTER(RmRawMaterial)),
                ("textures", C.POINTER(RmRawTexture)),
                ("sky_width", C.c_int32), ("sky_height", C.c_int32), ("sky_rgb", C.POINTER(C.c_float))]


class RmRenderArgs(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("direction", C.c_float * 3), ("up", C.c_float * 3),
                ("right", C.c_float * 3), ("accuracy", C.c_float), ("focus", C.c_float), ("CoC", C.c_float),
                ("exposure", C.c_float), ("P_Direct", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("spp", C.c_int32)]


import numpy as np

HITINFO_DTYPE = np.dtype([("shapeNormal", "<f4", 3), ("surfaceNormal", "<f4", 3), ("emission", "<f4", 3),
                          ("baseColor", "<f4", 3), ("position", "<f4", 3), ("specular", "<f4"), ("roughness", "<f4"),
                          ("metallic", "<f4"), ("opacity", "<f4"), ("eta", "<f4"), ("id", "<i4"), ("entering", "u1"),
                          ("_pad", "u1", 3)])
RADIANCE_DTYPE = np.dtype([("radiance", "<f4", 3), ("Var", "<f4")])
BVHNODE_DTYPE = np.dtype([("v0", "<f4", 3), ("v1", "<f4", 3), ("faceL", "<i4"), ("faceR", "<i4")])
assert HITINFO_DTYPE.itemsize == 88 and RADIANCE_DTYPE.itemsize == 16 and BVHNODE_DTYPE.itemsize == 32
