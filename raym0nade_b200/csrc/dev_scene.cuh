// dev_scene.cuh — the scene as it lies in HBM (see DESIGN.md "Data layout").
//
// Everything is flattened from the post-load RmSceneDesc by rm_scene_upload:
//   nodes    2 x float4 per BVH_Node, heap-indexed exactly like BVH::node (src/bvh.cpp:18-54):
//            the two children of u are nodes 2u, 2u+1 = one 64-byte aligned 64-byte block.
//            float4 a = {v0.x v0.y v0.z v1.x}, float4 b = {v1.y v1.z faceL faceR} (ints as bits).
//   tri      3 x float4 per face for traversal only: {v0.xyz e1.x} {e1.yz e2.xy} {e2.z |e1| cutout 0}
//            where e1 = v1-v0, e2 = v2-v0, |e1| = length(e1): the fp32 values
//            RayTriangleIntersection recomputes per test (src/geometry.cpp:65-70), hoisted.
//   shade    7 x float4 per face for shading: positions[9] uv[6] normals[9] material(int) pad[3]
//   texels   all mip levels of all textures in one byte blob, 16-byte aligned per level
#pragma once
#include <cstdint>
#include "dev_math.cuh"

namespace rm {

// float4 per face in the traversal stream `tri`: 3 used; 4 = each record padded to one aligned 64-byte block so its
// first 32 bytes come with a single 256-bit load (the L1 data pipe is bound by per-lane load count, not bytes)
#ifndef RM_TRI_STRIDE
#define RM_TRI_STRIDE 3
#endif
constexpr int kTriStride = RM_TRI_STRIDE;

constexpr int kSkyGuide = 4096;     // power of two: u * kSkyGuide is exact in fp32

struct DevTexture {
    int32_t width, height, channels, map_depth;
    uint32_t offset[8];             // byte offset of each level in the texel blob
};

struct DevMaterial {
    int32_t tex[4];                 // diffuse, specular, emissive, normals; -1 = empty
    float opacity, ior, roughness;
    float tc[3];                    // transmittingColor
    int32_t cutout;                 // hasFullyTransparentPart
    int32_t _pad;
};

struct DevLight {
    float center[3], color[3];
    float power;
    int32_t n_faces;
    int32_t face_offset;            // into light_pos / light_nrm (faces), light_cdf (floats)
    int32_t _pad[3];
};

struct DevScene {
    const float4 *nodes;
    const float4 *tri;
    const float4 *shade;
    const DevMaterial *materials;
    const DevTexture *textures;
    const uint8_t *texels;
    const DevLight *lights;
    const float *light_pos;         // [total light faces][9]
    const float *light_nrm;         // [total light faces][9]
    const float *light_cdf;         // [total light faces]
    const float *sky_data;          // [h*w][3], premultiplied by texel solid angle
    const float *sky_cdf;           // [h*w]
    const int32_t *sky_guide;       // [kSkyGuide + 1] lower_bound(sky_cdf, total * j / kSkyGuide): brackets of the CDF search
    const float *div255;            // [256] k / 255.0f, correctly rounded (RGBA8 decode without a division per channel)
    int32_t n_faces, n_nodes, n_materials, n_lights;
    int32_t sky_width, sky_height;
    int32_t any_cutout;             // some material has hasFullyTransparentPart
    int32_t root_is_leaf;
    // Secondary-ray tree (fast_bvh.cpp): same record layout, but an inner record names the block of its children in faceL
    // (explicit_children) and triangle indices are in the builder's order (face_map -> index into `shade` / the reference order)
    int32_t explicit_children;
    const int32_t *face_map;
    // 4-wide form of the secondary-ray tree (wide_bvh.cpp): `nodes` then holds 64-byte RmWideNode records, record 0 = the root
    int32_t wide;
};

// Camera / render arguments in device form (RenderArgs, include/render.h:8-15).
struct DevArgs {
    V3 position, direction, up, right;
    float accuracy, exposure, P_Direct;
    int32_t width, height, spp;
};

} // namespace rm
