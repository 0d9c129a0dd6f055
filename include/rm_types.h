/* rm_types.h — plain-C data formats shared by the product library, the host-side
 * scene preparation, the oracle and the tests.  Data layout only: no algorithm lives
 * here, so the oracle and the product stay independent implementations.
 *
 * Every struct mirrors a reference type; the citation names the reference file:line
 * (paths relative to the lemonchu/Raym0nade tree).
 */
#ifndef RM_TYPES_H
#define RM_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- raw (pre-load) scene: what assimp + the asset decoders hand to Model::Model ----
 * The reference builds this in Model::processMesh / processMaterial (src/model.cpp:88-168).
 * Synthetic scenes are generated directly in this form. */

/* ImageData level 0 (include/material.h:12-25).  channels is 4 (RGBA8: diffuse, specular,
 * emissive) or 3 (RGB8: normal maps); the reference strides by the fetch type, not by
 * `channels` (src/material.cpp:58), so only these two pairings are meaningful. */
typedef struct RmRawTexture {
    int32_t width, height, channels;
    int32_t _pad;
    const uint8_t *pixels; /* row-major, width*height*channels bytes */
} RmRawTexture;

/* Material (include/material.h:30-51).  Texture slots follow the four aiTextureType
 * values the hot path reads (src/material.cpp:349-383); -1 = empty slot.
 * hasFullyTransparentPart is derived from the diffuse texture's alpha by the loader
 * (src/material.cpp:330-333), it is not an input. */
typedef struct RmRawMaterial {
    int32_t tex_diffuse, tex_specular, tex_emissive, tex_normals;
    float opacity, ior, roughness;
    float transmitting_color[3];
} RmRawMaterial;

/* One aiMesh: a contiguous face range with one material (src/model.cpp:88-123).
 * Light objects are formed per mesh (checkLightObject, src/model.cpp:44-82). */
typedef struct RmRawMesh {
    int32_t face_begin, face_end, material;
} RmRawMesh;

typedef struct RmRawScene {
    int32_t n_faces;
    int32_t n_meshes;
    int32_t n_materials;
    int32_t n_textures;
    const float *positions;   /* [n_faces][3 corners][xyz]   Face::v      (include/component.h:15-17) */
    const float *uvs;         /* [n_faces][3 corners][uv]    VertexData   (include/component.h:8-13)  */
    const float *normals;     /* [n_faces][3 corners][xyz]   VertexData::normal                       */
    const RmRawMesh *meshes;
    const RmRawMaterial *materials;
    const RmRawTexture *textures;
    int32_t sky_width, sky_height; /* 0,0 = "null" sky (src/model.cpp:178,208)            */
    const float *sky_rgb;          /* [sky_height][sky_width][rgb] radiance, before SkyBox::Init */
} RmRawScene;

/* ---- render arguments: RenderArgs (include/render.h:8-15) minus threads/savePath ---- */
typedef struct RmRenderArgs {
    float position[3], direction[3], up[3], right[3];
    float accuracy, focus, CoC, exposure, P_Direct;
    int32_t width, height, spp;
} RmRenderArgs;

/* ---- per-pixel outputs in the reference's own AoS layouts ---- */

/* HitInfo (include/model.h:12-18), 88 bytes, field for field. */
typedef struct RmHitInfo {
    float shapeNormal[3], surfaceNormal[3], emission[3], baseColor[3], position[3];
    float specular, roughness, metallic, opacity, eta;
    int32_t id;
    uint8_t entering;
    uint8_t _pad[3];
} RmHitInfo;

/* RadianceData (include/image.h:9-13), 16 bytes. */
typedef struct RmRadiance {
    float radiance[3];
    float Var;
} RmRadiance;

/* BVH_Node (include/bvh.h:7-11), 32 bytes; heap-indexed (children of u are 2u, 2u+1),
 * leaf iff faceR != 0 (src/bvh.cpp:57). */
typedef struct RmBvhNode {
    float v0[3], v1[3];
    int32_t faceL, faceR;
} RmBvhNode;

#ifdef __cplusplus
}
#endif
#endif /* RM_TYPES_H */
