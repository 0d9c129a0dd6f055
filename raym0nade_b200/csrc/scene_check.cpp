// scene_check.cpp — rm_scene_validate: structural checks of a post-load scene before any of it reaches the device
// (pure host C++).  The kernels index with what the scene says - node children, leaf face ranges, material and texture
// indices, light face lists - and do not re-check per ray, so an inconsistent scene is refused here, with a message
// that names the offending element, instead of reading out of bounds on the GPU.  The reference has no such step: a
// Model that loaded is trusted (src/model.cpp:172-215).
#include <cmath>
#include <cstdint>
#include <vector>

#include "raym0nade_b200.h"
#include "rm_internal.h"

extern "C" int rm_scene_validate(const RmSceneDesc *sc) {
    if (!sc) return rm_fail(RM_ERR_INVALID, "scene: null descriptor");
    if (sc->n_faces <= 0 || sc->n_nodes < 2 || !sc->nodes || !sc->positions || !sc->uvs || !sc->normals || !sc->face_material)
        return rm_fail(RM_ERR_INVALID, "scene is missing geometry");
    if ((int64_t)sc->n_faces >= ((int64_t)1 << 27)) return rm_fail(RM_ERR_INVALID, "scene: more than 2^27 faces");
    if (sc->n_materials <= 0 || !sc->materials) return rm_fail(RM_ERR_INVALID, "scene: faces need at least one material");
    if (sc->n_textures < 0 || (sc->n_textures > 0 && !sc->textures)) return rm_fail(RM_ERR_INVALID, "scene: texture table missing");
    if (sc->n_lights < 0 || (sc->n_lights > 0 && !sc->lights)) return rm_fail(RM_ERR_INVALID, "scene: light table missing");
    if (sc->sky_width < 0 || sc->sky_height < 0 || (sc->sky_width == 0) != (sc->sky_height == 0))
        return rm_fail(RM_ERR_INVALID, "scene: bad sky size %d x %d", sc->sky_width, sc->sky_height);

    // The tree as BVH::dfs_rayHit walks it (src/bvh.cpp:56-88): node u is a leaf iff faceR != 0 and then owns faces
    // [faceL, faceR); otherwise its children are 2u and 2u+1.  Every node reachable from the root must be inside the
    // array and every leaf range inside the face array; all faces must be owned by some leaf.
    {
        std::vector<int32_t> stack;
        stack.push_back(1);
        int64_t owned = 0;
        while (!stack.empty()) {
            const int32_t u = stack.back();
            stack.pop_back();
            const RmBvhNode &nd = sc->nodes[u];
            if (nd.faceR != 0) {
                if (nd.faceL < 0 || nd.faceL >= nd.faceR || nd.faceR > sc->n_faces)
                    return rm_fail(RM_ERR_INVALID, "BVH node %d: leaf range [%d, %d) is outside the %d faces", u, nd.faceL, nd.faceR, sc->n_faces);
                // a leaf reference carries its face count in 4 bits (dev_trace.cuh leaf_ref); the reference's builder stops
                // splitting at <= 10 faces (src/bvh.cpp:22), so a longer leaf is not a tree it could have produced
                if (nd.faceR - nd.faceL > 15)
                    return rm_fail(RM_ERR_INVALID, "BVH node %d: leaf of %d faces (at most 15 are supported)", u, nd.faceR - nd.faceL);
                owned += nd.faceR - nd.faceL;
            } else {
                if (2 * int64_t(u) + 1 >= sc->n_nodes)
                    return rm_fail(RM_ERR_INVALID, "BVH node %d: inner node whose children lie beyond the %d nodes", u, sc->n_nodes);
                stack.push_back(2 * u + 1);
                stack.push_back(2 * u);
            }
        }
        if (owned < sc->n_faces) return rm_fail(RM_ERR_INVALID, "BVH: leaves own %lld of the %d faces", (long long)owned, sc->n_faces);
    }

    for (int i = 0; i < sc->n_faces; i++) {
        const int mat = sc->face_material[i];
        if (mat < 0 || mat >= sc->n_materials) return rm_fail(RM_ERR_INVALID, "face %d: material index out of range", i);
    }
    for (int i = 0; i < sc->n_materials; i++)
        for (int k = 0; k < 4; k++) {
            const int t = sc->materials[i].tex[k];
            if (t >= sc->n_textures) return rm_fail(RM_ERR_INVALID, "material %d: texture index out of range", i);
            // the fetch strides by its own type - RGBA8 for diffuse / specular / emissive, RGB8 for normals
            // (src/material.cpp:58) - so any other pairing would run past the end of the level
            if (t >= 0 && sc->textures[t].channels != (k == 3 ? 3 : 4))
                return rm_fail(RM_ERR_INVALID, "material %d: texture %d has %d channels in slot %d", i, t, sc->textures[t].channels, k);
        }
    for (int i = 0; i < sc->n_textures; i++) {
        const RmTextureDesc &t = sc->textures[i];
        if (t.map_depth < 1 || t.map_depth > 8 || (t.channels != 3 && t.channels != 4))
            return rm_fail(RM_ERR_INVALID, "texture %d: bad map_depth/channels", i);
        if (t.width <= 0 || t.height <= 0) return rm_fail(RM_ERR_INVALID, "texture %d: bad size %d x %d", i, t.width, t.height);
        // every level the mip selection can reach must have texels: a level with a zero dimension would be fetched modulo 0
        if ((t.width >> (t.map_depth - 1)) == 0 || (t.height >> (t.map_depth - 1)) == 0)
            return rm_fail(RM_ERR_INVALID, "texture %d: %d levels of a %d x %d image (level %d is empty)", i, t.map_depth, t.width, t.height, t.map_depth - 1);
        for (int l = 0; l < t.map_depth; l++)
            if (!t.levels[l]) return rm_fail(RM_ERR_INVALID, "texture %d: level %d is NULL", i, l);
    }
    for (int i = 0; i < sc->n_lights; i++) {
        const RmLightDesc &L = sc->lights[i];
        if (L.n_faces <= 0 || !L.face_positions || !L.face_normals || !L.face_cdf)
            return rm_fail(RM_ERR_INVALID, "light %d: no faces", i);
        // RandomDistribution prefix sums (src/component.cpp:12-18): the sampler scales a uniform draw by the last entry
        if (!(L.face_cdf[L.n_faces - 1] > 0.0f) || !std::isfinite(L.face_cdf[L.n_faces - 1]))
            return rm_fail(RM_ERR_INVALID, "light %d: face distribution sums to %g", i, double(L.face_cdf[L.n_faces - 1]));
    }
    if (sc->sky_width > 0) {
        if (!sc->sky_data || !sc->sky_cdf) return rm_fail(RM_ERR_INVALID, "sky size set but sky_data/sky_cdf missing");
        const float tot = sc->sky_cdf[size_t(sc->sky_width) * sc->sky_height - 1];
        if (!std::isfinite(tot)) return rm_fail(RM_ERR_INVALID, "sky: distribution sums to %g", double(tot));
    }
    return RM_OK;
}
