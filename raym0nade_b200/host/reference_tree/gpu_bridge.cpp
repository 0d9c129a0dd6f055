// gpu_bridge.cpp — the one translation unit a maintainer adds to the Raym0nade tree (src/gpu_bridge.cpp) to run the
// render on a B200: it views a loaded `Model` as the plain-C `RmSceneDesc` of include/raym0nade_b200.h and provides
// `render_multiThread_b200`, a drop-in body for `render_multiThread` (src/render.cpp:593-676).  It is written against the
// reference's own headers and compiled in this repository against the unmodified reference tree (`make -C oracle bridge`);
// tests/test_cpu_host.py checks that the descriptor it fills from a reference-built Model equals, bit for bit, the one
// rm_prepare_scene derives from the same raw scene.
//
// Two private members are read: BVH::node (include/bvh.h:13) and RandomDistribution::prefixSums (include/component.h:38).
// In the reference tree add `friend struct B200Bridge;` to `BVH` and `RandomDistribution`; this repository leaves the
// reference untouched and compiles this file with -fno-access-control instead.
#include <algorithm>
#include <iostream>

#include "gpu_bridge.h"

int nodeCount(int u, int n);                                  // src/bvh.cpp:44-46 (external linkage): size of BVH::node minus one

static_assert(sizeof(BVH_Node) == sizeof(RmBvhNode), "BVH_Node is handed over as it is");
static_assert(sizeof(vec3) == 3 * sizeof(float), "SkyBox::data is handed over as packed floats");

const RmSceneDesc *B200Bridge::fill(const Model &m) {
    const int n = int(m.faces.size());
    positions.clear(); uvs.clear(); normals.clear(); faceMaterial.clear();
    materials.clear(); textures.clear(); lights.clear(); lightArrays.clear(); texelStorage.clear();
    positions.reserve(size_t(n) * 9); uvs.reserve(size_t(n) * 6); normals.reserve(size_t(n) * 9);
    for (const Face &f : m.faces) {                       // already in post-build order: BVH::build permuted Model::faces (bvh.cpp:38)
        for (int c = 0; c < 3; c++) {
            positions.insert(positions.end(), {f.v[c].x, f.v[c].y, f.v[c].z});
            uvs.insert(uvs.end(), {f.data[c]->uv.x, f.data[c]->uv.y});
            normals.insert(normals.end(), {f.data[c]->normal.x, f.data[c]->normal.y, f.data[c]->normal.z});
        }
        faceMaterial.push_back(f.material->id);
    }
    // the four texture slots the path reads, in the order of RmMaterialDesc::tex (src/material.cpp:349-383)
    const int slots[4] = {aiTextureType_DIFFUSE, aiTextureType_SPECULAR, aiTextureType_EMISSIVE, aiTextureType_NORMALS};
    for (const Material &M : m.materials) {
        RmMaterialDesc d{};
        for (int k = 0; k < 4; k++) {
            const ImageData &img = M.texture[slots[k]];
            d.tex[k] = -1;
            if (img.empty()) continue;
            RmTextureDesc t{};
            t.width = img.width; t.height = img.height;
            t.channels = k == 3 ? 3 : 4;                  // the stride of the slot's fetch type, not ImageData::channels (src/material.cpp:58)
            t.map_depth = img.map_depth;
            // The slot's fetch strides by its own type (vec4 for diffuse / specular / emissive, vec3 for normals,
            // src/material.cpp:48-79) whatever ImageData::channels says; a PNG is stored as packed RGB8
            // (src/material.cpp:273-281), so a PNG in an RGBA slot is SHORTER than the w*h*4 bytes the fetch (and
            // rm_scene_upload) walks: the reference reads past the end of its vector there.  Such a level is handed over
            // as a bridge-owned copy of the claimed size - the bytes the reference owns verbatim (so every in-bounds
            // fetch returns what the reference's returns), zeros where the reference would over-read.
            for (int l = 0; l < img.map_depth; l++) {
                const size_t need = size_t(img.width >> l) * size_t(img.height >> l) * size_t(t.channels);
                if (img.data[l].size() >= need) { t.levels[l] = img.data[l].data(); continue; }
                texelStorage.emplace_back(need, uint8_t(0));
                std::copy(img.data[l].begin(), img.data[l].end(), texelStorage.back().begin());
                t.levels[l] = texelStorage.back().data();
            }
            d.tex[k] = int32_t(textures.size());
            textures.push_back(t);
        }
        d.opacity = M.opacity; d.ior = M.ior; d.roughness = M.roughness;
        for (int k = 0; k < 3; k++) d.transmitting_color[k] = M.transmittingColor[k];
        d.has_fully_transparent_part = M.hasFullyTransparentPart ? 1 : 0;
        materials.push_back(d);
    }
    lightArrays.reserve(m.lightObjects.size() * 2);
    for (const LightObject &L : m.lightObjects) {
        std::vector<float> lp, ln;
        for (const Face &f : L.faces)
            for (int c = 0; c < 3; c++) {
                lp.insert(lp.end(), {f.v[c].x, f.v[c].y, f.v[c].z});
                ln.insert(ln.end(), {f.data[c]->normal.x, f.data[c]->normal.y, f.data[c]->normal.z});
            }
        lightArrays.push_back(std::move(lp));
        const float *facePositions = lightArrays.back().data();
        lightArrays.push_back(std::move(ln));
        const float *faceNormals = lightArrays.back().data();
        RmLightDesc d{};
        for (int k = 0; k < 3; k++) { d.center[k] = L.center[k]; d.color[k] = L.color[k]; }
        d.power = L.power;
        d.n_faces = int32_t(L.faces.size());
        d.face_positions = facePositions; d.face_normals = faceNormals;
        d.face_cdf = L.faceDist.prefixSums.data();        // the running fp32 sums themselves (src/component.cpp:12-18)
        lights.push_back(d);
    }
    desc = RmSceneDesc{};
    desc.n_faces = n;
    desc.n_nodes = nodeCount(1, n) + 1;                   // src/bvh.cpp:49
    desc.nodes = reinterpret_cast<const RmBvhNode *>(m.bvh.node);
    desc.positions = positions.data(); desc.uvs = uvs.data(); desc.normals = normals.data(); desc.face_material = faceMaterial.data();
    desc.n_materials = int32_t(materials.size()); desc.materials = materials.data();
    desc.n_textures = int32_t(textures.size()); desc.textures = textures.data();
    desc.n_lights = int32_t(lights.size()); desc.lights = lights.data();
    if (!m.skyMap.empty()) {                              // SkyBox::Init has already premultiplied the texels (src/component.cpp:54-67)
        desc.sky_width = m.skyMap.width; desc.sky_height = m.skyMap.height;
        desc.sky_data = &m.skyMap.data[0].x;
        desc.sky_cdf = m.skyMap.dist.prefixSums.data();
    }
    return &desc;
}

// Drop-in body for render_multiThread: everything up to the filled Photo happens on the GPU; the exports that follow
// (src/render.cpp:635-676) are the reference's own code and stay as they are.
void render_multiThread_b200(Model &model, const RenderArgs &args, Photo &photo) {
    static RmContext *ctx = nullptr;
    if (!ctx && rm_context_create(0, nullptr, &ctx) != RM_OK) { std::cerr << rm_last_error() << std::endl; return; }
    B200Bridge bridge;
    if (rm_scene_upload(ctx, bridge.fill(model)) != RM_OK) { std::cerr << rm_last_error() << std::endl; return; }
    RmRenderArgs a{};                                         // RenderArgs field for field (include/render.h:8-15)
    for (int k = 0; k < 3; k++) {
        a.position[k] = args.position[k]; a.direction[k] = args.direction[k];
        a.up[k] = args.up[k]; a.right[k] = args.right[k];
    }
    a.accuracy = args.accuracy; a.focus = args.focus; a.CoC = args.CoC; a.exposure = args.exposure;
    a.P_Direct = args.P_Direct; a.width = args.width; a.height = args.height; a.spp = args.spp;
    static_assert(sizeof(HitInfo) == sizeof(RmHitInfo) && sizeof(RadianceData) == sizeof(RmRadiance), "Photo buffers are written in place");
    photo.exposure = args.exposure;
    if (rm_render(ctx, &a, /*seed*/ 0, reinterpret_cast<RmHitInfo *>(photo.Gbuffer),
                  reinterpret_cast<RmRadiance *>(photo.radiance_Dd), reinterpret_cast<RmRadiance *>(photo.radiance_Ds),
                  reinterpret_cast<RmRadiance *>(photo.radiance_Id), reinterpret_cast<RmRadiance *>(photo.radiance_Is)) != RM_OK)
        std::cerr << rm_last_error() << std::endl;
}
