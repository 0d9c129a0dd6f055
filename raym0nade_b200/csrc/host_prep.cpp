// host_prep.cpp — host-side scene preparation (pure C++, no CUDA).
//
// The north star keeps model loading and the BVH build on the host.  The reference does
// these derivations inside Model::Model (src/model.cpp:172-215); this file performs the
// same ones on a raw scene so that the staged data is what a loaded reference Model
// would hold:
//   build_bvh        BVH::build / dfs_build / nodeCount        src/bvh.cpp:18-54
//   build_mips       ImageData::generateMipmaps                 src/material.cpp:113-148
//   build_lights     Model::checkLightObject (+ averaging)      src/model.cpp:23-82
//   init_sky         SkyBox::Init                               src/component.cpp:54-67
// Float arithmetic follows the reference's operation order (fp32, no contraction) because
// the tree topology, the CDFs and the light powers feed bit-exact comparisons.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "raym0nade_b200.h"
#include "rm_internal.h"

namespace {

struct V3 { float x, y, z; };
inline V3 ld3(const float *p) { return {p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
// glm 1.0.0 spells `vec3 / scalar` as `v *= 1/scalar` (lib/glm/glm/detail/type_vec3.inl:582-585) but
// `vec3 /= scalar` as a true per-component division (type_vec3.inl:290-293).  Both occur in the
// reference and differ in the last bit, so they are two different functions here.
inline V3 div_recip(V3 a, float s) { float r = 1.0f / s; return {a.x * r, a.y * r, a.z * r}; }
inline V3 div_true(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // (x+y)+z like glm
inline V3 cross3(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float len3(V3 a) { return std::sqrt(dot3(a, a)); }
constexpr float kEps = 1e-4f;                 // eps_zero, include/geometry.h:13
const V3 kLum = {0.3f, 0.6f, 0.1f};           // RGB_Weight, include/geometry.h:16
const float kPi = 3.14159265358979323846f;

// ------------------------------------------------------------------ BVH build
struct BvhBuilder {
    const float *pos;                 // raw positions [n][9]
    std::vector<V3> center;           // Face::center(), src/component.cpp:37-39
    std::vector<int32_t> order;       // permutation being built
    std::vector<RmBvhNode> nodes;

    static int nodeCount(int u, int n) { return (n <= 10) ? u : nodeCount(u << 1 | 1, (n + 1) >> 1); }

    void leafBox(RmBvhNode &nd, int L, int R) {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = L; i < R; i++) {
            const float *p = pos + size_t(order[i]) * 9;
            for (int k = 0; k < 3; k++) {
                // Face::aabb: glm::min(v0, glm::min(v1, v2)); then Box + Box with std::fmin/fmax
                float mn = std::min(p[k], std::min(p[3 + k], p[6 + k]));
                float mx = std::max(p[k], std::max(p[3 + k], p[6 + k]));
                lo[k] = std::fmin(lo[k], mn);
                hi[k] = std::fmax(hi[k], mx);
            }
        }
        for (int k = 0; k < 3; k++) { nd.v0[k] = lo[k]; nd.v1[k] = hi[k]; }
    }

    void build(int u, int L, int R) {
        if (R - L <= 10) {
            nodes[u].faceL = L;
            nodes[u].faceR = R;
            leafBox(nodes[u], L, R);
            return;
        }
        V3 Em = {0, 0, 0}, Em2 = {0, 0, 0};
        for (int i = L; i < R; i++) {
            V3 m = center[order[i]];
            Em = Em + m;
            Em2 = Em2 + m * m;
        }
        V3 D = Em2 - div_recip(Em * Em, float(R - L));
        int axis = 0;
        if (D.y > D.x) axis = 1;
        if (D.z > D.x && D.z > D.y) axis = 2;
        int M = (R + L) / 2;
        const V3 *c = center.data();
        auto key = [c, axis](int32_t f) { return axis == 0 ? c[f].x : (axis == 1 ? c[f].y : c[f].z); };
        std::nth_element(order.begin() + L, order.begin() + M, order.begin() + R,
                         [&](int32_t a, int32_t b) { return key(a) < key(b); });
        build(u << 1, L, M);
        build(u << 1 | 1, M, R);
        const RmBvhNode &a = nodes[u << 1], &b = nodes[u << 1 | 1];
        for (int k = 0; k < 3; k++) {
            nodes[u].v0[k] = std::fmin(a.v0[k], b.v0[k]);
            nodes[u].v1[k] = std::fmax(a.v1[k], b.v1[k]);
        }
    }

    void run(const float *positions, int n) {
        pos = positions;
        center.resize(n);
        order.resize(n);
        for (int i = 0; i < n; i++) {
            const float *p = pos + size_t(i) * 9;
            center[i] = div_recip((ld3(p) + ld3(p + 3)) + ld3(p + 6), 3.0f);
            order[i] = i;
        }
        nodes.assign(size_t(nodeCount(1, n)) + 1, RmBvhNode{});
        if (n > 0) build(1, 0, n);
    }
};

// ------------------------------------------------------------------ textures
struct HostTexture {
    int width = 0, height = 0, channels = 0, map_depth = 0;
    std::vector<uint8_t> level[8];
};

void build_mips(HostTexture &t) {
    t.map_depth = 8;
    for (int level = 1; level < 8; ++level) {
        int pw = t.width >> (level - 1), ph = t.height >> (level - 1);
        int cw = t.width >> level, ch = t.height >> level;
        if (cw == 0 || ch == 0) { t.map_depth = level; break; }
        t.level[level].resize(size_t(cw) * ch * t.channels);
        const uint8_t *src = t.level[level - 1].data();
        uint8_t *dst = t.level[level].data();
        for (int y = 0; y < ch; ++y)
            for (int x = 0; x < cw; ++x)
                for (int c = 0; c < t.channels; ++c) {
                    int sum = 0;
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            int px = (x * 2 + dx) % pw, py = (y * 2 + dy) % ph;
                            sum += src[(size_t(py) * pw + px) * t.channels + c];
                        }
                    dst[(size_t(y) * cw + x) * t.channels + c] = uint8_t(sum / 4);
                }
    }
}

// Level-0 RGBA8 bilinear fetch with gamma decode, i.e. Material::getEmissiveColor(u, v, NAN)
// (src/material.cpp:365-372 -> ImageData::get depth 0 -> get_bilinear 48-79 -> gammaPow 337-346).
// Only used at load time for the light-object colour average.
V3 emissive_level0(const HostTexture &t, float u, float v) {
    v = 1.0f - v;
    auto wrap = [](int &x, int m) { if (x < 0 || x >= m) { x %= m; if (x < 0) x += m; } };
    // depth 0: level = nextLevel = 0 when map_depth == 1, else blend weight 0 with level 1 -> ret1*1 + ret2*0
    auto bilinear = [&](const std::vector<uint8_t> &data, int w, int h, float out[4]) {
        float x = u * float(w), y = v * float(h);
        int x0 = int(floorf(x)), y0 = int(floorf(y));
        float dx = x - float(x0), dy = y - float(y0);
        wrap(x0, w);
        wrap(y0, h);
        int x1 = (x0 + 1) % w, y1 = (y0 + 1) % h;
        for (int c = 0; c < 4; c++) {
            float c00 = float(data[(size_t(y0) * w + x0) * 4 + c]) / 255.0f;
            float c01 = float(data[(size_t(y0) * w + x1) * 4 + c]) / 255.0f;
            float c10 = float(data[(size_t(y1) * w + x0) * 4 + c]) / 255.0f;
            float c11 = float(data[(size_t(y1) * w + x1) * 4 + c]) / 255.0f;
            float c0 = c00 * (1.0f - dx) + c01 * dx;
            float c1 = c10 * (1.0f - dx) + c11 * dx;
            out[c] = c0 * (1.0f - dy) + c1 * dy;
        }
    };
    float r1[4], r2[4];
    int next = std::min(1, t.map_depth - 1);
    bilinear(t.level[0], t.width, t.height, r1);
    bilinear(t.level[next], t.width >> next, t.height >> next, r2);
    float c[3];
    for (int k = 0; k < 3; k++) c[k] = powf(r1[k] * (1.0f - 0.0f) + r2[k] * 0.0f, 2.2f);
    return {c[0], c[1], c[2]};
}

struct HostLight {
    V3 center{0, 0, 0}, color{0, 0, 0};
    float power = 0;
    std::vector<float> positions, normals, cdf;
};

} // namespace

struct RmPrepared {
    RmSceneDesc desc{};
    std::vector<RmBvhNode> nodes;
    std::vector<int32_t> perm, face_material;
    std::vector<float> positions, uvs, normals, sky_data, sky_cdf;
    std::vector<HostTexture> textures;
    std::vector<RmTextureDesc> texture_descs;
    std::vector<RmMaterialDesc> materials;
    std::vector<HostLight> lights;
    std::vector<RmLightDesc> light_descs;
    std::vector<void *> pinned;       // arrays page-locked by rm_prepared_pin
};

// `tree`: who builds the reference's tree over the raw positions - nullptr = the host recursion above; rm_prepare_scene_device
// (gpu_ref_bvh.cu) passes the device builder.  Everything else is the same host code either way.
using RmTreeFn = std::function<int(const float *, int, std::vector<RmBvhNode> &, std::vector<int32_t> &)>;

int rm_prepare_scene_impl(const RmRawScene *raw, RmPrepared **out, const RmTreeFn *tree) {
    if (!raw || !out) return rm_fail(RM_ERR_INVALID, "rm_prepare_scene: null argument");
    if (raw->n_faces <= 0) return rm_fail(RM_ERR_INVALID, "rm_prepare_scene: scene has no faces");
    for (int k = 0; k < raw->n_meshes; k++) {
        const RmRawMesh &m = raw->meshes[k];
        if (m.face_begin < 0 || m.face_end > raw->n_faces || m.face_begin > m.face_end || m.material < 0 ||
            m.material >= raw->n_materials)
            return rm_fail(RM_ERR_INVALID, "rm_prepare_scene: mesh %d has a bad face range or material", k);
    }
    std::unique_ptr<RmPrepared> P(new RmPrepared());
    const int n = raw->n_faces;

    // textures + mip chains
    P->textures.resize(raw->n_textures);
    for (int i = 0; i < raw->n_textures; i++) {
        const RmRawTexture &t = raw->textures[i];
        if (t.width <= 0 || t.height <= 0 || (t.channels != 3 && t.channels != 4) || !t.pixels)
            return rm_fail(RM_ERR_INVALID, "rm_prepare_scene: texture %d must be RGBA8 or RGB8 with positive size", i);
        HostTexture &h = P->textures[i];
        h.width = t.width; h.height = t.height; h.channels = t.channels;
        h.level[0].assign(t.pixels, t.pixels + size_t(t.width) * t.height * t.channels);
        build_mips(h);
    }
    // materials
    P->materials.resize(raw->n_materials);
    for (int i = 0; i < raw->n_materials; i++) {
        const RmRawMaterial &r = raw->materials[i];
        RmMaterialDesc &m = P->materials[i];
        const int tex[4] = {r.tex_diffuse, r.tex_specular, r.tex_emissive, r.tex_normals};
        for (int k = 0; k < 4; k++) {
            if (tex[k] >= raw->n_textures) return rm_fail(RM_ERR_INVALID, "rm_prepare_scene: material %d texture index out of range", i);
            m.tex[k] = tex[k] < 0 ? -1 : tex[k];
            // the reference fetches diffuse/specular/emissive as vec4 (stride 4) and normals as vec3 (stride 3)
            // regardless of `channels` (src/material.cpp:58): reject pairings it would misread
            if (tex[k] >= 0 && P->textures[tex[k]].channels != (k == 3 ? 3 : 4))
                return rm_fail(RM_ERR_INVALID, "rm_prepare_scene: material %d slot %d needs %s", i, k, k == 3 ? "RGB8" : "RGBA8");
        }
        m.opacity = r.opacity; m.ior = r.ior; m.roughness = r.roughness;
        for (int k = 0; k < 3; k++) m.transmitting_color[k] = r.transmitting_color[k];
        m.has_fully_transparent_part = 0;
        if (m.tex[0] >= 0) {   // ImageData::hasTransparentPart, src/material.cpp:102-107
            const auto &d = P->textures[m.tex[0]].level[0];
            for (size_t b = 3; b < d.size(); b += 4)
                if (d[b] < 255) { m.has_fully_transparent_part = 1; break; }
        }
    }
    // light objects: one candidate per mesh whose material has an emissive texture (src/model.cpp:120-122;
    // the sky is loaded after the meshes, 191-211, so the skyMap.empty() guard always passes)
    std::vector<int32_t> raw_material(n, 0);
    for (int k = 0; k < raw->n_meshes; k++) {
        const RmRawMesh &mesh = raw->meshes[k];
        for (int f = mesh.face_begin; f < mesh.face_end; f++) raw_material[f] = mesh.material;
        const RmMaterialDesc &mat = P->materials[mesh.material];
        if (mat.tex[2] < 0) continue;
        const HostTexture &et = P->textures[mat.tex[2]];
        HostLight L;
        V3 color = {0, 0, 0};
        std::vector<float> weights;
        std::vector<V3> centers;
        for (int f = mesh.face_begin; f < mesh.face_end; f++) {
            const float *p = raw->positions + size_t(f) * 9;
            const float *uv = raw->uvs + size_t(f) * 6;
            V3 avg = {0, 0, 0};                                  // getAverageEmissiveColor, src/model.cpp:23-42
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j) {
                    float a = float(i) / 7, b = float(j) / 7;
                    if (a + b > 1.0f) { a = 1.0f - a; b = 1.0f - b; }
                    float c = 1.0f - a - b;
                    float tu = a * uv[0] + b * uv[2] + c * uv[4];
                    float tv = a * uv[1] + b * uv[3] + c * uv[5];
                    avg = avg + emissive_level0(et, tu, tv);
                }
            avg = div_recip(avg, 64.0f);
            V3 v0 = ld3(p), v1 = ld3(p + 3), v2 = ld3(p + 6);
            float Clum = dot3(avg, kLum);
            float area = len3(cross3(v1 - v0, v2 - v0)) / 2.0f;
            if (Clum < kEps || area < kEps) continue;
            color = color + avg * area;
            weights.push_back(area * Clum);
            centers.push_back(div_recip((v0 + v1) + v2, 3.0f));
            L.positions.insert(L.positions.end(), p, p + 9);
            const float *nr = raw->normals + size_t(f) * 9;
            L.normals.insert(L.normals.end(), nr, nr + 9);
        }
        if (weights.empty()) continue;
        L.color = div_recip(color, dot3(color, kLum));
        for (size_t j = 0; j < weights.size(); j++) {
            L.power += weights[j];
            L.center = L.center + centers[j] * weights[j];
        }
        L.cdf.resize(weights.size());                             // RandomDistribution::Init
        L.cdf[0] = weights[0];
        for (size_t j = 1; j < weights.size(); j++) L.cdf[j] = L.cdf[j - 1] + weights[j];
        L.center = div_true(L.center, L.power);       // `center /= power`
        P->lights.push_back(std::move(L));
    }
    // sky
    if (raw->sky_rgb && raw->sky_width > 0 && raw->sky_height > 0) {
        const int w = raw->sky_width, h = raw->sky_height;
        P->sky_data.assign(raw->sky_rgb, raw->sky_rgb + size_t(w) * h * 3);
        P->sky_cdf.resize(size_t(w) * h);
        float run = 0.0f;
        for (int v = 0; v < h; v++)
            for (int u = 0; u < w; u++) {
                float phi = kPi * (float(v) + 0.5f) / float(h);
                float area = sinf(phi) * 2.0f * kPi / float(w * h);
                size_t id = size_t(v) * w + u;
                float *d = &P->sky_data[id * 3];
                d[0] *= area; d[1] *= area; d[2] *= area;
                float Clum = dot3({d[0], d[1], d[2]}, kLum);
                run = (id == 0) ? Clum : run + Clum;
                P->sky_cdf[id] = run;
            }
    }
    // BVH + permuted face streams
    if (tree) {
        const int rc = (*tree)(raw->positions, n, P->nodes, P->perm);
        if (rc) return rc;
        if (P->perm.size() != size_t(n) || P->nodes.size() < 2) return rm_fail(RM_ERR_STATE, "rm_prepare_scene: the tree builder returned no tree");
    } else {
        BvhBuilder B;
        B.run(raw->positions, n);
        P->nodes = std::move(B.nodes);
        P->perm = std::move(B.order);
    }
    P->positions.resize(size_t(n) * 9);
    P->uvs.resize(size_t(n) * 6);
    P->normals.resize(size_t(n) * 9);
    P->face_material.resize(n);
    for (int i = 0; i < n; i++) {
        size_t s = size_t(P->perm[i]);
        std::memcpy(&P->positions[size_t(i) * 9], raw->positions + s * 9, 36);
        std::memcpy(&P->uvs[size_t(i) * 6], raw->uvs + s * 6, 24);
        std::memcpy(&P->normals[size_t(i) * 9], raw->normals + s * 9, 36);
        P->face_material[i] = raw_material[s];
    }
    // descriptor
    P->texture_descs.resize(P->textures.size());
    for (size_t i = 0; i < P->textures.size(); i++) {
        const HostTexture &h = P->textures[i];
        RmTextureDesc &d = P->texture_descs[i];
        d.width = h.width; d.height = h.height; d.channels = h.channels; d.map_depth = h.map_depth;
        for (int l = 0; l < 8; l++) d.levels[l] = (l < h.map_depth) ? h.level[l].data() : nullptr;
    }
    P->light_descs.resize(P->lights.size());
    for (size_t i = 0; i < P->lights.size(); i++) {
        const HostLight &L = P->lights[i];
        RmLightDesc &d = P->light_descs[i];
        d.center[0] = L.center.x; d.center[1] = L.center.y; d.center[2] = L.center.z;
        d.color[0] = L.color.x; d.color[1] = L.color.y; d.color[2] = L.color.z;
        d.power = L.power;
        d.n_faces = int32_t(L.cdf.size());
        d.face_positions = L.positions.data();
        d.face_normals = L.normals.data();
        d.face_cdf = L.cdf.data();
    }
    RmSceneDesc &D = P->desc;
    D.n_faces = n;
    D.n_nodes = int32_t(P->nodes.size());
    D.n_materials = int32_t(P->materials.size());
    D.n_textures = int32_t(P->textures.size());
    D.n_lights = int32_t(P->lights.size());
    D.sky_width = P->sky_data.empty() ? 0 : raw->sky_width;
    D.sky_height = P->sky_data.empty() ? 0 : raw->sky_height;
    D.nodes = P->nodes.data();
    D.positions = P->positions.data();
    D.uvs = P->uvs.data();
    D.normals = P->normals.data();
    D.face_material = P->face_material.data();
    D.materials = P->materials.data();
    D.textures = P->texture_descs.data();
    D.lights = P->light_descs.data();
    D.sky_data = P->sky_data.empty() ? nullptr : P->sky_data.data();
    D.sky_cdf = P->sky_cdf.empty() ? nullptr : P->sky_cdf.data();
    *out = P.release();
    return RM_OK;
}

// The arrays a prepared scene hands to rm_scene_upload, for rm_prepared_pin (rm_api.cu: page-locking needs the CUDA runtime, which
// this file stays clear of)
void rm_prepared_spans(RmPrepared *P, std::vector<std::pair<void *, size_t>> &out) {
    auto add = [&](void *p, size_t bytes) { if (p && bytes) out.emplace_back(p, bytes); };
    add(P->nodes.data(), P->nodes.size() * sizeof(RmBvhNode));
    add(P->positions.data(), P->positions.size() * 4);
    add(P->uvs.data(), P->uvs.size() * 4);
    add(P->normals.data(), P->normals.size() * 4);
    add(P->face_material.data(), P->face_material.size() * 4);
    add(P->sky_data.data(), P->sky_data.size() * 4);
    add(P->sky_cdf.data(), P->sky_cdf.size() * 4);
    for (HostTexture &t : P->textures)
        for (int l = 0; l < t.map_depth; l++) add(t.level[l].data(), t.level[l].size());
}
std::vector<void *> &rm_prepared_pinned(RmPrepared *P) { return P->pinned; }

extern "C" {

int rm_prepare_scene(const RmRawScene *raw, RmPrepared **out) { return rm_prepare_scene_impl(raw, out, nullptr); }

const RmSceneDesc *rm_prepared_desc(const RmPrepared *p) { return p ? &p->desc : nullptr; }
const int32_t *rm_prepared_permutation(const RmPrepared *p) { return p ? p->perm.data() : nullptr; }
void rm_prepared_free(RmPrepared *p) { delete p; }

} // extern "C"
