// ref_harness.cpp — C-ABI driver around the UNMODIFIED reference translation units.
//
// TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library this builds
// (oracle/_ref/libraym_ref.so).  Nothing under raym0nade_b200/ links or calls it.
//
// The reference (lemonchu/Raym0nade) has no FFI; this file reaches its ordinary C++
// entry points: Model's public members (include/model.h:31-42), BVH::build
// (src/bvh.cpp:48-54), renderPixel (src/render.cpp:448), Photo::FXAA (src/image.cpp:363).
// It is compiled with -fno-access-control so it can also reach BVH::node,
// SkyBox::Init, Model::checkLightObject and BSDF::getBRDF for known-answer vectors.
// No reference source is copied: the reference objects are compiled from
// /root/reference/src where they lie (oracle/Makefile).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "render.h"
#include "image.h"
#include "sampling.h"
#include "../include/rm_types.h"

// ---- reference functions with external linkage that no header declares ----
void renderPixel(const Model &model, const RenderArgs &args, RenderData &renderData, Photo &image, int x, int y);
void initRayDiff(const vec3 &d, const RenderArgs &args, RayDifferential &base_diff);
void getHitInfo(const HitRecord &hit, const Ray &ray, const RayDifferential &base_diff,
                vec3 &hit_dPdx, vec3 &hit_dPdy, HitInfo &hitInfo);
void sampleIndirectLightFromFirstIntersection(const HitInfo &hitInfo, const vec3 &origin,
                                              const RayDifferential &base_diff, const Model &model,
                                              RenderData &renderData, std::vector<LightSample> &samples);
std::vector<LightSample> sampleRay(Ray ray, const RayDifferential &base_diff, const Model &model,
                                   RenderData &renderData, float roughnessFactor, bool excludeDirectLight, int depth);
bool TransparentTest(const Ray &ray, const HitRecord &hit);
vec3 getAbsorb(const vec3 &absorb, float dis);
void sampleSkyBox(const SkyBox &skyBox, const vec3 &shapeNormal, Generator &gen, vec3 &Dir, vec3 &light);

static_assert(sizeof(HitInfo) == sizeof(RmHitInfo), "HitInfo layout");
static_assert(sizeof(RadianceData) == sizeof(RmRadiance), "RadianceData layout");
static_assert(sizeof(BVH_Node) == sizeof(RmBvhNode), "BVH_Node layout");

// ---- per-thread work counters, fed by the --wrap'ed symbols in the counting build ----
struct Counters { uint64_t rays = 0, box = 0, tri = 0; };
static thread_local Counters tl_cnt;

#ifdef RM_REF_COUNT_RAYS
extern "C" void __real__ZNK3BVH6rayHitERK3RayR9HitRecord(const BVH *, const Ray &, HitRecord &);
extern "C" void __wrap__ZNK3BVH6rayHitERK3RayR9HitRecord(const BVH *self, const Ray &ray, HitRecord &hit) {
    tl_cnt.rays++;
    __real__ZNK3BVH6rayHitERK3RayR9HitRecord(self, ray, hit);
}
#endif
#ifdef RM_REF_COUNT_TESTS
extern "C" void __real__Z8rayInBoxRK3RayRK3BoxRfS5_(const Ray &, const Box &, float &, float &);
extern "C" void __wrap__Z8rayInBoxRK3RayRK3BoxRfS5_(const Ray &r, const Box &b, float &tL, float &tR) {
    tl_cnt.box++;
    __real__Z8rayInBoxRK3RayRK3BoxRfS5_(r, b, tL, tR);
}
extern "C" float __real__Z23RayTriangleIntersectionRK3RayRKN3glm3vecILi3EfLNS2_9qualifierE0EEES7_S7_(
        const Ray &, const vec3 &, const vec3 &, const vec3 &);
extern "C" float __wrap__Z23RayTriangleIntersectionRK3RayRKN3glm3vecILi3EfLNS2_9qualifierE0EEES7_S7_(
        const Ray &r, const vec3 &a, const vec3 &b, const vec3 &c) {
    tl_cnt.tri++;
    return __real__Z23RayTriangleIntersectionRK3RayRKN3glm3vecILi3EfLNS2_9qualifierE0EEES7_S7_(r, a, b, c);
}
#endif

namespace {

struct RefScene {
    Model model;            // never moved after construction: Faces hold raw pointers into it
    int n_faces = 0;
    int node_count = 0;
};

vec3 v3(const float *p) { return vec3(p[0], p[1], p[2]); }

int nodeCountLocal(int u, int n) { return (n <= 10) ? u : nodeCountLocal(u << 1 | 1, (n + 1) >> 1); }

RenderArgs toArgs(const RmRenderArgs *a, int threads) {
    RenderArgs r;
    r.position = v3(a->position);
    r.direction = v3(a->direction);
    r.up = v3(a->up);
    r.right = v3(a->right);
    r.accuracy = a->accuracy;
    r.focus = a->focus;
    r.CoC = a->CoC;
    r.exposure = a->exposure;
    r.P_Direct = a->P_Direct;
    r.width = a->width;
    r.height = a->height;
    r.spp = a->spp;
    r.threads = threads;
    return r;
}

// Run fn(y) over rows with `threads` workers, rows dealt dynamically from the LAST row
// down, as TaskQueue::getTask does (src/render.cpp:565-573).
template <typename F>
void parallelRows(int height, int threads, F fn) {
    std::atomic<int> next(height - 1);
    std::vector<std::thread> pool;
    for (int i = 0; i < threads; i++)
        pool.emplace_back([&, i]() {
            for (;;) {
                int y = next.fetch_sub(1);
                if (y < 0) break;
                fn(y, i);
            }
        });
    for (auto &t : pool) t.join();
}

// ---- deterministic replay support -------------------------------------------------
// std::mt19937 cannot be replaced without touching the reference, but its state can be
// pre-loaded so that its next 624 outputs are a chosen 32-bit sequence: write the
// un-tempered words into the state array and set the read index to 0.  libstdc++ lays
// mersenne_twister_engine out as { uint_fast32_t _M_x[624]; size_t _M_p; }.
uint32_t untemper(uint32_t y) {
    // inverse of: y ^= y>>11; y ^= (y<<7)&0x9d2c5680; y ^= (y<<15)&0xefc60000; y ^= y>>18
    y ^= y >> 18;
    y ^= (y << 15) & 0xefc60000u;
    uint32_t t = y;
    for (int i = 0; i < 5; i++) t = y ^ ((t << 7) & 0x9d2c5680u);
    y = t;
    t = y;
    for (int i = 0; i < 3; i++) t = y ^ (t >> 11);
    return t;
}

struct MtImage { uint_fast32_t x[624]; size_t p; };
static_assert(sizeof(MtImage) == sizeof(std::mt19937), "mt19937 layout");

void loadDraws(Generator &gen, const uint32_t *u32, int n) {
    MtImage img;
    for (int i = 0; i < 624; i++) img.x[i] = untemper(i < n ? u32[i] : 0u);
    img.p = 0;
    std::memcpy(static_cast<void *>(&gen.mt), &img, sizeof(img));
}
int drawsUsed(const Generator &gen) {
    MtImage img;
    std::memcpy(&img, static_cast<const void *>(&gen.mt), sizeof(img));
    return int(img.p);
}

} // namespace

extern "C" {

// ---------------------------------------------------------------- scene lifetime
void *ref_scene_create(const RmRawScene *raw) {
    auto *s = new RefScene();
    Model &m = s->model;
    s->n_faces = raw->n_faces;

    // materials + textures; mip chains by the reference's own generateMipmaps
    // (what loadImageFromFile does after decoding, src/material.cpp:288-298)
    m.materials.resize(raw->n_materials);
    for (int i = 0; i < raw->n_materials; i++) {
        const RmRawMaterial &rm = raw->materials[i];
        Material &mat = m.materials[i];
        const int slots[4] = {aiTextureType_DIFFUSE, aiTextureType_SPECULAR, aiTextureType_EMISSIVE, aiTextureType_NORMALS};
        const int tex[4] = {rm.tex_diffuse, rm.tex_specular, rm.tex_emissive, rm.tex_normals};
        for (int k = 0; k < 4; k++) {
            if (tex[k] < 0) continue;
            const RmRawTexture &t = raw->textures[tex[k]];
            ImageData &img = mat.texture[slots[k]];
            img.width = t.width;
            img.height = t.height;
            img.channels = t.channels;
            img.data[0].assign(t.pixels, t.pixels + size_t(t.width) * t.height * t.channels);
            img.generateMipmaps();
        }
        mat.id = i;                      // src/model.cpp:159
        mat.opacity = rm.opacity;        // the values loadMaterialProperties would set (src/material.cpp:300-328)
        mat.ior = rm.ior;
        mat.roughness = rm.roughness;
        mat.transmittingColor = v3(rm.transmitting_color);
        if (mat.texture[aiTextureType_DIFFUSE].hasTransparentPart())   // src/material.cpp:330-333
            mat.hasFullyTransparentPart = true;
    }

    // faces; one private VertexData triple per face so the original index can be
    // recovered from Face::data[0] after BVH::build permutes the array
    m.vertexDatas.reserve(size_t(raw->n_faces) * 3);
    m.faces.reserve(raw->n_faces);
    for (int f = 0; f < raw->n_faces; f++)
        for (int c = 0; c < 3; c++)
            m.vertexDatas.emplace_back(vec2(raw->uvs[(f * 3 + c) * 2], raw->uvs[(f * 3 + c) * 2 + 1]),
                                       v3(raw->normals + (size_t(f) * 3 + c) * 3));
    for (int k = 0; k < raw->n_meshes; k++) {
        const RmRawMesh &mesh = raw->meshes[k];
        const Material &mat = m.materials[mesh.material];
        size_t offset = m.faces.size();
        for (int f = mesh.face_begin; f < mesh.face_end; f++) {
            const float *p = raw->positions + size_t(f) * 9;
            m.faces.push_back({{v3(p), v3(p + 3), v3(p + 6)},
                               {&m.vertexDatas[size_t(f) * 3], &m.vertexDatas[size_t(f) * 3 + 1], &m.vertexDatas[size_t(f) * 3 + 2]},
                               &mat});
        }
        // src/model.cpp:120-122 (the sky is still empty at this point in the reference's load order)
        if (!mat.texture[aiTextureType_EMISSIVE].empty() && m.skyMap.empty()) {
            aiMesh fake{};
            fake.mNumFaces = unsigned(mesh.face_end - mesh.face_begin);
            m.checkLightObject(&m.faces[offset], &fake, mat);
        }
    }

    if (raw->sky_rgb && raw->sky_width > 0) {       // what SkyBox::load does after decoding (src/component.cpp:90-101)
        m.skyMap.width = raw->sky_width;
        m.skyMap.height = raw->sky_height;
        size_t n = size_t(raw->sky_width) * raw->sky_height;
        m.skyMap.data.reserve(n);
        for (size_t i = 0; i < n; i++) m.skyMap.data.emplace_back(raw->sky_rgb[i * 3], raw->sky_rgb[i * 3 + 1], raw->sky_rgb[i * 3 + 2]);
        m.skyMap.Init();
    }

    m.bvh.build(m.faces);                            // src/model.cpp:214
    s->node_count = nodeCountLocal(1, raw->n_faces) + 1;
    return s;
}

void ref_scene_destroy(void *h) { delete static_cast<RefScene *>(h); }
// the reference Model behind a scene handle (for oracle/bridge_harness.cpp)
const void *ref_scene_model(void *h) { return &static_cast<RefScene *>(h)->model; }

int ref_node_count(void *h) { return static_cast<RefScene *>(h)->node_count; }
int ref_light_count(void *h) { return int(static_cast<RefScene *>(h)->model.lightObjects.size()); }

// nodes: [node_count] raw BVH_Node images (slots the build never wrote are zeroed);
// perm[i] = original face index now stored at faces[i].
void ref_bvh_export(void *h, RmBvhNode *nodes, int32_t *perm) {
    auto *s = static_cast<RefScene *>(h);
    const Model &m = s->model;
    std::vector<char> written(s->node_count, 0);
    std::vector<int> stack{1};
    std::memset(nodes, 0, sizeof(RmBvhNode) * s->node_count);
    while (!stack.empty()) {
        int u = stack.back();
        stack.pop_back();
        std::memcpy(&nodes[u], &m.bvh.node[u], sizeof(RmBvhNode));
        if (!m.bvh.node[u].faceR) { stack.push_back(u << 1); stack.push_back(u << 1 | 1); }
    }
    for (int i = 0; i < s->n_faces; i++)
        perm[i] = int32_t((m.faces[i].data[0] - &m.vertexDatas[0]) / 3);
}

// light objects as the reference built them: per light {center[3], color[3], power, n_faces}
void ref_light_export(void *h, int idx, float *center_color_power7, int32_t *n_faces) {
    const LightObject &L = static_cast<RefScene *>(h)->model.lightObjects[idx];
    for (int k = 0; k < 3; k++) { center_color_power7[k] = L.center[k]; center_color_power7[3 + k] = L.color[k]; }
    center_color_power7[6] = L.power;
    *n_faces = int32_t(L.faces.size());
}
// faces_pos [n][9], cdf [n] (the RandomDistribution prefix sums)
void ref_light_faces(void *h, int idx, float *faces_pos, float *cdf) {
    const LightObject &L = static_cast<RefScene *>(h)->model.lightObjects[idx];
    for (size_t i = 0; i < L.faces.size(); i++) {
        for (int c = 0; c < 3; c++)
            for (int k = 0; k < 3; k++) faces_pos[i * 9 + c * 3 + k] = L.faces[i].v[c][k];
        cdf[i] = L.faceDist.prefixSums[i];
    }
}
// premultiplied sky texels [h*w*3] and the luminance prefix sums [h*w]
void ref_sky_export(void *h, float *data, float *cdf) {
    const SkyBox &sky = static_cast<RefScene *>(h)->model.skyMap;
    for (size_t i = 0; i < sky.data.size(); i++) {
        data[i * 3] = sky.data[i].x; data[i * 3 + 1] = sky.data[i].y; data[i * 3 + 2] = sky.data[i].z;
        cdf[i] = sky.dist.prefixSums[i];
    }
}
// one mip level of one material texture slot (slot = aiTextureType value); returns byte count
int64_t ref_texture_level(void *h, int material, int slot, int level, uint8_t *out, int32_t *map_depth) {
    const ImageData &img = static_cast<RefScene *>(h)->model.materials[material].texture[slot];
    *map_depth = img.map_depth;
    if (level < 0 || level >= MAX_MIPMAP_LEVEL) return 0;
    if (out) std::memcpy(out, img.data[level].data(), img.data[level].size());
    return int64_t(img.data[level].size());
}

// ---------------------------------------------------------------- tracing
// Primary rays exactly as renderPixel forms them (src/render.cpp:466-479).
// tri_idx = index into the post-build faces array, -1 on miss; t = hit.t_max (INF on miss).
void ref_trace_primary(void *h, const RmRenderArgs *a, int threads, int32_t *tri_idx, float *t, uint64_t *counters3) {
    const Model &m = static_cast<RefScene *>(h)->model;
    RenderArgs args = toArgs(a, threads);
    std::vector<Counters> per(threads);
    parallelRows(args.height, threads, [&](int y, int tid) {
        Counters before = tl_cnt;
        for (int x = 0; x < args.width; x++) {
            float rayX = float(x) - float(args.width) / 2.0f, rayY = float(y) - float(args.height) / 2.0f;
            vec3 d = args.direction + args.accuracy * (rayX * args.right + rayY * args.up);
            Ray ray = {args.position, normalize(d)};
            HitRecord hit = m.rayHit(ray);
            int id = y * args.width + x;
            tri_idx[id] = (hit.t_max == INFINITY) ? -1 : int32_t(hit.face - &m.faces[0]);
            t[id] = hit.t_max;
        }
        per[tid].rays += tl_cnt.rays - before.rays;
        per[tid].box += tl_cnt.box - before.box;
        per[tid].tri += tl_cnt.tri - before.tri;
    });
    if (counters3) {
        counters3[0] = counters3[1] = counters3[2] = 0;
        for (auto &c : per) { counters3[0] += c.rays; counters3[1] += c.box; counters3[2] += c.tri; }
    }
}

// Arbitrary closest-hit rays through Model::rayHit (src/model.cpp:332-341).
void ref_trace_closest(void *h, int64_t n, const float *org, const float *dir, int32_t *tri_idx, float *t, uint64_t *counters3) {
    const Model &m = static_cast<RefScene *>(h)->model;
    Counters before = tl_cnt;
    for (int64_t i = 0; i < n; i++) {
        HitRecord hit = m.rayHit({v3(org + i * 3), v3(dir + i * 3)});
        tri_idx[i] = (hit.t_max == INFINITY) ? -1 : int32_t(hit.face - &m.faces[0]);
        t[i] = hit.t_max;
    }
    if (counters3) { counters3[0] = tl_cnt.rays - before.rays; counters3[1] = tl_cnt.box - before.box; counters3[2] = tl_cnt.tri - before.tri; }
}

// Occlusion rays through Model::rayHit_test (src/model.cpp:343-354): out[i] = 1 if blocked.
void ref_trace_occluded(void *h, int64_t n, const float *org, const float *dir, const float *aim, uint8_t *out) {
    const Model &m = static_cast<RefScene *>(h)->model;
    for (int64_t i = 0; i < n; i++) out[i] = m.rayHit_test({v3(org + i * 3), v3(dir + i * 3)}, aim[i]) ? 1 : 0;
}

// ---------------------------------------------------------------- full render
// The reference's per-pixel estimator over every pixel: renderPixel (src/render.cpp:448-551)
// driven like render_multiThread (593-626): `threads` workers, RenderData seeded
// seed_base + thread index, rows dealt dynamically.  Wall-clock around the pixel loop only.
void ref_render(void *h, const RmRenderArgs *a, int threads, int seed_base,
                RmHitInfo *gbuffer, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is,
                double *seconds, uint64_t *counters3) {
    const Model &m = static_cast<RefScene *>(h)->model;
    RenderArgs args = toArgs(a, threads);
    Photo photo(args.width, args.height);
    photo.exposure = args.exposure;
    std::vector<RenderData> datas;
    datas.reserve(threads);
    for (int i = 0; i < threads; i++) datas.emplace_back(seed_base + i);
    std::vector<Counters> per(threads);
    auto t0 = std::chrono::steady_clock::now();
    parallelRows(args.height, threads, [&](int y, int tid) {
        Counters before = tl_cnt;
        for (int x = 0; x < args.width; x++) renderPixel(m, args, datas[tid], photo, x, y);
        per[tid].rays += tl_cnt.rays - before.rays;
        per[tid].box += tl_cnt.box - before.box;
        per[tid].tri += tl_cnt.tri - before.tri;
    });
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    size_t n = size_t(args.width) * args.height;
    if (gbuffer) std::memcpy(gbuffer, photo.Gbuffer, n * sizeof(RmHitInfo));
    if (Dd) std::memcpy(Dd, photo.radiance_Dd, n * sizeof(RmRadiance));
    if (Ds) std::memcpy(Ds, photo.radiance_Ds, n * sizeof(RmRadiance));
    if (Id) std::memcpy(Id, photo.radiance_Id, n * sizeof(RmRadiance));
    if (Is) std::memcpy(Is, photo.radiance_Is, n * sizeof(RmRadiance));
    if (counters3) {
        counters3[0] = counters3[1] = counters3[2] = 0;
        for (auto &c : per) { counters3[0] += c.rays; counters3[1] += c.box; counters3[2] += c.tri; }
    }
    delete[] photo.pixelarray;   // Photo::~Photo leaks it (src/image.cpp:22-28)
    photo.pixelarray = nullptr;
}

// ---------------------------------------------------------------- deterministic replay
// One indirect sample of one pixel with a caller-chosen 32-bit draw sequence standing in
// for mt19937's output (see loadDraws).  Runs the reference's
// sampleIndirectLightFromFirstIntersection (src/render.cpp:314-423) on the given G-buffer
// entry and returns its LightSamples as 7 floats each {bsdfPdf[3], light[3], weight}.
// Returns the number of samples; *draws_used = how many 32-bit draws the path consumed
// (>= 624 means the pre-loaded sequence ran out and the result is not a valid replay).
int ref_replay_indirect(void *h, const RmRenderArgs *a, int x, int y, const RmHitInfo *g_in,
                        const uint32_t *u32, int n_u32, float *samples7, int max_samples, int *draws_used) {
    const Model &m = static_cast<RefScene *>(h)->model;
    RenderArgs args = toArgs(a, 1);
    float rayX = float(x) - float(args.width) / 2.0f, rayY = float(y) - float(args.height) / 2.0f;
    vec3 d = args.direction + args.accuracy * (rayX * args.right + rayY * args.up);
    RayDifferential base_diff;
    initRayDiff(d, args, base_diff);
    HitInfo g;
    std::memcpy(static_cast<void *>(&g), g_in, sizeof(g));
    RenderData rd(0);
    loadDraws(rd.gen, u32, n_u32);
    std::vector<LightSample> samples;
    sampleIndirectLightFromFirstIntersection(g, args.position, base_diff, m, rd, samples);
    if (draws_used) *draws_used = drawsUsed(rd.gen);
    int n = 0;
    for (const auto &s : samples) {
        if (n >= max_samples) break;
        float *o = samples7 + n * 7;
        o[0] = s.bsdfPdf.x; o[1] = s.bsdfPdf.y; o[2] = s.bsdfPdf.z;
        o[3] = s.light.x; o[4] = s.light.y; o[5] = s.light.z; o[6] = s.weight;
        n++;
    }
    return int(samples.size());
}

// One direct-light sample (sampleCnt = 1) at a G-buffer entry: sampleDirectLight
// (src/sampling.cpp:467-527) with a pre-loaded draw sequence.
int ref_replay_direct(void *h, const RmRenderArgs *a, const RmHitInfo *g_in, const uint32_t *u32, int n_u32,
                      float *samples7, int *draws_used) {
    const Model &m = static_cast<RefScene *>(h)->model;
    HitInfo g;
    std::memcpy(static_cast<void *>(&g), g_in, sizeof(g));
    vec3 inDir = normalize(g.position - v3(a->position));     // src/render.cpp:431-432
    BSDF bsdf(-inDir, g);
    Generator gen(0);
    loadDraws(gen, u32, n_u32);
    std::vector<LightSample> samples = sampleDirectLight(bsdf, m, gen, 1);
    if (draws_used) *draws_used = drawsUsed(gen);
    if (!samples.empty()) {
        const auto &s = samples[0];
        samples7[0] = s.bsdfPdf.x; samples7[1] = s.bsdfPdf.y; samples7[2] = s.bsdfPdf.z;
        samples7[3] = s.light.x; samples7[4] = s.light.y; samples7[5] = s.light.z; samples7[6] = s.weight;
    }
    return int(samples.size());
}

// Generator::operator() (src/component.cpp:5-10) on chosen raw 32-bit draws.
void ref_uniform_from_u32(const uint32_t *u32, int n, float *out) {
    Generator gen(0);
    for (int base = 0; base < n; base += 624) {
        int m = std::min(624, n - base);
        loadDraws(gen, u32 + base, m);
        for (int i = 0; i < m; i++) out[base + i] = gen();
    }
}

// ---------------------------------------------------------------- G-buffer of one pixel set
// Primary hit + getHitInfo for every pixel, i.e. renderPixel with spp = 0
// (src/render.cpp:466-495 without the sample loops; baseColor is the restored value).
void ref_gbuffer(void *h, const RmRenderArgs *a, int threads, RmHitInfo *gbuffer) {
    RmRenderArgs a0 = *a;
    a0.spp = 0;
    ref_render(h, &a0, threads, 0, gbuffer, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

// ---------------------------------------------------------------- post pass
// Photo::FXAA (src/image.cpp:363-452) on a caller-supplied fp32 RGB image.
void ref_fxaa(const float *rgb_in, float *rgb_out, int width, int height) {
    Photo photo(width, height);
    std::memcpy(static_cast<void *>(photo.pixelarray), rgb_in, size_t(width) * height * sizeof(vec3));
    photo.FXAA();
    std::memcpy(rgb_out, photo.pixelarray, size_t(width) * height * sizeof(vec3));
    delete[] photo.pixelarray;
    photo.pixelarray = nullptr;
}

// Photo::shade [+ bloom] + gammaCorrection [+ FXAA] = Photo::postProcessing (src/image.cpp:470-479)
// without depth-of-field, on caller-supplied planes.
void ref_postprocess(const RmHitInfo *gbuffer, const RmRadiance *Dd, const RmRadiance *Ds, const RmRadiance *Id,
                     const RmRadiance *Is, int width, int height, float exposure, int shade_options, float *rgb_out) {
    Photo photo(width, height);
    size_t n = size_t(width) * height;
    photo.exposure = exposure;
    std::memcpy(static_cast<void *>(photo.Gbuffer), gbuffer, n * sizeof(RmHitInfo));
    std::memcpy(static_cast<void *>(photo.radiance_Dd), Dd, n * sizeof(RmRadiance));
    std::memcpy(static_cast<void *>(photo.radiance_Ds), Ds, n * sizeof(RmRadiance));
    std::memcpy(static_cast<void *>(photo.radiance_Id), Id, n * sizeof(RmRadiance));
    std::memcpy(static_cast<void *>(photo.radiance_Is), Is, n * sizeof(RmRadiance));
    photo.postProcessing(shade_options & ~Photo::DoDepthFieldBlur);
    std::memcpy(rgb_out, photo.pixelarray, n * sizeof(vec3));
    delete[] photo.pixelarray;
    photo.pixelarray = nullptr;
}

// Photo::postProcessing with every stage available, depth of field included (focus / CoC / cameraPosition as
// render_multiThread sets them, src/render.cpp:665-668)
void ref_postprocess_full(const RmHitInfo *gbuffer, const RmRadiance *Dd, const RmRadiance *Ds, const RmRadiance *Id,
                          const RmRadiance *Is, int width, int height, float exposure, int shade_options,
                          const float *camera_position, float focus, float CoC, float *rgb_out) {
    Photo photo(width, height);
    size_t n = size_t(width) * height;
    photo.exposure = exposure;
    photo.focus = focus;
    photo.CoC = CoC;
    photo.cameraPosition = vec3(camera_position[0], camera_position[1], camera_position[2]);
    std::memcpy(static_cast<void *>(photo.Gbuffer), gbuffer, n * sizeof(RmHitInfo));
    std::memcpy(static_cast<void *>(photo.radiance_Dd), Dd, n * sizeof(RmRadiance));
    std::memcpy(static_cast<void *>(photo.radiance_Ds), Ds, n * sizeof(RmRadiance));
    std::memcpy(static_cast<void *>(photo.radiance_Id), Id, n * sizeof(RmRadiance));
    std::memcpy(static_cast<void *>(photo.radiance_Is), Is, n * sizeof(RmRadiance));
    photo.postProcessing(shade_options);
    std::memcpy(rgb_out, photo.pixelarray, n * sizeof(vec3));
    delete[] photo.pixelarray;
    photo.pixelarray = nullptr;
}

// Photo::depthFeildBlur (src/image.cpp:285-356) on a caller-supplied rgb frame and G-buffer
void ref_depth_field_blur(const RmHitInfo *gbuffer, const float *rgb_in, float *rgb_out, int width, int height,
                          const float *camera_position, float focus, float CoC) {
    Photo photo(width, height);
    size_t n = size_t(width) * height;
    std::memcpy(static_cast<void *>(photo.Gbuffer), gbuffer, n * sizeof(RmHitInfo));
    std::memcpy(static_cast<void *>(photo.pixelarray), rgb_in, n * sizeof(vec3));
    photo.focus = focus;
    photo.CoC = CoC;
    photo.cameraPosition = vec3(camera_position[0], camera_position[1], camera_position[2]);
    photo.depthFeildBlur();
    std::memcpy(rgb_out, photo.pixelarray, n * sizeof(vec3));
    delete[] photo.pixelarray;
    photo.pixelarray = nullptr;
}

// Photo::spatialClamp (src/image.cpp:78-83) and Photo::filter (203-213) on caller-supplied planes, in place,
// in the order render_multiThread applies them (src/render.cpp:645, 654).  stages: bit 0 clamp, bit 1 filter.
void ref_denoise(const RmHitInfo *gbuffer, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is,
                 int width, int height, int stages) {
    Photo photo(width, height);
    size_t n = size_t(width) * height;
    RmRadiance *host[4] = {Dd, Ds, Id, Is};
    RadianceData *mine[4] = {photo.radiance_Dd, photo.radiance_Ds, photo.radiance_Id, photo.radiance_Is};
    std::memcpy(static_cast<void *>(photo.Gbuffer), gbuffer, n * sizeof(RmHitInfo));
    for (int k = 0; k < 4; k++) std::memcpy(static_cast<void *>(mine[k]), host[k], n * sizeof(RmRadiance));
    if (stages & 1) photo.spatialClamp();
    if (stages & 2) photo.filter();
    for (int k = 0; k < 4; k++) std::memcpy(host[k], mine[k], n * sizeof(RmRadiance));
    delete[] photo.pixelarray;
    photo.pixelarray = nullptr;
}

// ---------------------------------------------------------------- known-answer wrappers
// rays [n][6] (origin, direction), boxes [n][6], tlr [n][2] in/out   (src/geometry.cpp:40-61)
void ref_kat_ray_in_box(int64_t n, const float *rays, const float *boxes, float *tlr) {
    for (int64_t i = 0; i < n; i++) {
        Ray r = {v3(rays + i * 6), v3(rays + i * 6 + 3)};
        Box b(v3(boxes + i * 6), v3(boxes + i * 6 + 3));
        rayInBox(r, b, tlr[i * 2], tlr[i * 2 + 1]);
    }
}
// tris [n][9]   (src/geometry.cpp:63-87)
void ref_kat_ray_triangle(int64_t n, const float *rays, const float *tris, float *t) {
    for (int64_t i = 0; i < n; i++) {
        Ray r = {v3(rays + i * 6), v3(rays + i * 6 + 3)};
        t[i] = RayTriangleIntersection(r, v3(tris + i * 9), v3(tris + i * 9 + 3), v3(tris + i * 9 + 6));
    }
}
// (src/geometry.cpp:89-103)
void ref_kat_barycentric(int64_t n, const float *tris, const float *p, float *out) {
    for (int64_t i = 0; i < n; i++) {
        vec3 b = barycentric(v3(tris + i * 9), v3(tris + i * 9 + 3), v3(tris + i * 9 + 6), v3(p + i * 3));
        out[i * 3] = b.x; out[i * 3 + 1] = b.y; out[i * 3 + 2] = b.z;
    }
}
// Material::getDiffuseColor / getEmissiveColor / getNormal / getSurfaceData (src/material.cpp:349-383)
// which: 0 diffuse (4 out), 1 emissive (3 out), 2 normal (3 out), 3 surface data (roughness, metallic)
void ref_kat_material_fetch(void *h, int material, int which, int64_t n, const float *uvd, float *out) {
    const Material &mat = static_cast<RefScene *>(h)->model.materials[material];
    for (int64_t i = 0; i < n; i++) {
        float u = uvd[i * 3], v = uvd[i * 3 + 1], d = uvd[i * 3 + 2];
        float *o = out + i * 4;
        o[0] = o[1] = o[2] = o[3] = 0.0f;
        if (which == 0) { vec4 c = mat.getDiffuseColor(u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w; }
        else if (which == 1) { vec3 c = mat.getEmissiveColor(u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
        else if (which == 2) { vec3 c = mat.getNormal(u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
        else mat.getSurfaceData(u, v, o[0], o[1]);
    }
}
// SkyBox::get (src/component.cpp:120-140)
void ref_kat_sky_get(void *h, int64_t n, const float *dirs, float *out) {
    const SkyBox &sky = static_cast<RefScene *>(h)->model.skyMap;
    for (int64_t i = 0; i < n; i++) {
        vec3 c = sky.get(v3(dirs + i * 3));
        out[i * 3] = c.x; out[i * 3 + 1] = c.y; out[i * 3 + 2] = c.z;
    }
}
// BSDF::getBSDF / getBRDF / getBTDF (src/sampling.cpp:49-199) at a surface record.
// which: 0 getBSDF, 1 getBRDF, 2 getBTDF.  in_dirs = BSDF::inDir (pointing away from the surface).
void ref_kat_bsdf(int which, int64_t n, const RmHitInfo *surf, const float *in_dirs, const float *out_dirs, float *out) {
    for (int64_t i = 0; i < n; i++) {
        HitInfo g;
        std::memcpy(static_cast<void *>(&g), surf + i, sizeof(g));
        BSDF bsdf(v3(in_dirs + i * 3), g);
        vec3 L = v3(out_dirs + i * 3);
        vec3 c = which == 0 ? bsdf.getBSDF(L) : (which == 1 ? bsdf.getBRDF(L) : bsdf.getBTDF(L));
        out[i * 3] = c.x; out[i * 3 + 1] = c.y; out[i * 3 + 2] = c.z;
    }
}
// BSDF::preciseRefraction (src/sampling.cpp:271-308): out [n][4] = {dir[3], F}
void ref_kat_precise_refraction(int64_t n, const RmHitInfo *surf, const float *in_dirs, float *out) {
    for (int64_t i = 0; i < n; i++) {
        HitInfo g;
        std::memcpy(static_cast<void *>(&g), surf + i, sizeof(g));
        BSDF bsdf(v3(in_dirs + i * 3), g);
        vec3 o;
        float F;
        bsdf.preciseRefraction(o, F);
        out[i * 4] = o.x; out[i * 4 + 1] = o.y; out[i * 4 + 2] = o.z; out[i * 4 + 3] = F;
    }
}
// accumulateInwardRadiance (src/image.cpp:630-659): samples7 [n]{bsdfPdf,light,weight}; out [n][8] = {d.rad,d.Var,s.rad,s.Var}
void ref_kat_accumulate(int64_t n, const float *base_colors, const float *samples7, float *out) {
    for (int64_t i = 0; i < n; i++) {
        RadianceData d, s;
        LightSample ls(v3(samples7 + i * 7), v3(samples7 + i * 7 + 3), samples7[i * 7 + 6]);
        accumulateInwardRadiance(v3(base_colors + i * 3), ls, d, s);
        float *o = out + i * 8;
        o[0] = d.radiance.x; o[1] = d.radiance.y; o[2] = d.radiance.z; o[3] = d.Var;
        o[4] = s.radiance.x; o[5] = s.radiance.y; o[6] = s.radiance.z; o[7] = s.Var;
    }
}
// getAbsorb (src/render.cpp:83-87)
void ref_kat_absorb(int64_t n, const float *absorb, const float *dist, float *out) {
    for (int64_t i = 0; i < n; i++) {
        vec3 c = getAbsorb(v3(absorb + i * 3), dist[i]);
        out[i * 3] = c.x; out[i * 3 + 1] = c.y; out[i * 3 + 2] = c.z;
    }
}

int ref_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }
const char *ref_build_flavour(void) {
#if defined(RM_REF_COUNT_TESTS)
    return "count-tests";
#elif defined(RM_REF_COUNT_RAYS)
    return "count-rays";
#else
    return "plain";
#endif
}

} // extern "C"
