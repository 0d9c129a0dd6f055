// device_on_host.cpp — TEST-ONLY: the device headers of raym0nade_b200/csrc compiled as host code (cuda_on_host.h) and
// exposed with the signatures of the oracle's known-answer functions, so that the CPU suite holds the kernels' own source
// for the deterministic stages - rayInBox, RayTriangleIntersection, barycentric, BRDF / BTDF evaluation, the RNG float
// mapping, trilinear material fetches, the sky lookup - to the golden vectors the compiled reference produced
// (tests/golden/reference_vectors.npz).  The GPU suite checks the same functions as they run on the device; this copy
// catches an arithmetic regression in every CPU-only test run.  Built with -ffp-contract=off, no fast-math.
#include <cstdio>
#include "cuda_on_host.h"
#include "raym0nade_b200.h"
namespace rm { struct Philox4; static Philox4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t k0, uint32_t k1); }
#include "dev_bsdf.cuh"
#include "dev_texture.cuh"
namespace rm { static Philox4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t k0, uint32_t k1) { return philox4x32_10(c0, c1, c2, 0u, k0, k1); } }

#include "kernels_post.cuh"
#undef __shared__
#define __shared__                   // k_trace's `extern __shared__` stack: declared, never defined - the engine is called directly
#include "kernels_trace.cuh"
#undef __shared__
#define __shared__ static
#include "kernels_render.cuh"

#include <vector>

// the secondary-ray tree builder of the product (raym0nade_b200/csrc/fast_bvh.cpp, compiled into this test library as it is)
int rm_build_fast_bvh(const float *positions, int n, int depth_cap, int leaf_max, std::vector<RmBvhNode> &nodes, std::vector<int32_t> &order, int *depth_out);

using namespace rm;

namespace {
V3 ld(const float *p) { return mk3(p[0], p[1], p[2]); }

// the scene streams of dev_scene.cuh formed on the host the way rm_scene_upload forms them (rm_api.cu): texture table +
// one texel blob with 16-byte aligned levels, materials, the k/255 table, the sky
struct HostScene {
    std::vector<DevTexture> textures;
    std::vector<uint8_t> texels;
    std::vector<DevMaterial> materials;
    std::vector<float4> nodes, tri, shade;
    std::vector<DevLight> lights;
    std::vector<float> lpos, lnrm, lcdf;
    std::vector<int32_t> guide;
    int levels = 2;                          // traversal stack entries = tree depth (rm_scene_upload)
    std::vector<RmBvhNode> fast_nodes;
    std::vector<int32_t> fast_order;
    std::vector<float4> fast_tri;
    // scene_fast of rm_scene_upload: the secondary-ray tree (pair blocks, explicit children), the triangle records in the
    // builder's order, face_map back to the reference order; depth cap 22, leaves <= 3 as the context's defaults
    bool use_secondary_tree(const RmSceneDesc *sc, int depth_cap = 22, int leaf_max = 3) {
        int depth = 0;
        if (rm_build_fast_bvh(sc->positions, sc->n_faces, depth_cap, leaf_max, fast_nodes, fast_order, &depth) != RM_OK) return false;
        fast_tri.resize(size_t(sc->n_faces) * kTriStride);
        for (int i = 0; i < sc->n_faces; i++)                      // k_permute_tris
            for (int k = 0; k < kTriStride; k++) fast_tri[size_t(i) * kTriStride + k] = tri[size_t(fast_order[i]) * kTriStride + k];
        S.nodes = reinterpret_cast<const float4 *>(fast_nodes.data());
        S.tri = fast_tri.data();
        S.face_map = fast_order.data();
        S.explicit_children = 1;
        S.root_is_leaf = fast_nodes[1].faceR != 0;
        levels = std::min(std::max(depth, 2), 40);
        return true;
    }
    float lut[256];
    DevScene S{};
    explicit HostScene(const RmSceneDesc *sc) {
        textures.resize(std::max(sc->n_textures, 1));
        size_t bytes = 0;
        for (int i = 0; i < sc->n_textures; i++) {
            const RmTextureDesc &t = sc->textures[i];
            DevTexture &d = textures[i];
            d.width = t.width; d.height = t.height; d.channels = t.channels; d.map_depth = t.map_depth;
            for (int l = 0; l < 8; l++) {
                d.offset[l] = 0;
                if (l >= t.map_depth) continue;
                const size_t n = size_t(t.width >> l) * (t.height >> l) * t.channels, off = (bytes + 15) & ~size_t(15);
                d.offset[l] = uint32_t(off);
                bytes = off + n;
                texels.resize(bytes);
                if (n) std::memcpy(&texels[off], t.levels[l], n);
            }
        }
        texels.resize(((bytes + 15) & ~size_t(15)) + 16);
        materials.resize(std::max(sc->n_materials, 1));
        for (int i = 0; i < sc->n_materials; i++) {
            const RmMaterialDesc &m = sc->materials[i];
            DevMaterial &d = materials[i];
            for (int k = 0; k < 4; k++) d.tex[k] = m.tex[k] < 0 ? -1 : m.tex[k];
            d.opacity = m.opacity; d.ior = m.ior; d.roughness = m.roughness;
            for (int k = 0; k < 3; k++) d.tc[k] = m.transmitting_color[k];
            d.cutout = m.has_fully_transparent_part ? 1 : 0;
            d._pad = 0;
        }
        for (int k = 0; k < 256; k++) lut[k] = float(k) / 255.0f;
        // traversal: node pairs as they are (padded to an even count), the 48-byte triangle records and the 112-byte shading
        // records of k_pack_faces (rm_api.cu), restated here with the same fp32 operations
        nodes.assign((size_t(sc->n_nodes + 2) & ~size_t(1)) * 2, make_float4(0, 0, 0, 0));
        std::memcpy(nodes.data(), sc->nodes, sizeof(RmBvhNode) * sc->n_nodes);
        tri.resize(size_t(sc->n_faces) * kTriStride);
        shade.resize(size_t(sc->n_faces) * 7);
        bool any_cutout = false;
        for (int i = 0; i < sc->n_faces; i++) {
            const float *v = sc->positions + size_t(i) * 9, *u = sc->uvs + size_t(i) * 6, *nn = sc->normals + size_t(i) * 9;
            const int m = sc->face_material[i];
            const V3 e1 = mk3(fsub(v[3], v[0]), fsub(v[4], v[1]), fsub(v[5], v[2])), e2 = mk3(fsub(v[6], v[0]), fsub(v[7], v[1]), fsub(v[8], v[2]));
            float4 *t = &tri[size_t(i) * kTriStride];
            t[0] = make_float4(v[0], v[1], v[2], e1.x);
            t[1] = make_float4(e1.y, e1.z, e2.x, e2.y);
            t[2] = make_float4(e2.z, length(e1), materials[m].cutout ? 1.0f : 0.0f, 0.0f);
            any_cutout |= materials[m].cutout != 0;
            float4 *s = &shade[size_t(i) * 7];
            s[0] = make_float4(v[0], v[1], v[2], v[3]);
            s[1] = make_float4(v[4], v[5], v[6], v[7]);
            s[2] = make_float4(v[8], u[0], u[1], u[2]);
            s[3] = make_float4(u[3], u[4], u[5], nn[0]);
            s[4] = make_float4(nn[1], nn[2], nn[3], nn[4]);
            s[5] = make_float4(nn[5], nn[6], nn[7], nn[8]);
            s[6] = make_float4(__int_as_float(m), 0.0f, 0.0f, 0.0f);
        }
        S.nodes = nodes.data(); S.tri = tri.data(); S.shade = shade.data();
        S.n_faces = sc->n_faces; S.n_nodes = sc->n_nodes;
        S.any_cutout = any_cutout ? 1 : 0;
        S.root_is_leaf = sc->nodes[1].faceR != 0;
        S.explicit_children = 0; S.face_map = nullptr;
        levels = 1;
        while ((int64_t(1) << levels) < int64_t(sc->n_nodes)) levels++;
        levels = std::min(std::max(levels, 2), 40);
        // lights and the sky-CDF brackets, as rm_scene_upload stages them
        lights.resize(std::max(sc->n_lights, 1));
        for (int i = 0; i < sc->n_lights; i++) {
            const RmLightDesc &L = sc->lights[i];
            DevLight &d = lights[i];
            for (int k = 0; k < 3; k++) { d.center[k] = L.center[k]; d.color[k] = L.color[k]; }
            d.power = L.power; d.n_faces = L.n_faces; d.face_offset = int32_t(lcdf.size());
            lpos.insert(lpos.end(), L.face_positions, L.face_positions + size_t(L.n_faces) * 9);
            lnrm.insert(lnrm.end(), L.face_normals, L.face_normals + size_t(L.n_faces) * 9);
            lcdf.insert(lcdf.end(), L.face_cdf, L.face_cdf + L.n_faces);
        }
        guide.assign(kSkyGuide + 1, 0);
        const size_t nsky = size_t(sc->sky_width) * sc->sky_height;
        if (nsky) {
            const float *cdf = sc->sky_cdf, tot = cdf[nsky - 1];
            for (int j = 0; j <= kSkyGuide; j++) guide[j] = int32_t(std::lower_bound(cdf, cdf + nsky, tot * (float(j) / float(kSkyGuide))) - cdf);
        }
        S.lights = lights.data(); S.light_pos = lpos.data(); S.light_nrm = lnrm.data(); S.light_cdf = lcdf.data();
        S.n_lights = sc->n_lights; S.sky_guide = guide.data();
        S.textures = textures.data(); S.texels = texels.data(); S.materials = materials.data(); S.div255 = lut;
        S.n_materials = sc->n_materials;
        S.sky_width = sc->sky_width; S.sky_height = sc->sky_height; S.sky_data = sc->sky_data; S.sky_cdf = sc->sky_cdf;
    }
};
}  // namespace

template <class Job>
void run_engine(const HostScene &H, Job job, int n, unsigned long long *out_counts, int smem_levels = 0, int w_leaf = 1) {
    int cursor = 0;
    TraceTune tune{28, 1, w_leaf, std::min(H.levels, smem_levels > 0 ? smem_levels : H.levels)};      // launch_trace's clamp
    std::vector<int2> stack(size_t(tune.smem_levels) * 32);
    unsigned long long total[3] = {0, 0, 0};
    std::mutex m;
    rm_host_launch_warp([&](int lane) {
        TraceCounters cnt = {0, 0, 0};
        Job mine = job;
        trace_engine<Job, true>(H.S, mine, n, &cursor, stack.data() + lane, 32, cnt, tune);
        std::lock_guard<std::mutex> l(m);
        total[0] += cnt.rays; total[1] += cnt.box; total[2] += cnt.tri;
    });
    if (out_counts) for (int k = 0; k < 3; k++) out_counts[k] = total[k];
}

extern "C" {

// variant 0: the general slab test; 1: the fast path the engine takes for rays without a parallel axis (others: general)
void doh_ray_in_box(int64_t n, const float *rays, const float *boxes, float *tlr, int variant) {
    for (int64_t i = 0; i < n; i++) {
        const RaySetup r = setup_ray(ld(rays + i * 6), ld(rays + i * 6 + 3));
        const float *b = boxes + i * 6;
        const float4 a = make_float4(b[0], b[1], b[2], b[3]), c = make_float4(b[4], b[5], 0.0f, 0.0f);
        if (variant == 1 && (r.flags & 7u) == 0) ray_in_box_fast(r, a, c, tlr[i * 2], tlr[i * 2 + 1]);
        else ray_in_box(r, a, c, tlr[i * 2], tlr[i * 2 + 1]);
    }
}

// the traversal record of a triangle as k_pack_faces forms it (rm_api.cu): {v0, e1.x} {e1.yz, e2.xy} {e2.z, |e1|, cutout, 0}
void doh_ray_triangle(int64_t n, const float *rays, const float *tris, float *t) {
    for (int64_t i = 0; i < n; i++) {
        const float *v = tris + i * 9;
        const V3 e1 = mk3(fsub(v[3], v[0]), fsub(v[4], v[1]), fsub(v[5], v[2])), e2 = mk3(fsub(v[6], v[0]), fsub(v[7], v[1]), fsub(v[8], v[2]));
        const float4 q0 = make_float4(v[0], v[1], v[2], e1.x), q1 = make_float4(e1.y, e1.z, e2.x, e2.y), q2 = make_float4(e2.z, length(e1), 0.0f, 0.0f);
        t[i] = ray_triangle(setup_ray(ld(rays + i * 6), ld(rays + i * 6 + 3)), q0, q1, q2);
    }
}

void doh_barycentric(int64_t n, const float *tris, const float *p, float *out) {
    for (int64_t i = 0; i < n; i++) {
        const V3 b = barycentric(ld(tris + i * 9), ld(tris + i * 9 + 3), ld(tris + i * 9 + 6), ld(p + i * 3));
        out[i * 3] = b.x; out[i * 3 + 1] = b.y; out[i * 3 + 2] = b.z;
    }
}

// which: 0 getBSDF, 1 getBRDF, 2 getBTDF (src/sampling.cpp:49-199)
void doh_bsdf(int which, int64_t n, const RmHitInfo *surf, const float *in_dirs, const float *out_dirs, float *out) {
    for (int64_t i = 0; i < n; i++) {
        Bsdf B;
        B.inDir = ld(in_dirs + i * 3);
        B.s = load_hitinfo(surf + i);
        const V3 L = ld(out_dirs + i * 3);
        const V3 c = which == 0 ? get_bsdf(B, L) : (which == 1 ? get_brdf(B, L) : get_btdf(B, L));
        out[i * 3] = c.x; out[i * 3 + 1] = c.y; out[i * 3 + 2] = c.z;
    }
}

void doh_uniform_from_u32(const uint32_t *u32, int n, float *out) {
    for (int i = 0; i < n; i++) out[i] = uniform_from_u32(u32[i]);
}

// which: 0 getDiffuseColor, 1 getEmissiveColor, 2 getNormal, 3 getSurfaceData (src/material.cpp:349-383); uvd = (u, v, duv)
void doh_material_fetch(const RmSceneDesc *sc, int material, int which, int64_t n, const float *uvd, float *out) {
    HostScene H(sc);
    const DevMaterial &m = H.materials[material];
    for (int64_t i = 0; i < n; i++) {
        const float u = uvd[i * 3], v = uvd[i * 3 + 1], d = uvd[i * 3 + 2];
        float *o = out + i * 4;
        o[0] = o[1] = o[2] = o[3] = 0.0f;
        if (which == 0) { const V4 c = mat_diffuse(H.S, m, u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w; }
        else if (which == 1) { const V3 c = mat_emissive(H.S, m, u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
        else if (which == 2) { const V3 c = mat_normal(H.S, m, u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
        else mat_surface(H.S, m, u, v, o[0], o[1]);
    }
}

void doh_sky_get(const RmSceneDesc *sc, int64_t n, const float *dirs, float *out) {
    HostScene H(sc);
    for (int64_t i = 0; i < n; i++) {
        const V3 c = sky_get(H.S, ld(dirs + i * 3));
        out[i * 3] = c.x; out[i * 3 + 1] = c.y; out[i * 3 + 2] = c.z;
    }
}

// ---- whole kernels, launched on the host thread by thread (cuda_on_host.h) with the launch shapes of rm_render.cu

void doh_fxaa(const float *in, float *out, int w, int h) {                 // rm_fxaa_device
    std::vector<int> list(size_t(w) * h + 1, 0);
    int count = 0;
    rm_host_launch_blocks(k_fxaa, dim3((w + kFxTileW - 1) / kFxTileW, (h + kFxTileH - 1) / kFxTileH), dim3(kFxTileW, kFxTileH), in, out, w, h, list.data(), &count);
    rm_host_launch(k_fxaa_edges, dim3(2), dim3(256), in, out, w, h, (const int *)list.data(), (const int *)&count);
}

// stages: bit 0 Photo::spatialClamp (rm_spatial_clamp), bit 1 Photo::filter (rm_filter), in the reference's order; planes in place
void doh_denoise(const RmHitInfo *G, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is, int w, int h, int stages) {
    const int npix = w * h;
    std::vector<RmRadiance> alt[4];
    Planes4 cur{{Dd, Ds, Id, Is}}, other;
    for (int k = 0; k < 4; k++) { alt[k].resize(npix); other.p[k] = alt[k].data(); }
    auto swap = [&] { std::swap(cur, other); };
    if (stages & 1) {
        rm_host_launch_blocks(k_spatial_clamp, dim3((w + kClW - 1) / kClW, (h + kClH - 1) / kClH, 4), dim3(kClW, kClH), cur, other, w, h);
        swap();
    }
    if (stages & 2) {
        std::vector<float4> pm(npix), ns(npix);
        std::vector<float> op(npix);
        FilterG F{pm.data(), ns.data(), op.data()};
        rm_host_launch(k_filter_pack, dim3((npix + 255) / 256), dim3(256), G, F, npix);
        const dim3 block(32, 4);
        rm_host_launch(k_filter_var, dim3((w + 31) / 32, (h + 3) / 4, 4), block, cur, other, w, h);
        swap();
        for (int step = 1; step <= 16; step *= 2) {
            rm_host_launch(k_atrous, dim3((w + 31) / 32, (h + 3) / 4, 1), block, G, F, cur, other, w, h, step);
            swap();
        }
    }
    if (cur.p[0] != Dd)                                                     // an odd number of passes: the result sits in the scratch set
        for (int k = 0; k < 4; k++) std::memcpy(other.p[k], cur.p[k], size_t(npix) * sizeof(RmRadiance));
}

// Photo::postProcessing without depth of field (rm_postprocess): shade -> [bloom] -> gamma -> [FXAA]
void doh_postprocess(const RmHitInfo *G, const RmRadiance *Dd, const RmRadiance *Ds, const RmRadiance *Id, const RmRadiance *Is, int w, int h,
                     float exposure, int options, float *rgb_out) {
    const int npix = w * h;
    const bool bloom = (options & 256) != 0;
    std::vector<float> rgb(size_t(npix) * 3), tmp(size_t(npix) * 3);
    rm_host_launch(k_shade_gamma, dim3((npix + 255) / 256), dim3(256), G, Dd, Ds, Id, Is, npix, exposure, options, !bloom, rgb.data());
    if (bloom) {
        std::vector<float> glow[2] = {std::vector<float>(size_t(npix) * 3), std::vector<float>(size_t(npix) * 3)};
        rm_host_launch(k_bloom_bright, dim3((npix + 255) / 256), dim3(256), (const float *)rgb.data(), glow[0].data(), npix);
        int cur = 0;
        for (int step = 1; step <= 16; step *= 2, cur ^= 1)
            rm_host_launch(k_bloom_pass, dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), (const float *)glow[cur].data(), glow[cur ^ 1].data(), rgb.data(), w, h, step);
        rm_host_launch(k_gamma, dim3((npix + 255) / 256), dim3(256), rgb.data(), npix);
    }
    if (options & 512) { doh_fxaa(rgb.data(), tmp.data(), w, h); rgb.swap(tmp); }
    std::memcpy(rgb_out, rgb.data(), rgb.size() * sizeof(float));
}

// ---- the traversal engine itself: one emulated warp (32 real threads voting and shuffling through cuda_on_host.h) pulls all
// the rays from the cursor, exactly as each warp of k_trace does on the device.  out_counts = {rays, box tests, triangle tests}.
// rm_trace_primary: the primary ray of every pixel as renderPixel forms it + Model::rayHit
void doh_trace_primary(const RmSceneDesc *sc, const RmRenderArgs *a, int32_t *tri_idx, float *t, unsigned long long *counts) {
    HostScene H(sc);
    PrimaryJob job;
    job.A.position = ld(a->position); job.A.direction = ld(a->direction); job.A.up = ld(a->up); job.A.right = ld(a->right);
    job.A.accuracy = a->accuracy; job.A.exposure = a->exposure; job.A.P_Direct = a->P_Direct;
    job.A.width = a->width; job.A.height = a->height; job.A.spp = a->spp;
    job.tiles_x = (a->width + 7) / 8;
    job.tri_idx = tri_idx; job.t_out = t;
    run_engine(H, job, job.tiles_x * ((a->height + 3) / 4) * 32, counts);
}

// rm_trace_closest / rm_trace_occluded: Model::rayHit / rayHit_test over a list of rays
void doh_trace_closest(const RmSceneDesc *sc, int n, const float *org, const float *dir, int32_t *tri_idx, float *t) {
    HostScene H(sc);
    ClosestJob job;
    job.org = org; job.dir = dir; job.aim_in = nullptr; job.tri_idx = tri_idx; job.t_out = t;
    run_engine(H, job, n, nullptr);
}
void doh_trace_occluded(const RmSceneDesc *sc, int n, const float *org, const float *dir, const float *aim, uint8_t *out) {
    HostScene H(sc);
    OccludedJob job;
    job.org = org; job.dir = dir; job.aim_in = aim; job.out = out;
    run_engine(H, job, n, nullptr);
}

// The same per-ray seam through the SECONDARY-RAY tree (what the estimator's bounce and shadow rays traverse): tune_fast of
// the context - vote weight 2 for the leaf step, 14 stack entries in "shared memory", deeper ones in the local spill array.
int doh_trace_closest_secondary(const RmSceneDesc *sc, int n, const float *org, const float *dir, int32_t *tri_idx, float *t, int smem_levels,
                                int32_t *shape4) {
    HostScene H(sc);
    if (!H.use_secondary_tree(sc)) return -1;
    ClosestJob job;
    job.org = org; job.dir = dir; job.aim_in = nullptr; job.tri_idx = tri_idx; job.t_out = t;
    run_engine(H, job, n, nullptr, smem_levels, 2);
    if (shape4) { shape4[0] = int32_t(H.fast_nodes.size() / 2); shape4[1] = H.levels; shape4[2] = smem_levels; shape4[3] = H.S.root_is_leaf; }
    return 0;
}
int doh_trace_occluded_secondary(const RmSceneDesc *sc, int n, const float *org, const float *dir, const float *aim, uint8_t *out, int smem_levels) {
    HostScene H(sc);
    if (!H.use_secondary_tree(sc)) return -1;
    OccludedJob job;
    job.org = org; job.dir = dir; job.aim_in = aim; job.out = out;
    run_engine(H, job, n, nullptr, smem_levels, 2);
    return 0;
}

// rm_gbuffer: k_gbuffer over the primary hits {tri_idx, t} - getHitInfo with ray differentials, normal mapping, trilinear
// material fetches, sky emission on a miss, the red nudge.  gbuffer = what the samplers read; sav_base = the un-nudged base colour.
void doh_gbuffer(const RmSceneDesc *sc, const RmRenderArgs *a, const int32_t *tri_idx, const float *t, RmHitInfo *gbuffer, float *sav_base, int32_t *n_ind) {
    HostScene H(sc);
    DevArgs A;
    A.position = ld(a->position); A.direction = ld(a->direction); A.up = ld(a->up); A.right = ld(a->right);
    A.accuracy = a->accuracy; A.exposure = a->exposure; A.P_Direct = a->P_Direct; A.width = a->width; A.height = a->height; A.spp = a->spp;
    const int npix = a->width * a->height;
    FrameBuffers Fb{gbuffer, sav_base, n_ind, nullptr};
    int glass = 0;
    const int spp_d = int(float(a->spp) * a->P_Direct);                    // src/render.cpp:500
    rm_host_launch(k_gbuffer, dim3((npix + 127) / 128), dim3(128), H.S, A, (const int *)tri_idx, t, Fb, spp_d, a->spp - spp_d, &glass);
}

// accumulateInwardRadiance (src/image.cpp:615-659) as the shading kernels apply it to one sample: accum_split into a zeroed
// {diffuse, specular} pair of {rgb, second moment}; s7 = {bsdfPdf rgb, light rgb, weight}
void doh_accumulate(int64_t n, const float *base, const float *s7, float *out) {
    for (int64_t i = 0; i < n; i++) {
        float *o = out + i * 8;
        for (int k = 0; k < 8; k++) o[k] = 0.0f;
        accum_split(o, o + 4, ld(base + i * 3), ld(s7 + i * 7), ld(s7 + i * 7 + 3), s7[i * 7 + 6]);
    }
}

// One direct-light sample per pixel (sample 0 of spp_direct = 1) exactly as k_direct_gen draws it - the pixel's surface
// record from the G-buffer, the per-light weights, nee_sample on the pixel's own Philox stream - then Model::rayHit_test on
// the shadow ray through the engine (ShadowJob over the items, on the reference's tree or the secondary-ray tree).
// out7[p] = {bsdfPdf rgb, light rgb, weight}; status[p]: 0 = no sample (miss, emissive pixel, no light, rejected), 1 = drawn
// but occluded, 2 = visible.
int doh_replay_direct(const RmSceneDesc *sc, const RmRenderArgs *a, const RmHitInfo *gbuffer, unsigned long long seed, float *out7, uint8_t *status,
                      int secondary_tree) {
    HostScene H(sc);
    const V3 cam = ld(a->position);
    const int npix = a->width * a->height;
    std::vector<ShadowItem> items(npix);
    for (int p = 0; p < npix; p++) {
        status[p] = 0;
        NeeOut n;
        n.valid = false;
        Bsdf B;
        B.s = load_hitinfo(gbuffer + p);
        if (isfinite_any(B.s.position) && !(length(B.s.emission) > 0.0f)) {
            B.inDir = -normalize(B.s.position - cam);
            float lw[kMaxLights], total = 0.0f;
            bool go = true;
            if (H.S.sky_width == 0) { total = light_weights(H.S, B, lw); go = total != 0.0f; }
            if (go) {
                Rng gen;
                gen.init(seed, unsigned(p), 0u, kStreamDirect);
                n = nee_sample(H.S, B, gen, lw, total, 1);
            }
        }
        if (!n.valid) { n.dir = splat3(0.0f); n.aim = CUDART_NAN_F; n.bsdf = n.light = splat3(0.0f); n.weight = 0.0f; }
        else status[p] = 1;
        write_shadow(&items[p], p | 0x80000000, B.s.position, n, n.bsdf, n.light, n.weight);
        float *o = out7 + size_t(p) * 7;
        o[0] = n.bsdf.x; o[1] = n.bsdf.y; o[2] = n.bsdf.z; o[3] = n.light.x; o[4] = n.light.y; o[5] = n.light.z; o[6] = n.weight;
    }
    if (secondary_tree && !H.use_secondary_tree(sc)) return -1;
    ShadowJob job;
    job.sq = items.data();
    run_engine(H, job, npix, nullptr, secondary_tree ? 14 : 0, secondary_tree ? 2 : 1);
    for (int p = 0; p < npix; p++)
        if (status[p] == 1 && items[p].vis != 0.0f) status[p] = 2;
    return 0;
}

// The direct-light half of rm_render_samples + rm_resolve as KERNELS, with the launch sequence of rm_render.cu: k_direct_gen
// (thread per pixel, or - `warp_per_pixel` - a warp per pixel as on environment-lit scenes) fills one contiguous block of the
// shadow queue per pixel, the engine runs ShadowJob over the queue, k_accum_direct folds the visible samples into the pixel's
// accumulators in sample order, k_finalise applies exposure and the variance formula.  Blocks are emulated as real threads with
// a CTA barrier and one warp context per 32 lanes.  spp direct samples per pixel in one wave.
int doh_direct_planes(const RmSceneDesc *sc, const RmRenderArgs *a, const RmHitInfo *gbuffer_in, unsigned long long seed, int spp, int warp_per_pixel,
                      int secondary_tree, RmRadiance *Dd, RmRadiance *Ds) {
    HostScene H(sc);
    DevArgs A;
    A.position = ld(a->position); A.direction = ld(a->direction); A.up = ld(a->up); A.right = ld(a->right);
    A.accuracy = a->accuracy; A.exposure = a->exposure; A.P_Direct = a->P_Direct; A.width = a->width; A.height = a->height; A.spp = spp;
    const int npix = a->width * a->height;
    std::vector<RmHitInfo> g(gbuffer_in, gbuffer_in + npix), g_out(npix);
    std::vector<float> sav(size_t(npix) * 3, 0.0f), rad(size_t(npix) * 16, 0.0f), clum_sum(size_t(npix) * 2, 0.0f), clum_max(npix, 0.0f), hold_clum(npix, -1.0f),
        hold(size_t(npix) * 8, 0.0f);
    std::vector<int> n_ind(npix, 0), dir_base(npix, -1), lock(npix, 0);
    FrameBuffers Fb{g.data(), sav.data(), n_ind.data(), dir_base.data()};
    Accum Ac{rad.data(), clum_sum.data(), clum_max.data(), hold_clum.data(), hold.data(), lock.data()};
    const int s_cap = npix * spp;
    std::vector<ShadowItem> sq(size_t(s_cap) + 1);
    int s_count = 0, overflow = 0;
    const dim3 grid(2), block(kShadeBlock);
    if (warp_per_pixel)
        rm_host_launch_blocks(k_direct_gen<true>, grid, block, H.S, A, Fb, spp, npix, 0, 1, spp, seed, sq.data(), &s_count, s_cap, &overflow);
    else
        rm_host_launch_blocks(k_direct_gen<false>, grid, block, H.S, A, Fb, spp, npix, 0, 1, spp, seed, sq.data(), &s_count, s_cap, &overflow);
    if (overflow || s_count > s_cap) return -2;
    if (secondary_tree && !H.use_secondary_tree(sc)) return -1;
    ShadowJob job;
    job.sq = sq.data();
    run_engine(H, job, s_count, nullptr, secondary_tree ? 14 : 0, secondary_tree ? 2 : 1);
    rm_host_launch(k_accum_direct, dim3((npix + 255) / 256), dim3(256), Fb, Ac, (const ShadowItem *)sq.data(), spp, npix);
    std::vector<RmRadiance> Id(npix), Is(npix);
    rm_host_launch(k_finalise, dim3((npix + 255) / 256), dim3(256), Ac, Fb, 0, npix, a->exposure, Dd, Ds, Id.data(), Is.data(), g_out.data());
    return s_count;
}

// glass pixels of the frame (rm_render.cu, k_glass_list): the pixels whose indirect sample count is 16x the base
static __global__ void k_glass_list_host(const int *n_ind, int npix, int base, int *list, int *count) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool glass = p < npix && base > 0 && n_ind[p] > base;
    int slot = alloc_slot(count, glass);
    if (glass) list[slot] = p;
}

// The indirect half of rm_render_samples + rm_resolve with the round loop of rm_render.cu restated around the kernels
// themselves: primary hits -> k_gbuffer -> [k_plan, k_regen, PathJob through the engine, k_surface, k_bounce, k_nee, k_shadow_gate,
// ShadowJob through the engine, k_accum_shadow]* -> k_publish_max, k_commit_hold (hold-back never dropped, like the
// "disable_clamp" option of the replay tests), k_finalise.  `spp` indirect samples per opaque pixel (16x on glass), no direct ones.
int doh_indirect_planes(const RmSceneDesc *sc, const RmRenderArgs *a, unsigned long long seed, int spp, int secondary_tree, RmHitInfo *g_out,
                        RmRadiance *Id, RmRadiance *Is, int32_t *rounds_out) {
    HostScene H(sc);
    DevArgs A;
    A.position = ld(a->position); A.direction = ld(a->direction); A.up = ld(a->up); A.right = ld(a->right);
    A.accuracy = a->accuracy; A.exposure = a->exposure; A.P_Direct = 0.0f; A.width = a->width; A.height = a->height; A.spp = spp;
    const int npix = a->width * a->height;
    // frame: primary hits and G-buffer
    std::vector<int> tri(npix, -1);
    std::vector<float> t(npix, 0.0f);
    {
        PrimaryJob job;
        job.A = A; job.tiles_x = (a->width + 7) / 8; job.tri_idx = tri.data(); job.t_out = t.data();
        run_engine(H, job, job.tiles_x * ((a->height + 3) / 4) * 32, nullptr);
    }
    std::vector<RmHitInfo> g(npix);
    std::vector<float> sav(size_t(npix) * 3, 0.0f);
    std::vector<int> n_ind(npix, 0), dir_base(npix, -1), glass_list(npix, 0);
    FrameBuffers Fb{g.data(), sav.data(), n_ind.data(), dir_base.data()};
    int C[C_TOTAL] = {0};
    const int base = spp;                                                   // spp_direct = 0
    rm_host_launch(k_gbuffer, dim3((npix + 127) / 128), dim3(128), H.S, A, (const int *)tri.data(), (const float *)t.data(), Fb, 0, base, C + 4);
    rm_host_launch_blocks(k_glass_list_host, dim3((npix + 127) / 128), dim3(128), (const int *)n_ind.data(), npix, base, glass_list.data(), C + 5);
    const int n_glass = C[5];
    // accumulators (reset_accum)
    std::vector<float> rad(size_t(npix) * 16, 0.0f), clum_sum(size_t(npix) * 2, 0.0f), clum_max(npix, 0.0f), hold_clum(npix, -1.0f), hold(size_t(npix) * 8, 0.0f);
    std::vector<int> lock(npix, 0);
    Accum Ac{rad.data(), clum_sum.data(), clum_max.data(), hold_clum.data(), hold.data(), lock.data()};
    // item space and queues (rm_render_samples)
    const int n_a = base, n_b_total = n_glass > 0 ? 16 * base : 0;
    const long long items_a = (long long)npix * n_a, items_b = n_b_total > n_a ? (long long)n_glass * (n_b_total - n_a) : 0, total_items = items_a + items_b;
    const int q_cap = int(std::max(1024LL, total_items)), s_cap = 6 * q_cap;
    const size_t words[20] = {1, 1, 1, 3, 3, 12, 3, 3, 1, 1, 1, size_t(kMediumSlots), size_t(4 * kMediumSlots), 1, 1, 22, 6, 7, 1, 1};
    std::vector<std::vector<uint32_t>> store[2];
    PathQueue Q[2];
    for (int w = 0; w < 2; w++) {
        store[w].resize(20);
        for (int k = 0; k < 20; k++) store[w][k].assign(size_t(q_cap) * words[k], 0u);
        auto at = [&](int k) { return static_cast<void *>(store[w][k].data()); };
        PathQueue &q = Q[w];
        q.cap = q_cap;
        q.pixel = (int *)at(0); q.sample = (unsigned *)at(1); q.drawn = (unsigned *)at(2);
        q.o = (float *)at(3); q.d = (float *)at(4); q.diff = (float *)at(5);
        q.T = (float *)at(6); q.B0 = (float *)at(7); q.W = (float *)at(8); q.rough = (float *)at(9);
        q.flags = (int *)at(10); q.med_id = (int *)at(11); q.med = (float *)at(12);
        q.hit_t = (float *)at(13); q.hit_face = (int *)at(14);
        q.surf = (float *)at(15); q.hdP = (float *)at(16);
        q.dec = (float *)at(17); q.skey = (int *)at(18); q.srank = (int *)at(19);
    }
    std::vector<int> sorted(q_cap);
    std::vector<ShadowItem> sq(size_t(s_cap) + 1);
    ItemSpace I;
    I.items_a = items_a; I.total = total_items; I.npix = npix; I.n_glass = std::max(n_glass, 1); I.n_a = n_a;
    I.glass_list = glass_list.data(); I.s_begin = 0; I.s_stride = 1;
    // the tree the bounce and shadow rays traverse
    HostScene Hsec(sc);
    if (secondary_tree && !Hsec.use_secondary_tree(sc)) return -1;
    const HostScene &T = secondary_tree ? Hsec : H;
    const int sl = secondary_tree ? 14 : 0, wl = secondary_tree ? 2 : 1;
    auto trace_shadow = [&] {
        ShadowJob job;
        job.sq = sq.data();
        run_engine(T, job, std::min(C[C_SQ_RUN], s_cap), nullptr, sl, wl);
        rm_host_launch_blocks(k_accum_shadow, dim3(2), dim3(256), Fb, Ac, (const ShadowItem *)sq.data(), (const int *)(C + C_SQ_RUN), s_cap);
    };
    const dim3 grid(2), block(kShadeBlock);
    const int shadow_threshold = std::max(1, q_cap / 2);
    int cur = 0, rounds = 0;
    const bool trace = std::getenv("RM_DOH_TRACE") != nullptr;
#define DOH_STEP(name) do { if (trace) std::fprintf(stderr, "round %d %s: q=%d/%d nee=%d sq=%d run=%d\n", rounds, name, C[0], C[1], C[C_NEE], C[C_SQ], C[C_SQ_RUN]); } while (0)
    if (total_items > 0) {
        for (;;) {
            PathQueue Qin = Q[cur], Qout = Q[cur ^ 1];
            rm_host_launch(k_plan, dim3(1), dim3(1), C, cur, Qin.cap, total_items);
            DOH_STEP("planned");
            rm_host_launch_blocks(k_regen, grid, block, H.S, A, Fb, I, (const int *)C, seed, Qin, C + cur);
            DOH_STEP("regenerated");
            PathJob pj;
            pj.Q = Qin;
            run_engine(T, pj, std::min(C[cur], Qin.cap), nullptr, sl, wl);
            DOH_STEP("traced");
            rm_host_launch_blocks(k_surface, grid, block, H.S, Fb, Ac, Qin, C, cur);
            DOH_STEP("surfaced");
            rm_host_launch_blocks(k_decide, grid, block, seed, Qin, (const int *)(C + cur), C + C_BINS);
            rm_host_launch_blocks(k_sort_offsets, dim3(1), dim3(kSortBins), (const int *)(C + C_BINS), C + C_OFFS);
            rm_host_launch_blocks(k_sort_scatter, grid, dim3(256), Qin, (const int *)(C + cur), (const int *)(C + C_OFFS), sorted.data());
            DOH_STEP("decided + sorted");
            rm_host_launch_blocks(k_continue<false>, grid, block, seed, Qin, Qout, C + (cur ^ 1), (const int *)sorted.data(), (const int *)(C + C_OFFS));
            rm_host_launch_blocks(k_continue<true>, grid, block, seed, Qin, Qout, C + (cur ^ 1), (const int *)sorted.data(), (const int *)(C + C_OFFS));
            DOH_STEP("continued");
            rm_host_launch_blocks(k_nee, grid, block, H.S, Fb, seed, Qin, (const int *)sorted.data(), (const int *)(C + C_OFFS), sq.data(), C + C_SQ, s_cap, C + C_OVERFLOW);
            DOH_STEP("nee drawn");
            rm_host_launch(k_shadow_gate, dim3(1), dim3(1), C, shadow_threshold, s_cap, 0);
            trace_shadow();
            DOH_STEP("shadows");
            cur ^= 1;
            rounds++;
            const long long handed = (long long)(unsigned)C[C_ITEM_LO] | ((long long)C[C_ITEM_HI] << 32);
            if ((handed >= total_items && C[cur] == 0) || rounds > 64) break;
        }
        rm_host_launch(k_plan, dim3(1), dim3(1), C, cur, q_cap, total_items);
        rm_host_launch(k_shadow_gate, dim3(1), dim3(1), C, shadow_threshold, s_cap, 1);
        trace_shadow();
    }
    if (C[C_OVERFLOW]) return -2;
    rm_host_launch(k_publish_max, dim3((npix + 255) / 256), dim3(256), Ac, npix);
    rm_host_launch(k_commit_hold, dim3((npix + 255) / 256), dim3(256), Ac, Fb, npix, true);
    std::vector<RmRadiance> Dd(npix), Ds(npix);
    rm_host_launch(k_finalise, dim3((npix + 127) / 128), dim3(128), Ac, Fb, 0, npix, a->exposure, Dd.data(), Ds.data(), Id, Is, g_out);
    if (rounds_out) *rounds_out = rounds;
    return n_glass;
}

// Photo::depthFeildBlur as dof_device runs it (rm_render.cu): k_dof_prepare, the reference's stable sort by camera distance on
// the host, k_dof_tile_lists (per-tile source lists, order kept), k_dof_gather (every destination replays its sources in order)
int doh_depth_field_blur(const RmHitInfo *G, const float *rgb_in, const float *cam3, float focus, float CoC, int w, int h, float *rgb_out) {
    const int npix = w * h;
    std::vector<float> depth(npix);
    std::vector<float2> src(npix);
    rm_host_launch(k_dof_prepare, dim3((npix + 127) / 128), dim3(128), G, ld(cam3), focus, CoC, npix, depth.data(), src.data());
    struct Px { int idx; float depth; };
    std::vector<Px> px(npix);
    int reach = 0;
    for (int i = 0; i < npix; i++) {
        px[i] = {i, depth[i]};
        if (src[i].x == src[i].x) reach = std::max(reach, int(src[i].x));
    }
    std::stable_sort(px.begin(), px.end(), [](const Px &a, const Px &b) { return a.depth < b.depth; });
    std::vector<int> sorted(npix);
    for (int i = 0; i < npix; i++) sorted[i] = px[i].idx;
    const int tiles_x = (w + kDofTile - 1) / kDofTile, tiles_y = (h + kDofTile - 1) / kDofTile, tiles = tiles_x * tiles_y;
    const long long side = kDofTile + 2LL * reach;
    const long long cap = std::min<long long>(npix, side * side);
    std::vector<int> lists(size_t(cap) * tiles), counts(tiles, 0);
    rm_host_launch_blocks(k_dof_tile_lists, dim3(tiles), dim3(256), (const int *)sorted.data(), npix, w, h, reach, tiles_x, int(cap), lists.data(), counts.data());
    rm_host_launch(k_dof_gather, dim3(tiles), dim3(kDofTile, kDofTile), rgb_in, (const float2 *)src.data(), (const int *)lists.data(), (const int *)counts.data(),
                   int(cap), w, h, tiles_x, rgb_out);
    return reach;
}

}
