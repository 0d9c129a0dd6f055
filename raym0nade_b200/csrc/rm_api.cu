// rm_api.cu — C ABI: context, scene staging, the per-ray seam and the primary-ray stage.
//
// No CPU fallback: every entry point that computes needs an sm_100 device and fails with
// RM_ERR_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <new>

#include "rm_context.cuh"
#include "kernels_trace.cuh"
#include "wide_bvh.h"
#include "gpu_bvh.h"

using namespace rm;

rm::DevArgs to_dev_args(const RmRenderArgs *a) {
    DevArgs d;
    d.position = {a->position[0], a->position[1], a->position[2]};
    d.direction = {a->direction[0], a->direction[1], a->direction[2]};
    d.up = {a->up[0], a->up[1], a->up[2]};
    d.right = {a->right[0], a->right[1], a->right[2]};
    d.accuracy = a->accuracy;
    d.exposure = a->exposure;
    d.P_Direct = a->P_Direct;
    d.width = a->width;
    d.height = a->height;
    d.spp = a->spp;
    return d;
}

int rm_check_args(const RmRenderArgs *a) {
    if (!a) return rm_fail(RM_ERR_INVALID, "render args are NULL");
    if (a->width <= 0 || a->height <= 0 || (int64_t)a->width * a->height > (int64_t)1 << 28)
        return rm_fail(RM_ERR_INVALID, "render args: bad image size %d x %d", a->width, a->height);
    if (a->spp < 0) return rm_fail(RM_ERR_INVALID, "render args: negative spp");
    return RM_OK;
}

static int upload(DevBuf &b, const void *src, size_t bytes, cudaStream_t st, int64_t &total) {
    int rc = b.alloc(bytes);
    if (rc) return rc;
    if (bytes) RM_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
    total += (int64_t)bytes;
    return RM_OK;
}

// test hook "seam_secondary_tree": the per-ray seam through the binary (1) or the 4-wide (2) secondary-ray tree; the seam
// itself (0) is the reference's tree in the reference's order
static int seam_pick(const RmContext *ctx) {          // 0 the reference's tree, 1 the binary secondary-ray tree, 2 its 4-wide form
    if (ctx->seam_tree == 2 && ctx->have_wide) return 2;
    if (ctx->seam_tree && ctx->have_fast) return 1;
    return ctx->seam_tree && ctx->have_wide ? 2 : 0;
}
static const DevScene &seam_scene(const RmContext *ctx) { const int k = seam_pick(ctx); return k == 2 ? ctx->scene_wide : (k ? ctx->scene_fast : ctx->scene); }
static int seam_levels(const RmContext *ctx) { const int k = seam_pick(ctx); return k == 2 ? ctx->stack_levels_wide : (k ? ctx->stack_levels_fast : ctx->stack_levels); }
static TraceTune seam_tune(const RmContext *ctx) { const int k = seam_pick(ctx); return k == 2 ? ctx->tune_wide : (k ? ctx->tune_fast : ctx->tune); }

extern "C" {

int rm_context_create(int device, void *stream, RmContext **out) {
    if (!out) return rm_fail(RM_ERR_INVALID, "rm_context_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return rm_fail(RM_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return rm_fail(RM_ERR_INVALID, "rm_context_create: device %d out of range (have %d)", device, n);
    RM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return rm_fail(RM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    RmContext *ctx = new (std::nothrow) RmContext();
    if (!ctx) return rm_fail(RM_ERR_INVALID, "out of host memory");
    ctx->device = device;
    ctx->stream = static_cast<cudaStream_t>(stream);
    int rc = ctx->b_counters.alloc(16 * sizeof(unsigned long long));   // 3 kernel kinds x {rays, box, tri}
    if (!rc) rc = ctx->b_cursor.alloc(4 * sizeof(int));
    if (rc) { delete ctx; return rc; }
    ctx->sm_count = prop.multiProcessorCount;
    cudaMemsetAsync(ctx->b_counters.p, 0, ctx->b_counters.bytes, ctx->stream);
    *out = ctx;
    return RM_OK;
}

void rm_context_destroy(RmContext *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rm_render_state_free(ctx);
    rm_comm_state_free(ctx);
    delete ctx;
}

int rm_context_synchronize(RmContext *ctx) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "context is NULL");
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    return RM_OK;
}

// Per face: the traversal record {v0, e1, e2, |e1|, cut-out flag} - the fp32 values RayTriangleIntersection forms per test
// (src/geometry.cpp:65-70), with the reference's operation order - and the shading record (positions, uvs, normals, material).
__global__ void k_pack_faces(const float *__restrict__ pos, const float *__restrict__ uv, const float *__restrict__ nrm, const int *__restrict__ mat,
                             const DevMaterial *__restrict__ mats, int n, float4 *__restrict__ tri, float4 *__restrict__ shade) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = pos + size_t(i) * 9;
    float v[9];
#pragma unroll
    for (int k = 0; k < 9; k++) v[k] = p[k];
    const V3 e1 = mk3(fsub(v[3], v[0]), fsub(v[4], v[1]), fsub(v[5], v[2])), e2 = mk3(fsub(v[6], v[0]), fsub(v[7], v[1]), fsub(v[8], v[2]));
    const float len = length(e1);                                  // glm::length(edge1)
    const int m = mat[i];
    float4 *t = tri + size_t(i) * kTriStride;
    t[0] = make_float4(v[0], v[1], v[2], e1.x);
    t[1] = make_float4(e1.y, e1.z, e2.x, e2.y);
    t[2] = make_float4(e2.z, len, mats[m].cutout ? 1.0f : 0.0f, 0.0f);
    if (kTriStride > 3) t[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float *u = uv + size_t(i) * 6, *nn = nrm + size_t(i) * 9;
    float4 *s = shade + size_t(i) * 7;
    s[0] = make_float4(v[0], v[1], v[2], v[3]);
    s[1] = make_float4(v[4], v[5], v[6], v[7]);
    s[2] = make_float4(v[8], u[0], u[1], u[2]);
    s[3] = make_float4(u[3], u[4], u[5], nn[0]);
    s[4] = make_float4(nn[1], nn[2], nn[3], nn[4]);
    s[5] = make_float4(nn[5], nn[6], nn[7], nn[8]);
    s[6] = make_float4(__int_as_float(m), 0.0f, 0.0f, 0.0f);
}

// the traversal records in the order of the secondary-ray tree's leaves
__global__ void k_permute_tris(const float4 *__restrict__ tri, const int *__restrict__ order, int n, float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *s = tri + size_t(order[i]) * kTriStride;
    float4 *d = out + size_t(i) * kTriStride;
#pragma unroll
    for (int k = 0; k < kTriStride; k++) d[k] = s[k];
}

static uint64_t hash_words(const void *p, size_t bytes) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    uint64_t h = 0x9E3779B97F4A7C15ull ^ bytes;
    for (size_t i = 0; i < bytes; i += 8) {          // memcpy: float arrays are only 4-byte aligned; the tail word is zero-padded
        uint64_t w = 0;
        std::memcpy(&w, b + i, bytes - i < 8 ? bytes - i : 8);
        h ^= w; h *= 0x100000001B3ull; h ^= h >> 29;
    }
    return h;
}

int rm_scene_upload(RmContext *ctx, const RmSceneDesc *sc) {
    if (!ctx || !sc) return rm_fail(RM_ERR_INVALID, "rm_scene_upload: null argument");
    int rc = rm_scene_validate(sc);          // tree, index and pointer consistency (scene_check.cpp): nothing unchecked reaches the device
    if (rc) return rc;
    RM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int64_t total = 0;
    const int n = sc->n_faces;
    // RM_TIMING=1: wall-clock per phase of the upload on stderr (synchronises the stream at every mark)
    static const bool timing = std::getenv("RM_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "rm_scene_upload: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    mark("validate");
    // the buffers of the previous scene are reused (and reallocated on growth) from here on: until this upload has
    // finished there is no scene, so a failure half-way cannot leave a later call traversing freed or half-written memory
    ctx->has_scene = ctx->have_primary = ctx->have_gbuffer = ctx->have_resolved = false;

    // nodes: same 32-byte images, heap-indexed; the array is padded to an even count so that
    // every child pair (2u, 2u+1) is a whole 64-byte block
    // (straight from the caller's array - a DMA when it is page-locked, rm_prepared_pin -, the padding record or two zeroed on the device)
    {
        const size_t padded = size_t(sc->n_nodes + 2) & ~size_t(1);
        if ((rc = ctx->b_nodes.alloc(padded * sizeof(RmBvhNode)))) return rc;
        RM_CUDA(cudaMemcpyAsync(ctx->b_nodes.p, sc->nodes, sizeof(RmBvhNode) * size_t(sc->n_nodes), cudaMemcpyHostToDevice, st));
        RM_CUDA(cudaMemsetAsync(ctx->b_nodes.as<RmBvhNode>() + sc->n_nodes, 0, (padded - size_t(sc->n_nodes)) * sizeof(RmBvhNode), st));
        total += int64_t(sizeof(RmBvhNode) * size_t(sc->n_nodes));
    }

    // materials (needed first for the per-triangle cut-out flag)
    std::vector<DevMaterial> mats(std::max(sc->n_materials, 1));
    bool any_cutout = false;
    for (int i = 0; i < sc->n_materials; i++) {
        const RmMaterialDesc &m = sc->materials[i];
        DevMaterial &d = mats[i];
        for (int k = 0; k < 4; k++) {
            if (m.tex[k] >= sc->n_textures) return rm_fail(RM_ERR_INVALID, "material %d: texture index out of range", i);
            d.tex[k] = m.tex[k] < 0 ? -1 : m.tex[k];
        }
        d.opacity = m.opacity; d.ior = m.ior; d.roughness = m.roughness;
        for (int k = 0; k < 3; k++) d.tc[k] = m.transmitting_color[k];
        d.cutout = m.has_fully_transparent_part ? 1 : 0;
        d._pad = 0;
        any_cutout |= d.cutout != 0;
    }
    if ((rc = upload(ctx->b_mats, mats.data(), mats.size() * sizeof(DevMaterial), st, total))) return rc;

    // traversal records (48 B) and shading records (112 B) are formed on the device from the caller's arrays as they
    // are (k_pack_faces); the material indices were checked by rm_scene_validate
    if ((rc = upload(ctx->b_raw[0], sc->positions, size_t(n) * 36, st, total))) return rc;
    if ((rc = upload(ctx->b_raw[1], sc->uvs, size_t(n) * 24, st, total))) return rc;
    if ((rc = upload(ctx->b_raw[2], sc->normals, size_t(n) * 36, st, total))) return rc;
    if ((rc = upload(ctx->b_raw[3], sc->face_material, size_t(n) * 4, st, total))) return rc;
    if ((rc = ctx->b_tri.alloc(size_t(n) * 16 * kTriStride)) || (rc = ctx->b_shade.alloc(size_t(n) * 112))) return rc;
    k_pack_faces<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_raw[0].as<float>(), ctx->b_raw[1].as<float>(), ctx->b_raw[2].as<float>(), ctx->b_raw[3].as<int>(),
                                                  ctx->b_mats.as<DevMaterial>(), n, ctx->b_tri.as<float4>(), ctx->b_shade.as<float4>());
    ctx->launches++;
    RM_CUDA(cudaGetLastError());
    mark("faces: H2D + records");
    for (int a = 0; a < 3; a++) ctx->scene_lo[a] = sc->nodes[1].v0[a];

    // The secondary-ray tree.  Default: built on the device from the positions just uploaded (gpu_sah_bvh.cu: top-down sweep
    // SAH, every level a few scans over all triangles, then the 4-wide collapse - ~8 ms per million triangles, so every upload
    // rebuilds it and nothing is cached; gpu_bvh.cu: Morton sort + PLOC clustering, 3 ms for a tree of ~14 % more visits).  The host
    // builders (fast_bvh.cpp: binned SAH; wide_bvh.cpp: its collapse) remain for "tree_builder" 0, for the binary form of the
    // tree ("secondary_tree" 1 / the seam test hook, asked for before the upload) and as the fallback should the device
    // tree come out deeper than the traversal stack allows; they are cached by geometry hash.
    {
        bool device_tree = false;
        ctx->have_fast = false;
        // The sweep-SAH build is deferred to the first call that needs the tree (rm_ensure_secondary_tree: rm_render_samples with
        // samples to draw, rm_tree_info): a scene that is only ever hit by primary rays (previews, G-buffers, BASELINE's configs[4])
        // never pays for it, and the positions it works from stay on the device anyway (b_raw[0], what rm_scene_refit replaces).
        ctx->wide_pending = ctx->tree_builder_mode == 3 && ctx->lazy_tree && ctx->seam_tree == 0 && !ctx->want_binary_tree;
        if (ctx->wide_pending) {
            ctx->have_wide = false;
            device_tree = true;
        } else if (ctx->tree_builder_mode >= 1) {
            int wlevels = 0, wnodes = 0;
            const RmBvhNode &rootbox = sc->nodes[1];           // the reference tree's root box = the scene bounds
            if (ctx->tree_builder_mode == 3) rc = rm_gpu_build_wide_sah(ctx, ctx->b_raw[0].as<float>(), n, ctx->fast_depth_cap, rootbox.v0, &wlevels, &wnodes);
            else rc = rm_gpu_build_wide(ctx, ctx->b_raw[0].as<float>(), n, rootbox.v0, rootbox.v1, &wlevels, &wnodes);
            if (rc) return rc;
            mark("device tree build");
            if (3 * wlevels <= ctx->tune_wide.smem_levels + rm::kStackSpillWide) {
                ctx->stack_levels_wide = std::max(3 * wlevels, 2);
                ctx->have_wide = true;
                ctx->wide_nodes = wnodes;
                ctx->wide_levels = wlevels;
                ctx->wide_built_by = ctx->tree_builder_mode == 3 ? 3 : 1;
                device_tree = true;
            }
        }
        // "tree_builder" 2: the device tree serves at once; the host SAH builder refines it in the background
        ctx->use_refined = false;
        if (device_tree && ctx->tree_builder_mode == 2) {
            const uint64_t key = ctx->tree_cache ? hash_words(sc->positions, size_t(n) * 36) : 0;      // (only ever compared while caching is on)
            if (ctx->refine && !(ctx->tree_cache && ctx->refine->started && ctx->refine->key == key && ctx->refine->n == n)) {
                ctx->refine->stop();                           // a build for other geometry (or caching is off): it unwinds within milliseconds
                ctx->refine.reset();
            }
            if (ctx->tree_cache && ctx->refined_installed && ctx->refined_key == key && ctx->refined_n == n) ctx->use_refined = true;
            else if (!ctx->refine) {
                ctx->refined_installed = false;
                ctx->refine.reset(new RefineJob());
                RefineJob *J = ctx->refine.get();
                J->pos.assign(sc->positions, sc->positions + size_t(n) * 9);
                J->n = n;
                J->key = key;
                J->limit = ctx->tune_wide.smem_levels + rm::kStackSpillWide;
                J->depth_cap = ctx->fast_depth_cap;
                // (the thread itself is launched by rm_start_refinement, from the first render that is long enough to profit)
            }
        }
        if (!device_tree || ctx->want_binary_tree) {
            const uint64_t key = hash_words(sc->positions, size_t(n) * 36);
            if (!(ctx->tree_cache && ctx->fast_key_valid && ctx->fast_key == key && ctx->fast_n == n && (device_tree || ctx->host_wide_valid))) {
                std::vector<RmBvhNode> fnodes;
                std::vector<int32_t> forder;
                int fdepth = 0;
                if ((rc = rm_build_fast_bvh(sc->positions, n, ctx->fast_depth_cap, ctx->fast_leaf_max, fnodes, forder, &fdepth))) return rc;
                if ((rc = upload(ctx->b_nodes_fast, fnodes.data(), fnodes.size() * sizeof(RmBvhNode), st, total))) return rc;
                if ((rc = upload(ctx->b_facemap, forder.data(), forder.size() * 4, st, total))) return rc;
                RM_CUDA(cudaStreamSynchronize(st));            // the host vectors die at scope exit
                ctx->stack_levels_fast = std::min(std::max(fdepth, 2), 40);
                ctx->fast_root_is_leaf = fnodes[1].faceR != 0;
                ctx->host_wide_valid = false;
                if (!device_tree) {
                    ctx->have_wide = false;
                    if (ctx->fast_leaf_max <= 3) {
                        std::vector<RmWideNode> wnodes;
                        std::vector<int32_t> worder;
                        int wdepth = 0;
                        if ((rc = rm_build_wide_bvh(fnodes, forder, n, wnodes, worder, &wdepth))) return rc;
                        if (3 * wdepth <= ctx->tune_wide.smem_levels + rm::kStackSpillWide) {       // up to three deferred children per level
                            if ((rc = upload(ctx->b_nodes_wide, wnodes.data(), wnodes.size() * sizeof(RmWideNode), st, total))) return rc;
                            if ((rc = upload(ctx->b_facemap_wide, worder.data(), worder.size() * 4, st, total))) return rc;
                            RM_CUDA(cudaStreamSynchronize(st));
                            ctx->stack_levels_wide = std::max(3 * wdepth, 2);
                            ctx->have_wide = true;
                            ctx->host_wide_valid = true;
                            ctx->wide_nodes = int(wnodes.size());
                            ctx->wide_levels = wdepth;
                        }
                    }
                }
                ctx->fast_key = key; ctx->fast_n = n; ctx->fast_key_valid = true;
            } else if (!device_tree) ctx->have_wide = ctx->host_wide_valid;
            ctx->have_fast = true;
            if ((rc = ctx->b_tri_fast.alloc(size_t(n) * 16 * kTriStride))) return rc;
            k_permute_tris<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_tri.as<float4>(), ctx->b_facemap.as<int>(), n, ctx->b_tri_fast.as<float4>());
            ctx->launches++;
        }
        if (ctx->have_wide) {
            if ((rc = ctx->b_tri_wide.alloc(size_t(n) * 16 * kTriStride))) return rc;
            k_permute_tris<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_tri.as<float4>(), ctx->b_facemap_wide.as<int>(), n, ctx->b_tri_wide.as<float4>());
            ctx->launches++;
        }
        if (ctx->use_refined) {          // same geometry as the refined tree on the device: its triangle records follow the new materials
            k_permute_tris<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_tri.as<float4>(), ctx->b_facemap_wide2.as<int>(), n, ctx->b_tri_wide2.as<float4>());
            ctx->launches++;
        }
        RM_CUDA(cudaGetLastError());
        mark("refine job / host trees");
    }

    // textures: one blob, each level 16-byte aligned, copied level by level straight from the caller's memory
    std::vector<DevTexture> texs(std::max(sc->n_textures, 1));
    size_t blob_bytes = 0;
    for (int i = 0; i < sc->n_textures; i++) {
        const RmTextureDesc &t = sc->textures[i];
        if (t.map_depth < 1 || t.map_depth > 8 || (t.channels != 3 && t.channels != 4))
            return rm_fail(RM_ERR_INVALID, "texture %d: bad map_depth/channels", i);
        DevTexture &d = texs[i];
        d.width = t.width; d.height = t.height; d.channels = t.channels; d.map_depth = t.map_depth;
        for (int l = 0; l < 8; l++) {
            d.offset[l] = 0;
            if (l >= t.map_depth) continue;
            const size_t bytes = size_t(t.width >> l) * (t.height >> l) * t.channels;
            const size_t off = (blob_bytes + 15) & ~size_t(15);
            if (off + bytes > 0xFFFFFFFFull) return rm_fail(RM_ERR_INVALID, "texture data exceeds 4 GiB");
            blob_bytes = off + bytes;
            d.offset[l] = uint32_t(off);
        }
    }
    blob_bytes = (blob_bytes + 15) & ~size_t(15);
    if ((rc = upload(ctx->b_texs, texs.data(), texs.size() * sizeof(DevTexture), st, total))) return rc;
    if ((rc = ctx->b_texels.alloc(blob_bytes))) return rc;
    for (int i = 0; i < sc->n_textures; i++)
        for (int l = 0; l < sc->textures[i].map_depth; l++) {
            const RmTextureDesc &t = sc->textures[i];
            const size_t bytes = size_t(t.width >> l) * (t.height >> l) * t.channels;
            if (!bytes) continue;
            RM_CUDA(cudaMemcpyAsync(ctx->b_texels.as<uint8_t>() + texs[i].offset[l], t.levels[l], bytes, cudaMemcpyHostToDevice, st));
            total += int64_t(bytes);
        }

    // lights
    std::vector<DevLight> lights(std::max(sc->n_lights, 1));
    std::vector<float> lpos, lnrm, lcdf;
    for (int i = 0; i < sc->n_lights; i++) {
        const RmLightDesc &L = sc->lights[i];
        DevLight &d = lights[i];
        for (int k = 0; k < 3; k++) { d.center[k] = L.center[k]; d.color[k] = L.color[k]; }
        d.power = L.power;
        d.n_faces = L.n_faces;
        d.face_offset = int32_t(lcdf.size());
        lpos.insert(lpos.end(), L.face_positions, L.face_positions + size_t(L.n_faces) * 9);
        lnrm.insert(lnrm.end(), L.face_normals, L.face_normals + size_t(L.n_faces) * 9);
        lcdf.insert(lcdf.end(), L.face_cdf, L.face_cdf + L.n_faces);
    }
    if ((rc = upload(ctx->b_lights, lights.data(), lights.size() * sizeof(DevLight), st, total))) return rc;
    if ((rc = upload(ctx->b_lpos, lpos.data(), lpos.size() * 4, st, total))) return rc;
    if ((rc = upload(ctx->b_lnrm, lnrm.data(), lnrm.size() * 4, st, total))) return rc;
    if ((rc = upload(ctx->b_lcdf, lcdf.data(), lcdf.size() * 4, st, total))) return rc;

    // sky
    size_t nsky = size_t(sc->sky_width) * sc->sky_height;
    if (nsky && (!sc->sky_data || !sc->sky_cdf)) return rm_fail(RM_ERR_INVALID, "sky size set but sky_data/sky_cdf missing");
    if ((rc = upload(ctx->b_sky, sc->sky_data, nsky * 12, st, total))) return rc;
    if ((rc = upload(ctx->b_skycdf, sc->sky_cdf, nsky * 4, st, total))) return rc;
    // brackets for the sky CDF search (dev_bsdf.cuh cdf_sample_guided): guide[j] = lower_bound(cdf, total * (j / G))
    std::vector<int32_t> guide(kSkyGuide + 1, 0);
    if (nsky) {
        const float *cdf = sc->sky_cdf, tot = cdf[nsky - 1];
        for (int j = 0; j <= kSkyGuide; j++) {
            const float x = tot * (float(j) / float(kSkyGuide));
            guide[j] = int32_t(std::lower_bound(cdf, cdf + nsky, x) - cdf);
        }
    }
    if ((rc = upload(ctx->b_skyguide, guide.data(), guide.size() * 4, st, total))) return rc;
    // k / 255.0f, correctly rounded: the RGBA8 decode table (dev_texture.cuh)
    float lut[256];
    for (int k = 0; k < 256; k++) lut[k] = float(k) / 255.0f;
    if ((rc = upload(ctx->b_lut, lut, sizeof(lut), st, total))) return rc;
    RM_CUDA(cudaStreamSynchronize(st));    // host staging vectors die at scope exit
    mark("textures, lights, sky");

    DevScene &S = ctx->scene;
    S.nodes = ctx->b_nodes.as<float4>();
    S.tri = ctx->b_tri.as<float4>();
    S.shade = ctx->b_shade.as<float4>();
    S.materials = ctx->b_mats.as<DevMaterial>();
    S.textures = ctx->b_texs.as<DevTexture>();
    S.texels = ctx->b_texels.as<uint8_t>();
    S.lights = ctx->b_lights.as<DevLight>();
    S.light_pos = ctx->b_lpos.as<float>();
    S.light_nrm = ctx->b_lnrm.as<float>();
    S.light_cdf = ctx->b_lcdf.as<float>();
    S.sky_data = ctx->b_sky.as<float>();
    S.sky_cdf = ctx->b_skycdf.as<float>();
    S.sky_guide = ctx->b_skyguide.as<int32_t>();
    S.div255 = ctx->b_lut.as<float>();
    S.n_faces = n;
    S.n_nodes = sc->n_nodes;
    S.n_materials = sc->n_materials;
    S.n_lights = sc->n_lights;
    S.sky_width = sc->sky_width;
    S.sky_height = sc->sky_height;
    S.any_cutout = any_cutout ? 1 : 0;
    S.root_is_leaf = sc->nodes[1].faceR != 0;
    S.explicit_children = 0;
    S.face_map = nullptr;
    S.wide = 0;
    // deferred children per ray <= inner levels of the heap-indexed tree (node indices < n_nodes)
    int levels = 1;
    while ((int64_t(1) << levels) < int64_t(sc->n_nodes)) levels++;
    ctx->stack_levels = std::min(std::max(levels, 2), 40);
    ctx->scene_fast = S;
    ctx->scene_fast.nodes = ctx->b_nodes_fast.as<float4>();
    ctx->scene_fast.tri = ctx->b_tri_fast.as<float4>();
    ctx->scene_fast.face_map = ctx->b_facemap.as<int32_t>();
    ctx->scene_fast.explicit_children = 1;
    ctx->scene_fast.root_is_leaf = ctx->fast_root_is_leaf ? 1 : 0;
    if (!ctx->have_fast) ctx->scene_fast = S;        // not built: whoever asks for it gets the reference's tree
    ctx->scene_wide = S;
    if (ctx->have_wide) {
        ctx->scene_wide.nodes = ctx->b_nodes_wide.as<float4>();
        ctx->scene_wide.tri = ctx->b_tri_wide.as<float4>();
        ctx->scene_wide.face_map = ctx->b_facemap_wide.as<int32_t>();
        ctx->scene_wide.root_is_leaf = 0;
        ctx->scene_wide.wide = 1;
        ctx->scene_wide_first = ctx->scene_wide;
        if (ctx->use_refined) {
            ctx->scene_wide.nodes = ctx->b_nodes_wide2.as<float4>();
            ctx->scene_wide.tri = ctx->b_tri_wide2.as<float4>();
            ctx->scene_wide.face_map = ctx->b_facemap_wide2.as<int32_t>();
            ctx->stack_levels_wide = std::max(3 * ctx->refined_levels, 2);
        }
    }
    ctx->scene_h2d_bytes = total;
    ctx->scene_bytes = 0;
    for (const DevBuf *b : {&ctx->b_nodes, &ctx->b_tri, &ctx->b_shade, &ctx->b_mats, &ctx->b_texs, &ctx->b_texels, &ctx->b_lights, &ctx->b_lpos,
                            &ctx->b_lnrm, &ctx->b_lcdf, &ctx->b_sky, &ctx->b_skycdf, &ctx->b_skyguide, &ctx->b_lut})
        ctx->scene_bytes += int64_t(b->bytes);
    ctx->has_scene = true;
    ctx->have_primary = ctx->have_gbuffer = ctx->have_resolved = false;
    return RM_OK;
}

// The deferred build of the secondary-ray tree (see rm_scene_upload): from the positions on the device, whatever rm_scene_refit
// has made of them since.
int rm_ensure_secondary_tree(RmContext *ctx) {
    if (!ctx || !ctx->wide_pending || !ctx->has_scene) return RM_OK;
    RM_CUDA(cudaSetDevice(ctx->device));          // (rm_tree_info and rm_set_option reach this without having done so)
    ctx->wide_pending = false;
    cudaStream_t st = ctx->stream;
    const int n = ctx->scene.n_faces;
    int rc, wlevels = 0, wnodes = 0;
    if ((rc = rm_gpu_build_wide_sah(ctx, ctx->b_raw[0].as<float>(), n, ctx->fast_depth_cap, ctx->scene_lo, &wlevels, &wnodes))) return rc;
    // (the builder's depth rule keeps the binary tree within the cap + 2 levels, and a wide level spans at least one binary level)
    if (3 * wlevels > ctx->tune_wide.smem_levels + rm::kStackSpillWide)
        return rm_fail(RM_ERR_STATE, "rm_ensure_secondary_tree: %d levels exceed the traversal stack", wlevels);
    ctx->stack_levels_wide = std::max(3 * wlevels, 2);
    ctx->wide_nodes = wnodes;
    ctx->wide_levels = wlevels;
    ctx->wide_built_by = 3;
    ctx->host_wide_valid = false;
    if ((rc = ctx->b_tri_wide.alloc(size_t(n) * 16 * kTriStride))) return rc;
    k_permute_tris<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_tri.as<float4>(), ctx->b_facemap_wide.as<int>(), n, ctx->b_tri_wide.as<float4>());
    ctx->launches += 2;
    RM_CUDA(cudaGetLastError());
    ctx->have_wide = true;
    ctx->scene_wide = ctx->scene;
    ctx->scene_wide.nodes = ctx->b_nodes_wide.as<float4>();
    ctx->scene_wide.tri = ctx->b_tri_wide.as<float4>();
    ctx->scene_wide.face_map = ctx->b_facemap_wide.as<int32_t>();
    ctx->scene_wide.root_is_leaf = 0;
    ctx->scene_wide.wide = 1;
    ctx->scene_wide_first = ctx->scene_wide;
    return RM_OK;
}

// After rm_scene_refit (gpu_ref_bvh.cu) has put new positions into b_raw[0]: the traversal and shading records are formed anew and
// re-permuted for the 4-wide tree(s).  The binary form of the secondary-ray tree is not refitted: whoever asks for it afterwards
// gets the reference's tree.
int rm_repack_faces(RmContext *ctx) {
    cudaStream_t st = ctx->stream;
    const int n = ctx->scene.n_faces;
    k_pack_faces<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_raw[0].as<float>(), ctx->b_raw[1].as<float>(), ctx->b_raw[2].as<float>(), ctx->b_raw[3].as<int>(),
                                                  ctx->b_mats.as<DevMaterial>(), n, ctx->b_tri.as<float4>(), ctx->b_shade.as<float4>());
    ctx->launches++;
    if (ctx->have_wide) {
        k_permute_tris<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_tri.as<float4>(), ctx->b_facemap_wide.as<int>(), n, ctx->b_tri_wide.as<float4>());
        ctx->launches++;
        if (ctx->refined_installed) {
            k_permute_tris<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_tri.as<float4>(), ctx->b_facemap_wide2.as<int>(), n, ctx->b_tri_wide2.as<float4>());
            ctx->launches++;
        }
    }
    ctx->have_fast = false;
    ctx->scene_fast = ctx->scene;
    RM_CUDA(cudaGetLastError());
    RM_CUDA(cudaStreamSynchronize(st));
    return RM_OK;
}

// The pending refinement of the secondary-ray tree starts here, not in rm_scene_upload: a build costs ~0.4 s of several host threads
// and pays back ~1.5 % of the traversal that follows it, so it is only worth starting when the scene is going to be rendered for
// longer than that - always when trees are cached across uploads (a static scene re-rendered), and with the cache off (a scene that
// changes every frame) only for a frame of at least kRefineMinSamples pixel-samples (~0.6 s on a B200).  Eight ranks of a host
// rendering 0.3 s frames used to spend each frame waiting for the previous frame's useless build to end.
constexpr int64_t kRefineMinSamples = 500000000;
void rm_start_refinement(RmContext *ctx, int64_t pixel_samples) {
    if (!ctx->refine || ctx->refine->started) return;
    if (!ctx->tree_cache && pixel_samples < kRefineMinSamples) { ctx->refine.reset(); return; }
    RefineJob *J = ctx->refine.get();
    J->started = true;
    J->state.store(1);
    J->th = std::thread([J] {
        std::vector<RmBvhNode> fnodes;
        std::vector<int32_t> forder;
        int fdepth = 0;
        bool ok = rm_build_fast_bvh_cancellable(J->pos.data(), J->n, J->depth_cap, 3, fnodes, forder, &fdepth, &J->cancel) == RM_OK && !J->cancel.load() &&
                  rm_build_wide_bvh(fnodes, forder, J->n, J->wnodes, J->worder, &J->wdepth) == RM_OK && 3 * J->wdepth <= J->limit;
        J->pos = std::vector<float>();
        J->state.store(ok && !J->cancel.load() ? 2 : 3);
    });
}

// The background build has finished: stage its tree next to the device builder's and point bounce / shadow rays at it.
// Called where the render loop waits for the device anyway (rm_render_samples: on entry and between batches of rounds).
int rm_install_refined_tree(RmContext *ctx) {
    if (!ctx || !ctx->refine || !ctx->has_scene) return RM_OK;
    RefineJob *J = ctx->refine.get();
    if (!J->started) return RM_OK;
    const int state = J->state.load();
    if (state == 1) return RM_OK;
    if (J->th.joinable()) J->th.join();
    std::unique_ptr<RefineJob> job = std::move(ctx->refine);
    if (state != 2 || job->discard || job->n != ctx->scene.n_faces || !ctx->have_wide) return RM_OK;
    RM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = job->n;
    int rc;
    int64_t total = 0;
    if ((rc = upload(ctx->b_nodes_wide2, job->wnodes.data(), job->wnodes.size() * sizeof(RmWideNode), st, total))) return rc;
    if ((rc = upload(ctx->b_facemap_wide2, job->worder.data(), job->worder.size() * 4, st, total))) return rc;
    if ((rc = ctx->b_tri_wide2.alloc(size_t(n) * 16 * kTriStride))) return rc;
    k_permute_tris<<<(n + 255) / 256, 256, 0, st>>>(ctx->b_tri.as<float4>(), ctx->b_facemap_wide2.as<int>(), n, ctx->b_tri_wide2.as<float4>());
    ctx->launches++;
    RM_CUDA(cudaGetLastError());
    RM_CUDA(cudaStreamSynchronize(st));            // the host vectors die with the job
    ctx->scene_wide.nodes = ctx->b_nodes_wide2.as<float4>();
    ctx->scene_wide.tri = ctx->b_tri_wide2.as<float4>();
    ctx->scene_wide.face_map = ctx->b_facemap_wide2.as<int32_t>();
    ctx->refined_levels = job->wdepth;
    ctx->refined_nodes = int(job->wnodes.size());
    ctx->stack_levels_wide = std::max(3 * job->wdepth, 2);
    ctx->refined_installed = true;
    ctx->use_refined = true;
    ctx->refined_key = job->key;
    ctx->refined_n = n;
    return RM_OK;
}

// {1 when the 4-wide tree came from the device builder (0: host), its 64-byte records, its levels, 1 when bounce / shadow rays use it}
int rm_tree_info(const RmContext *ctx_in, int32_t out[4]) {
    if (!ctx_in || !out) return rm_fail(RM_ERR_INVALID, "rm_tree_info: null argument");
    if (!ctx_in->has_scene) return rm_fail(RM_ERR_STATE, "rm_tree_info: no scene uploaded");
    RmContext *ctx = const_cast<RmContext *>(ctx_in);          // a deferred build happens now: the caller asks what the rays will traverse
    const int rc_build = rm_ensure_secondary_tree(ctx);
    if (rc_build) return rc_build;
    out[0] = ctx->have_wide && !ctx->host_wide_valid ? (ctx->use_refined ? 2 : ctx->wide_built_by == 3 ? 3 : 1) : 0;
    out[1] = ctx->use_refined ? ctx->refined_nodes : ctx->wide_nodes;
    out[2] = ctx->use_refined ? ctx->refined_levels : ctx->wide_levels;
    out[3] = ctx->have_wide && !ctx->exact_secondary && (ctx->secondary_tree == 2 || !ctx->have_fast) ? 1 : 0;
    return RM_OK;
}

// Page-lock (on = 1) / release (on = 0) the arrays of a prepared scene, so that every later rm_scene_upload of it is a DMA
// straight out of them instead of a staged copy of pageable memory.  A few tens of milliseconds once; worth it for a scene
// that is uploaded more than once (an animation, the end-to-end loop of bench.py) or by several ranks of one host at the same
// time (eight staged 100 MB copies share the host's memory bandwidth).  Release before rm_prepared_free.
int rm_prepared_pin(RmPrepared *p, int32_t on) {
    if (!p) return rm_fail(RM_ERR_INVALID, "rm_prepared_pin: null argument");
    std::vector<void *> &pinned = rm_prepared_pinned(p);
    if (!on) {
        for (void *q : pinned) cudaHostUnregister(q);
        pinned.clear();
        return RM_OK;
    }
    if (!pinned.empty()) return RM_OK;
    std::vector<std::pair<void *, size_t>> spans;
    rm_prepared_spans(p, spans);
    for (auto &sp : spans) {
        if (sp.second < (size_t(1) << 16)) continue;             // small arrays are not worth a registration
        const cudaError_t e = cudaHostRegister(sp.first, sp.second, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
        if (e == cudaSuccess) pinned.push_back(sp.first);
        else if (cudaHostRegister(sp.first, sp.second, cudaHostRegisterPortable) == cudaSuccess) pinned.push_back(sp.first);
        else { cudaGetLastError(); return rm_fail(RM_ERR_CUDA, "rm_prepared_pin: cudaHostRegister(%zu bytes) failed: %s", sp.second, cudaGetErrorString(e)); }
    }
    cudaGetLastError();
    return RM_OK;
}

int64_t rm_scene_device_bytes(const RmContext *ctx) { return ctx ? ctx->scene_bytes : 0; }
int64_t rm_scene_h2d_bytes(const RmContext *ctx) { return ctx ? ctx->scene_h2d_bytes : 0; }

// ------------------------------------------------------------------------ per-ray seam
int rm_trace_closest(RmContext *ctx, int64_t n, const float *org, const float *dir, int32_t *tri_idx, float *t) {
    if (!ctx || !ctx->has_scene) return rm_fail(RM_ERR_STATE, "rm_trace_closest: no scene uploaded");
    if (n < 0 || n > 0x7fffff00LL || (n > 0 && (!org || !dir || !tri_idx || !t))) return rm_fail(RM_ERR_INVALID, "rm_trace_closest: bad arguments");
    if (n == 0) return RM_OK;
    RM_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->b_io[0].alloc(n * 12)) || (rc = ctx->b_io[1].alloc(n * 12)) || (rc = ctx->b_io[2].alloc(n * 4)) || (rc = ctx->b_io[3].alloc(n * 4))) return rc;
    cudaStream_t st = ctx->stream;
    RM_CUDA(cudaMemcpyAsync(ctx->b_io[0].p, org, n * 12, cudaMemcpyHostToDevice, st));
    RM_CUDA(cudaMemcpyAsync(ctx->b_io[1].p, dir, n * 12, cudaMemcpyHostToDevice, st));
    auto *cnt = ctx->b_counters.as<unsigned long long>();
    ClosestJob job;
    job.org = ctx->b_io[0].as<float>(); job.dir = ctx->b_io[1].as<float>(); job.aim_in = nullptr;
    job.tri_idx = ctx->b_io[2].as<int>(); job.t_out = ctx->b_io[3].as<float>();
    RM_CUDA(cudaMemsetAsync(ctx->b_cursor.p, 0, 4, st));
    const int grid = ctx->sm_count * kTraceCtasPerSm;
    launch_trace(seam_scene(ctx), seam_levels(ctx), ctx->count_tests, grid, st, job, int(n), nullptr, ctx->b_cursor.as<int>(), cnt, seam_tune(ctx));
    ctx->launches++;
    RM_CUDA(cudaGetLastError());
    RM_CUDA(cudaMemcpyAsync(tri_idx, ctx->b_io[2].p, n * 4, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaMemcpyAsync(t, ctx->b_io[3].p, n * 4, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    return RM_OK;
}

int rm_trace_occluded(RmContext *ctx, int64_t n, const float *org, const float *dir, const float *aim, uint8_t *out) {
    if (!ctx || !ctx->has_scene) return rm_fail(RM_ERR_STATE, "rm_trace_occluded: no scene uploaded");
    if (n < 0 || n > 0x7fffff00LL || (n > 0 && (!org || !dir || !aim || !out))) return rm_fail(RM_ERR_INVALID, "rm_trace_occluded: bad arguments");
    if (n == 0) return RM_OK;
    RM_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->b_io[0].alloc(n * 12)) || (rc = ctx->b_io[1].alloc(n * 12)) || (rc = ctx->b_io[2].alloc(n * 4)) || (rc = ctx->b_io[3].alloc(n * 4))) return rc;
    cudaStream_t st = ctx->stream;
    RM_CUDA(cudaMemcpyAsync(ctx->b_io[0].p, org, n * 12, cudaMemcpyHostToDevice, st));
    RM_CUDA(cudaMemcpyAsync(ctx->b_io[1].p, dir, n * 12, cudaMemcpyHostToDevice, st));
    RM_CUDA(cudaMemcpyAsync(ctx->b_io[2].p, aim, n * 4, cudaMemcpyHostToDevice, st));
    auto *cnt = ctx->b_counters.as<unsigned long long>();
    OccludedJob job;
    job.org = ctx->b_io[0].as<float>(); job.dir = ctx->b_io[1].as<float>(); job.aim_in = ctx->b_io[2].as<float>();
    job.out = ctx->b_io[3].as<unsigned char>();
    RM_CUDA(cudaMemsetAsync(ctx->b_cursor.p, 0, 4, st));
    const int grid = ctx->sm_count * kTraceCtasPerSm;
    launch_trace(seam_scene(ctx), seam_levels(ctx), ctx->count_tests, grid, st, job, int(n), nullptr, ctx->b_cursor.as<int>(), cnt, seam_tune(ctx));
    ctx->launches++;
    RM_CUDA(cudaGetLastError());
    RM_CUDA(cudaMemcpyAsync(out, ctx->b_io[3].p, n, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    return RM_OK;
}

// ------------------------------------------------------------------------ primary rays (K1)
int rm_trace_primary(RmContext *ctx, const RmRenderArgs *args, int32_t *tri_idx, float *t) {
    if (!ctx || !ctx->has_scene) return rm_fail(RM_ERR_STATE, "rm_trace_primary: no scene uploaded");
    int rc = rm_check_args(args);
    if (rc) return rc;
    RM_CUDA(cudaSetDevice(ctx->device));
    const size_t npix = size_t(args->width) * args->height;
    if ((rc = ctx->b_tri_idx.alloc(npix * 4)) || (rc = ctx->b_t.alloc(npix * 4))) return rc;
    cudaStream_t st = ctx->stream;
    auto *cnt = ctx->b_counters.as<unsigned long long>();
    PrimaryJob job;
    job.A = to_dev_args(args);
    job.tiles_x = (args->width + 7) / 8;
    job.tri_idx = ctx->b_tri_idx.as<int>();
    job.t_out = ctx->b_t.as<float>();
    const int n_rays = job.tiles_x * ((args->height + 3) / 4) * 32;
    RM_CUDA(cudaMemsetAsync(ctx->b_cursor.p, 0, 4, st));
    const int grid = ctx->sm_count * kTraceCtasPerSm;
    ctx->timed_begin(RM_KIND_PRIMARY);
    launch_trace(ctx->scene, ctx->stack_levels, ctx->count_tests, grid, st, job, n_rays, nullptr, ctx->b_cursor.as<int>(), cnt, ctx->tune);
    ctx->timed_end();
    ctx->launches++;
    RM_CUDA(cudaGetLastError());
    ctx->width = args->width;
    ctx->height = args->height;
    ctx->frame_args = *args;
    ctx->have_primary = true;
    ctx->have_gbuffer = ctx->have_resolved = false;
    if (tri_idx) RM_CUDA(cudaMemcpyAsync(tri_idx, ctx->b_tri_idx.p, npix * 4, cudaMemcpyDeviceToHost, st));
    if (t) RM_CUDA(cudaMemcpyAsync(t, ctx->b_t.p, npix * 4, cudaMemcpyDeviceToHost, st));
    if (tri_idx || t) RM_CUDA(cudaStreamSynchronize(st));
    return RM_OK;
}

// ------------------------------------------------------------------------ counters / options
int rm_stats_reset(RmContext *ctx) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "context is NULL");
    RM_CUDA(cudaSetDevice(ctx->device));
    RM_CUDA(cudaMemsetAsync(ctx->b_counters.p, 0, ctx->b_counters.bytes, ctx->stream));
    ctx->launches = 0;
    ctx->ev_kind.clear();
    return RM_OK;
}

int rm_stats_read(RmContext *ctx, uint64_t out[4]) {
    if (!ctx || !out) return rm_fail(RM_ERR_INVALID, "rm_stats_read: null argument");
    RM_CUDA(cudaSetDevice(ctx->device));
    unsigned long long h[9];
    RM_CUDA(cudaMemcpyAsync(h, ctx->b_counters.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 3; k++) out[k] = h[k] + h[3 + k] + h[6 + k];
    out[3] = ctx->launches;
    return RM_OK;
}

int rm_stats_kernels(RmContext *ctx, uint64_t counters[9], double ms[4], uint64_t timed_launches[4]) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "rm_stats_kernels: null argument");
    RM_CUDA(cudaSetDevice(ctx->device));
    unsigned long long h[9];
    RM_CUDA(cudaMemcpyAsync(h, ctx->b_counters.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (counters) for (int k = 0; k < 9; k++) counters[k] = h[k];
    double t[4] = {0, 0, 0, 0};
    uint64_t n[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < ctx->ev_kind.size(); i++) {
        float e = 0.0f;
        if (cudaEventElapsedTime(&e, ctx->ev_pool[2 * i], ctx->ev_pool[2 * i + 1]) == cudaSuccess) { t[ctx->ev_kind[i]] += e; n[ctx->ev_kind[i]]++; }
    }
    ctx->ev_kind.clear();
    for (int k = 0; k < 4; k++) { if (ms) ms[k] = t[k]; if (timed_launches) timed_launches[k] = n[k]; }
    return RM_OK;
}

int rm_set_option(RmContext *ctx, const char *name, int64_t value) {
    if (!ctx || !name) return rm_fail(RM_ERR_INVALID, "rm_set_option: null argument");
    if (!std::strcmp(name, "count_tests")) { ctx->count_tests = value != 0; return RM_OK; }
    if (!std::strcmp(name, "exact_secondary")) { ctx->exact_secondary = value != 0; return RM_OK; }
    // test hook: rm_trace_closest / rm_trace_occluded through the secondary-ray tree (the seam itself is the reference's tree)
    if (!std::strcmp(name, "seam_secondary_tree")) {
        ctx->seam_tree = int(std::min<int64_t>(std::max<int64_t>(value, 0), 2));
        if (ctx->seam_tree == 1) ctx->want_binary_tree = true;           // takes effect at the next rm_scene_upload
        return ctx->seam_tree ? rm_ensure_secondary_tree(ctx) : RM_OK;   // a deferred build of the 4-wide tree happens now
    }
    // bounce and shadow rays: 1 = the binary secondary-ray tree (host-built; set before rm_scene_upload), 2 = the 4-wide quantised tree (default)
    if (!std::strcmp(name, "secondary_tree")) { ctx->secondary_tree = value == 1 ? 1 : 2; if (value == 1) ctx->want_binary_tree = true; return RM_OK; }
    // 3 (default): on the device at every upload by the sweep-SAH builder (gpu_sah_bvh.cu); 1: on the device by PLOC (gpu_bvh.cu);
    // 2: PLOC, then refined in the background by the host builder (the render loop swaps the better tree in when it is ready);
    // 0: on the host, cached by geometry hash
    if (!std::strcmp(name, "tree_builder")) { ctx->tree_builder_mode = int(std::min<int64_t>(std::max<int64_t>(value, 0), 3)); ctx->fast_key_valid = false; return RM_OK; }
    // 1 (default): the sweep-SAH build of the secondary-ray tree waits for the first call that needs the tree; 0: rm_scene_upload builds it
    if (!std::strcmp(name, "lazy_tree")) { ctx->lazy_tree = value != 0; return RM_OK; }
    // block until the secondary-ray tree is in place: a deferred build runs, a background refinement (if any) finishes and its tree is installed
    if (!std::strcmp(name, "tree_wait")) {
        const int rc_build = rm_ensure_secondary_tree(ctx);
        if (rc_build) return rc_build;
        rm_start_refinement(ctx, kRefineMinSamples);
        if (ctx->refine && ctx->refine->th.joinable()) ctx->refine->th.join();
        return rm_install_refined_tree(ctx);
    }
    // 0: a tree built for earlier geometry is never reused (every upload rebuilds; what bench.py's end-to-end loop asks for)
    if (!std::strcmp(name, "tree_cache")) { ctx->tree_cache = value != 0; if (!ctx->tree_cache) ctx->fast_key_valid = false; return RM_OK; }
    if (!std::strcmp(name, "fast_leaf_max")) { ctx->fast_leaf_max = int(std::min<int64_t>(std::max<int64_t>(value, 1), 15)); ctx->fast_key_valid = false; return RM_OK; }
    if (!std::strcmp(name, "fast_depth_cap")) { ctx->fast_depth_cap = int(std::min<int64_t>(std::max<int64_t>(value, 8), 26)); ctx->fast_key_valid = false; return RM_OK; }
    if (!std::strcmp(name, "time_kernels")) { ctx->time_kernels = value != 0; ctx->ev_kind.clear(); return RM_OK; }
    if (!std::strcmp(name, "trace_refill")) { ctx->tune.refill_live = ctx->tune_fast.refill_live = ctx->tune_wide.refill_live = int(std::min<int64_t>(std::max<int64_t>(value, 1), 32)); return RM_OK; }
    if (!std::strcmp(name, "trace_w_inner")) { ctx->tune.w_inner = ctx->tune_fast.w_inner = ctx->tune_wide.w_inner = int(std::max<int64_t>(value, 1)); return RM_OK; }
    if (!std::strcmp(name, "trace_w_leaf")) { ctx->tune.w_leaf = ctx->tune_fast.w_leaf = ctx->tune_wide.w_leaf = int(std::max<int64_t>(value, 1)); return RM_OK; }
    if (!std::strcmp(name, "smem_levels")) {        // stack entries per thread in shared memory (0 = the whole tree depth); deeper entries spill to local memory
        ctx->tune.smem_levels = ctx->tune_fast.smem_levels = ctx->tune_wide.smem_levels = int(std::min<int64_t>(std::max<int64_t>(value, 0), 40));
        return RM_OK;
    }
    if (!std::strcmp(name, "stack_levels")) {       // perf experiments only: never below the tree depth rm_scene_upload derived
        ctx->stack_levels = int(std::min<int64_t>(std::max<int64_t>(value, ctx->stack_levels), 40));
        return RM_OK;
    }
    if (!std::strcmp(name, "wave_paths")) { ctx->wave_paths = int(std::min<int64_t>(std::max<int64_t>(value, 1 << 16), 1 << 25)); return RM_OK; }
    if (!std::strcmp(name, "max_depth")) { ctx->max_depth = int(std::min<int64_t>(std::max<int64_t>(value, 1), 16)); return RM_OK; }
    if (!std::strcmp(name, "direct_warp")) { ctx->direct_warp = int(std::min<int64_t>(std::max<int64_t>(value, 0), 2)); return RM_OK; }
    if (!std::strcmp(name, "compact_pixels")) { ctx->compact_pixels = value != 0; ctx->have_gbuffer = false; return RM_OK; }
    if (!std::strcmp(name, "fxaa_rows")) { ctx->fxaa_rows = int(std::min<int64_t>(std::max<int64_t>(value < 0 ? 16 : value, 0), 256)); ctx->fxaa_auto = value < 0; return RM_OK; }
    if (!std::strcmp(name, "disable_clamp")) { ctx->disable_clamp = value != 0; return RM_OK; }
    return rm_fail(RM_ERR_INVALID, "rm_set_option: unknown option '%s'", name);
}

} // extern "C"
