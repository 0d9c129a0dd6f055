/* raym0nade_b200.h — C ABI of the B200-native path-tracing hot path.
 *
 * The reference (lemonchu/Raym0nade) has no plugin / FFI layer; its seams are ordinary
 * C++ calls.  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference tree).  Conventions: plain C structs, POD only,
 * no exceptions across the boundary; every call returns RM_OK (0) or a negative error
 * and leaves a message for rm_last_error(); one context per CUDA device; calls on one
 * context are serialised by the caller; all buffers are caller-owned.  Pointers named
 * `d_*` are DEVICE pointers, all others are host pointers.
 *
 * There is no CPU fallback: every compute entry point fails with RM_ERR_CUDA when no
 * sm_100 device is present.
 */
#ifndef RAYM0NADE_B200_H
#define RAYM0NADE_B200_H

#include <stddef.h>
#include <stdint.h>
#include "rm_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RM_OK 0
#define RM_ERR_INVALID (-1)
#define RM_ERR_CUDA (-2)
#define RM_ERR_STATE (-3)

/* ---------------------------------------------------------------------------------
 * Post-load scene: exactly what a loaded reference `Model` holds (include/model.h:31-37)
 * after Model::Model has run (src/model.cpp:172-215).  An integration fills this from
 * its own Model; rm_prepare_scene() below builds it from a raw scene.
 * ------------------------------------------------------------------------------- */

/* ImageData with its mip chain (include/material.h:12-25; generateMipmaps src/material.cpp:113-148). */
typedef struct RmTextureDesc {
    int32_t width, height, channels, map_depth;
    const uint8_t *levels[8];          /* level l: (width>>l)*(height>>l)*channels bytes */
} RmTextureDesc;

/* Material (include/material.h:30-51). tex[]: 0 diffuse, 1 specular, 2 emissive, 3 normals; -1 = empty. */
typedef struct RmMaterialDesc {
    int32_t tex[4];
    float opacity, ior, roughness;
    float transmitting_color[3];
    int32_t has_fully_transparent_part;
    int32_t _pad;
} RmMaterialDesc;

/* LightObject (include/component.h:42-50; built by checkLightObject src/model.cpp:44-82). */
typedef struct RmLightDesc {
    float center[3], color[3];
    float power;
    int32_t n_faces;
    const float *face_positions;       /* [n_faces][3][3] */
    const float *face_normals;         /* [n_faces][3][3] vertex normals of those faces */
    const float *face_cdf;             /* [n_faces] RandomDistribution prefix sums (src/component.cpp:12-18) */
} RmLightDesc;

typedef struct RmSceneDesc {
    int32_t n_faces, n_nodes, n_materials, n_textures, n_lights;
    int32_t sky_width, sky_height;     /* 0 = no sky */
    int32_t _pad;
    const RmBvhNode *nodes;            /* BVH::node, heap-indexed from 1 (src/bvh.cpp:18-54); never-written slots zero */
    const float *positions;            /* [n_faces][3][3] Model::faces in POST-BUILD order (BVH::build permutes, bvh.cpp:38) */
    const float *uvs;                  /* [n_faces][3][2] */
    const float *normals;              /* [n_faces][3][3] */
    const int32_t *face_material;      /* [n_faces] Material::id */
    const RmMaterialDesc *materials;
    const RmTextureDesc *textures;
    const RmLightDesc *lights;
    const float *sky_data;             /* [h][w][3] SkyBox::data AFTER SkyBox::Init premultiplied it (src/component.cpp:54-67) */
    const float *sky_cdf;              /* [h*w] SkyBox::dist prefix sums */
} RmSceneDesc;

/* ---------------------------------------------------------------------------------
 * Host-side scene preparation — the load-time derivations of Model::Model that the
 * north star keeps on the host: BVH::build (src/bvh.cpp:18-54), generateMipmaps
 * (src/material.cpp:113-148), checkLightObject (src/model.cpp:23-82), SkyBox::Init
 * (src/component.cpp:54-67), hasTransparentPart (src/material.cpp:102-107,330-333).
 * Pure host C++, no CUDA.
 * ------------------------------------------------------------------------------- */
typedef struct RmPrepared RmPrepared;
int rm_prepare_scene(const RmRawScene *raw, RmPrepared **out);
const RmSceneDesc *rm_prepared_desc(const RmPrepared *p);
const int32_t *rm_prepared_permutation(const RmPrepared *p); /* perm[i] = raw face index stored at post-build slot i */
void rm_prepared_free(RmPrepared *p);
/* Page-lock (on = 1) or release (on = 0) the arrays of a prepared scene (cudaHostRegister; needs a CUDA device): every later
 * rm_scene_upload of it is then a DMA straight out of them.  For scenes uploaded more than once, or by several ranks of one host
 * at a time.  Release before rm_prepared_free.  Nothing like it in the reference (no device). */
int rm_prepared_pin(RmPrepared *p, int32_t on);

/* ---------------------------------------------------------------------------------
 * Context and scene upload
 * ------------------------------------------------------------------------------- */
typedef struct RmContext RmContext;

const char *rm_last_error(void);
const char *rm_version(void);

/* One context per CUDA device.  stream = a cudaStream_t the caller owns (0 = default
 * stream); every kernel and copy of this context is issued on it. */
int rm_context_create(int device, void *stream, RmContext **out);
void rm_context_destroy(RmContext *ctx);
int rm_context_synchronize(RmContext *ctx);

/* Structural check of a post-load scene (host only, no GPU): every node reachable from the root lies inside the node
 * array, leaf ranges inside the face array and together own every face (the tree as BVH::dfs_rayHit reads it,
 * src/bvh.cpp:56-88), material / texture indices and slot-channel pairings (src/material.cpp:58), texture levels, light
 * face lists and distributions, sky buffers.  rm_scene_upload runs it first; the reference trusts a loaded Model. */
int rm_scene_validate(const RmSceneDesc *scene);

/* Flatten a post-load scene into the SoA device layout and stage it to HBM.
 * Replaces nothing in the reference (it has no device); it is what `const Model&`
 * is to render_multiThread (src/render.cpp:593). */
int rm_scene_upload(RmContext *ctx, const RmSceneDesc *scene);
/* bytes of HBM the staged scene occupies */
int64_t rm_scene_device_bytes(const RmContext *ctx);
/* Bytes the last rm_scene_upload copied host -> device (the caller's arrays as they are; the traversal and shading
 * records are formed from them on the device). */
int64_t rm_scene_h2d_bytes(const RmContext *ctx);

/* ---------------------------------------------------------------------------------
 * The reference's tree built and refitted on the device (csrc/gpu_ref_bvh.cu)
 * ------------------------------------------------------------------------------- */
/* nodeCount(1, n) + 1 (src/bvh.cpp:44-49): the length of BVH::node for n faces. */
int32_t rm_tree_node_count(int32_t n_faces);
/* BVH::build (src/bvh.cpp:18-54) on the device: the reference's rule - leaves of at most 10 faces, split at the median of the
 * face centres along the axis of their largest variance, heap-indexed nodes - applied level by level (prefix sums for the
 * variances, one radix sort per level).  positions: raw faces [n_faces][3][3] (host).  nodes (host, n_nodes =
 * rm_tree_node_count(n_faces) records) receives BVH::node; perm[i] (host) = the raw face that the build moves to slot i, i.e.
 * what BVH::build does to Model::faces in place (bvh.cpp:38).  Same tree shape and boxes as the reference's; which of two faces
 * with equal centre coordinates lands left of a median, and the order inside a leaf, are std::nth_element's in the reference
 * and a sort's here. */
int rm_tree_build(RmContext *ctx, const float *positions, int32_t n_faces, RmBvhNode *nodes, int32_t n_nodes, int32_t *perm);
/* rm_prepare_scene with the tree built by rm_tree_build on `ctx`'s device; everything else is the same host code. */
int rm_prepare_scene_device(RmContext *ctx, const RmRawScene *raw, RmPrepared **out);
/* Vertices moved, topology kept: positions [n_faces][3][3] (host) in the uploaded scene's POST-BUILD order replace the staged
 * ones (n_faces must be the uploaded scene's); the boxes of the reference's tree are recomputed on the device (leaf boxes from their faces, inner boxes bottom-up:
 * dfs_build's box arithmetic, bvh.cpp:21-24,41) and the secondary-ray tree is refitted and re-quantised bottom-up.  uvs, normals,
 * materials, lights and sky stay as uploaded (a moved emissive face keeps its old light-object entry). */
int rm_scene_refit(RmContext *ctx, const float *positions, int32_t n_faces);

/* ---------------------------------------------------------------------------------
 * Per-ray seam: Model::rayHit / Model::rayHit_test (include/model.h:41-42,
 * src/model.cpp:332-354) over batches of rays.
 * ------------------------------------------------------------------------------- */
/* Closest hit.  org/dir: [n][3].  tri_idx: post-build face index or -1; t: hit.t_max (INF on miss). */
int rm_trace_closest(RmContext *ctx, int64_t n, const float *org, const float *dir, int32_t *tri_idx, float *t);
/* Occlusion test, rayHit_test semantics: out[i] = 1 iff something blocks the ray before aim[i]. */
int rm_trace_occluded(RmContext *ctx, int64_t n, const float *org, const float *dir, const float *aim, uint8_t *out);

/* ---------------------------------------------------------------------------------
 * Per-pixel seam: renderPixel (src/render.cpp:448-551) split into its stages.
 * ------------------------------------------------------------------------------- */
/* Primary rays exactly as renderPixel forms them (466-479) + Model::rayHit.
 * Leaves {tri_idx, t} in the context and, if non-NULL, copies them to the host. */
int rm_trace_primary(RmContext *ctx, const RmRenderArgs *args, int32_t *tri_idx, float *t);

/* G-buffer: getHitInfo (src/render.cpp:62-79) at the primary hit, sky emission on a
 * miss (480-485).  Requires rm_trace_primary for the same args.  gbuffer (host, may be
 * NULL) receives width*height HitInfo records in the reference's AoS layout with the
 * restored baseColor (550). */
int rm_gbuffer(RmContext *ctx, const RmRenderArgs *args, RmHitInfo *gbuffer);

/* Sample loops of renderPixel (498-547): the direct-light samples s with
 * s = sample_begin + k*sample_stride < spp_direct and the indirect samples likewise
 * below spp_indirect, each weighted with the GLOBAL 1/spp factors, accumulated into the
 * context's fp32 accumulators.  sample_begin = rank, sample_stride = world size shards a
 * render across GPUs by interleaved sample index.  Resets the accumulators first when
 * `reset` is non-zero.  Requires rm_gbuffer. */
int rm_render_samples(RmContext *ctx, const RmRenderArgs *args, int32_t sample_begin, int32_t sample_stride,
                      uint64_t seed, int32_t reset);

/* Multi-GPU reduction of a sharded render, in three steps (single-GPU callers skip all of
 * this: rm_resolve does the local equivalent).  The caller owns the collective
 * (ncclAllReduce / torch.distributed) and runs it on the context's stream:
 *   1. rm_accum_view: firefly-clamp side data (render.cpp:534-547).  `d_sum` = 2 floats per
 *      pixel {sum of sample luminances, number of indirect samples} to be SUMMED across
 *      ranks; `d_max` = 1 float per pixel (luminance of the rank's held-back sample) to be
 *      MAX-reduced across ranks.  Sizes in floats.
 *   2. rm_accum_after_reduce: every rank commits its held-back sample against the global
 *      totals (only the rank owning the global maximum can drop it).
 *   3. rm_accum_radiance: `d_rad` = 16 floats per pixel (Dd, Ds, Id, Is as {rgb, second
 *      moment}) to be SUMMED across ranks; then rm_resolve on the rank(s) holding the sum. */
int rm_accum_view(RmContext *ctx, float **d_sum, int64_t *n_sum, float **d_max, int64_t *n_max);
int rm_accum_after_reduce(RmContext *ctx, int32_t rank, int32_t world);
int rm_accum_radiance(RmContext *ctx, float **d_rad, int64_t *n_rad);

/* The same exchange done by the library over NCCL (bound at run time: libnccl.so.2), for hosts that do not bring their
 * own collective layer.  One process per GPU, one context per process.  rm_comm_unique_id on rank 0 -> the host hands
 * the 128 bytes to every rank (MPI, sockets, a file) -> rm_comm_init on every rank (collective) -> after each rank's
 * rm_render_samples(ctx, args, rank, world, ...), rm_reduce(ctx, root) performs steps 1-3 on the context's stream
 * (collective) and rank `root` calls rm_resolve.  Nothing like it exists in the reference (single process, threads). */
int rm_comm_unique_id(uint8_t id[128]);
int rm_comm_init(RmContext *ctx, const uint8_t id[128], int32_t rank, int32_t world);
int rm_reduce(RmContext *ctx, int32_t root);
int rm_comm_destroy(RmContext *ctx);
/* The exchange with its last step scattered (ncclReduceScatter): afterwards rank r holds the summed frame for ITS slice of
 * the pixels only - [r * per, (r + 1) * per), per = ceil(width * height / world) - and every rank finalises and downloads
 * its own slice in parallel: rm_resolve_slice writes pixels [first, first + count) of the WHOLE-frame host arrays it is
 * given (e.g. one shared pinned frame all ranks of a box map); rm_frame_slice reports the slice.  Without a preceding
 * rm_reduce_scatter the slice is the whole frame. */
int rm_reduce_scatter(RmContext *ctx);
int rm_frame_slice(RmContext *ctx, int64_t *first_pixel, int64_t *pixels);
int rm_resolve_slice(RmContext *ctx, const RmRenderArgs *args, RmHitInfo *gbuffer, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is);

/* Progressive checkpoint / resume (the reference has none): the un-finalised accumulators
 * of the current frame as one opaque blob of rm_checkpoint_bytes(args) bytes.  Save after any rm_render_samples call;
 * after rm_checkpoint_load (same scene uploaded, same args) further rm_render_samples calls with reset = 0 add the
 * remaining sample shards, and rm_resolve gives the frame a single uninterrupted render would have given. */
int64_t rm_checkpoint_bytes(const RmRenderArgs *args);
int rm_checkpoint_save(RmContext *ctx, void *host, int64_t bytes);
int rm_checkpoint_load(RmContext *ctx, const RmRenderArgs *args, const void *host, int64_t bytes);

/* Resolve: firefly clamp + diffuse/specular split + exposure/variance finalise
 * (render.cpp:510-549) -> the four RadianceData planes (host, reference AoS layout). */
int rm_resolve(RmContext *ctx, const RmRenderArgs *args, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is);

/* Copy the resolved Photo buffers (restored G-buffer + the four planes) to host memory; any may be NULL. */
int rm_download_resolved(RmContext *ctx, RmHitInfo *gbuffer, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is);

/* Whole render_multiThread pixel loop (src/render.cpp:593-626) on one GPU:
 * primary + G-buffer + all samples + resolve, results to host buffers (any may be NULL). */
int rm_render(RmContext *ctx, const RmRenderArgs *args, uint64_t seed,
              RmHitInfo *gbuffer, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is);

/* ---------------------------------------------------------------------------------
 * Post pass
 * ------------------------------------------------------------------------------- */
/* Photo::FXAA (include/image.h:54, src/image.cpp:363-452): fp32 RGB in -> fp32 RGB out, host buffers. */
int rm_fxaa(RmContext *ctx, const float *rgb_in, float *rgb_out, int32_t width, int32_t height);
/* Same on device buffers (no copies): d_rgb_in/d_rgb_out [h][w][3]. */
int rm_fxaa_device(RmContext *ctx, const float *d_rgb_in, float *d_rgb_out, int32_t width, int32_t height);
/* Stage host Photo buffers (Gbuffer + radiance_Dd/Ds/Id/Is, include/image.h:38-47) as the context's resolved
 * frame, so the image-space passes below can run on planes produced elsewhere (e.g. the reference's CPU render). */
int rm_upload_resolved(RmContext *ctx, const RmRenderArgs *args, const RmHitInfo *gbuffer, const RmRadiance *Dd,
                       const RmRadiance *Ds, const RmRadiance *Id, const RmRadiance *Is);
/* Photo::spatialClamp (include/image.h, src/image.cpp:30-82): 7x7 separable luminance blur per plane and a
 * rescale of pixels brighter than 36x their neighbourhood; in place on the context's resolved planes
 * (device).  render_multiThread applies it before the "Raw" exports (src/render.cpp:645). */
int rm_spatial_clamp(RmContext *ctx, const RmRenderArgs *args);
/* Photo::filter (src/image.cpp:84-213): filterVar, then five edge-stopping a-trous passes (steps 1,2,4,8,16)
 * over each of the four planes, guided by the resolved G-buffer; in place on the context's planes (device).
 * Fetch the result with rm_download_resolved or shade it with rm_postprocess. */
int rm_filter(RmContext *ctx, const RmRenderArgs *args);
/* Photo::depthFeildBlur (src/image.cpp:285-356) on an rgb frame (host buffers) with the context's resolved G-buffer; focus,
 * CoC and the camera position come from `args` (src/render.cpp:665-668).  Every destination pixel replays, in the reference's
 * depth order, the sources whose disc reaches it: same arithmetic, same order, no atomics. */
int rm_depth_field_blur(RmContext *ctx, const RmRenderArgs *args, const float *rgb_in, float *rgb_out);
/* Photo::postProcessing (src/image.cpp:470-479): shade (215-246) [+ depthFeildBlur (285-356) when shade_options has
 * DoDepthFieldBlur = 1024] [+ bloom (248-283) when it has DoBloom = 256] + gammaCorrection (454-468) [+ FXAA when it has
 * DoFXAA = 512] on the context's resolved planes -> host RGB. */
int rm_postprocess(RmContext *ctx, const RmRenderArgs *args, int32_t shade_options, float *rgb_out);

/* ---------------------------------------------------------------------------------
 * Work counters (device-side), reset by rm_stats_reset:
 *   [0] rays = BVH::rayHit invocations   [1] box tests (rayInBox)   [2] triangle tests
 *   [3] kernel launches issued by this context
 * Box/triangle counts are only collected when rm_set_option("count_tests", 1). */
int rm_stats_reset(RmContext *ctx);
int rm_stats_read(RmContext *ctx, uint64_t out[4]);
/* Host-only diagnostic (no GPU): build the secondary-ray tree (see "exact_secondary" below) for `positions` [n][9] and verify
 * its invariants; out = {pair blocks, depth, leaves, largest leaf}. */
int rm_secondary_tree_stats(const float *positions, int32_t n, int32_t depth_cap, int32_t leaf_max, int32_t out[4]);
/* The same for the 4-wide, 8-bit quantised form of that tree that bounce and shadow rays traverse by default
 * (csrc/wide_bvh.cpp): every triangle in one leaf, every decoded child box encloses the vertices beneath it;
 * out = {nodes (64-byte records), levels, leaves, children per node x 100}. */
int rm_wide_tree_stats(const float *positions, int32_t n, int32_t depth_cap, int32_t out[4]);
/* The 4-wide tree of the uploaded scene: out = {its builder - 3: on the device by the top-down sweep SAH of csrc/gpu_sah_bvh.cu
 * (the default, "tree_builder" 3), 1: on the device by Morton sort + PLOC clustering (csrc/gpu_bvh.cu, "tree_builder" 1), 2: the
 * PLOC tree replaced by the host builder's in the background ("tree_builder" 2), 0: on the host -, records, levels, 1 if bounce
 * and shadow rays traverse it}. */
int rm_tree_info(const RmContext *ctx, int32_t out[4]);
/* Options (integers).  "exact_secondary" 0|1 (default 0): 1 sends the estimator's bounce and shadow rays through the
 * reference's own BVH in the reference's visit order, like primary rays and the per-ray seam always are; 0 lets them use the
 * library's second tree over the same triangles (same box / triangle tests, binned-SAH topology - the closest accepted hit is
 * the same, far fewer tests per ray) - "secondary_tree" 1: as the binary tree, 2 (default): collapsed to 4-wide nodes with
 * quantised child boxes and a conservative slab test.  "tree_builder" 3 (default) | 1 | 2 | 0: who builds that tree - the
 * device by a top-down sweep SAH at every upload, the device by PLOC clustering, PLOC followed by a background refinement on
 * the host, the host; "lazy_tree" 1 (default) | 0: the sweep-SAH build waits for the first call that needs the tree
 * (rm_render_samples with samples to draw, rm_tree_info) instead of running inside rm_scene_upload.
 * "count_tests", "time_kernels": counters / per-kind device timing.  The others
 * ("trace_refill", "wave_paths", "stack_levels", ...) are tuning and test hooks, see rm_api.cu. */
int rm_set_option(RmContext *ctx, const char *name, int64_t value);
/* Per-kernel-kind breakdown.  Kinds: 0 primary / batched per-ray kernels, 1 closest hit over the
 * path queue, 2 occlusion over the shadow queue, 3 shading.  counters[3*kind + {0,1,2}] =
 * {rays, box tests, triangle tests} for kinds 0..2.  ms / timed_launches: device time (cudaEvent
 * pairs on the context's stream) and launch count per kind since the last call, collected only
 * while rm_set_option("time_kernels", 1).  Synchronises the stream. */
int rm_stats_kernels(RmContext *ctx, uint64_t counters[9], double ms[4], uint64_t timed_launches[4]);

#ifdef __cplusplus
}
#endif
#endif /* RAYM0NADE_B200_H */
