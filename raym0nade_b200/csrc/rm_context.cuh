// rm_context.cuh — the context object behind the C ABI (host side, CUDA runtime).
#pragma once
#include <cstdint>
#include <atomic>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#include "raym0nade_b200.h"
#include "rm_internal.h"
#include "dev_scene.cuh"
#include "dev_trace.cuh"
#include "wide_bvh.h"

#define RM_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return rm_fail(RM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Owning device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n) {
        if (n <= bytes && p) return RM_OK;
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) { p = nullptr; bytes = 0; return rm_fail(RM_ERR_CUDA, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e)); }
        bytes = n;
        return RM_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// Background refinement of the secondary-ray tree ("tree_builder" 2): the device builder's tree serves from the first ray; a
// host thread meanwhile runs the binned-SAH builder + collapse over a private copy of the positions, and the render loop
// swaps its (better) tree in when it is ready.
struct RefineJob {
    std::thread th;
    std::atomic<int> state{0};             // 1 running, 2 ready, 3 failed
    std::vector<float> pos;
    int n = 0;
    uint64_t key = 0;
    std::vector<RmWideNode> wnodes;
    std::vector<int32_t> worder;
    int wdepth = 0;
    bool discard = false;                  // the vertices moved while it ran (rm_scene_refit): its tree is not installed
    bool started = false;                  // the thread is launched by the first render that is long enough to profit (rm_start_refinement)
    std::atomic<bool> cancel{false};       // asks a running build to unwind (a new scene arrived, the context goes away)
    int limit = 0, depth_cap = 22;         // stack entries the traversal allows; depth cap of the binary tree
    void stop() { cancel.store(true); if (th.joinable()) th.join(); }
};

struct RmContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool has_scene = false;
    bool count_tests = false;
    bool exact_secondary = false;      // true: bounce and shadow rays also traverse the reference's tree in the reference's order
    bool disable_clamp = false;        // debugging aid: never drop the held-back sample
    uint64_t launches = 0;

    // optional per-kernel-kind device timing (cudaEvent pairs on the context's stream)
    bool time_kernels = false;
    std::vector<cudaEvent_t> ev_pool;      // [2*i], [2*i+1] = begin / end of timed launch i
    std::vector<int> ev_kind;
    void timed_begin(int kind) {
        if (!time_kernels) return;
        size_t i = ev_kind.size();
        while (ev_pool.size() < 2 * (i + 1)) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
        ev_kind.push_back(kind);
        cudaEventRecord(ev_pool[2 * i], stream);
    }
    void timed_end() {
        if (!time_kernels) return;
        cudaEventRecord(ev_pool[2 * (ev_kind.size() - 1) + 1], stream);
    }

    // scene
    rm::DevScene scene{};
    rm::DevScene scene_fast{};             // the same scene with the secondary-ray tree (fast_bvh.cpp) in place of the reference's
    rm::DevScene scene_wide_first{};       // the device builder's tree (what scene_wide points at until the refined one is installed)
    bool use_refined = false;
    rm::DevScene scene_wide{};             // ... and with that tree collapsed to 4-wide quantised nodes (wide_bvh.cpp): what bounce and shadow rays traverse
    int stack_levels_fast = 24, stack_levels_wide = 42;
    int secondary_tree = 2;                // 1: the binary secondary-ray tree, 2: its 4-wide form (rm_set_option "secondary_tree")
    bool have_wide = false, have_fast = false, host_wide_valid = false, want_binary_tree = false;
    int wide_nodes = 0, wide_levels = 0;
    DevBuf b_nodes_fast, b_tri_fast, b_facemap;
    DevBuf b_nodes_wide, b_tri_wide, b_facemap_wide;
    DevBuf b_nodes_wide2, b_tri_wide2, b_facemap_wide2;   // the refined tree (tree_builder 2)
    std::unique_ptr<RefineJob> refine;
    bool refined_installed = false, tree_cache = true;
    uint64_t refined_key = 0;
    int refined_n = 0, refined_levels = 0, refined_nodes = 0;
    DevBuf b_build[16];                    // scratch of the device tree builder (gpu_bvh.cu), kept across uploads
    int tree_builder_mode = 3;             // the secondary-ray tree: 3 (default) built on the device at every upload by the sweep-SAH builder (gpu_sah_bvh.cu);
                                           // 1: on the device by Morton sort + PLOC (gpu_bvh.cu; a 3 ms build, ~14 % more node visits per ray); 2: PLOC, then
                                           // refined by the host builder in the background; 0: on the host (fast_bvh.cpp + wide_bvh.cpp), cached by geometry hash
    int wide_built_by = 0;                 // the builder mode that made the tree in b_nodes_wide
    bool lazy_tree = true, wide_pending = false;      // the sweep-SAH build waits for the first call that needs the tree (rm_ensure_secondary_tree)
    float scene_lo[3] = {0, 0, 0};         // lower corner of the scene bounds (the reference tree's root box)
    int fast_depth_cap = 22;               // depth cap of the secondary-ray tree = its traversal stack entries (8 CTAs x 128 threads x 8 B x depth of shared memory per SM)
    int fast_leaf_max = 3;                 // triangles per leaf of the secondary-ray tree (A/B of 2..8 and caps 20..24: profiles/r01f_ab16_secondary_tree.txt)
    bool fast_root_is_leaf = false, fast_key_valid = false;
    int seam_tree = 0;                     // test hook: which tree rm_trace_closest / rm_trace_occluded traverse (0 = the reference's)
    uint64_t fast_key = 0;
    int fast_n = 0;
    DevBuf b_nodes, b_tri, b_shade, b_mats, b_texs, b_texels, b_lights, b_lpos, b_lnrm, b_lcdf, b_sky, b_skycdf, b_skyguide, b_lut;
    DevBuf b_raw[4];                       // the caller's positions / uvs / normals / face materials as uploaded (input of k_pack_faces)
    int64_t scene_bytes = 0;               // device-resident bytes of the uploaded scene
    int64_t scene_h2d_bytes = 0;           // bytes copied host -> device by the last rm_scene_upload

    // counters: rays, box, tri (device)
    DevBuf b_counters;
    DevBuf b_cursor;                       // int[4]: work cursors of the persistent trace kernels
    int sm_count = 148;
    int stack_levels = 24;                 // traversal stack entries per ray = tree depth of the uploaded scene (rm_scene_upload)
    rm::TraceTune tune{28, 1, 1, 0};          // see dev_trace.cuh; adjustable through rm_set_option for perf experiments
    rm::TraceTune tune_wide{28, 1, 2, 14};    // the 4-wide tree: same vote, same shared-memory stack share
    rm::TraceTune tune_fast{28, 1, 2, 14};    // the secondary-ray tree has short leaves: the vote leans towards the leaf step (profiles/r01g_ab17_votes.txt);
                                           // 14 stack entries in shared memory, deeper ones (rare) in local memory (profiles/r01g_ab18_stack_spill.txt)
    int wave_paths = 1 << 25;              // path-queue capacity of the wavefront loop (vertices in flight per round).  32 M keeps every
                                           // launch long enough that kernel tails and launch gaps stay ~1 % (4 M: -10 %, 16 M: -1 %,
                                           // profiles/r01c_ab7_wave_size.txt, r01d_ab8_wave_size.txt) for 38 GB of queues out of 180 GB
    int max_depth = 16;                    // perf experiments only: bounce limit of the wavefront loop (16 = the reference's maxRayDepth)
    int direct_warp = 0;                   // k_direct_gen mapping: 0 = a warp per pixel for environment-lit scenes, a thread per pixel otherwise; 1 / 2 force one (A/B)
    bool compact_pixels = true;            // the per-pixel stages run over the list of sampled pixels (rm_gbuffer); 0: over every pixel (A/B, rm_set_option "compact_pixels")
    bool fxaa_auto = true;                 // the strip height is chosen per frame so that all strips are resident at once; rm_set_option("fxaa_rows", n) pins it
    int fxaa_rows = 16;                    // rows per warp strip of k_fxaa_strip; 0 = the two-pass tiled form (what frames with width % 4 != 0 always get)

    // per-frame state
    int width = 0, height = 0;
    bool have_primary = false, have_gbuffer = false, have_resolved = false;
    RmRenderArgs frame_args{};
    DevBuf b_tri_idx, b_t;                 // primary hits
    DevBuf b_gbuffer;                      // RmHitInfo AoS [npix]
    DevBuf b_io[4];                        // staging for the batched per-ray seam / fxaa
    // wavefront + accumulators live in rm_render.cu's state
    void *render_state = nullptr;
    // multi-GPU: NCCL communicator of this context (rm_comm.cu)
    void *comm = nullptr;
    int comm_rank = 0, comm_world = 1;

    ~RmContext() {
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        for (DevBuf *b : {&b_nodes, &b_tri, &b_shade, &b_mats, &b_texs, &b_texels, &b_lights, &b_lpos, &b_lnrm, &b_lcdf,
                          &b_sky, &b_skycdf, &b_skyguide, &b_lut, &b_nodes_fast, &b_tri_fast, &b_facemap, &b_nodes_wide, &b_tri_wide, &b_facemap_wide, &b_raw[0], &b_raw[1], &b_raw[2], &b_raw[3], &b_counters, &b_cursor, &b_tri_idx, &b_t, &b_gbuffer, &b_io[0], &b_io[1], &b_io[2], &b_io[3]})
            b->release();
        if (refine) refine->stop();
        for (DevBuf &b : b_build) b.release();
        for (DevBuf *b : {&b_nodes_wide2, &b_tri_wide2, &b_facemap_wide2}) b->release();
    }
};

enum { RM_KIND_PRIMARY = 0, RM_KIND_PATHS = 1, RM_KIND_SHADOW = 2, RM_KIND_SHADE = 3 };
rm::DevArgs to_dev_args(const RmRenderArgs *a);
int rm_check_args(const RmRenderArgs *a);
// implemented in rm_render.cu
void rm_render_state_free(RmContext *ctx);
extern "C" int rm_accum_mark_slice(RmContext *ctx, int64_t first_pixel, int64_t pixels);       // after rm_reduce_scatter: only this slice of the accumulators holds the frame
// implemented in fast_bvh.cpp
int rm_build_fast_bvh(const float *positions, int n, int depth_cap, int leaf_max, std::vector<RmBvhNode> &nodes, std::vector<int32_t> &order, int *depth_out);
int rm_build_fast_bvh_cancellable(const float *positions, int n, int depth_cap, int leaf_max, std::vector<RmBvhNode> &nodes, std::vector<int32_t> &order, int *depth_out,
                                  const std::atomic<bool> *cancel);
// implemented in wide_bvh.cpp (declared in wide_bvh.h)
// implemented in rm_api.cu: the deferred build of the secondary-ray tree, if one is pending
extern "C" int rm_ensure_secondary_tree(RmContext *ctx);
// implemented in rm_api.cu: swaps the background-refined secondary-ray tree in once it is ready (no-op otherwise)
extern "C" int rm_install_refined_tree(RmContext *ctx);
// implemented in rm_api.cu: launches the pending background refinement if `pixel_samples` of rendering are about to follow that make
// it worth its host time (always when trees are cached across uploads), drops it otherwise
extern "C" void rm_start_refinement(RmContext *ctx, int64_t pixel_samples);
// implemented in rm_api.cu: after rm_scene_refit (gpu_ref_bvh.cu) - the traversal and shading records formed anew from ctx->b_raw[0] and re-permuted for the 4-wide tree(s)
extern "C" int rm_repack_faces(RmContext *ctx);
// implemented in rm_comm.cu
void rm_comm_state_free(RmContext *ctx);
// implemented in host_prep.cpp: the arrays of a prepared scene / the ones currently page-locked (rm_prepared_pin)
void rm_prepared_spans(RmPrepared *P, std::vector<std::pair<void *, size_t>> &out);
std::vector<void *> &rm_prepared_pinned(RmPrepared *P);
