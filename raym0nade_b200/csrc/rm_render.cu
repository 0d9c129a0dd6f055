// rm_render.cu — C ABI for the per-pixel stages: G-buffer, sample loops (wavefront), resolve, post pass.
//
// Host orchestration only; the kernels are in kernels_render.cuh / kernels_post.cuh.  The
// whole render is issued asynchronously on the context's stream: queue lengths live in
// device memory and every kernel is a grid-stride loop over them, so the host never waits
// between waves or bounces.
#include <algorithm>
#include <new>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "rm_context.cuh"
#include "kernels_render.cuh"
#include "kernels_post.cuh"

using namespace rm;

namespace {

struct RenderState {
    int npix = 0;
    // frame
    DevBuf gbuffer, sav_base, n_ind, glass_list, dir_base, active_list, active_tmp;
    int n_active = 0;                   // pixels with a primary hit on a non-emissive surface (FrameBuffers::active_list)
    // accumulators
    DevBuf rad, clum_sum, clum_max, hold_clum, hold, lock;
    bool accum_valid = false, hold_committed = false;
    int64_t slice_first = 0, slice_pixels = -1;      // >= 0: after rm_reduce_scatter only these pixels of `rad` hold the frame's sums
    // queues
    int q_cap = 0, s_cap = 0;
    DevBuf q[2][20];
    DevBuf shadow, sorted;    // shadow items; the round's live vertices ordered by (mode, material) (k_decide / k_sort_*)
    DevBuf counts;            // int[C_COUNT]: device-side pipeline state, see the C_* indices in kernels_render.cuh
    int *h_counts = nullptr;  // pinned mirror of `counts`
    // resolved planes + output images
    DevBuf planes[4], g_out, rgb[2];
    DevBuf planes_alt[4];     // second set of planes: every image-space pass reads one set and writes the other
    DevBuf f_pm, f_ns, f_op;  // packed neighbour fields of the G-buffer for the edge-stopping filter
    DevBuf glow[2];           // bloom
    DevBuf dof_depth, dof_src, dof_sorted, dof_lists, dof_counts, dof_keys, dof_iota, dof_temp, dof_stat;      // depth of field
    DevBuf fx_list;           // FXAA: the edge pixels of the frame (count + indices)
    int n_glass = 0;
    int sm_count = 148;
};

RenderState *state(RmContext *ctx) {
    if (!ctx->render_state) {
        auto *s = new (std::nothrow) RenderState();
        if (s) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess) s->sm_count = prop.multiProcessorCount;
        }
        ctx->render_state = s;
    }
    return static_cast<RenderState *>(ctx->render_state);
}

PathQueue make_queue(RenderState *R, int which) {
    PathQueue Q;
    DevBuf *b = R->q[which];
    Q.cap = R->q_cap;
    Q.pixel = b[0].as<int>(); Q.sample = b[1].as<unsigned>(); Q.drawn = b[2].as<unsigned>();
    Q.o = b[3].as<float>(); Q.d = b[4].as<float>(); Q.diff = b[5].as<float>();
    Q.T = b[6].as<float>(); Q.B0 = b[7].as<float>(); Q.W = b[8].as<float>(); Q.rough = b[9].as<float>();
    Q.flags = b[10].as<int>(); Q.med_id = b[11].as<int>(); Q.med = b[12].as<float>();
    Q.hit_t = b[13].as<float>(); Q.hit_face = b[14].as<int>();
    Q.surf = b[15].as<float>(); Q.hdP = b[16].as<float>();
    Q.dec = b[17].as<float>(); Q.skey = b[18].as<int>(); Q.srank = b[19].as<int>();
    return Q;
}

int alloc_queues(RenderState *R, int q_cap, int s_cap) {
    int rc;
    if (q_cap > R->q_cap) {
        const size_t c = size_t(q_cap);
        const size_t words[20] = {1, 1, 1, 3, 3, 12, 3, 3, 1, 1, 1, size_t(kMediumSlots), size_t(4 * kMediumSlots), 1, 1, 22, 6, 7, 1, 1};
        for (int w = 0; w < 2; w++)
            for (int k = 0; k < 20; k++)
                if ((rc = R->q[w][k].alloc(c * words[k] * 4))) return rc;
        if ((rc = R->sorted.alloc(c * sizeof(int)))) return rc;
        R->q_cap = q_cap;
    }
    if (s_cap > R->s_cap) {
        if ((rc = R->shadow.alloc(size_t(s_cap) * sizeof(ShadowItem)))) return rc;
        R->s_cap = s_cap;
    }
    return RM_OK;
}

FrameBuffers frame(RenderState *R) {
    FrameBuffers F;
    F.gbuffer = R->gbuffer.as<RmHitInfo>();
    F.sav_base = R->sav_base.as<float>();
    F.n_ind = R->n_ind.as<int>();
    F.dir_base = R->dir_base.as<int>();
    F.active_list = R->n_active >= 0 ? R->active_list.as<int>() : nullptr;
    F.n_active = R->n_active;
    return F;
}

Accum accum(RenderState *R) {
    Accum A;
    A.rad = R->rad.as<float>();
    A.clum_sum = R->clum_sum.as<float>();
    A.clum_max = R->clum_max.as<float>();
    A.hold_clum = R->hold_clum.as<float>();
    A.hold = R->hold.as<float>();
    A.lock = R->lock.as<int>();
    return A;
}

__global__ void k_fill(float *p, float v, size_t n) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = v;
}

__global__ void k_glass_list(const int *__restrict__ n_ind, int npix, int base, int *list, int *count) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool glass = p < npix && base > 0 && n_ind[p] > base;
    int slot = alloc_slot(count, glass);
    if (glass) list[slot] = p;
}

// what k_direct_gen / k_regen ask of a pixel before they sample it (src/render.cpp:480-489): a hit, and not on a light
struct IsSampledPixel {
    const RmHitInfo *g;
    __host__ __device__ bool operator()(int p) const {
        const float *f = reinterpret_cast<const float *>(g + p);
        const bool hit = isfinite(f[12]) || isfinite(f[13]) || isfinite(f[14]);         // position
        // length(emission) > 0  <=>  its squared length, summed in glm's order, > 0 (a correctly rounded root is positive iff its argument is)
        const float e2 = (f[6] * f[6] + f[7] * f[7]) + f[8] * f[8];
        return hit && !(e2 > 0.0f);
    }
};

constexpr size_t kSlicePad = 64;         // pixels of padding behind the radiance accumulators: a reduce-scatter over <= 64 ranks needs world * ceil(npix / world)

bool same_args(const RmRenderArgs &a, const RmRenderArgs &b) { return std::memcmp(&a, &b, sizeof(RmRenderArgs)) == 0; }

int spp_direct_of(const RmRenderArgs *a) { return int(float(a->spp) * a->P_Direct); }        // src/render.cpp:500

} // namespace

void rm_render_state_free(RmContext *ctx) {
    auto *R = static_cast<RenderState *>(ctx->render_state);
    if (!R) return;
    for (DevBuf *b : {&R->gbuffer, &R->sav_base, &R->n_ind, &R->glass_list, &R->dir_base, &R->active_list, &R->active_tmp, &R->rad, &R->clum_sum, &R->clum_max, &R->hold_clum, &R->hold,
                      &R->lock, &R->shadow, &R->counts, &R->planes[0], &R->planes[1], &R->planes[2], &R->planes[3], &R->g_out, &R->rgb[0], &R->rgb[1],
                      &R->planes_alt[0], &R->planes_alt[1], &R->planes_alt[2], &R->planes_alt[3], &R->f_pm, &R->f_ns, &R->f_op, &R->glow[0], &R->glow[1], &R->dof_depth, &R->dof_src, &R->dof_sorted, &R->dof_lists, &R->dof_counts, &R->dof_keys, &R->dof_iota, &R->dof_temp, &R->dof_stat, &R->fx_list})
        b->release();
    for (int w = 0; w < 2; w++)
        for (int k = 0; k < 20; k++) R->q[w][k].release();
    R->sorted.release();
    if (R->h_counts) cudaFreeHost(R->h_counts);
    delete R;
    ctx->render_state = nullptr;
}

extern "C" {

// ------------------------------------------------------------------------ G-buffer (K2)
int rm_gbuffer(RmContext *ctx, const RmRenderArgs *args, RmHitInfo *gbuffer) {
    if (!ctx || !ctx->has_scene) return rm_fail(RM_ERR_STATE, "rm_gbuffer: no scene uploaded");
    int rc = rm_check_args(args);
    if (rc) return rc;
    RM_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->have_primary || !same_args(ctx->frame_args, *args))
        if ((rc = rm_trace_primary(ctx, args, nullptr, nullptr))) return rc;
    RenderState *R = state(ctx);
    if (!R) return rm_fail(RM_ERR_INVALID, "out of host memory");
    const int npix = args->width * args->height;
    if ((rc = R->gbuffer.alloc(size_t(npix) * sizeof(RmHitInfo))) || (rc = R->sav_base.alloc(size_t(npix) * 12)) ||
        (rc = R->n_ind.alloc(size_t(npix) * 4)) || (rc = R->glass_list.alloc(size_t(npix) * 4)) || (rc = R->dir_base.alloc(size_t(npix) * 4)) ||
        (rc = R->active_list.alloc(size_t(npix) * 4)) || (rc = R->counts.alloc(C_TOTAL * 4)))
        return rc;
    size_t sel_bytes = 0;
    cub::DeviceSelect::If(nullptr, sel_bytes, thrust::counting_iterator<int>(0), R->active_list.as<int>(), R->counts.as<int>() + (C_TOTAL - 1), npix,
                          IsSampledPixel{R->gbuffer.as<RmHitInfo>()}, ctx->stream);
    if ((rc = R->active_tmp.alloc(sel_bytes))) return rc;
    R->npix = npix;
    cudaStream_t st = ctx->stream;
    RM_CUDA(cudaMemsetAsync(R->counts.p, 0, C_TOTAL * 4, st));
    const int spp_d = spp_direct_of(args), base = args->spp - spp_d;
    int *counts = R->counts.as<int>();
    k_gbuffer<<<(npix + 127) / 128, 128, 0, st>>>(ctx->scene, to_dev_args(args), ctx->b_tri_idx.as<int>(), ctx->b_t.as<float>(), frame(R), spp_d, base, counts + 4);
    k_glass_list<<<(npix + 127) / 128, 128, 0, st>>>(R->n_ind.as<int>(), npix, base, R->glass_list.as<int>(), counts + 5);
    // the sampled pixels in pixel order (an ordered stream compaction: neighbouring pixels stay neighbours in the item space);
    // a pixel outside the list keeps dir_base = -1 for good
    RM_CUDA(cub::DeviceSelect::If(R->active_tmp.p, sel_bytes, thrust::counting_iterator<int>(0), R->active_list.as<int>(), counts + (C_TOTAL - 1), npix,
                                  IsSampledPixel{R->gbuffer.as<RmHitInfo>()}, st));
    RM_CUDA(cudaMemsetAsync(R->dir_base.p, 0xff, size_t(npix) * 4, st));
    ctx->launches += 2;
    RM_CUDA(cudaGetLastError());
    int h[8], h_active = 0;
    RM_CUDA(cudaMemcpyAsync(&h_active, counts + (C_TOTAL - 1), 4, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaMemcpyAsync(h, counts, sizeof(h), cudaMemcpyDeviceToHost, st));
    if (gbuffer) RM_CUDA(cudaMemcpyAsync(gbuffer, R->gbuffer.p, size_t(npix) * sizeof(RmHitInfo), cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    R->n_glass = h[5];
    R->n_active = ctx->compact_pixels ? h_active : -1;
    ctx->have_gbuffer = true;
    ctx->have_resolved = false;
    R->accum_valid = false;
    return RM_OK;
}

// ------------------------------------------------------------------------ sample loops
static int reset_accum(RmContext *ctx, RenderState *R) {
    const size_t n = size_t(R->npix);
    int rc;
    if ((rc = R->rad.alloc((n + kSlicePad) * 64)) || (rc = R->clum_sum.alloc(n * 8)) || (rc = R->clum_max.alloc(n * 4)) || (rc = R->hold_clum.alloc(n * 4)) ||
        (rc = R->hold.alloc(n * 32)) || (rc = R->lock.alloc(n * 4)))
        return rc;
    cudaStream_t st = ctx->stream;
    RM_CUDA(cudaMemsetAsync(R->rad.p, 0, (n + kSlicePad) * 64, st));       // the tail: padding of the reduce-scatter slices
    RM_CUDA(cudaMemsetAsync(R->clum_sum.p, 0, n * 8, st));
    RM_CUDA(cudaMemsetAsync(R->clum_max.p, 0, n * 4, st));
    RM_CUDA(cudaMemsetAsync(R->hold.p, 0, n * 32, st));
    RM_CUDA(cudaMemsetAsync(R->lock.p, 0, n * 4, st));
    k_fill<<<R->sm_count * 4, 256, 0, st>>>(R->hold_clum.as<float>(), -1.0f, n);
    ctx->launches++;
    R->accum_valid = true;
    R->hold_committed = false;
    R->slice_first = 0; R->slice_pixels = -1;
    return RM_OK;
}

int rm_render_samples(RmContext *ctx, const RmRenderArgs *args, int32_t sample_begin, int32_t sample_stride, uint64_t seed, int32_t reset) {
    if (!ctx || !ctx->has_scene) return rm_fail(RM_ERR_STATE, "rm_render_samples: no scene uploaded");
    int rc = rm_check_args(args);
    if (rc) return rc;
    if (sample_begin < 0 || sample_stride < 1) return rm_fail(RM_ERR_INVALID, "rm_render_samples: bad sample range");
    if (ctx->scene.sky_width == 0 && ctx->scene.n_lights > kMaxLights)
        return rm_fail(RM_ERR_INVALID, "rm_render_samples: %d light objects, at most %d supported", ctx->scene.n_lights, kMaxLights);
    RM_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->have_gbuffer || !same_args(ctx->frame_args, *args))
        if ((rc = rm_gbuffer(ctx, args, nullptr))) return rc;
    RenderState *R = state(ctx);
    if (reset || !R->accum_valid)
        if ((rc = reset_accum(ctx, R))) return rc;
    if (R->hold_committed) return rm_fail(RM_ERR_STATE, "rm_render_samples: accumulators were already resolved; pass reset=1");
    const int npix = R->npix;
    const int spp_d = spp_direct_of(args), base = args->spp - spp_d;
    auto local_count = [&](int total) { return total > sample_begin ? (total - sample_begin + sample_stride - 1) / sample_stride : 0; };
    const int n_d = local_count(spp_d), n_a = local_count(base);
    const int n_b_total = (R->n_glass > 0) ? local_count(16 * base) : 0;     // local samples below 16*base; the first n_a are phase A's

    // queue sizes: the path queue holds `wave_paths` vertices and is kept full by k_regen; a vertex that
    // terminates emits at most 6 shadow rays; a direct wave is one shadow item per (pixel, sample)
    const long long target = ctx->wave_paths;
    const int S_all = int(std::max(1LL, target / npix));
    const int n_first = R->n_active >= 0 ? R->n_active : npix;          // pixels of the per-pixel stages
    const long long items_a = (long long)n_first * n_a;
    const long long items_b = n_b_total > n_a ? (long long)R->n_glass * (n_b_total - n_a) : 0;
    const long long total_items = items_a + items_b;
    const long long q_cap = std::max(1024LL, std::min(target, total_items));
    const int S_dir = int(std::max<long long>(S_all, std::min<long long>(6 * q_cap / npix, 64)));
    const long long max_direct = n_d > 0 ? (long long)npix * std::min(S_dir, n_d) : 0;
    if (q_cap > 0x7fffffffLL / 8 || max_direct > 0x7fffffffLL / 2) return rm_fail(RM_ERR_INVALID, "rm_render_samples: wave too large");
    if ((rc = alloc_queues(R, int(q_cap), int(std::max({max_direct, 6 * q_cap, 1LL}))))) return rc;
    if (!R->h_counts) RM_CUDA(cudaMallocHost(&R->h_counts, C_COUNT * sizeof(int)));

    cudaStream_t st = ctx->stream;
    int *C = R->counts.as<int>();
    auto *cnt = ctx->b_counters.as<unsigned long long>();
    const DevArgs A = to_dev_args(args);
    const FrameBuffers Fb = frame(R);
    const Accum Ac = accum(R);
    ShadowItem *sq = R->shadow.as<ShadowItem>();
    int *sorted = R->sorted.as<int>();
    const int grid = R->sm_count * 8;
    const int tgrid = R->sm_count * kTraceCtasPerSm;
    const bool ct = ctx->count_tests;
    // bounce and shadow rays: the secondary-ray tree unless the caller asked for the reference's traversal order throughout
    if (n_d + n_a > 0 && !ctx->exact_secondary && (rc = rm_ensure_secondary_tree(ctx))) return rc;          // a deferred build of that tree happens now
    rm_start_refinement(ctx, int64_t(npix) * std::max(n_d + n_a, 1));          // a pending background refinement: worth it for this much rendering?
    if ((rc = rm_install_refined_tree(ctx))) return rc;          // a background-refined tree that became ready since the last call
    const bool use_wide = !ctx->exact_secondary && ctx->have_wide && (ctx->secondary_tree == 2 || !ctx->have_fast);
    const bool use_ref = ctx->exact_secondary || (!use_wide && !ctx->have_fast);
    const DevScene &sec_scene = use_ref ? ctx->scene : (use_wide ? ctx->scene_wide : ctx->scene_fast);     // (a member: follows a later install)
    int sec_levels = use_ref ? ctx->stack_levels : (use_wide ? ctx->stack_levels_wide : ctx->stack_levels_fast);
    const TraceTune sec_tune = use_ref ? ctx->tune : (use_wide ? ctx->tune_wide : ctx->tune_fast);

    // rayHit_test over the first *n_dev items of the shadow queue, then the coalesced accumulation pass
    // (C_CUR_SHADOW must be 0)
    auto trace_shadow = [&](const int *n_dev, int direct_samples) {
        ShadowJob job;
        job.sq = sq;
        ctx->timed_begin(RM_KIND_SHADOW);
        launch_trace(sec_scene, sec_levels, ct, tgrid, st, job, R->s_cap, n_dev, C + C_CUR_SHADOW, cnt + 6, sec_tune);
        ctx->timed_end();
        if (direct_samples > 0) k_accum_direct<<<(std::max(n_first, 1) + 255) / 256, 256, 0, st>>>(Fb, Ac, sq, direct_samples, npix);
        else k_accum_shadow<<<grid, 256, 0, st>>>(Fb, Ac, sq, n_dev, R->s_cap);
        ctx->launches += 2;
    };

    // ---- direct light at the primary hit: waves of S_dir samples per pixel, one shadow item each (as many as the
    // shadow queue holds anyway for the indirect rounds: fewer, fuller launches)
    for (int k0 = 0; k0 < n_d; k0 += S_dir) {
        const int S = std::min(S_dir, n_d - k0);
        RM_CUDA(cudaMemsetAsync(C + C_SQ, 0, 4, st));
        RM_CUDA(cudaMemsetAsync(C + C_CUR_SHADOW, 0, 4, st));
        // environment-lit scenes: a warp per pixel (the sky CDF search dominates and the lanes share the surface);
        // light objects: a thread per pixel (the per-light weights - one BSDF evaluation each - are formed once per pixel);
        // A/B in profiles/r01e_ab15_direct_gen_mapping.txt
        if (S >= 16 && (ctx->direct_warp == 1 || (ctx->direct_warp == 0 && ctx->scene.sky_width != 0)))
            k_direct_gen<true><<<R->sm_count * kCtasDirect, kShadeBlock, 0, st>>>(ctx->scene, A, Fb, S, npix, sample_begin + k0 * sample_stride, sample_stride, spp_d,
                                                                                 seed, sq, C + C_SQ, R->s_cap, C + C_OVERFLOW);
        else
            k_direct_gen<false><<<R->sm_count * kCtasDirect, kShadeBlock, 0, st>>>(ctx->scene, A, Fb, S, npix, sample_begin + k0 * sample_stride, sample_stride, spp_d,
                                                                                  seed, sq, C + C_SQ, R->s_cap, C + C_OVERFLOW);
        ctx->launches++;
        trace_shadow(C + C_SQ, S);
    }

    // ---- indirect paths.  One round = every vertex in the path queue advances by one bounce:
    //   plan + regen (top the queue up with fresh first vertices) -> closest hit -> surface -> bounce
    //   (-> next queue, NEE requests) -> NEE -> visibility -> accumulate.
    // The host issues rounds in batches and looks at the device counters between batches: the loop ends when
    // every item has been handed out and the queue has drained.
    if (total_items > 0) {
        ItemSpace I;
        I.items_a = items_a; I.total = total_items; I.npix = npix; I.n_glass = std::max(R->n_glass, 1); I.n_a = n_a;
        I.glass_list = R->glass_list.as<int>(); I.s_begin = sample_begin; I.s_stride = sample_stride;
        RM_CUDA(cudaMemsetAsync(C, 0, 2 * sizeof(int), st));                       // both path queues empty
        RM_CUDA(cudaMemsetAsync(C + C_NEE, 0, (C_COUNT - C_NEE) * sizeof(int), st)); // item cursor = 0
        RM_CUDA(cudaMemsetAsync(C + C_SQ, 0, sizeof(int), st));
        const int shadow_threshold = std::max(1, R->q_cap / 2);
        int cur = 0;
        const int batch = 4;
        const int max_rounds = ctx->max_depth < kMaxRayDepth ? ctx->max_depth : 0x7fffffff;   // perf experiments only
        int rounds = 0;
        bool done = false;
        while (!done) {
            for (int b = 0; b < batch && rounds < max_rounds; b++, rounds++) {
                PathQueue Qin = make_queue(R, cur), Qout = make_queue(R, cur ^ 1);
                k_plan<<<1, 1, 0, st>>>(C, cur, Qin.cap, total_items);
                k_regen<<<R->sm_count * kCtasRegen, kShadeBlock, 0, st>>>(ctx->scene, A, Fb, I, C, seed, Qin, C + cur);
                PathJob pj;
                pj.Q = Qin;
                ctx->timed_begin(RM_KIND_PATHS);
                launch_trace(sec_scene, sec_levels, ct, tgrid, st, pj, Qin.cap, C + cur, C + C_CUR_PATH, cnt + 3, sec_tune);
                ctx->timed_end();
                ctx->timed_begin(RM_KIND_SHADE);
                k_surface<<<R->sm_count * kCtasSurface, kShadeBlock, 0, st>>>(ctx->scene, Fb, Ac, Qin, C, cur);
                k_decide<<<R->sm_count * RM_CTAS_DECIDE, kShadeBlock, 0, st>>>(seed, Qin, C + cur, C + C_BINS);
                k_sort_offsets<<<1, kSortBins, 0, st>>>(C + C_BINS, C + C_OFFS);
                k_sort_scatter<<<R->sm_count * 8, 256, 0, st>>>(Qin, C + cur, C + C_OFFS, sorted);
                k_continue<false><<<R->sm_count * kCtasBounce, kShadeBlock, 0, st>>>(seed, Qin, Qout, C + (cur ^ 1), sorted, C + C_OFFS);
                k_continue<true><<<R->sm_count * kCtasBounce, kShadeBlock, 0, st>>>(seed, Qin, Qout, C + (cur ^ 1), sorted, C + C_OFFS);
                k_nee<<<R->sm_count * kCtasNee, kShadeBlock, 0, st>>>(ctx->scene, Fb, seed, Qin, sorted, C + C_OFFS, sq, C + C_SQ, R->s_cap, C + C_OVERFLOW);
                ctx->timed_end();
                k_shadow_gate<<<1, 1, 0, st>>>(C, shadow_threshold, R->s_cap, 0);
                ctx->launches += 11;
                trace_shadow(C + C_SQ_RUN, 0);
                cur ^= 1;
            }
            RM_CUDA(cudaMemcpyAsync(R->h_counts, C, C_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
            RM_CUDA(cudaStreamSynchronize(st));
            if (use_wide && ctx->refine) {                  // the host waits here anyway: swap the refined tree in if it has arrived
                if ((rc = rm_install_refined_tree(ctx))) return rc;
                sec_levels = ctx->stack_levels_wide;
            }
            const long long handed = (long long)(unsigned)R->h_counts[C_ITEM_LO] | ((long long)R->h_counts[C_ITEM_HI] << 32);
            done = (handed >= total_items && R->h_counts[cur] == 0) || rounds >= max_rounds;
        }
        k_plan<<<1, 1, 0, st>>>(C, cur, R->q_cap, total_items);          // retires a queue traced in the last round
        k_shadow_gate<<<1, 1, 0, st>>>(C, shadow_threshold, R->s_cap, 1);
        trace_shadow(C + C_SQ_RUN, 0);
        ctx->launches += 2;
    }
    RM_CUDA(cudaGetLastError());
    ctx->have_resolved = false;
    return RM_OK;
}

int rm_accum_view(RmContext *ctx, float **d_sum, int64_t *n_sum, float **d_max, int64_t *n_max) {
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "rm_accum_view: nothing rendered yet");
    RenderState *R = state(ctx);
    if (!R->accum_valid) return rm_fail(RM_ERR_STATE, "rm_accum_view: nothing rendered yet");
    RM_CUDA(cudaSetDevice(ctx->device));
    // publish the local held-back luminance into the max-reduced plane
    if (!R->hold_committed) {
        k_publish_max<<<(R->npix + 255) / 256, 256, 0, ctx->stream>>>(accum(R), R->npix);
        ctx->launches++;
    }
    if (d_sum) *d_sum = R->clum_sum.as<float>();
    if (n_sum) *n_sum = int64_t(R->npix) * 2;
    if (d_max) *d_max = R->clum_max.as<float>();
    if (n_max) *n_max = R->npix;
    return RM_OK;
}

// After the firefly side data {clum_sum (sum), clum_max (max)} has been reduced across ranks:
// commit every rank's held-back sample against the global totals.  Afterwards the 16-float
// radiance accumulators (returned through d_sum/n_sum of a second rm_accum_view call... see
// rm_accum_radiance) can be summed across ranks.
int rm_accum_after_reduce(RmContext *ctx, int32_t rank, int32_t world) {
    (void)rank; (void)world;
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "rm_accum_after_reduce: nothing rendered yet");
    RenderState *R = state(ctx);
    if (!R->accum_valid) return rm_fail(RM_ERR_STATE, "rm_accum_after_reduce: nothing rendered yet");
    RM_CUDA(cudaSetDevice(ctx->device));
    if (!R->hold_committed) {
        k_commit_hold<<<(R->npix + 255) / 256, 256, 0, ctx->stream>>>(accum(R), frame(R), R->npix, ctx->disable_clamp);
        ctx->launches++;
        R->hold_committed = true;
    }
    return RM_OK;
}

int rm_accum_radiance(RmContext *ctx, float **d_rad, int64_t *n_rad) {
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "rm_accum_radiance: nothing rendered yet");
    RenderState *R = state(ctx);
    if (!R->accum_valid || !R->hold_committed) return rm_fail(RM_ERR_STATE, "rm_accum_radiance: call rm_accum_after_reduce first");
    if (d_rad) *d_rad = R->rad.as<float>();
    if (n_rad) *n_rad = int64_t(R->npix) * 16;
    return RM_OK;
}

// ------------------------------------------------------------------------ progressive checkpoint / resume
// The un-finalised accumulators of the current frame as one blob: a 64-byte header, then per plane-of-floats
// rad[npix][16], clum_sum[npix][2], clum_max[npix], hold_clum[npix], hold[npix][8].  A render stopped after some sample
// shards (rm_render_samples with reset = 0 adds more) can be resumed later, in another context or on another GPU.
namespace {
struct CheckpointHeader {
    char magic[4];
    int32_t version, width, height, spp;
    float P_Direct;
    int32_t hold_committed;
    int32_t _pad[9];
};
static_assert(sizeof(CheckpointHeader) == 64, "checkpoint header is 64 bytes");
constexpr int kCkFloatsPerPixel = 16 + 2 + 1 + 1 + 8;
} // namespace

int64_t rm_checkpoint_bytes(const RmRenderArgs *args) {
    if (rm_check_args(args)) return -1;
    return int64_t(sizeof(CheckpointHeader)) + int64_t(args->width) * args->height * kCkFloatsPerPixel * 4;
}

int rm_checkpoint_save(RmContext *ctx, void *host, int64_t bytes) {
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "rm_checkpoint_save: nothing rendered yet");
    RenderState *R = state(ctx);
    if (!R->accum_valid) return rm_fail(RM_ERR_STATE, "rm_checkpoint_save: nothing rendered yet");
    const RmRenderArgs &a = ctx->frame_args;
    if (!host || bytes != rm_checkpoint_bytes(&a)) return rm_fail(RM_ERR_INVALID, "rm_checkpoint_save: buffer must hold rm_checkpoint_bytes() bytes");
    RM_CUDA(cudaSetDevice(ctx->device));
    CheckpointHeader h{};
    std::memcpy(h.magic, "RMCK", 4);
    h.version = 1; h.width = a.width; h.height = a.height; h.spp = a.spp; h.P_Direct = a.P_Direct;
    h.hold_committed = R->hold_committed ? 1 : 0;
    std::memcpy(host, &h, sizeof(h));
    const size_t n = size_t(R->npix);
    char *dst = static_cast<char *>(host) + sizeof(h);
    const DevBuf *src[5] = {&R->rad, &R->clum_sum, &R->clum_max, &R->hold_clum, &R->hold};
    const size_t words[5] = {16, 2, 1, 1, 8};
    for (int k = 0; k < 5; k++) {
        RM_CUDA(cudaMemcpyAsync(dst, src[k]->p, n * words[k] * 4, cudaMemcpyDeviceToHost, ctx->stream));
        dst += n * words[k] * 4;
    }
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    return RM_OK;
}

int rm_checkpoint_load(RmContext *ctx, const RmRenderArgs *args, const void *host, int64_t bytes) {
    if (!ctx || !ctx->has_scene) return rm_fail(RM_ERR_STATE, "rm_checkpoint_load: no scene uploaded");
    int rc = rm_check_args(args);
    if (rc) return rc;
    if (!host || bytes != rm_checkpoint_bytes(args)) return rm_fail(RM_ERR_INVALID, "rm_checkpoint_load: size does not match these args");
    CheckpointHeader h;
    std::memcpy(&h, host, sizeof(h));
    if (std::memcmp(h.magic, "RMCK", 4) || h.version != 1) return rm_fail(RM_ERR_INVALID, "rm_checkpoint_load: not a checkpoint");
    if (h.width != args->width || h.height != args->height || h.spp != args->spp || h.P_Direct != args->P_Direct)
        return rm_fail(RM_ERR_INVALID, "rm_checkpoint_load: checkpoint was made for %dx%d spp %d P_Direct %g", h.width, h.height, h.spp, h.P_Direct);
    RM_CUDA(cudaSetDevice(ctx->device));
    // the frame state the samplers read (primary hits, G-buffer) is a pure function of scene + args: recompute it
    if (!ctx->have_gbuffer || !same_args(ctx->frame_args, *args))
        if ((rc = rm_gbuffer(ctx, args, nullptr))) return rc;
    RenderState *R = state(ctx);
    if ((rc = reset_accum(ctx, R))) return rc;
    const size_t n = size_t(R->npix);
    const char *src = static_cast<const char *>(host) + sizeof(h);
    DevBuf *dst[5] = {&R->rad, &R->clum_sum, &R->clum_max, &R->hold_clum, &R->hold};
    const size_t words[5] = {16, 2, 1, 1, 8};
    for (int k = 0; k < 5; k++) {
        RM_CUDA(cudaMemcpyAsync(dst[k]->p, src, n * words[k] * 4, cudaMemcpyHostToDevice, ctx->stream));
        src += n * words[k] * 4;
    }
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    R->hold_committed = h.hold_committed != 0;
    ctx->have_resolved = false;
    return RM_OK;
}

int rm_resolve(RmContext *ctx, const RmRenderArgs *args, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is) {
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "rm_resolve: nothing rendered yet");
    int rc = rm_check_args(args);
    if (rc) return rc;
    RenderState *R = state(ctx);
    if (!R->accum_valid || R->npix != args->width * args->height) return rm_fail(RM_ERR_STATE, "rm_resolve: accumulators do not match these args");
    if (R->slice_pixels >= 0) return rm_fail(RM_ERR_STATE, "rm_resolve: after rm_reduce_scatter this rank holds a slice of the frame only; call rm_resolve_slice");
    RM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int npix = R->npix;
    if (!R->hold_committed) {          // single-GPU path: local totals are the global totals
        k_publish_max<<<(npix + 255) / 256, 256, 0, st>>>(accum(R), npix);
        k_commit_hold<<<(npix + 255) / 256, 256, 0, st>>>(accum(R), frame(R), npix, ctx->disable_clamp);
        ctx->launches += 2;
        R->hold_committed = true;
    }
    for (int k = 0; k < 4; k++)
        if ((rc = R->planes[k].alloc(size_t(npix) * sizeof(RmRadiance)))) return rc;
    if ((rc = R->g_out.alloc(size_t(npix) * sizeof(RmHitInfo)))) return rc;
    k_finalise<<<(npix + 127) / 128, 128, 0, st>>>(accum(R), frame(R), 0, npix, args->exposure, R->planes[0].as<RmRadiance>(), R->planes[1].as<RmRadiance>(),
                                                   R->planes[2].as<RmRadiance>(), R->planes[3].as<RmRadiance>(), R->g_out.as<RmHitInfo>());
    ctx->launches++;
    RM_CUDA(cudaGetLastError());
    RmRadiance *host[4] = {Dd, Ds, Id, Is};
    for (int k = 0; k < 4; k++)
        if (host[k]) { RM_CUDA(cudaMemcpyAsync(host[k], R->planes[k].p, size_t(npix) * sizeof(RmRadiance), cudaMemcpyDeviceToHost, st)); }
    int h[8];
    RM_CUDA(cudaMemcpyAsync(h, R->counts.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    if (h[3]) return rm_fail(RM_ERR_STATE, "rm_resolve: a shadow queue overflowed during rendering (results incomplete)");
    ctx->have_resolved = true;
    return RM_OK;
}

int rm_accum_mark_slice(RmContext *ctx, int64_t first_pixel, int64_t pixels) {
    RenderState *R = state(ctx);
    if (!R || !R->accum_valid) return rm_fail(RM_ERR_STATE, "rm_reduce_scatter: nothing rendered yet");
    R->slice_first = std::min<int64_t>(first_pixel, R->npix);
    R->slice_pixels = std::max<int64_t>(0, std::min<int64_t>(pixels, R->npix - R->slice_first));
    return RM_OK;
}

int rm_frame_slice(RmContext *ctx, int64_t *first_pixel, int64_t *pixels) {
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "rm_frame_slice: nothing rendered yet");
    RenderState *R = state(ctx);
    if (first_pixel) *first_pixel = R->slice_pixels >= 0 ? R->slice_first : 0;
    if (pixels) *pixels = R->slice_pixels >= 0 ? R->slice_pixels : R->npix;
    return RM_OK;
}

// rm_resolve for the slice of the frame this rank holds after rm_reduce_scatter (the whole frame when there was none): the
// host pointers address WHOLE-frame arrays - e.g. one pinned / shared-memory frame every rank of the box maps - and only
// this rank's pixels [first, first + count) of them are written; any may be NULL.
int rm_resolve_slice(RmContext *ctx, const RmRenderArgs *args, RmHitInfo *gbuffer, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is) {
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "rm_resolve_slice: nothing rendered yet");
    int rc = rm_check_args(args);
    if (rc) return rc;
    RenderState *R = state(ctx);
    if (!R->accum_valid || R->npix != args->width * args->height) return rm_fail(RM_ERR_STATE, "rm_resolve_slice: accumulators do not match these args");
    RM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int npix = R->npix;
    if (!R->hold_committed) {          // no exchange happened: local totals are the global totals
        k_publish_max<<<(npix + 255) / 256, 256, 0, st>>>(accum(R), npix);
        k_commit_hold<<<(npix + 255) / 256, 256, 0, st>>>(accum(R), frame(R), npix, ctx->disable_clamp);
        ctx->launches += 2;
        R->hold_committed = true;
    }
    const int first = R->slice_pixels >= 0 ? int(R->slice_first) : 0, count = R->slice_pixels >= 0 ? int(R->slice_pixels) : npix;
    for (int k = 0; k < 4; k++)
        if ((rc = R->planes[k].alloc(size_t(npix) * sizeof(RmRadiance)))) return rc;
    if ((rc = R->g_out.alloc(size_t(npix) * sizeof(RmHitInfo)))) return rc;
    if (count > 0) {
        k_finalise<<<(count + 127) / 128, 128, 0, st>>>(accum(R), frame(R), first, first + count, args->exposure, R->planes[0].as<RmRadiance>(), R->planes[1].as<RmRadiance>(),
                                                        R->planes[2].as<RmRadiance>(), R->planes[3].as<RmRadiance>(), R->g_out.as<RmHitInfo>());
        ctx->launches++;
        RM_CUDA(cudaGetLastError());
        RmRadiance *host[4] = {Dd, Ds, Id, Is};
        for (int k = 0; k < 4; k++)
            if (host[k]) RM_CUDA(cudaMemcpyAsync(host[k] + first, R->planes[k].as<RmRadiance>() + first, size_t(count) * sizeof(RmRadiance), cudaMemcpyDeviceToHost, st));
        if (gbuffer) RM_CUDA(cudaMemcpyAsync(gbuffer + first, R->g_out.as<RmHitInfo>() + first, size_t(count) * sizeof(RmHitInfo), cudaMemcpyDeviceToHost, st));
    }
    int h[8];
    RM_CUDA(cudaMemcpyAsync(h, R->counts.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    if (h[3]) return rm_fail(RM_ERR_STATE, "rm_resolve_slice: a shadow queue overflowed during rendering (results incomplete)");
    ctx->have_resolved = R->slice_pixels < 0;         // the image-space passes need the whole frame on one device
    return RM_OK;
}

int rm_download_resolved(RmContext *ctx, RmHitInfo *gbuffer, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is) {
    if (!ctx || !ctx->render_state || !ctx->have_resolved) return rm_fail(RM_ERR_STATE, "rm_download_resolved: call rm_resolve first");
    RenderState *R = state(ctx);
    RM_CUDA(cudaSetDevice(ctx->device));
    const size_t npix = size_t(R->npix);
    RmRadiance *host[4] = {Dd, Ds, Id, Is};
    for (int k = 0; k < 4; k++)
        if (host[k]) RM_CUDA(cudaMemcpyAsync(host[k], R->planes[k].p, npix * sizeof(RmRadiance), cudaMemcpyDeviceToHost, ctx->stream));
    if (gbuffer) RM_CUDA(cudaMemcpyAsync(gbuffer, R->g_out.p, npix * sizeof(RmHitInfo), cudaMemcpyDeviceToHost, ctx->stream));
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    return RM_OK;
}

int rm_render(RmContext *ctx, const RmRenderArgs *args, uint64_t seed, RmHitInfo *gbuffer, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is) {
    int rc;
    if ((rc = rm_trace_primary(ctx, args, nullptr, nullptr))) return rc;
    if ((rc = rm_gbuffer(ctx, args, nullptr))) return rc;
    if ((rc = rm_render_samples(ctx, args, 0, 1, seed, 1))) return rc;
    if ((rc = rm_resolve(ctx, args, Dd, Ds, Id, Is))) return rc;
    if (gbuffer) {
        RenderState *R = state(ctx);
        RM_CUDA(cudaMemcpyAsync(gbuffer, R->g_out.p, size_t(R->npix) * sizeof(RmHitInfo), cudaMemcpyDeviceToHost, ctx->stream));
        RM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return RM_OK;
}

// ------------------------------------------------------------------------ image-space passes over the resolved planes
namespace {
Planes4 planes_of(DevBuf *b) {
    Planes4 P;
    for (int k = 0; k < 4; k++) P.p[k] = b[k].as<RmRadiance>();
    return P;
}
void swap_planes(RenderState *R) {
    for (int k = 0; k < 4; k++) std::swap(R->planes[k], R->planes_alt[k]);
}
int post_ready(RmContext *ctx, const RmRenderArgs *args, const char *who, RenderState **out) {
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "%s: nothing resolved yet", who);
    int rc = rm_check_args(args);
    if (rc) return rc;
    RenderState *R = state(ctx);
    if (!ctx->have_resolved || R->npix != args->width * args->height) return rm_fail(RM_ERR_STATE, "%s: call rm_resolve for these args first", who);
    RM_CUDA(cudaSetDevice(ctx->device));
    for (int k = 0; k < 4; k++)
        if ((rc = R->planes_alt[k].alloc(size_t(R->npix) * sizeof(RmRadiance)))) return rc;
    *out = R;
    return RM_OK;
}
} // namespace

// Stage host-side Photo buffers (G-buffer + the four planes) as the context's resolved frame, so the image-space
// passes can run on planes produced elsewhere (e.g. by the reference's own CPU render).
int rm_upload_resolved(RmContext *ctx, const RmRenderArgs *args, const RmHitInfo *gbuffer, const RmRadiance *Dd, const RmRadiance *Ds,
                       const RmRadiance *Id, const RmRadiance *Is) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "context is NULL");
    int rc = rm_check_args(args);
    if (rc) return rc;
    if (!gbuffer || !Dd || !Ds || !Id || !Is) return rm_fail(RM_ERR_INVALID, "rm_upload_resolved: NULL buffer");
    RM_CUDA(cudaSetDevice(ctx->device));
    RenderState *R = state(ctx);
    if (!R) return rm_fail(RM_ERR_INVALID, "out of host memory");
    const size_t npix = size_t(args->width) * args->height;
    const RmRadiance *host[4] = {Dd, Ds, Id, Is};
    for (int k = 0; k < 4; k++) {
        if ((rc = R->planes[k].alloc(npix * sizeof(RmRadiance)))) return rc;
        RM_CUDA(cudaMemcpyAsync(R->planes[k].p, host[k], npix * sizeof(RmRadiance), cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = R->g_out.alloc(npix * sizeof(RmHitInfo)))) return rc;
    RM_CUDA(cudaMemcpyAsync(R->g_out.p, gbuffer, npix * sizeof(RmHitInfo), cudaMemcpyHostToDevice, ctx->stream));
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (R->npix != int(npix)) { R->accum_valid = false; ctx->have_gbuffer = false; ctx->have_primary = false; }
    R->npix = int(npix);
    ctx->have_resolved = true;
    return RM_OK;
}

int rm_spatial_clamp(RmContext *ctx, const RmRenderArgs *args) {
    RenderState *R = nullptr;
    int rc = post_ready(ctx, args, "rm_spatial_clamp", &R);
    if (rc) return rc;
    dim3 grid((args->width + kClW - 1) / kClW, (args->height + kClH - 1) / kClH, 4), block(kClW, kClH);
    k_spatial_clamp<<<grid, block, 0, ctx->stream>>>(planes_of(R->planes), planes_of(R->planes_alt), args->width, args->height);
    ctx->launches++;
    RM_CUDA(cudaGetLastError());
    swap_planes(R);
    return RM_OK;
}

int rm_filter(RmContext *ctx, const RmRenderArgs *args) {
    RenderState *R = nullptr;
    int rc = post_ready(ctx, args, "rm_filter", &R);
    if (rc) return rc;
    const int npix = R->npix, w = args->width, h = args->height;
    if ((rc = R->f_pm.alloc(size_t(npix) * 16)) || (rc = R->f_ns.alloc(size_t(npix) * 16)) || (rc = R->f_op.alloc(size_t(npix) * 4))) return rc;
    cudaStream_t st = ctx->stream;
    FilterG F;
    F.pm = R->f_pm.as<float4>(); F.ns = R->f_ns.as<float4>(); F.opacity = R->f_op.as<float>();
    k_filter_pack<<<(npix + 255) / 256, 256, 0, st>>>(R->g_out.as<RmHitInfo>(), F, npix);
    dim3 block(32, 4), grid((w + 31) / 32, (h + 3) / 4, 4);
    k_filter_var<<<grid, block, 0, st>>>(planes_of(R->planes), planes_of(R->planes_alt), w, h);
    swap_planes(R);
    ctx->launches += 2;
    dim3 grid1((w + 31) / 32, (h + 3) / 4, 1);
    for (int step = 1; step <= 16; step *= 2) {
        k_atrous<<<grid1, block, 0, st>>>(R->g_out.as<RmHitInfo>(), F, planes_of(R->planes), planes_of(R->planes_alt), w, h, step);
        swap_planes(R);
        ctx->launches++;
    }
    RM_CUDA(cudaGetLastError());
    return RM_OK;
}

// Photo::depthFeildBlur (src/image.cpp:285-356) with focus / CoC / cameraPosition from the render arguments
// (src/render.cpp:665-668): d_in -> d_out, both device rgb frames of the resolved frame's size.
static int dof_device(RmContext *ctx, RenderState *R, const RmRenderArgs *args, const float *d_in, float *d_out) {
    int rc;
    cudaStream_t st = ctx->stream;
    const int npix = R->npix;
    const int w = args->width, h = args->height;
    if ((rc = R->dof_depth.alloc(size_t(npix) * 4)) || (rc = R->dof_src.alloc(size_t(npix) * 8)) || (rc = R->dof_sorted.alloc(size_t(npix) * 4))) return rc;
    const V3 cam = {args->position[0], args->position[1], args->position[2]};
    k_dof_prepare<<<(npix + 127) / 128, 128, 0, st>>>(R->g_out.as<RmHitInfo>(), cam, args->focus, args->CoC, npix, R->dof_depth.as<float>(), R->dof_src.as<float2>());
    // The visiting order is the reference's stable sort of the pixels by camera distance (src/image.cpp:300-303).  Distances are
    // >= +0, so their bit patterns order like the floats and a stable LSD radix sort of (bits, pixel) gives that very order -
    // as long as no distance is NaN.  A pixel without a hit has a NaN distance, and with NaNs in the range the reference's
    // comparator is no ordering at all: where the other pixels end up is then a property of libstdc++'s merge sort, which only
    // that routine reproduces - such frames take the host path below (same comparator, same library routine).
    if ((rc = R->dof_stat.alloc(64))) return rc;
    int *stat = R->dof_stat.as<int>();              // [0] NaN distances, [1] largest circle of confusion (whole pixels)
    RM_CUDA(cudaMemsetAsync(stat, 0, 8, st));
    k_dof_stats<<<R->sm_count * 4, 256, 0, st>>>(R->dof_depth.as<float>(), R->dof_src.as<float2>(), npix, stat);
    int h_stat[2] = {0, 0};
    RM_CUDA(cudaMemcpyAsync(h_stat, stat, 8, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    ctx->launches += 2;
    const int reach = h_stat[1];
    if (h_stat[0] == 0) {
        if ((rc = R->dof_keys.alloc(size_t(npix) * 4)) || (rc = R->dof_iota.alloc(size_t(npix) * 4))) return rc;
        k_iota<<<R->sm_count * 4, 256, 0, st>>>(R->dof_iota.as<int>(), npix);
        size_t temp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, temp, R->dof_depth.as<unsigned>(), R->dof_keys.as<unsigned>(), R->dof_iota.as<int>(), R->dof_sorted.as<int>(), npix, 0, 32, st);
        if ((rc = R->dof_temp.alloc(temp))) return rc;
        RM_CUDA(cub::DeviceRadixSort::SortPairs(R->dof_temp.p, temp, R->dof_depth.as<unsigned>(), R->dof_keys.as<unsigned>(), R->dof_iota.as<int>(), R->dof_sorted.as<int>(), npix,
                                                0, 32, st));
        ctx->launches += 2;
    } else {
        std::vector<float> depth(npix);
        RM_CUDA(cudaMemcpyAsync(depth.data(), R->dof_depth.p, size_t(npix) * 4, cudaMemcpyDeviceToHost, st));
        RM_CUDA(cudaStreamSynchronize(st));
        struct Px { int idx; float depth; };
        std::vector<Px> px(npix);
        for (int i = 0; i < npix; i++) px[i] = {i, depth[i]};
        std::stable_sort(px.begin(), px.end(), [](const Px &a, const Px &b) { return a.depth < b.depth; });
        std::vector<int> sorted(npix);
        for (int i = 0; i < npix; i++) sorted[i] = px[i].idx;
        RM_CUDA(cudaMemcpyAsync(R->dof_sorted.p, sorted.data(), size_t(npix) * 4, cudaMemcpyHostToDevice, st));
        RM_CUDA(cudaStreamSynchronize(st));        // `sorted` dies at scope exit
    }
    const int tiles_x = (w + kDofTile - 1) / kDofTile, tiles_y = (h + kDofTile - 1) / kDofTile, tiles = tiles_x * tiles_y;
    const long long side = kDofTile + 2LL * reach;
    const long long cap = std::min<long long>(npix, side * side);
    if (cap * tiles > (1LL << 31)) return rm_fail(RM_ERR_INVALID, "depth of field: a reach of %d px needs %lld list entries", reach, cap * tiles);
    if ((rc = R->dof_lists.alloc(size_t(cap) * tiles * 4)) || (rc = R->dof_counts.alloc(size_t(tiles) * 4))) return rc;
    k_dof_tile_lists<<<tiles, 256, 0, st>>>(R->dof_sorted.as<int>(), npix, w, h, reach, tiles_x, int(cap), R->dof_lists.as<int>(), R->dof_counts.as<int>());
    k_dof_gather<<<tiles, dim3(kDofTile, kDofTile), 0, st>>>(d_in, R->dof_src.as<float2>(), R->dof_lists.as<int>(), R->dof_counts.as<int>(), int(cap), w, h, tiles_x, d_out);
    ctx->launches += 3;
    RM_CUDA(cudaGetLastError());
    return RM_OK;
}

// ------------------------------------------------------------------------ post pass
int rm_fxaa_device(RmContext *ctx, const float *d_rgb_in, float *d_rgb_out, int32_t width, int32_t height) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "context is NULL");
    if (!d_rgb_in || !d_rgb_out || width <= 0 || height <= 0) return rm_fail(RM_ERR_INVALID, "rm_fxaa_device: bad arguments");
    if (d_rgb_in == d_rgb_out) return rm_fail(RM_ERR_INVALID, "rm_fxaa_device: in-place operation is not supported");
    RM_CUDA(cudaSetDevice(ctx->device));
    RenderState *R = state(ctx);
    if (!R) return rm_fail(RM_ERR_INVALID, "out of host memory");
    // one launch when every row is a whole number of 16-byte vectors (kernels_post.cuh: k_fxaa_strip)
    if (ctx->fxaa_rows && (width & 3) == 0 && ((reinterpret_cast<uintptr_t>(d_rgb_in) | reinterpret_cast<uintptr_t>(d_rgb_out)) & 15) == 0) {
        // A warp (= a CTA) walks one strip of 128 x rows pixels: first the rows (bulk copies in and out, bound by memory), then its
        // edge pixels (dependent gathers, bound by latency).  Short strips, several per resident warp, so that the two phases of
        // different warps overlap on an SM and the block scheduler evens out the finish; at least 4 rows (a strip re-reads two
        // halo rows).  "fxaa_rows" pins the height.
        static int ctas_per_sm = 0;
        if (!ctas_per_sm) {
            RM_CUDA(cudaFuncSetAttribute(k_fxaa_strip, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            RM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_fxaa_strip, 32 * kFxStripWarps, kFxStripWarps * kFxWarpBytes));
            if (ctas_per_sm < 1) ctas_per_sm = 1;
        }
        const int spans = (width + 127) / 128, slots = ctas_per_sm * R->sm_count * kFxStripWarps;
        int rows = ctx->fxaa_rows;
        if (ctx->fxaa_auto) {
            rows = 8;
            while (rows > 4 && int64_t(spans) * ((height + rows - 1) / rows) < 2 * int64_t(slots)) rows--;
        }
        const int strips = spans * ((height + rows - 1) / rows);
        const int grid = (strips + kFxStripWarps - 1) / kFxStripWarps;
        k_fxaa_strip<<<grid, 32 * kFxStripWarps, kFxStripWarps * kFxWarpBytes, ctx->stream>>>(d_rgb_in, d_rgb_out, width, height, rows);
        ctx->launches += 1;
        RM_CUDA(cudaGetLastError());
        return RM_OK;
    }
    int rc;
    if ((rc = R->fx_list.alloc(size_t(width) * height * 4 + 16))) return rc;
    int *list = R->fx_list.as<int>() + 4, *count = R->fx_list.as<int>();           // [0] the number of edge pixels, [4..] their indices
    RM_CUDA(cudaMemsetAsync(count, 0, 4, ctx->stream));
    dim3 grid((width + kFxTileW - 1) / kFxTileW, (height + kFxTileH - 1) / kFxTileH), block(kFxTileW, kFxTileH);
    k_fxaa<<<grid, block, 0, ctx->stream>>>(d_rgb_in, d_rgb_out, width, height, list, count);
    // one thread per edge pixel for up to ~1.2 M of them (a 4K frame has 0.4 - 1 M): the pass is a chain of dependent gathers per
    // pixel, so it wants all of them in flight at once; blocks beyond the list's length exit at once
    k_fxaa_edges<<<R->sm_count * 32, 256, 0, ctx->stream>>>(d_rgb_in, d_rgb_out, width, height, list, count);
    ctx->launches += 2;
    RM_CUDA(cudaGetLastError());
    return RM_OK;
}

int rm_fxaa(RmContext *ctx, const float *rgb_in, float *rgb_out, int32_t width, int32_t height) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "context is NULL");
    if (!rgb_in || !rgb_out || width <= 0 || height <= 0) return rm_fail(RM_ERR_INVALID, "rm_fxaa: bad arguments");
    RM_CUDA(cudaSetDevice(ctx->device));
    RenderState *R = state(ctx);
    if (!R) return rm_fail(RM_ERR_INVALID, "out of host memory");
    const size_t bytes = size_t(width) * height * 12;
    int rc;
    if ((rc = R->rgb[0].alloc(bytes)) || (rc = R->rgb[1].alloc(bytes))) return rc;
    RM_CUDA(cudaMemcpyAsync(R->rgb[0].p, rgb_in, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = rm_fxaa_device(ctx, R->rgb[0].as<float>(), R->rgb[1].as<float>(), width, height))) return rc;
    RM_CUDA(cudaMemcpyAsync(rgb_out, R->rgb[1].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    return RM_OK;
}

// Photo::depthFeildBlur on a caller-supplied rgb frame (host buffers), with the context's resolved G-buffer
int rm_depth_field_blur(RmContext *ctx, const RmRenderArgs *args, const float *rgb_in, float *rgb_out) {
    RenderState *R = nullptr;
    int rc = post_ready(ctx, args, "rm_depth_field_blur", &R);
    if (rc) return rc;
    if (!rgb_in || !rgb_out) return rm_fail(RM_ERR_INVALID, "rm_depth_field_blur: NULL buffer");
    const size_t bytes = size_t(R->npix) * 12;
    if ((rc = R->rgb[0].alloc(bytes)) || (rc = R->rgb[1].alloc(bytes))) return rc;
    RM_CUDA(cudaMemcpyAsync(R->rgb[0].p, rgb_in, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = dof_device(ctx, R, args, R->rgb[0].as<float>(), R->rgb[1].as<float>()))) return rc;
    RM_CUDA(cudaMemcpyAsync(rgb_out, R->rgb[1].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RM_CUDA(cudaStreamSynchronize(ctx->stream));
    return RM_OK;
}

int rm_postprocess(RmContext *ctx, const RmRenderArgs *args, int32_t shade_options, float *rgb_out) {
    if (!ctx || !ctx->render_state) return rm_fail(RM_ERR_STATE, "rm_postprocess: nothing resolved yet");
    int rc = rm_check_args(args);
    if (rc) return rc;
    RenderState *R = state(ctx);
    if (!ctx->have_resolved || R->npix != args->width * args->height) return rm_fail(RM_ERR_STATE, "rm_postprocess: call rm_resolve for these args first");
    if (!rgb_out) return rm_fail(RM_ERR_INVALID, "rm_postprocess: rgb_out is NULL");
    RM_CUDA(cudaSetDevice(ctx->device));
    const int npix = R->npix;
    const size_t bytes = size_t(npix) * 12;
    if ((rc = R->rgb[0].alloc(bytes)) || (rc = R->rgb[1].alloc(bytes))) return rc;
    cudaStream_t st = ctx->stream;
    // Photo::postProcessing (src/image.cpp:470-479): shade -> [depth of field] -> [bloom] -> gamma -> [FXAA]
    const bool bloom = (shade_options & 256) != 0, dof = (shade_options & 1024) != 0;
    const bool gamma_later = bloom || dof;
    k_shade_gamma<<<(npix + 255) / 256, 256, 0, st>>>(R->g_out.as<RmHitInfo>(), R->planes[0].as<RmRadiance>(), R->planes[1].as<RmRadiance>(),
                                                       R->planes[2].as<RmRadiance>(), R->planes[3].as<RmRadiance>(), npix, args->exposure, shade_options,
                                                       !gamma_later, R->rgb[0].as<float>());
    ctx->launches++;
    int out = 0;
    if (dof) {
        if ((rc = dof_device(ctx, R, args, R->rgb[0].as<float>(), R->rgb[1].as<float>()))) return rc;
        out = 1;
    }
    if (bloom) {
        if ((rc = R->glow[0].alloc(bytes)) || (rc = R->glow[1].alloc(bytes))) return rc;
        float *img = R->rgb[out].as<float>();
        k_bloom_bright<<<(npix + 255) / 256, 256, 0, st>>>(img, R->glow[0].as<float>(), npix);
        dim3 block(32, 8), grid((args->width + 31) / 32, (args->height + 7) / 8);
        int cur = 0;
        for (int step = 1; step <= 16; step *= 2, cur ^= 1)
            k_bloom_pass<<<grid, block, 0, st>>>(R->glow[cur].as<float>(), R->glow[cur ^ 1].as<float>(), img, args->width, args->height, step);
        ctx->launches += 6;
    }
    if (gamma_later) {
        k_gamma<<<(npix + 255) / 256, 256, 0, st>>>(R->rgb[out].as<float>(), npix);
        ctx->launches++;
    }
    if (shade_options & 512) {          // DoFXAA
        if ((rc = rm_fxaa_device(ctx, R->rgb[out].as<float>(), R->rgb[out ^ 1].as<float>(), args->width, args->height))) return rc;
        out ^= 1;
    }
    RM_CUDA(cudaGetLastError());
    RM_CUDA(cudaMemcpyAsync(rgb_out, R->rgb[out].p, bytes, cudaMemcpyDeviceToHost, st));
    RM_CUDA(cudaStreamSynchronize(st));
    return RM_OK;
}

} // extern "C"
