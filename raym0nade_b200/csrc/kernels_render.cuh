// kernels_render.cuh — the per-pixel estimator as a wavefront pipeline.
//
// renderPixel (src/render.cpp:448-551) is split into stages; every stage restates the
// reference's procedure (including its idiosyncrasies) per sample:
//   k_gbuffer        466-495   primary hit -> HitInfo, sky emission on a miss, red nudge
//   k_direct_gen     425-437 + sampleDirectLight (src/sampling.cpp:467-527): a pixel's direct NEE samples -> one contiguous
//                              block of the shadow queue; k_accum_direct sums the visible ones per pixel
//   k_regen          314-423   first vertex of an indirect path from the cached primary hit; tops the
//                              path queue up with fresh (pixel, sample) items whenever paths have ended
//   k_trace<PathJob> Model::rayHit over the path queue (closest hit)
//   k_surface        128-166   miss -> sky, getHitInfo, emissive hit -> the surface record of the vertex
//   k_bounce         143-312   one sampleRay level: Fresnel / Russian roulette / reflect or refract
//   k_nee            207-221   sampleDirectLight at a terminating vertex: 1..6 shadow items
//   k_trace<ShadowJob> + k_accum_shadow   Model::rayHit_test + accumulateInwardRadiance
//   k_commit_hold / k_finalise 510-549   firefly clamp (top-1 hold-back), exposure, variance finalise
// One stage = one small kernel: the shading code is branchy and, fused into one kernel, it was bound by
// instruction-cache misses (234 KB of SASS, 90 % of stall samples "no instruction", profiles/).
// Recursion is unrolled into a linear chain: a path carries the product of the per-level
// bsdfPdf*absorb (T), the first-vertex bsdfPdf (B0, kept apart because
// accumulateInwardRadiance splits on it, src/image.cpp:630-659) and the product of the
// 1/(1+fails) re-weights (W).
#pragma once
#include "dev_bsdf.cuh"
#include "kernels_trace.cuh"

namespace rm {

// shading-stage launch shape: kShadeCtasPerSm CTAs of kShadeBlock threads per SM (register budget = 64K / their product)
#ifndef RM_SHADE_BLOCK
#define RM_SHADE_BLOCK 256
#endif
#ifndef RM_SHADE_CTAS
#define RM_SHADE_CTAS 3
#endif
// CTA-wide lock step per batch / per phase of the shading stages: the warps of a CTA then run the same stretch of a large,
// branchy kernel together and share its instruction-cache lines.  It paid while one sampleRay level was a single kernel (round 1,
// k_bounce: 234 KB of SASS); with the level split into k_decide / k_continue / k_nee the barriers cost more than they save (17 - 20 %
// of the stall samples of k_surface / k_direct_gen sat on them; shade 103.1 -> 101.4 ms per 128 spp without, profiles/r02x_*), so it
// is off unless RM_LOCKSTEP_ON is defined.
#ifndef RM_LOCKSTEP_ON
#define RM_LOCKSTEP() ((void)0)
#else
#define RM_LOCKSTEP() __syncthreads()
#endif
constexpr int kShadeBlock = RM_SHADE_BLOCK;
constexpr int kShadeCtasPerSm = RM_SHADE_CTAS;
// per stage (register budget vs resident warps is a per-kernel trade: profiles/r01e_ab14_per_stage_ctas.txt)
#ifndef RM_CTAS_BOUNCE
#define RM_CTAS_BOUNCE RM_SHADE_CTAS
#endif
#ifndef RM_CTAS_SURFACE
#define RM_CTAS_SURFACE RM_SHADE_CTAS
#endif
#ifndef RM_CTAS_NEE
#define RM_CTAS_NEE RM_SHADE_CTAS
#endif
#ifndef RM_CTAS_REGEN
#define RM_CTAS_REGEN RM_SHADE_CTAS
#endif
#ifndef RM_CTAS_DIRECT
#define RM_CTAS_DIRECT RM_SHADE_CTAS
#endif
constexpr int kCtasBounce = RM_CTAS_BOUNCE, kCtasSurface = RM_CTAS_SURFACE, kCtasNee = RM_CTAS_NEE, kCtasRegen = RM_CTAS_REGEN, kCtasDirect = RM_CTAS_DIRECT;

constexpr int kMediumSlots = 16;    // nested-dielectric entries per path besides the implicit air entry: a path gains at most one per level,
                                    // maxRayDepth levels (the reference's multimap is unbounded, src/render.cpp:13-42)
constexpr int kMaxRayDepth = 16;    // maxRayDepth, src/render.cpp:125

// ------------------------------------------------------------------ path queue (SoA)
struct PathQueue {
    int cap;
    int *pixel;
    unsigned *sample, *drawn;
    float *o, *d;            // [3][cap]
    float *diff;             // [12][cap]
    float *T, *B0;           // [3][cap]
    float *W, *rough;
    int *flags;              // depth | exclude << 8 | n_medium << 16 | (k_decide:) mode << 24 | doDirect << 26 | nee_pass_absorb << 27
    int *med_id;             // [kMediumSlots][cap]   the medium stack lives here, in insertion order; the kernels walk it in place
    float *med;              // [4][kMediumSlots][cap]  ior, absorb rgb
    float *hit_t;            // INF = this entry is finished (miss, emissive hit) - the later stages skip it
    int *hit_face;           // k_trace: the face hit; from k_surface on: the vertex's DENSE index v (below)
    // Per-vertex records of one round.  Most of a queue's slots are dead by the time they are shaded (in the bench scene 60 % of
    // a round's rays leave the scene), so these are not indexed by queue slot but by a dense vertex index v that k_surface hands
    // out window by window (one atomic per 1024-slot window; the live slots of a window get consecutive v): written and read
    // as whole sectors instead of 40 %-full ones.
    float *surf;             // [22][cap]  the vertex's HitInfo (written by k_surface), field order of RmHitInfo; at [k * cap + v]
    float *hdP;              // [6][cap]   dPdx, dPdy at the hit (calc_dPdxy); at v
    float *dec;              // [7][cap]   decision record of k_decide: P_reflect | NEE scaling, P_RR, F, absorb rgb, relative eta; at v
    int *skey, *srank;       // sort key of the vertex and its position inside the key's bin
};

// one NEE / terminal sample waiting for its visibility test
struct __align__(16) ShadowItem {
    float o[3], aim;
    float d[3];
    int pixel;               // bit 31: direct-light sample (goes to Dd/Ds, no firefly hold-back)
    float b[3], weight;      // LightSample::bsdfPdf, weight (final)
    float l[3], vis;         // LightSample::light (final); vis = 1 when rayHit_test found the light unoccluded
};

// per-pixel accumulators
struct Accum {
    float *rad;              // [npix][16]  Dd{rgb,Var} Ds Id Is   (summed across GPUs)
    float *clum_sum;         // [npix][2]   sum of Clum, number of indirect LightSamples (summed across GPUs)
    float *clum_max;         // [npix]      luminance of the held-back sample (max-reduced across GPUs)
    float *hold_clum;        // [npix]      local held-back luminance (-1 = none)
    float *hold;             // [npix][8]   held-back sample: bsdfPdf rgb, light rgb, weight
    int *lock;               // [npix]
};

struct FrameBuffers {
    RmHitInfo *gbuffer;      // AoS, baseColor = nudged value the samplers use
    float *sav_base;         // [npix][3] un-nudged baseColor (restored at resolve, src/render.cpp:550)
    int *n_ind;              // [npix] spp_indirect of the pixel (0 when nothing is sampled)
    int *dir_base;           // [npix] first shadow-queue slot of the pixel's direct samples in the current direct wave (-1: none)
    // The pixels that are sampled at all - a primary hit on a non-emissive surface - in pixel order (rm_gbuffer); the per-pixel
    // stages (k_direct_gen, k_accum_direct, the first path vertices of k_regen) run over this list, so that a frame that is
    // 40 % background does not leave 40 % of their lanes idle.  nullptr: every pixel (the per-pixel checks stay in place).
    const int *active_list;
    int n_active;
};

// pow_s (src/geometry.cpp:22-28): there `pow` resolves to the double version.  Out of line: one copy of the (large)
// double-precision pow per kernel instead of three - the shading stages are instruction-fetch sensitive.
RM_NI float pow_s(float a, float k) { return (float)pow((double)a, (double)k); }

// getAbsorb (src/render.cpp:83-87)
RM_DI V3 get_absorb(V3 absorb, float dis) {
    float C = lum(absorb);
    if (!(C < fsub(1.0f, kEps))) return splat3(1.0f);
    float k = fmul(32.0f, dis);
    V3 a = absorb;
    if (a.x < 0.0f) a.x = 0.0f;
    if (a.y < 0.0f) a.y = 0.0f;
    if (a.z < 0.0f) a.z = 0.0f;
    return mk3(pow_s(a.x, k), pow_s(a.y, k), pow_s(a.z, k));
}

// device-side pipeline state (ints): queue lengths, cursors
enum { C_Q0 = 0, C_Q1 = 1, C_SQ = 2, C_OVERFLOW = 3, C_GLASS = 4, C_GLASS_LIST = 5, C_CUR_PATH = 6, C_CUR_SHADOW = 7,
       C_NEE = 8, C_PLAN_TAKE = 9, C_ITEM_LO = 10, C_ITEM_HI = 11, C_PLAN_LO = 12, C_PLAN_HI = 13, C_SQ_RUN = 14, C_VERT = 15, C_COUNT = 16 };

// the sort of a round's live vertices by mode, see k_decide
constexpr int kKeyReflect = 0;                        // 0: reflection off an opaque surface, 1: off a dielectric
constexpr int kKeyRefract = 2, kKeyNee = 3;           // 3..8: NEE with 1..6 light samples
constexpr int kSortBins = 16;
// pipeline state beyond C_COUNT: per-bin counts of this round, then the exclusive offsets (kSortBins + 1)
enum { C_BINS = C_COUNT, C_OFFS = C_COUNT + kSortBins, C_TOTAL = C_COUNT + 2 * kSortBins + 16 };

// warp-aggregated slot allocation
RM_DI int alloc_slot(int *counter, bool want) {
    unsigned mask = __ballot_sync(0xffffffffu, want);   // every lane of the warp calls this (loops are padded to whole warps)
    if (!want) return -1;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1));
}

// CTA-local compaction.  A path queue mixes live entries with finished ones (a miss, an emissive hit); run one
// thread per queue slot and most lanes idle through the shading code.  Instead a CTA scans a window of
// kWindow consecutive slots, packs the indices of the live ones into shared memory (ballot + popc per warp, one
// shared atomic per warp) and then works through that dense list 256 at a time, so its warps run full.  The
// gathers stay inside the window's few KB per SoA array.  Returns the number of live entries; s_idx / s_n are
// shared.  Every thread of the CTA calls this.
constexpr int kWindow = 1024;
template <class Pred>
RM_DI int cta_compact(int base, int n, int *s_idx, int *s_n, Pred live_at) {
    const int lane = threadIdx.x & 31;
    __syncthreads();                           // the previous window's list is no longer read
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < kWindow; k += blockDim.x) {
        const int i = base + k;
        const bool live = i < n && live_at(i);
        const unsigned m = __ballot_sync(0xffffffffu, live);
        int wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(s_n, __popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (live) s_idx[wbase + __popc(m & ((1u << lane) - 1u))] = i;
    }
    __syncthreads();
    return *s_n;
}

// ------------------------------------------------------------------ accumulation
// LOCAL = false: r4 points at the pixel's accumulators in global memory, shared by many threads (atomicAdd);
// LOCAL = true:  r4 is a running sum the calling thread owns (k_accum_direct)
template <bool LOCAL>
RM_DI void accum_basic(float *r4, V3 inrad, float weight) {          // accumulateInwardRadiance_basic, src/image.cpp:615-628
    if (!isfinite_any(inrad)) return;
    if (!isfinite(weight)) return;
    const float v0 = fmul(inrad.x, weight), v1 = fmul(inrad.y, weight), v2 = fmul(inrad.z, weight), v3 = fmul(dot(inrad, inrad), weight);
    if (LOCAL) { r4[0] = fadd(r4[0], v0); r4[1] = fadd(r4[1], v1); r4[2] = fadd(r4[2], v2); r4[3] = fadd(r4[3], v3); }
    else { atomicAdd(r4 + 0, v0); atomicAdd(r4 + 1, v1); atomicAdd(r4 + 2, v2); atomicAdd(r4 + 3, v3); }
}

// accumulateInwardRadiance (src/image.cpp:630-659): split into demodulated diffuse + specular
template <bool LOCAL>
RM_DI void accum_split_t(float *rd, float *rs, V3 baseColor, V3 b, V3 l, float w) {
    if (length(l) < kEps) return;
    V3 base0 = normalize(baseColor);
    if (length(baseColor) < kEps) { accum_basic<LOCAL>(rs, l * b, w); return; }
    const V3 White = normalize(splat3(1.0f));
    float XdotY = dot(base0, White);
    if (XdotY > 0.99f) { accum_basic<LOCAL>(rd, div_true(l * b, baseColor), w); return; }
    V3 perp = normalize(cross(base0, White));
    V3 bp = b - perp * dot(perp, b);
    float d1 = dot(bp, White), d2 = dot(bp, base0);
    float AplusB = fdiv(fadd(d1, d2), fadd(1.0f, XdotY));
    float AminusB = fdiv(fsub(d1, d2), fsub(1.0f, XdotY));
    float Bc = fdiv(fsub(AplusB, AminusB), 2.0f);
    V3 base_part = Bc * base0;
    accum_basic<LOCAL>(rd, div_recip(l * Bc, length(baseColor)), w);
    accum_basic<LOCAL>(rs, l * (b - base_part), w);
}
RM_NI void accum_split(float *rd, float *rs, V3 baseColor, V3 b, V3 l, float w) { accum_split_t<false>(rd, rs, baseColor, b, l, w); }

// One finished indirect LightSample of pixel p.  The reference drops a sample when it alone
// exceeds 16/17 of the pixel's total luminance (src/render.cpp:534-547); at most one sample per
// pixel can qualify, so the running maximum is held back un-accumulated (exact, single pass).
RM_DI void add_indirect(const Accum &A, const FrameBuffers &Fb, int p, V3 b, V3 l, float w) {
    const RmHitInfo *g = Fb.gbuffer + p;
    const float *gf = reinterpret_cast<const float *>(g);
    V3 base = gf[18] < kEps ? splat3(0.0f) : mk3(gf[9], gf[10], gf[11]);
    float clum = fmul(lum(b * l), w);
    atomicAdd(A.clum_sum + 2 * p, clum);
    atomicAdd(A.clum_sum + 2 * p + 1, 1.0f);
    volatile float *hc = A.hold_clum + p;
    if (!(clum > *hc)) { accum_split(A.rad + 16 * p + 8, A.rad + 16 * p + 12, base, b, l, w); return; }
    bool done = false, spill = false;
    V3 ob = b, ol = l;
    float ow = w;
    while (!done) {
        if (atomicCAS(A.lock + p, 0, 1) == 0) {
            __threadfence();
            float cur = *hc;
            volatile float *h = A.hold + 8 * p;
            if (clum > cur) {
                if (cur >= 0.0f) { ob = mk3(h[0], h[1], h[2]); ol = mk3(h[3], h[4], h[5]); ow = h[6]; spill = true; }
                h[0] = b.x; h[1] = b.y; h[2] = b.z; h[3] = l.x; h[4] = l.y; h[5] = l.z; h[6] = w;
                *hc = clum;
            } else spill = true;
            __threadfence();
            atomicExch(A.lock + p, 0);
            done = true;
        }
    }
    if (spill) accum_split(A.rad + 16 * p + 8, A.rad + 16 * p + 12, base, ob, ol, ow);
}

RM_DI void add_direct(const Accum &A, const FrameBuffers &Fb, int p, V3 b, V3 l, float w) {
    const float *gf = reinterpret_cast<const float *>(Fb.gbuffer + p);
    accum_split(A.rad + 16 * p, A.rad + 16 * p + 4, mk3(gf[9], gf[10], gf[11]), b, l, w);
}

// ------------------------------------------------------------------ K2: G-buffer
__global__ void __launch_bounds__(128) k_gbuffer(DevScene S, DevArgs A, const int *__restrict__ tri_idx, const float *__restrict__ t_in,
                                                 FrameBuffers Fb, int spp_direct, int spp_indirect_base, int *glass_count) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.width * A.height) return;
    int x = p % A.width, y = p / A.width;
    V3 d = primary_d(A, x, y);
    V3 Dir = normalize(d);
    Surface s = default_surface();
    int face = tri_idx[p];
    int n_ind = 0;
    V3 sav = splat3(0.0f);
    if (face < 0) {
        s.position = splat3(CUDART_NAN_F);
        if (S.sky_width) s.emission = sky_get(S, Dir);
    } else {
        float t = t_in[p];
        s.position = A.position + t * Dir;
        RayDiff bd = init_ray_diff(d, A);
        V3 dPdx, dPdy;
        get_hit_info(S, face, t, Dir, bd, dPdx, dPdy, s);
        bool opaque = s.opacity > fsub(1.0f, kEps);
        sav = s.baseColor;
        // near-black / near-grey base colours get +0.04 red so that baseColor and white stay
        // separable (src/render.cpp:492-495; the literals there are doubles)
        float C0 = lum(s.baseColor);
        bool nudge = (double)C0 < 2e-2;
        if (!nudge && C0 < 0.8f) nudge = (double)length(div_recip(s.baseColor, C0) - splat3(1.0f)) < 2e-2;
        if (nudge) s.baseColor.x = (float)((double)s.baseColor.x + 4e-2);
        n_ind = spp_indirect_base * (opaque ? 1 : 16);
        if (!opaque && spp_indirect_base > 0) atomicAdd(glass_count, 1);
    }
    store_hitinfo(Fb.gbuffer + p, s);
    Fb.sav_base[3 * p] = sav.x; Fb.sav_base[3 * p + 1] = sav.y; Fb.sav_base[3 * p + 2] = sav.z;
    Fb.n_ind[p] = n_ind;
}

// ------------------------------------------------------------------ NEE (sampleDirectLight for ONE sample)
// Emits at most one ShadowItem.  scale multiplies bsdfPdf (the `scaling` step of sampleRay);
// chain: light is multiplied by the path's carried throughput when `indirect`.
struct NeeOut { bool valid; V3 dir; float aim; V3 bsdf, light; float weight; };

RM_DI NeeOut nee_sample(const DevScene &S, const Bsdf &B, Rng &gen, const float *lw, float total, int sampleCnt) {
    NeeOut o;
    o.valid = false;
    const V3 pos = B.s.position;
    if (S.sky_width == 0) {
        int li = pick_light(lw, S.n_lights, fmul(total, gen()));
        if (li < 0) return o;
        float P_light = fdiv(lw[li], total);
        const DevLight &L = S.lights[li];
        V3 lightPos;
        int fails = 0;
        sample_light_face(S, L, pos, gen, lightPos, fails);
        if (!isfinite_any(lightPos)) return o;
        o.dir = normalize(lightPos - pos);
        float distance = length(lightPos - pos);
        o.aim = fsub(distance, kEps);
        o.bsdf = get_bsdf(B, o.dir);
        clamp_lum(o.bsdf);
        V3 color = mk3(L.color[0], L.color[1], L.color[2]);
        V3 li3 = fmul(fmul(2.0f, kPi), L.power) * color;
        o.light = div_recip(div_recip(li3, fadd(fmul(distance, distance), 1e-3f)), P_light);
        o.weight = fdiv(1.0f, float(sampleCnt * (fails + 1)));
        o.valid = true;
    } else {
        V3 light;
        sample_sky(S, B.s.surfaceNormal, gen, o.dir, light);
        if (!isfinite_any(o.dir)) return o;
        o.aim = CUDART_INF_F;
        o.bsdf = get_bsdf(B, o.dir);
        clamp_lum(o.bsdf);
        o.light = light;
        o.weight = fdiv(1.0f, float(sampleCnt));
        o.valid = true;
    }
    return o;
}

RM_DI void write_shadow(ShadowItem *dst_item, int pixel_tag, V3 org, const NeeOut &n, V3 b, V3 l, float w) {
    ShadowItem it;
    it.o[0] = org.x; it.o[1] = org.y; it.o[2] = org.z; it.aim = n.aim;
    it.d[0] = n.dir.x; it.d[1] = n.dir.y; it.d[2] = n.dir.z; it.pixel = pixel_tag;
    it.b[0] = b.x; it.b[1] = b.y; it.b[2] = b.z; it.weight = w;
    it.l[0] = l.x; it.l[1] = l.y; it.l[2] = l.z; it.vis = 0.0f;
    float4 *dst = reinterpret_cast<float4 *>(dst_item);
    const float4 *src = reinterpret_cast<const float4 *>(&it);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
}

RM_DI void push_shadow(ShadowItem *q, int *count, int cap, int *overflow, bool want, int pixel_tag, V3 org, const NeeOut &n, V3 b, V3 l, float w) {
    int slot = alloc_slot(count, want);
    if (!want) return;
    if (slot >= cap) { atomicExch(overflow, 1); return; }
    write_shadow(q + slot, pixel_tag, org, n, b, l, w);
}

// ------------------------------------------------------------------ direct light at the primary hit
// One thread per pixel draws the wave's S direct samples (sample index s = s_begin + k*s_stride, k < S): the
// G-buffer record and the per-light-object weights (getLightObjectWeight, one BSDF evaluation per light object,
// src/sampling.cpp:406-417) depend on the pixel only and are formed once instead of once per sample.  Every
// sample keeps its own random stream, so the samples are the ones a per-sample loop draws.
// WARP = true (waves of >= 16 samples per pixel): one WARP per pixel, one lane per sample - the surface record, the
// material branches of the BSDF and the light weights are uniform across the warp and the lanes' 64-byte items are
// consecutive (coalesced).  WARP = false (few samples per pixel, e.g. previews): one THREAD per pixel loops over them.
template <bool WARP>
__global__ void __launch_bounds__(kShadeBlock, kCtasDirect) k_direct_gen(DevScene S, DevArgs A, FrameBuffers Fb, int n_samples, int npix, int s_begin,
                                                    int s_stride, int spp_direct, unsigned long long seed,
                                                    ShadowItem *sq, int *s_count, int s_cap, int *overflow) {
    const int lane = threadIdx.x & 31;
    const int per_cta = WARP ? (blockDim.x >> 5) : blockDim.x;          // pixels a CTA takes per batch
    const int n_loop = Fb.active_list ? Fb.n_active : npix;
    for (int base = blockIdx.x * per_cta; base < n_loop; base += gridDim.x * per_cta) {
        RM_LOCKSTEP();                       // CTA-wide lock step per batch: shared instruction-cache lines (see k_bounce)
        const int at = base + (WARP ? (threadIdx.x >> 5) : threadIdx.x);
        const int p = at < n_loop ? (Fb.active_list ? Fb.active_list[at] : at) : npix;
        bool go = false;
        Bsdf B;
        B.s = default_surface();
        B.inDir = splat3(0.0f);
        float lw[kMaxLights];
        float total = 0.0f;
        if (p < npix) {
            B.s = load_hitinfo(Fb.gbuffer + p);
            if (isfinite_any(B.s.position) && !(length(B.s.emission) > 0.0f)) {
                B.inDir = -normalize(B.s.position - A.position);
                go = true;
                if (S.sky_width == 0) { total = light_weights(S, B, lw); if (total == 0.0f) go = false; }
            }
        }
        // a pixel's n_samples items are one contiguous block of the shadow queue: k_accum_direct sums them per pixel in
        // sample order without atomics, and a warp of the visibility pass gets rays that share their origin
        int first = 0;
        if (WARP) {                          // go is warp-uniform
            if (go) {
                if (lane == 0) first = atomicAdd(s_count, n_samples);
                first = __shfl_sync(0xffffffffu, first, 0);
            }
        } else {
            const unsigned m = __ballot_sync(0xffffffffu, go);
            if (m) {
                if (lane == __ffs(m) - 1) first = atomicAdd(s_count, __popc(m) * n_samples);
                first = __shfl_sync(0xffffffffu, first, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u)) * n_samples;
            }
        }
        if (go && first + n_samples > s_cap) { atomicExch(overflow, 1); go = false; }
        if (p < npix && (!WARP || lane == 0)) Fb.dir_base[p] = go ? first : -1;
        if (!go) continue;
#pragma unroll 1
        for (int k = WARP ? lane : 0; k < n_samples; k += WARP ? 32 : 1) {
            const int s = s_begin + k * s_stride;
            NeeOut n;
            n.valid = false;
            if (s < spp_direct) {
                Rng gen;
                gen.init(seed, (unsigned)p, (unsigned)s, kStreamDirect);
                n = nee_sample(S, B, gen, lw, total, spp_direct);
            }
            if (!n.valid) { n.dir = splat3(0.0f); n.aim = CUDART_NAN_F; n.bsdf = n.light = splat3(0.0f); n.weight = 0.0f; }      // null item
            write_shadow(sq + first + k, p | 0x80000000, B.s.position, n, n.bsdf, n.light, n.weight);
        }
    }
}

// accumulateInwardRadiance over a direct wave: one thread per pixel adds its visible samples in sample order into
// running sums it owns and folds them into the pixel's Dd / Ds accumulators once (the same pixel is touched by no
// other thread while this kernel runs) - 8 read-modify-writes per pixel and wave instead of 8 atomics per sample.
__global__ void __launch_bounds__(256) k_accum_direct(FrameBuffers Fb, Accum Ac, const ShadowItem *__restrict__ sq, int n_samples, int npix) {
    const int at = blockIdx.x * blockDim.x + threadIdx.x;
    if (at >= (Fb.active_list ? Fb.n_active : npix)) return;
    const int p = Fb.active_list ? Fb.active_list[at] : at;
    const int first = Fb.dir_base[p];
    if (first < 0) return;
    const float *gf = reinterpret_cast<const float *>(Fb.gbuffer + p);
    const V3 base = mk3(gf[9], gf[10], gf[11]);
    float rd[4] = {0.0f, 0.0f, 0.0f, 0.0f}, rs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
    for (int k = 0; k < n_samples; k++) {
        const float4 *src = reinterpret_cast<const float4 *>(sq + first + k);
        const float4 d = src[3];
        if (d.w == 0.0f) continue;
        const float4 c = src[2];
        accum_split_t<true>(rd, rs, base, mk3(c.x, c.y, c.z), mk3(d.x, d.y, d.z), c.w);
    }
    float *r = Ac.rad + 16 * size_t(p);
#pragma unroll
    for (int j = 0; j < 4; j++) { r[j] = fadd(r[j], rd[j]); r[4 + j] = fadd(r[4 + j], rs[j]); }
}

// ------------------------------------------------------------------ path state I/O
RM_DI void store_path(const PathQueue &Q, int i, int pixel, unsigned sample, unsigned drawn, V3 o, V3 d, const RayDiff &df,
                      V3 T, V3 B0, float W, float rough, int depth, bool exclude, int n_medium) {
    const int c = Q.cap;
    Q.pixel[i] = pixel; Q.sample[i] = sample; Q.drawn[i] = drawn;
    Q.o[i] = o.x; Q.o[c + i] = o.y; Q.o[2 * c + i] = o.z;
    Q.d[i] = d.x; Q.d[c + i] = d.y; Q.d[2 * c + i] = d.z;
    const V3 *dv = &df.dPdx;
#pragma unroll
    for (int k = 0; k < 4; k++) { Q.diff[(3 * k) * c + i] = dv[k].x; Q.diff[(3 * k + 1) * c + i] = dv[k].y; Q.diff[(3 * k + 2) * c + i] = dv[k].z; }
    Q.T[i] = T.x; Q.T[c + i] = T.y; Q.T[2 * c + i] = T.z;
    Q.B0[i] = B0.x; Q.B0[c + i] = B0.y; Q.B0[2 * c + i] = B0.z;
    Q.W[i] = W; Q.rough[i] = rough;
    Q.flags[i] = depth | (exclude ? 256 : 0) | (n_medium << 16);       // the caller writes the n_medium entries of the medium stack
}

// ------------------------------------------------------------------ work items of the indirect sample loop
// item j in [0, items_a): sample k = j / npix of pixel j % npix                      (every pixel, k < n_a; with
//                        FrameBuffers::active_list: sample j / n_active of pixel active_list[j % n_active])
// item j in [items_a, total): j' = j - items_a, sample n_a + j' / n_glass of glass pixel list[j' % n_glass]
// (glass pixels take 16x the indirect samples, src/render.cpp:498-501); local sample k is the global
// sample index s_begin + k * s_stride (interleaved over GPUs).
struct ItemSpace {
    long long items_a, total;
    int npix, n_glass, n_a;
    const int *glass_list;
    int s_begin, s_stride;
};

// One thread: how many fresh items fit into the path queue this round (the queue is topped up after
// k_bounce compacted the surviving paths into it), and where they start.
__global__ void k_plan(int *C, int q_slot, int cap, long long total) {
    long long cur = (long long)(unsigned)C[C_ITEM_LO] | ((long long)C[C_ITEM_HI] << 32);
    int free_slots = cap - min(C[q_slot], cap);
    long long left = total - cur;
    int take = (int)(left < (long long)free_slots ? left : (long long)free_slots);
    if (take < 0) take = 0;
    C[C_PLAN_TAKE] = take;
    C[C_PLAN_LO] = (int)(unsigned)(cur & 0xffffffffLL);
    C[C_PLAN_HI] = (int)(cur >> 32);
    cur += take;
    C[C_ITEM_LO] = (int)(unsigned)(cur & 0xffffffffLL);
    C[C_ITEM_HI] = (int)(cur >> 32);
    C[C_CUR_PATH] = 0;
    C[C_VERT] = 0;                                           // dense vertex indices of this round (k_surface)
    if (C[C_SQ_RUN]) { C[C_SQ] = 0; C[C_SQ_RUN] = 0; }       // the shadow queue was traced last round: start it afresh
}

// Shadow items pile up over rounds (their results only feed the accumulators) and are traced once enough of them
// wait for a full-width launch - or when the render ends (flush).
__global__ void k_shadow_gate(int *C, int threshold, int cap, int flush) {
    const int n = min(C[C_SQ], cap);
    C[C_SQ_RUN] = (n >= threshold || (flush && n > 0)) ? n : 0;
    C[C_CUR_SHADOW] = 0;
}

// ------------------------------------------------------------------ first vertex of an indirect path
// sampleIndirectLightFromFirstIntersection (src/render.cpp:314-423) up to the new ray.
__global__ void __launch_bounds__(kShadeBlock, kCtasRegen) k_regen(DevScene S, DevArgs A, FrameBuffers Fb, ItemSpace I, const int *__restrict__ C, unsigned long long seed,
                                               PathQueue Q, int *q_count) {
    const int n_items = C[C_PLAN_TAKE];
    const long long first = (long long)(unsigned)C[C_PLAN_LO] | ((long long)C[C_PLAN_HI] << 32);
    for (int base = blockIdx.x * blockDim.x; base < n_items; base += gridDim.x * blockDim.x) {
        RM_LOCKSTEP();                       // CTA-wide lock step per batch (see k_bounce)
        const int i = base + threadIdx.x;
        bool want = false;
        int p = 0;
        unsigned s = 0;
        Rng gen;
        V3 newDir = splat3(0.0f), bsdfPdf = splat3(CUDART_NAN_F), pos = splat3(0.0f);
        RayDiff next;
        float W = 1.0f, rough = 0.0f;
        bool entered = false;                 // the first vertex refracted into a dielectric: the path starts with one medium entry
        int med_id = 0;
        float med_ior = 1.0f;
        V3 med_ab = splat3(1.0f);
        if (i < n_items) {
            const long long j = first + i;
            int k;
            if (j < I.items_a) {
                const int per = Fb.active_list ? Fb.n_active : I.npix;
                k = int(j / per);
                p = int(j % per);
                if (Fb.active_list) p = Fb.active_list[p];
            }
            else { const long long jj = j - I.items_a; k = I.n_a + int(jj / I.n_glass); p = I.glass_list[int(jj % I.n_glass)]; }
            s = unsigned(I.s_begin + k * I.s_stride);
            int n_ind = Fb.n_ind[p];
            Surface g = load_hitinfo(Fb.gbuffer + p);
            if ((int)s < n_ind && !(length(g.emission) > 0.0f)) {
                gen.init(seed, (unsigned)p, s, kStreamIndirect);
                int x = p % A.width, y = p / A.width;
                RayDiff bd = init_ray_diff(primary_d(A, x, y), A);
                V3 inDir = normalize(g.position - A.position);
                float hit_t = length(g.position - A.position);
                Bsdf B;
                B.inDir = -inDir;
                B.s = g;
                rough = fmul(g.roughness, 1.0f);
                float ior = B.s.eta;
                B.s.eta = fdiv(1.0f, B.s.eta);
                float P_reflect = 1.0f, F = 0.0f;
                V3 refr;
                if (B.s.opacity < kEps) {
                    precise_refraction(B, refr, F);
                    P_reflect = fadd(0.24f, fmul(fsub(1.0f, 0.24f), F));
                }
                int fails = 0;
                if (gen() < P_reflect) {
                    V3 dPdx, dPdy, dDdx, dDdy;
                    calc_dPdxy(inDir, hit_t, B.s.shapeNormal, bd, dPdx, dPdy);
                    calc_dDdxy(inDir, B.s.surfaceNormal, bd, dDdx, dDdy);
                    next.dPdx = dPdx; next.dPdy = dPdy; next.dDdx = dDdx; next.dDdy = dDdy;
                    sample_reflection(B, gen, newDir, bsdfPdf, fails);
                    bsdfPdf = div_true(bsdfPdf, P_reflect);
                } else {
                    sample_btdf(B, gen, newDir, bsdfPdf, fails);
                    bsdfPdf = bsdfPdf * fsub(1.0f, F);
                    bsdfPdf = div_true(bsdfPdf, fsub(1.0f, P_reflect));
                    next = bd;
                    // (leaving a dielectric here erases from an empty stack: nothing to do)
                    if (B.s.entering) { entered = true; med_id = B.s.id; med_ior = ior; med_ab = B.s.baseColor; }
                }
                if (isfinite_any(newDir)) {
                    want = true;
                    pos = B.s.position;
                    if (fails > 0) W = fdiv(1.0f, float(1 + fails));
                }
            }
        }
        int slot = alloc_slot(q_count, want);
        if (want && slot < Q.cap) {
            store_path(Q, slot, p, s, gen.drawn, pos, newDir, next, splat3(1.0f), bsdfPdf, W, rough, 1, true, entered ? 1 : 0);
            if (entered) {
                const int c = Q.cap;
                Q.med_id[slot] = med_id;
                Q.med[slot] = med_ior; Q.med[c + slot] = med_ab.x; Q.med[2 * c + slot] = med_ab.y; Q.med[3 * c + slot] = med_ab.z;
            }
        }
    }
}

// ------------------------------------------------------------------ closest hit over the path queue
struct PathJob {
    static constexpr bool kOcclusion = false;
    PathQueue Q;
    RM_DI bool load(int i, V3 &o, V3 &d, float &aim) const {
        const int c = Q.cap;
        o = mk3(Q.o[i], Q.o[c + i], Q.o[2 * c + i]);
        d = mk3(Q.d[i], Q.d[c + i], Q.d[2 * c + i]);
        aim = CUDART_INF_F;
        return true;
    }
    RM_DI void hit(int i, float t, int face) const { Q.hit_t[i] = t; Q.hit_face[i] = face; }
    RM_DI void visibility(int, bool) const {}
};

// ------------------------------------------------------------------ the surface record of a path vertex
RM_DI void store_surface(const PathQueue &Q, int i, const Surface &s, V3 dPdx, V3 dPdy) {
    const int c = Q.cap;
    float *f = Q.surf + i;
    f[0 * c] = s.shapeNormal.x; f[1 * c] = s.shapeNormal.y; f[2 * c] = s.shapeNormal.z;
    f[3 * c] = s.surfaceNormal.x; f[4 * c] = s.surfaceNormal.y; f[5 * c] = s.surfaceNormal.z;
    f[6 * c] = s.emission.x; f[7 * c] = s.emission.y; f[8 * c] = s.emission.z;
    f[9 * c] = s.baseColor.x; f[10 * c] = s.baseColor.y; f[11 * c] = s.baseColor.z;
    f[12 * c] = s.position.x; f[13 * c] = s.position.y; f[14 * c] = s.position.z;
    f[15 * c] = s.specular; f[16 * c] = s.roughness; f[17 * c] = s.metallic; f[18 * c] = s.opacity; f[19 * c] = s.eta;
    f[20 * c] = __int_as_float(s.id); f[21 * c] = __int_as_float(s.entering ? 1 : 0);
    float *h = Q.hdP + i;
    h[0 * c] = dPdx.x; h[1 * c] = dPdx.y; h[2 * c] = dPdx.z; h[3 * c] = dPdy.x; h[4 * c] = dPdy.y; h[5 * c] = dPdy.z;
}

RM_DI Surface load_surface(const PathQueue &Q, int i) {
    const int c = Q.cap;
    const float *f = Q.surf + i;
    Surface s;
    s.shapeNormal = mk3(f[0 * c], f[1 * c], f[2 * c]);
    s.surfaceNormal = mk3(f[3 * c], f[4 * c], f[5 * c]);
    s.emission = mk3(f[6 * c], f[7 * c], f[8 * c]);
    s.baseColor = mk3(f[9 * c], f[10 * c], f[11 * c]);
    s.position = mk3(f[12 * c], f[13 * c], f[14 * c]);
    s.specular = f[15 * c]; s.roughness = f[16 * c]; s.metallic = f[17 * c]; s.opacity = f[18 * c]; s.eta = f[19 * c];
    s.id = __float_as_int(f[20 * c]);
    s.entering = __float_as_int(f[21 * c]) != 0;
    return s;
}

// Medium::absorb() (src/render.cpp:33-38) over the path's medium stack
RM_DI V3 medium_absorb(const PathQueue &Q, int i, int n) {
    const int c = Q.cap;
    V3 a = splat3(1.0f);
    for (int k = 0; k < n; k++) a = a * mk3(Q.med[(4 * k + 1) * c + i], Q.med[(4 * k + 2) * c + i], Q.med[(4 * k + 3) * c + i]);
    return a;
}

// sampleRay up to the surface (src/render.cpp:128-166): a miss returns the sky (unless direct light is
// excluded), a hit builds the HitInfo, an emissive hit returns its emission.  Finished entries get
// hit_t = INF.  Thread 0 also resets the counters the later stages of this round append to.
__global__ void __launch_bounds__(kShadeBlock, kCtasSurface) k_surface(DevScene S, FrameBuffers Fb, Accum Ac, PathQueue Q, int *C, int q_slot) {
    const int n = min(C[q_slot], Q.cap);
    if (blockIdx.x == 0) {                   // the counters the later stages of this round append to
        if (threadIdx.x == 0) { C[q_slot ^ 1] = 0; C[C_NEE] = 0; }
        for (int k = threadIdx.x; k < kSortBins; k += blockDim.x) C[C_BINS + k] = 0;
    }
    const int c = Q.cap;
    __shared__ int s_idx[kWindow];
    __shared__ int s_n, s_vbase;
    const bool sky = S.sky_width != 0;
    for (int base = blockIdx.x * kWindow; base < n; base += gridDim.x * kWindow) {
      // live here: a hit, or a miss that returns the sky (src/render.cpp:129-133)
      const int n_live = cta_compact(base, n, s_idx, &s_n, [&](int i) { return Q.hit_t[i] != CUDART_INF_F || (sky && !(Q.flags[i] & 256)); });
      if (threadIdx.x == 0) s_vbase = n_live ? atomicAdd(C + C_VERT, n_live) : 0;       // this window's block of dense vertex indices
      __syncthreads();
      const int vbase = s_vbase;
      for (int j0 = 0; j0 < n_live; j0 += blockDim.x) {
        RM_LOCKSTEP();                       // CTA-wide lock step per batch (see k_bounce)
        const int j = j0 + threadIdx.x;
        if (j >= n_live) continue;
        const int i = s_idx[j];
        const int p = Q.pixel[i];
        const int fl = Q.flags[i];
        const bool exclude = (fl & 256) != 0;
        const V3 dir = mk3(Q.d[i], Q.d[c + i], Q.d[2 * c + i]);
        const float t = Q.hit_t[i];
        bool add = false;                       // this vertex ends the path with a light sample (1, light, 1)
        V3 light = splat3(0.0f);
        if (t == CUDART_INF_F) {
            // miss (src/render.cpp:129-133)
            if (!(exclude || S.sky_width == 0)) { light = sky_get(S, dir) * splat3(1.0f); add = true; }
        } else {
            const V3 org = mk3(Q.o[i], Q.o[c + i], Q.o[2 * c + i]);
            RayDiff bd;
            bd.dPdx = mk3(Q.diff[0 * c + i], Q.diff[1 * c + i], Q.diff[2 * c + i]);
            bd.dPdy = mk3(Q.diff[3 * c + i], Q.diff[4 * c + i], Q.diff[5 * c + i]);
            bd.dDdx = mk3(Q.diff[6 * c + i], Q.diff[7 * c + i], Q.diff[8 * c + i]);
            bd.dDdy = mk3(Q.diff[9 * c + i], Q.diff[10 * c + i], Q.diff[11 * c + i]);
            Surface sf = default_surface();
            sf.position = org + dir * t;
            V3 dPdx, dPdy;
            get_hit_info(S, Q.hit_face[i], t, dir, bd, dPdx, dPdy, sf);
            if (length(sf.emission) > kEps) {
                // emissive surface (src/render.cpp:162-166)
                if (!exclude) {
                    light = sf.emission * get_absorb(medium_absorb(Q, i, (fl >> 16) & 31), t);
                    add = true;
                }
                Q.hit_t[i] = CUDART_INF_F;
            } else {
                store_surface(Q, vbase + j, sf, dPdx, dPdy);
                Q.hit_face[i] = vbase + j;
            }
        }
        if (add) {
            const V3 T = mk3(Q.T[i], Q.T[c + i], Q.T[2 * c + i]), B0 = mk3(Q.B0[i], Q.B0[c + i], Q.B0[2 * c + i]);
            const float inv_spp = fdiv(1.0f, float(Fb.n_ind[p]));
            add_indirect(Ac, Fb, p, B0, light * T, fmul(fmul(1.0f, Q.W[i]), inv_spp));
        }
      }
    }
}

// ------------------------------------------------------------------ one sampleRay level
__constant__ int c_sampleCount[kMaxRayDepth + 1] = {0, 1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 6};

// sampleRay from the surface on (src/render.cpp:143-312) is three kinds of work that share nothing but the vertex:
// a reflected continuation (sample_reflection), a refracted one (sample_btdf) or a termination in next-event estimation
// (sampleDirectLight, 1..6 light samples).  Which one a vertex takes is decided by its Fresnel term and two draws.  Run
// together in one kernel the three leave most lanes of a warp idle (ncu, round 1: 12.6 of 32 lanes per instruction), so
// the level is split the way the reference branches (src/render.cpp:172-284):
//   k_decide     roughness regularisation, medium scan + calcEta, Fresnel, Russian roulette -> the vertex's MODE, written
//                with the few numbers the next stage needs (the decision record), and a sort key
//   k_sort_*     counting sort of the live vertices by key = mode (reflection split by opaque / dielectric surface, NEE by
//                its number of light samples)
//   k_continue   <reflect> and <refract>: dense over their segment of the sorted queue -> the next path queue
//   k_nee        dense over the NEE segment -> shadow items
// Every warp of the dense kernels then runs one mode (full warps up to the rejection-loop tails).  The sorted queue holds
// path-queue slots, so the dense kernels gather: a bin is therefore filled in blocks - the vertices of one 1024-slot window
// of the path queue that share a key sit next to each other in it - and a warp's gathers stay inside a few KB per SoA
// array.  (Keys that also carried the material id, 256 bins, cut the blocks to a handful of vertices: ncu showed 14.6 GB
// of DRAM reads per k_continue launch for 3.6 GB of records, profiles/r02d_*.)
enum { kBounceDead = 0, kBounceNee = 1, kBounceReflect = 2, kBounceRefract = 3 };
// PathQueue::flags: depth | exclude << 8 | n_medium << 16 | mode << 24 | doDirect << 26 | nee_pass_absorb << 27
// decision record PathQueue::dec [7][cap]: P_reflect (NEE: the `scaling` factor), P_RR, F, absorb rgb, relative eta

#ifndef RM_CTAS_DECIDE
#define RM_CTAS_DECIDE 4
#endif
__global__ void __launch_bounds__(kShadeBlock, RM_CTAS_DECIDE) k_decide(unsigned long long seed, PathQueue Q, const int *__restrict__ in_count, int *bins) {
    const int n = min(*in_count, Q.cap);
    const int c = Q.cap;
    __shared__ int s_idx[kWindow];
    __shared__ int s_n;
    __shared__ int s_cnt[kSortBins], s_base[kSortBins];
    for (int base = blockIdx.x * kWindow; base < n; base += gridDim.x * kWindow) {
      if (threadIdx.x < kSortBins) s_cnt[threadIdx.x] = 0;
      // live here: a hit on a non-emissive surface (k_surface marked the others finished)
      const int n_live = cta_compact(base, n, s_idx, &s_n, [&](int i) { return Q.hit_t[i] != CUDART_INF_F; });
#pragma unroll 1
      for (int j0 = 0; j0 < n_live; j0 += blockDim.x) {
        RM_LOCKSTEP();
        const int j = j0 + threadIdx.x;
        const int i = j < n_live ? s_idx[j] : -1;
        int key = -1;
        if (i >= 0) {
            const int p = Q.pixel[i];
            const int fl = Q.flags[i];
            const int depth = fl & 255, n_med = (fl >> 16) & 31;
            const float t = Q.hit_t[i];
            Rng gen;
            gen.init(seed, (unsigned)p, Q.sample[i], kStreamIndirect, Q.drawn[i]);
            const int v = Q.hit_face[i];                        // dense vertex index (k_surface)
            const float *f = Q.surf + v;
            const float s_rough = f[16 * c], opacity = f[18 * c], s_eta = f[19 * c];
            const int id = __float_as_int(f[20 * c]);
            const bool entering = __float_as_int(f[21 * c]) != 0;
            // the medium stack as the reference's Medium::ior() / absorb() see it, and as they would after erase(id)
            float eta_all = 1.0f, eta_excl = 1.0f;
            V3 ab = splat3(1.0f);
            int n_excl = 0;
            for (int k = 0; k < n_med; k++) {
                const float ior = Q.med[(4 * k) * c + i];
                eta_all = fmaxf(eta_all, ior);
                ab = ab * mk3(Q.med[(4 * k + 1) * c + i], Q.med[(4 * k + 2) * c + i], Q.med[(4 * k + 3) * c + i]);
                if (Q.med_id[k * c + i] != id) { eta_excl = fmaxf(eta_excl, ior); n_excl++; }
            }
            // roughness regularisation along the path (src/render.cpp:143-146): both the carried factor and the vertex take the maximum
            const float rough = fmaxf(fmaxf(Q.rough[i], fmul(1.0f, s_rough)), s_rough);
            const float P_RR = fadd(1.0f, fmul(fsub(0.5f, 1.0f), fsqrt(rough)));
            bool doDirect = P_RR < 0.9f;
            const V3 absorb = get_absorb(ab, t);
            float P_reflect = 1.0f, F = 0.0f, eta_rel = s_eta;
            int n_after = n_med;
            if (opacity < kEps) {
                // calcEta (src/render.cpp:89-99); the base air entry is implicit (ior 1, absorb 1)
                const float eta2 = entering ? fmaxf(eta_all, s_eta) : eta_excl;
                if (!entering) n_after = n_excl + 1;            // erase(id) then insert(id, ...)
                eta_rel = fdiv(eta_all, eta2);
                Bsdf B;
                B.s = default_surface();
                B.inDir = -mk3(Q.d[i], Q.d[c + i], Q.d[2 * c + i]);
                B.s.surfaceNormal = mk3(f[3 * c], f[4 * c], f[5 * c]);
                B.s.eta = eta_rel;
                V3 refr;
                precise_refraction(B, refr, F);
                if (n_after == 0) P_reflect = fadd(0.24f, fmul(fsub(1.0f, 0.24f), F));
                else P_reflect = F;
                P_reflect = fmaxf(fsub(P_reflect, 1e-3f), 0.0f);
            }
            int mode;
            float first = P_reflect;                            // decision record slot 0
            bool pass_absorb = false;
            if (gen() <= P_reflect) {
                doDirect = doDirect && entering && n_after == 0;
                if (doDirect && (depth >= kMaxRayDepth || gen() > P_RR)) {
                    mode = kBounceNee;
                    first = fdiv(1.0f, P_reflect);
                    if (depth < kMaxRayDepth) first = fdiv(first, fsub(1.0f, P_RR));
                } else mode = kBounceReflect;
            } else {
                doDirect = doDirect && !entering && n_after == 1;
                if (doDirect && (depth >= kMaxRayDepth || gen() > P_RR)) {
                    mode = kBounceNee;
                    pass_absorb = true;
                    first = fdiv(1.0f, fsub(1.0f, P_reflect));
                    if (depth < kMaxRayDepth) first = fdiv(first, fsub(1.0f, P_RR));
                } else mode = kBounceRefract;
            }
            Q.flags[i] = (fl & 0x00ffffff) | (mode << 24) | (doDirect ? 1 << 26 : 0) | (pass_absorb ? 1 << 27 : 0);
            Q.drawn[i] = gen.drawn;
            Q.rough[i] = rough;
            float *d = Q.dec + v;
            d[0] = first; d[c] = P_RR; d[2 * c] = F; d[3 * c] = absorb.x; d[4 * c] = absorb.y; d[5 * c] = absorb.z; d[6 * c] = eta_rel;
            key = mode == kBounceReflect ? kKeyReflect + (opacity < kEps ? 1 : 0) : (mode == kBounceRefract ? kKeyRefract : kKeyNee + c_sampleCount[depth] - 1);
        }
        if (key >= 0) { Q.skey[i] = key; Q.srank[i] = atomicAdd(&s_cnt[key], 1); }      // position among the window's vertices of this key
      }
      // one block of each bin for this window
      __syncthreads();
      if (threadIdx.x < kSortBins) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(bins + threadIdx.x, s_cnt[threadIdx.x]) : 0;
      __syncthreads();
      for (int j = threadIdx.x; j < n_live; j += blockDim.x) {
        const int i = s_idx[j];
        Q.srank[i] += s_base[Q.skey[i]];
      }
    }
}

// exclusive offsets of the bins (one block)
__global__ void __launch_bounds__(kSortBins) k_sort_offsets(const int *__restrict__ bins, int *offs) {
    __shared__ int s[kSortBins];
    const int k = threadIdx.x;
    const int v = bins[k];
    s[k] = v;
    __syncthreads();
    for (int o = 1; o < kSortBins; o <<= 1) {
        const int add = k >= o ? s[k - o] : 0;
        __syncthreads();
        s[k] += add;
        __syncthreads();
    }
    offs[k] = s[k] - v;
    if (k == kSortBins - 1) offs[kSortBins] = s[k];
}

// the sorted queue: position -> path-queue slot
__global__ void __launch_bounds__(256) k_sort_scatter(PathQueue Q, const int *__restrict__ in_count, const int *__restrict__ offs, int *__restrict__ sorted) {
    const int n = min(*in_count, Q.cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (Q.hit_t[i] != CUDART_INF_F) sorted[offs[Q.skey[i]] + Q.srank[i]] = i;
}

// A sampled continuation: reflection (src/render.cpp:222-236) or refraction (264-283), dense over its segment of the
// sorted queue; the surviving vertex goes into the next path queue.
template <bool REFRACT>
__global__ void __launch_bounds__(kShadeBlock, kCtasBounce) k_continue(unsigned long long seed, PathQueue Qin, PathQueue Qout, int *out_count,
                                                                        const int *__restrict__ sorted, const int *__restrict__ offs) {
    const int lo = offs[REFRACT ? kKeyRefract : kKeyReflect], hi = offs[REFRACT ? kKeyNee : kKeyRefract];
    const int c = Qin.cap;
    for (int base = lo + blockIdx.x * blockDim.x; base < hi; base += gridDim.x * blockDim.x) {
        RM_LOCKSTEP();
        const int j = base + threadIdx.x;
        bool cont = false;
        int i = 0, p = 0, depth = 0, fl = 0, fails = 0;
        unsigned sample = 0;
        Rng gen;
        V3 T = splat3(1.0f), B0 = splat3(0.0f), newDir = splat3(0.0f);
        float W = 1.0f, rough = 0.0f, ior = 1.0f;
        RayDiff next;
        Bsdf B;
        B.s = default_surface();
        B.inDir = splat3(0.0f);
        if (j < hi) {
            i = sorted[j];
            p = Qin.pixel[i];
            sample = Qin.sample[i];
            fl = Qin.flags[i];
            depth = fl & 255;
            const V3 dir = mk3(Qin.d[i], Qin.d[c + i], Qin.d[2 * c + i]);
            T = mk3(Qin.T[i], Qin.T[c + i], Qin.T[2 * c + i]);
            B0 = mk3(Qin.B0[i], Qin.B0[c + i], Qin.B0[2 * c + i]);
            W = Qin.W[i];
            rough = Qin.rough[i];
            gen.init(seed, (unsigned)p, sample, kStreamIndirect, Qin.drawn[i]);
            const int v = Qin.hit_face[i];                      // dense vertex index: surface, hdP and decision records
            const float *d = Qin.dec + v;
            const float P_reflect = d[0], P_RR = d[c], F = d[2 * c];
            const V3 absorb = mk3(d[3 * c], d[4 * c], d[5 * c]);
            B.inDir = -dir;
            B.s = load_surface(Qin, v);
            ior = B.s.eta;
            B.s.roughness = rough;
            B.s.eta = d[6 * c];
            const bool doDirect = (fl >> 26) & 1;
            V3 bsdfPdf = splat3(CUDART_NAN_F);
            if (!REFRACT) {
                sample_reflection(B, gen, newDir, bsdfPdf, fails);
                bsdfPdf = div_true(bsdfPdf, P_reflect);
                if (doDirect) bsdfPdf = div_true(bsdfPdf, P_RR);
                RayDiff bd;                                   // only the direction differentials enter calc_dDdxy
                bd.dPdx = bd.dPdy = splat3(0.0f);
                bd.dDdx = mk3(Qin.diff[6 * c + i], Qin.diff[7 * c + i], Qin.diff[8 * c + i]);
                bd.dDdy = mk3(Qin.diff[9 * c + i], Qin.diff[10 * c + i], Qin.diff[11 * c + i]);
                V3 dDdx, dDdy;
                calc_dDdxy(dir, B.s.surfaceNormal, bd, dDdx, dDdy);
                next.dPdx = mk3(Qin.hdP[v], Qin.hdP[c + v], Qin.hdP[2 * c + v]);
                next.dPdy = mk3(Qin.hdP[3 * c + v], Qin.hdP[4 * c + v], Qin.hdP[5 * c + v]);
                next.dDdx = dDdx; next.dDdy = dDdy;
            } else {
                // the incoming differentials pass through a refraction unchanged (src/render.cpp:268,273)
                if (fabsf(fsub(B.s.eta, 1.0f)) < kEps) { newDir = dir; bsdfPdf = splat3(1.0f); }
                else { sample_btdf(B, gen, newDir, bsdfPdf, fails); bsdfPdf = bsdfPdf * fsub(1.0f, F); }
                next.dPdx = mk3(Qin.diff[0 * c + i], Qin.diff[1 * c + i], Qin.diff[2 * c + i]);
                next.dPdy = mk3(Qin.diff[3 * c + i], Qin.diff[4 * c + i], Qin.diff[5 * c + i]);
                next.dDdx = mk3(Qin.diff[6 * c + i], Qin.diff[7 * c + i], Qin.diff[8 * c + i]);
                next.dDdy = mk3(Qin.diff[9 * c + i], Qin.diff[10 * c + i], Qin.diff[11 * c + i]);
                bsdfPdf = div_true(bsdfPdf, fsub(1.0f, P_reflect));
                if (doDirect) bsdfPdf = div_true(bsdfPdf, P_RR);
            }
            if (isfinite_any(newDir) && depth != kMaxRayDepth) {
                cont = true;
                bsdfPdf = bsdfPdf * absorb;
                T = T * bsdfPdf;
                if (fails > 0) W = fmul(W, fdiv(1.0f, float(1 + fails)));
            }
        }
        const int slot = alloc_slot(out_count, cont);
        if (cont && slot < Qout.cap) {
            // the medium stack the continuing path carries: calcEta's erase(id) + insert(id, ...) on leaving a dielectric, then
            // the refraction's own insert (entering) or erase (leaving)
            const bool glass = B.s.opacity < kEps;
            const bool drop = glass && !B.s.entering;
            const bool append = glass && (REFRACT ? B.s.entering : !B.s.entering);
            const int n_med = (fl >> 16) & 31;
            const int oc = Qout.cap;
            int n_out = 0;
            for (int k = 0; k < n_med; k++) {
                const int mid = Qin.med_id[k * c + i];
                if (drop && mid == B.s.id) continue;
                Qout.med_id[n_out * oc + slot] = mid;
#pragma unroll
                for (int q = 0; q < 4; q++) Qout.med[(4 * n_out + q) * oc + slot] = Qin.med[(4 * k + q) * c + i];
                n_out++;
            }
            if (append && n_out < kMediumSlots) {
                Qout.med_id[n_out * oc + slot] = B.s.id;
                Qout.med[(4 * n_out) * oc + slot] = ior;
                Qout.med[(4 * n_out + 1) * oc + slot] = B.s.baseColor.x; Qout.med[(4 * n_out + 2) * oc + slot] = B.s.baseColor.y; Qout.med[(4 * n_out + 3) * oc + slot] = B.s.baseColor.z;
                n_out++;
            }
            store_path(Qout, slot, p, sample, gen.drawn, B.s.position, newDir, next, T, B0, W, rough, depth + 1, (fl >> 26) & 1, n_out);
        }
    }
}

// ------------------------------------------------------------------ NEE at a terminating vertex
// sampleDirectLight(bsdf, model, gen, sampleCount[depth]) (src/sampling.cpp:467-527) for every vertex of the NEE segment of
// the sorted queue (ordered by sample count, then material).  A vertex reserves its 1..6 shadow-queue slots up front; a
// sample that comes out invalid leaves a null item (aim = NaN) that the visibility pass skips.
__global__ void __launch_bounds__(kShadeBlock, kCtasNee) k_nee(DevScene S, FrameBuffers Fb, unsigned long long seed, PathQueue Q, const int *__restrict__ sorted,
                                             const int *__restrict__ offs, ShadowItem *sq, int *s_count, int s_cap, int *overflow) {
    const int lo = offs[kKeyNee], hi = offs[kSortBins];
    const int c = Q.cap;
    const int lane = threadIdx.x & 31;
    for (int base = lo + blockIdx.x * blockDim.x; base < hi; base += gridDim.x * blockDim.x) {
        RM_LOCKSTEP();                       // CTA-wide lock step per batch: the warps share the kernel's instruction-cache lines
        const int r = base + threadIdx.x;
        int cnt = 0, i = 0, fl = 0;
        if (r < hi) {
            i = sorted[r];
            fl = Q.flags[i];
            cnt = c_sampleCount[fl & 255];
        }
        // reserve cnt consecutive shadow slots per vertex: warp scan + one atomic per warp
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int first = 0;
        if (lane == 31 && incl > 0) first = atomicAdd(s_count, incl);
        first = __shfl_sync(0xffffffffu, first, 31) + incl - cnt;
        if (cnt == 0) continue;
        const bool pass_absorb = (fl >> 27) & 1;
        const int v = Q.hit_face[i];                            // dense vertex index: surface and decision records
        const float *d = Q.dec + v;
        const float nee_factor = d[0];
        const V3 absorb = mk3(d[3 * c], d[4 * c], d[5 * c]);
        const int p = Q.pixel[i];
        Rng gen;
        gen.init(seed, (unsigned)p, Q.sample[i], kStreamIndirect, Q.drawn[i]);
        Bsdf B;
        B.inDir = -mk3(Q.d[i], Q.d[c + i], Q.d[2 * c + i]);
        B.s = load_surface(Q, v);
        // the light samples see the regularised roughness and the relative eta of this vertex
        B.s.roughness = Q.rough[i];
        B.s.eta = d[6 * c];
        const V3 T = mk3(Q.T[i], Q.T[c + i], Q.T[2 * c + i]), B0 = mk3(Q.B0[i], Q.B0[c + i], Q.B0[2 * c + i]);
        const float W = Q.W[i];
        const float inv_spp = fdiv(1.0f, float(Fb.n_ind[p]));
        float lw[kMaxLights];
        float lw_total = 0.0f;
        bool any = true;
        if (S.sky_width == 0) { lw_total = light_weights(S, B, lw); any = lw_total != 0.0f; }
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            NeeOut ne;
            ne.valid = false;
            if (any) ne = nee_sample(S, B, gen, lw, lw_total, cnt);
            V3 b = B0, l = splat3(0.0f);
            float w = 0.0f;
            if (ne.valid) {
                V3 sb = ne.bsdf * nee_factor;                 // scaling()
                V3 lt = ne.light * sb;                        // passBsdf: light *= bsdfPdf
                if (pass_absorb) lt = lt * absorb;            // refract branch: passBsdf(samples, absorb) then one more level
                l = lt * T;
                w = fmul(fmul(ne.weight, W), inv_spp);
            } else {
                ne.dir = splat3(0.0f);
                ne.aim = CUDART_NAN_F;                        // null item
            }
            const int slot = first + k;
            if (slot >= s_cap) { atomicExch(overflow, 1); break; }
            write_shadow(sq + slot, p, B.s.position, ne, b, l, w);
        }
    }
}

// ------------------------------------------------------------------ visibility + accumulation
// rayHit_test over the shadow queue marks each item (its `vis` word); a second, fully coalesced pass
// accumulates the visible ones, so the divergent traversal never carries the accumulation code.
struct ShadowJob {
    static constexpr bool kOcclusion = true;
    ShadowItem *sq;
    RM_DI bool load(int i, V3 &o, V3 &d, float &aim) const {
        const float4 *src = reinterpret_cast<const float4 *>(sq + i);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        o = mk3(a.x, a.y, a.z); aim = a.w;
        d = mk3(b.x, b.y, b.z);
        return !isnan(aim);                    // null item (k_nee): vis stays 0
    }
    RM_DI void hit(int, float, int) const {}
    RM_DI void visibility(int i, bool occluded) const { sq[i].vis = occluded ? 0.0f : 1.0f; }
};

__global__ void __launch_bounds__(256) k_accum_shadow(FrameBuffers Fb, Accum Ac, const ShadowItem *__restrict__ sq, const int *__restrict__ s_count, int s_cap) {
    const int n = min(*s_count, s_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 *src = reinterpret_cast<const float4 *>(sq + i);
        const float4 d = src[3];
        if (d.w == 0.0f) continue;
        const float4 b = src[1], c = src[2];
        const int tag = __float_as_int(b.w);
        const V3 bs = mk3(c.x, c.y, c.z), li = mk3(d.x, d.y, d.z);
        if (tag < 0) add_direct(Ac, Fb, tag & 0x7fffffff, bs, li, c.w);
        else add_indirect(Ac, Fb, tag, bs, li, c.w);
    }
}

// ------------------------------------------------------------------ resolve
// Commit the held-back sample against the (cross-GPU) totals, then calcVar (src/render.cpp:510-516).
__global__ void k_commit_hold(Accum Ac, FrameBuffers Fb, int npix, bool never_drop) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float hc = Ac.hold_clum[p];
    if (!(hc >= 0.0f)) return;
    float total = Ac.clum_sum[2 * p], gmax = Ac.clum_max[p];
    bool drop = false;
    if (hc == gmax && !never_drop) drop = fdiv(hc, fadd(fsub(total, hc), kEps)) > 16.0f;      // clampThreshold
    if (!drop) {
        const float *h = Ac.hold + 8 * p;
        const float *gf = reinterpret_cast<const float *>(Fb.gbuffer + p);
        V3 base = gf[18] < kEps ? splat3(0.0f) : mk3(gf[9], gf[10], gf[11]);
        accum_split(Ac.rad + 16 * p + 8, Ac.rad + 16 * p + 12, base, mk3(h[0], h[1], h[2]), mk3(h[3], h[4], h[5]), h[6]);
    }
    Ac.hold_clum[p] = -1.0f;
}

__global__ void k_publish_max(Accum Ac, int npix) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npix) Ac.clum_max[p] = Ac.hold_clum[p];
}

__global__ void k_finalise(Accum Ac, FrameBuffers Fb, int p_begin, int p_end, float exposure, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is,
                           RmHitInfo *g_out) {
    int p = p_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p_end) return;
    RmRadiance *planes[4] = {Dd, Ds, Id, Is};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float *r = Ac.rad + 16 * p + 4 * k;
        V3 rad = mk3(r[0], r[1], r[2]) * exposure;
        float var = fmul(r[3], fmul(exposure, exposure));
        var = fsub(var, dot(rad, rad));
        if (var < 0.0f) var = 0.0f;
        planes[k][p].radiance[0] = rad.x; planes[k][p].radiance[1] = rad.y; planes[k][p].radiance[2] = rad.z;
        planes[k][p].Var = var;
    }
    // the reference restores the un-nudged baseColor only when the pixel produced indirect samples (529-530, 550)
    RmHitInfo g = Fb.gbuffer[p];
    if (Ac.clum_sum[2 * p + 1] > 0.0f) { g.baseColor[0] = Fb.sav_base[3 * p]; g.baseColor[1] = Fb.sav_base[3 * p + 1]; g.baseColor[2] = Fb.sav_base[3 * p + 2]; }
    g_out[p] = g;
}

} // namespace rm
