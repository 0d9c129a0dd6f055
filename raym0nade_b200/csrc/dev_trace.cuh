// dev_trace.cuh — ray / scene intersection on the device.
//
// Exact replica of the reference's traversal semantics, expressed iteratively:
//   rayInBox                 src/geometry.cpp:40-61  (quirky slab test: |d|<1e-4 is "parallel",
//                            tR += 1e-4 per axis, early return leaves partial state)
//   RayTriangleIntersection  src/geometry.cpp:63-87  (Moller-Trumbore, degenerate test |a|/|e1| < 1e-4)
//   BVH::dfs_rayHit          src/bvh.cpp:56-92       (near child first by tL, RIGHT child first on ties,
//                            deferred child re-tested against the shrunken t_max)
//   Model::rayHit / rayHit_test / TransparentTest    src/model.cpp:217-230,332-354
// The recursion becomes an explicit stack of {child ref, tL}; a deferred child is re-tested
// with `tL < t_max` when popped, which is exactly when the reference evaluates that condition.
// The same tree, the same arithmetic (dev_math.cuh), the same visit order => the same hit.
#pragma once
#include "dev_scene.cuh"
#include "dev_texture.cuh"

namespace rm {

struct TraceCounters { unsigned long long rays, box, tri; };

// One 256-bit read-only load (sm_100 LDG.E.256): half the L1 lookups of two 128-bit loads.  p is 32-byte aligned.
RM_DI void ldg256(const float4 *p, float4 &a, float4 &b) {
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
#else
    a = p[0]; b = p[1];          // the headers compiled as host code (tests/tools/cuda_on_host.h)
#endif
}

// A reference to a BVH child: inner node index u >= 1, or a leaf encoded as ~(faceL<<4 | count).
RM_DI int leaf_ref(int faceL, int faceR) { return ~((faceL << 4) | (faceR - faceL)); }
constexpr int kTraceDone = int(0x80000000u);     // "stack empty": never a valid leaf ref (faceL < 2^27)

struct RaySetup {
    V3 o, d;
    float inv[3];      // 1.0f / d[i]                    (src/geometry.cpp:48)
    unsigned flags;    // bit i: |d[i]| < eps_zero (src/geometry.cpp:42); bit 4+i: !(inv[i] >= 0)
};

// __frcp_rn is the correctly rounded reciprocal, i.e. bit-identical to the reference's 1.0f / x
RM_DI RaySetup setup_ray(V3 o, V3 d) {
    RaySetup r;
    r.o = o; r.d = d;
    r.inv[0] = __frcp_rn(d.x); r.inv[1] = __frcp_rn(d.y); r.inv[2] = __frcp_rn(d.z);
    r.flags = (fabsf(d.x) < kEps ? 1u : 0u) | (fabsf(d.y) < kEps ? 2u : 0u) | (fabsf(d.z) < kEps ? 4u : 0u) |
              (r.inv[0] >= 0.0f ? 0u : 16u) | (r.inv[1] >= 0.0f ? 0u : 32u) | (r.inv[2] >= 0.0f ? 0u : 64u);
    return r;
}

// One axis of rayInBox, branch-free; `live` carries the early-return state.
RM_DI void slab_axis(bool par, bool neg, float o, float inv, float b0, float b1, float &tL, float &tR, bool &live) {
    float t0 = fmul(fsub(b0, o), inv), t1 = fmul(fsub(b1, o), inv);
    float nL = fmaxf(tL, neg ? t1 : t0);
    float nR = fadd(fminf(tR, neg ? t0 : t1), kEps);
    bool outside = o < b0 || o > b1;
    bool upd = live && !par;
    bool kill = live && par && outside;
    tL = upd ? nL : tL;
    tR = upd ? nR : (kill ? -1.0f : tR);
    live = live && (par ? !outside : !(nL > nR));
}

RM_DI void ray_in_box(const RaySetup &r, float4 a, float4 b, float &tL, float &tR) {
    bool live = true;
    slab_axis(r.flags & 1u, r.flags & 16u, r.o.x, r.inv[0], a.x, a.w, tL, tR, live);
    slab_axis(r.flags & 2u, r.flags & 32u, r.o.y, r.inv[1], a.y, b.x, tL, tR, live);
    slab_axis(r.flags & 4u, r.flags & 64u, r.o.z, r.inv[2], a.z, b.y, tL, tR, live);
}

// The same test for a ray without a "parallel" axis (all |d[i]| >= eps_zero - almost every ray): the
// near / far plane of each axis is picked once per ray (lo/hi hold the selectors), and the early
// return becomes a predicate on the remaining updates.
RM_DI void slab_fast(float o, float inv, float bn, float bf, float &tL, float &tR, bool &live) {
    const float nL = fmaxf(tL, fmul(fsub(bn, o), inv));
    const float nR = fadd(fminf(tR, fmul(fsub(bf, o), inv)), kEps);
    if (live) { tL = nL; tR = nR; }
    live = live && !(nL > nR);
}
RM_DI void ray_in_box_fast(const RaySetup &r, float4 a, float4 b, float &tL, float &tR) {
    const bool nx = r.flags & 16u, ny = r.flags & 32u, nz = r.flags & 64u;
    bool live = true;
    slab_fast(r.o.x, r.inv[0], nx ? a.w : a.x, nx ? a.x : a.w, tL, tR, live);
    slab_fast(r.o.y, r.inv[1], ny ? b.x : a.y, ny ? a.y : b.x, tL, tR, live);
    slab_fast(r.o.z, r.inv[2], nz ? b.y : a.z, nz ? a.z : b.y, tL, tR, live);
}

// ------------------------------------------------------------------------------------------
// The 4-wide secondary-ray tree (wide_bvh.cpp): one 64-byte record per node = two 256-bit loads, holding the node's
// quantisation grid (origin o, step s) and, per child, its box in 8-bit grid units.  A child plane is o + s*q; along the
// ray that is t = q * (s * inv) + (o - org) * inv: one fused multiply-add per plane after a per-node, per-axis setup.
// This is NOT the reference's rayInBox - it has no counterpart there - it is a conservative slab test (the decoded box
// encloses the child's true box, the interval is widened by a few ulps and by rayInBox's own 1e-4), so every triangle the ray can hit is still
// reached and tested with the reference's RayTriangleIntersection.
struct WideRay {
    float inv[3];      // 1 / d[i], with |d[i]| clamped away from zero (box test only)
};
RM_DI WideRay setup_wide_ray(V3 d) {
    // rayInBox calls an axis with |d| < eps_zero "parallel" and only asks whether the origin lies inside the slab
    // (src/geometry.cpp:42-47).  A huge reciprocal says the same thing through the ordinary slab arithmetic: both planes map
    // to -huge / +huge when the origin is between them, to two values of one sign (an empty interval) when it is not.
    WideRay w;
    const float dx = fabsf(d.x) < kEps ? copysignf(1e-30f, d.x) : d.x, dy = fabsf(d.y) < kEps ? copysignf(1e-30f, d.y) : d.y,
                dz = fabsf(d.z) < kEps ? copysignf(1e-30f, d.z) : d.z;
    w.inv[0] = __frcp_rn(dx); w.inv[1] = __frcp_rn(dy); w.inv[2] = __frcp_rn(dz);
    return w;
}
// byte C of w as a float: the byte is dropped into the mantissa of 2^23 (one PRMT), then 2^23 is subtracted (exact).  The PRMT is
// spelled out so that the SELECTOR is its immediate and 2^23's bit pattern comes from the kernel's parameters (TraceTune::q8_magic -
// a value ptxas cannot see): given two constants, ptxas makes 2^23 the immediate and moves every one of the 24 selectors of a node
// step into a register first (IMAD.U32 Rx, RZ, RZ, URy before each PRMT - a tenth of the step's instructions, cuobjdump -sass).
template <int C> RM_DI float q8(unsigned w, unsigned magic) {
#ifdef __CUDA_ARCH__
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "n"(0x7440 | C));
    return __uint_as_float(r) - 8388608.0f;
#else
    (void)magic;
    return float((w >> (8 * C)) & 0xffu);
#endif
}

// returns t or +INF.  Straight-line: the early returns of the reference become one select at the end.  Under SIMT
// an early return only saves work when every lane of the warp takes it, and the divergent returns cost more than
// the arithmetic they skip; the value is a pure function of the inputs, so the result is the reference's.
RM_DI float ray_triangle(const RaySetup &r, float4 q0, float4 q1, float4 q2) {
    V3 v0 = mk3(q0.x, q0.y, q0.z);
    V3 e1 = mk3(q0.w, q1.x, q1.y);
    V3 e2 = mk3(q1.z, q1.w, q2.x);
    V3 h = cross(r.d, e2);
    float a = dot(e1, h);
    // degenerate test |a| / |e1| < eps_zero.  The correctly rounded quotient can only fall below
    // eps_zero when |a| < 1.0001e-4 * |e1| (rounding moves either side by < 1e-7 relative), so the
    // division is evaluated only in that sliver; NaN / zero |e1| take the same side as the reference.
    float aa = fabsf(a);
    bool miss = false;
    if (aa < fmul(1.0001e-4f, q2.y)) miss = fdiv(aa, q2.y) < kEps;
    float f = __frcp_rn(a);
    V3 s = r.o - v0;
    float u = fmul(f, dot(s, h));
    V3 q = cross(s, e1);
    float v = fmul(f, dot(r.d, q));
    float t = fmul(f, dot(e2, q));
    miss = miss || u < 0.0f || u > 1.0f || v < 0.0f || fadd(u, v) > 1.0f;
    return miss ? CUDART_INF_F : t;
}

// barycentric (src/geometry.cpp:89-103): returns (gamma, alpha, beta)
RM_DI V3 barycentric(V3 v0, V3 v1, V3 v2, V3 P) {
    V3 v0v1 = v1 - v0, v0v2 = v2 - v0;
    V3 n = cross(v0v1, v0v2);
    float denom = dot(n, n);
    V3 v0P = P - v0;
    float alpha = fdiv(dot(cross(v0P, v0v2), n), denom);
    float beta = fdiv(dot(cross(v0v1, v0P), n), denom);
    float gamma = fsub(fsub(1.0f, alpha), beta);
    return mk3(gamma, alpha, beta);
}

struct FaceShade {
    V3 v[3];
    V2 uv[3];
    V3 n[3];
    int material;
};

RM_DI FaceShade load_face(const DevScene &S, int face) {
    const float4 *p = S.shade + size_t(face) * 7;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4), f = __ldg(p + 5), g = __ldg(p + 6);
    FaceShade F;
    F.v[0] = mk3(a.x, a.y, a.z); F.v[1] = mk3(a.w, b.x, b.y); F.v[2] = mk3(b.z, b.w, c.x);
    F.uv[0] = mk2(c.y, c.z); F.uv[1] = mk2(c.w, d.x); F.uv[2] = mk2(d.y, d.z);
    F.n[0] = mk3(d.w, e.x, e.y); F.n[1] = mk3(e.z, e.w, f.x); F.n[2] = mk3(f.y, f.z, f.w);
    F.material = __float_as_int(g.x);
    return F;
}

RM_DI V2 interp_uv(const FaceShade &F, V3 bary) {
    return (bary.x * F.uv[0] + bary.y * F.uv[1]) + bary.z * F.uv[2];
}

// TransparentTest (src/model.cpp:217-230): true when the hit texel is an alpha cut-out
RM_DI bool transparent_test(const DevScene &S, const RaySetup &r, float t, int face) {
    float cut = __ldg(&S.tri[size_t(face) * kTriStride + 2].z);
    if (cut == 0.0f) return false;                       // material without hasFullyTransparentPart
    FaceShade F = load_face(S, S.face_map ? __ldg(S.face_map + face) : face);
    V3 P = r.o + r.d * t;
    V3 bary = barycentric(F.v[0], F.v[1], F.v[2], P);
    V2 uv = interp_uv(F, bary);
    return mat_diffuse_alpha0(S, S.materials[F.material], uv.x, uv.y) < kEps;
}

// ------------------------------------------------------------------------------------------
// Persistent-warp trace engine.
//
// Every lane of a warp owns one ray at a time and keeps its whole traversal state in registers
// (+ its column of the shared-memory stack).  The warp advances in lock-step iterations; in each one
// either the lanes sitting on an inner node test their two children, or the lanes holding a leaf test
// their next triangle - the warp votes (ballot + popc) for the step with more lanes ready.  A lane
// whose ray is finished goes idle; when fewer than kRefillLive lanes are left the warp drops back to
// the fetch section and the idle lanes (found with a ballot, ranked with popc) take the next rays of
// the warp's current chunk - chunks of kTraceChunk consecutive rays are handed out from a global
// cursor with one atomicAdd per chunk - while the busy lanes keep their state.
//
// The per-ray semantics are exactly the reference's:
//   BVH::dfs_rayHit          src/bvh.cpp:56-92    near child first by tL, RIGHT child first on ties,
//                            a deferred child is re-tested `tL < t_max` when it is popped
//   Model::rayHit            src/model.cpp:332-341  (<= 8 re-traces through alpha cut-outs)
//   Model::rayHit_test       src/model.cpp:343-354
// A Job supplies the rays and consumes the results:
//   static constexpr bool kOcclusion;
//   bool load(int i, V3 &o, V3 &d, float &aim);          false = no ray at this index
//   void hit(int i, float t, int face);                  closest-hit result (kOcclusion == false)
//   void visibility(int i, bool occluded);               rayHit_test result (kOcclusion == true)
constexpr int kTraceChunk = 64;
#ifndef RM_LEAF_REPS
#define RM_LEAF_REPS 1              // leaf sub-steps (of two triangles each) a warp runs before it votes again
#endif

struct TraceTune {
    int refill_live;     // refill idle lanes once fewer than this many lanes are busy
    int w_inner, w_leaf; // vote weights: the inner step runs when n_inner * w_inner >= n_leaf * w_leaf
    int smem_levels;     // stack entries per thread held in shared memory; deeper ones (rare) go to a small local array
    unsigned q8_magic = 0x4B000000u;      // the bit pattern of 2^23, handed over as a kernel parameter so that ptxas cannot fold it (see q8)
};
constexpr int kStackSpill = 28;   // local spill entries: smem_levels + kStackSpill >= any tree depth we build (<= 40)
constexpr int kStackSpillWide = 82; // the 4-wide tree defers up to three children per level (rm_scene_upload checks 3 * levels against it)

template <class Job, bool COUNT, bool WIDE = false>
RM_DI void trace_engine(const DevScene &S, Job &job, const int n, int *cursor, int2 *stack, const int stride, TraceCounters &cnt,
                        const TraceTune tune) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    int root;
    if (WIDE) root = 0;                          // record 0; a scene that is one leaf still has a root record with that one child
    else if (S.root_is_leaf) {
        float4 rb = __ldg(S.nodes + 3);
        root = leaf_ref(__float_as_int(rb.z), __float_as_int(rb.w));
    } else root = 1;
    // a rayHit_test may stop at the first accepted triangle below aimDepth when no material has alpha
    // cut-outs: "the closest accepted t is < aim" <=> "some accepted t is < aim", and the traversal state is
    // identical up to that triangle, so the boolean is the reference's
    const bool anyhit = Job::kOcclusion && !S.any_cutout;
    const bool expl = S.explicit_children != 0;
    // A ray's stack rarely holds more than ~14 deferred children even in a 23-level tree (scripts/bvh_quality_proto.cpp), so
    // only the first tune.smem_levels entries live in shared memory - which leaves more of the SM's 256 KB as L1 - and the
    // rest in a local array that is almost never touched.
    const int cap = tune.smem_levels;
    int2 spill[WIDE ? kStackSpillWide : kStackSpill];
    auto push = [&](int sp_, int2 e) { if (sp_ < cap) stack[sp_ * stride] = e; else spill[sp_ - cap] = e; };
    auto peek = [&](int sp_) { return sp_ < cap ? stack[sp_ * stride] : spill[sp_ - cap]; };

    bool active = false;
    int chunk_next = 0, chunk_end = 0;          // warp-uniform: rays of the current chunk not handed out yet
    bool exhausted = false;                     // warp-uniform: the global cursor ran past n
    RaySetup r;
    WideRay wr;
    float t_min = kEps, t = CUDART_INF_F, aim = CUDART_INF_F, t_reset = CUDART_INF_F;
    int face = -1, cur = kTraceDone, sp = 0, pass = 0, idx = 0, ti = -1, tend = 0;

    for (;;) {
        // ---- report: the lanes whose BVH::rayHit ended since the last refill resolve cut-outs and hand their result over
        // together - run per lane as soon as it ends, this block would execute on most iterations with one or two lanes in it
        if (active && cur == kTraceDone) {
            bool again = false;
            if (!Job::kOcclusion) {
                if (t != CUDART_INF_F && S.any_cutout && transparent_test(S, r, t, face)) {
                    t_min = fadd(t, kEps); t = CUDART_INF_F; face = -1; pass++;
                    again = pass < 8;
                }
                if (!again) job.hit(idx, t, (S.face_map && face >= 0) ? __ldg(S.face_map + face) : face);
            } else {
                bool occluded = !(t >= aim);
                if (occluded && S.any_cutout && transparent_test(S, r, t, face)) {
                    t_min = fadd(t, kEps); t = t_reset; face = -1; pass++;
                    again = pass < 8;                 // 8 cut-outs in a row: the reference reports "blocked"
                }
                if (!again) job.visibility(idx, occluded);
            }
            if (again) { cur = root; sp = 0; ti = -1; cnt.rays++; }
            else active = false;
        }
        // ---- fetch: idle lanes take the next rays
        unsigned idle = __ballot_sync(FULL, !active);
        while (idle != 0u) {
            if (chunk_next >= chunk_end) {
                if (exhausted) break;
                int base = 0;
                if (lane == 0) base = atomicAdd(cursor, kTraceChunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; break; }
                chunk_next = base;
                chunk_end = min(base + kTraceChunk, n);
            }
            const int avail = chunk_end - chunk_next;
            const int rank = __popc(idle & lt_mask);
            if (!active && rank < avail) {
                idx = chunk_next + rank;
                V3 o, d;
                if (job.load(idx, o, d, aim)) {
                    r = setup_ray(o, d);
                    if (WIDE) wr = setup_wide_ray(d);
                    t_min = kEps;
                    t_reset = Job::kOcclusion ? fadd(aim, kEps) : CUDART_INF_F;
                    t = t_reset;
                    face = -1; cur = root; sp = 0; pass = 0; ti = -1;
                    active = true;
                    cnt.rays++;
                }
            }
            chunk_next += min(avail, __popc(idle));
            idle = __ballot_sync(FULL, !active);
        }
        unsigned live = __ballot_sync(FULL, active);
        if (live == 0u) break;
        const int need = (exhausted && chunk_next >= chunk_end) ? 1 : tune.refill_live;

        // ---- trace: one step per iteration - either every lane that sits on an inner node tests its two
        // children, or every lane that holds a leaf tests its next triangle - whichever has more lanes
        // ready; the others wait a turn.  Lanes therefore never idle through a whole descent or a whole
        // leaf of their neighbours, and both step bodies run with most of their lanes busy.
        live = __ballot_sync(FULL, active && cur != kTraceDone);
        do {
            const bool stepping = active && cur != kTraceDone;        // a lane whose ray has ended waits for the next report / refill
            const bool wantI = stepping && cur >= 0;
            const int nI = __popc(__ballot_sync(FULL, wantI));
            const int nL = __popc(live) - nI;
            if (nI * tune.w_inner >= nL * tune.w_leaf) {
                if (WIDE) {
                  if (wantI) {
                    // ---- one 4-wide node: 64 bytes, up to four children
                    const float4 *nd = S.nodes + (size_t(cur) << 2);
                    float4 h0, h1, h2, h3;
                    ldg256(nd, h0, h1);
                    ldg256(nd + 2, h2, h3);
                    // per axis: t(q) = q * a + b with a = s * inv, b = (o - org) * inv
                    const float ax = h0.w * wr.inv[0], ay = h1.x * wr.inv[1], az = h1.y * wr.inv[2];
                    // (o - org) first: the difference is formed at the precision of the coordinates, as the reference's (b - o) * inv is;
                    // o * inv - org * inv would carry the rounding of two large products into a small t
                    const float bx = (h0.x - r.o.x) * wr.inv[0], by = (h0.y - r.o.y) * wr.inv[1], bz = (h0.z - r.o.z) * wr.inv[2];
                    const unsigned qlx = __float_as_uint(h1.z), qly = __float_as_uint(h1.w), qlz = __float_as_uint(h2.x);
                    const unsigned qhx = __float_as_uint(h2.y), qhy = __float_as_uint(h2.z), qhz = __float_as_uint(h2.w);
                    // the near plane of an axis is the lower one when the ray runs upwards along it
                    const bool ux = wr.inv[0] >= 0.0f, uy = wr.inv[1] >= 0.0f, uz = wr.inv[2] >= 0.0f;
                    const unsigned nx = ux ? qlx : qhx, fx = ux ? qhx : qlx, ny = uy ? qly : qhy, fy = uy ? qhy : qly, nz = uz ? qlz : qhz, fz = uz ? qhz : qlz;
                    const unsigned meta = __float_as_uint(h3.z);
                    const int child_base = __float_as_int(h3.x), tri_base = __float_as_int(h3.y);
                    unsigned key[4];
                    const unsigned magic = tune.q8_magic;
#define RM_WIDE_CHILD(c)                                                                                                                                          \
                    {                                                                                                                                             \
                        const float tn = fmaxf(fmaxf(fmaf(q8<c>(nx, magic), ax, bx), fmaf(q8<c>(ny, magic), ay, by)), fmaxf(fmaf(q8<c>(nz, magic), az, bz), t_min)); \
                        const float tf = fminf(fminf(fmaf(q8<c>(fx, magic), ax, bx), fmaf(q8<c>(fy, magic), ay, by)), fminf(fmaf(q8<c>(fz, magic), az, bz), t));     \
                        /* a few ulps, and the slack rayInBox gives its far plane */                                                                             \
                        const bool hitc = ((meta >> (8 * c)) & 0xffu) != 0u && tn * 0.999998f <= tf * 1.000002f + kEps;                                           \
                        /* sort key: the entry distance (positive, so its bit pattern orders like the float) with the slot in the low bits */                    \
                        key[c] = hitc ? ((__float_as_uint(tn) & ~3u) | unsigned(c)) : 0xffffffffu;                                                                \
                    }
                    RM_WIDE_CHILD(0) RM_WIDE_CHILD(1) RM_WIDE_CHILD(2) RM_WIDE_CHILD(3)
#undef RM_WIDE_CHILD
                    if (COUNT) cnt.box += (meta & 0xffu ? 1 : 0) + (meta & 0xff00u ? 1 : 0) + (meta & 0xff0000u ? 1 : 0) + (meta & 0xff000000u ? 1 : 0);     // child boxes tested
                    // 5-comparator network: ascending by entry distance, misses last
#define RM_CSWAP(a, b) { const unsigned lo_ = min(key[a], key[b]), hi_ = max(key[a], key[b]); key[a] = lo_; key[b] = hi_; }
                    RM_CSWAP(0, 1) RM_CSWAP(2, 3) RM_CSWAP(0, 2) RM_CSWAP(1, 3) RM_CSWAP(1, 2)
#undef RM_CSWAP
                    auto ref_of = [&](unsigned k) {
                        const unsigned m = (meta >> (8 * (k & 3u))) & 0xffu;
                        return (m & 0x80u) ? child_base + int(m & 0x7fu) : ~(((tri_base + int(m >> 2)) << 4) | int(m & 3u));
                    };
                    // far children first onto the stack, the nearest one is visited next
#pragma unroll
                    for (int c = 3; c >= 1; c--)
                        if (key[c] != 0xffffffffu) { push(sp, make_int2(ref_of(key[c]), int(key[c] & ~3u))); sp++; }
                    if (key[0] != 0xffffffffu) cur = ref_of(key[0]);
                    else {
                        cur = kTraceDone;
                        while (sp > 0) {
                            sp--;
                            const int2 e = peek(sp);
                            if (__int_as_float(e.y) < t) { cur = e.x; break; }
                        }
                    }
                    ti = -1;
                  }
                } else if (wantI) {
                    const float4 *nd = S.nodes + (size_t(cur) << 2);        // children 2u, 2u+1: one 64-byte block
                    float4 a0, b0, a1, b1;
                    ldg256(nd, a0, b0);
                    ldg256(nd + 2, a1, b1);
                    float tL0 = t_min, tR0 = t, tL1 = t_min, tR1 = t;
                    if (r.flags & 7u) { ray_in_box(r, a0, b0, tL0, tR0); ray_in_box(r, a1, b1, tL1, tR1); }
                    else { ray_in_box_fast(r, a0, b0, tL0, tR0); ray_in_box_fast(r, a1, b1, tL1, tR1); }
                    if (COUNT) cnt.box += 2;
                    const int fr0 = __float_as_int(b0.w), fr1 = __float_as_int(b1.w);
                    // an inner child's own children: blocks 2u, 2u+1 of the implicit heap (the reference's tree), or the
                    // block its record names (the secondary-ray tree)
                    const int ref0 = fr0 ? leaf_ref(__float_as_int(b0.z), fr0) : (expl ? __float_as_int(b0.z) : (cur << 1));
                    const int ref1 = fr1 ? leaf_ref(__float_as_int(b1.z), fr1) : (expl ? __float_as_int(b1.z) : (cur << 1 | 1));
                    const bool ok0 = tL0 < tR0, ok1 = tL1 < tR1;
                    const bool zero_first = tL0 < tL1;
                    const int first = zero_first ? ref0 : ref1, second = zero_first ? ref1 : ref0;
                    const bool okF = zero_first ? ok0 : ok1, okS = zero_first ? ok1 : ok0;
                    const float tLS = zero_first ? tL1 : tL0;
                    if (okF) {
                        if (okS) { push(sp, make_int2(second, __float_as_int(tLS))); sp++; }
                        cur = first;
                    } else if (okS && tLS < t) cur = second;
                    else {
                        cur = kTraceDone;
                        while (sp > 0) {
                            sp--;
                            const int2 e = peek(sp);
                            if (__int_as_float(e.y) < t) { cur = e.x; break; }
                        }
                    }
                    ti = -1;
                }
            } else if (stepping && !wantI) {
                if (ti < 0) { const int x = ~cur; ti = x >> 4; tend = ti + (x & 15); }        // first visit of this leaf
#if RM_LEAF_REPS > 1
#pragma unroll 1
              for (int rep = 0; rep < RM_LEAF_REPS && ti >= 0; rep++) {
#endif
                // up to two triangles of the leaf per step, their six loads issued together
                const float4 *q = S.tri + size_t(ti) * kTriStride;
                const bool two = ti + 1 < tend;
#if RM_TRI_STRIDE == 4
                float4 qa, qb, qd, qe;
                ldg256(q, qa, qb);
                const float4 qc = __ldg(q + 2);
                float4 qf = qc;
                qd = qa; qe = qb;
                if (two) { ldg256(q + 4, qd, qe); qf = __ldg(q + 6); }
#else
                const float4 qa = __ldg(q), qb = __ldg(q + 1), qc = __ldg(q + 2);
                float4 qd = qa, qe = qb, qf = qc;
                if (two) { qd = __ldg(q + 3); qe = __ldg(q + 4); qf = __ldg(q + 5); }
#endif
                // both tests are evaluated straight-line and interleaved (a lane without a second triangle repeats the first)
                const float tt = ray_triangle(r, qa, qb, qc);
                const float t2 = ray_triangle(r, qd, qe, qf);
                if (COUNT) cnt.tri++;
                bool stop = false;
                if (t_min < tt && tt < t) {
                    t = tt;
                    face = ti;
                    stop = anyhit && tt < aim;
                }
                if (two && !stop) {
                    ti++;
                    if (COUNT) cnt.tri++;
                    if (t_min < t2 && t2 < t) {
                        t = t2;
                        face = ti;
                        stop = anyhit && t2 < aim;
                    }
                }
                ti++;
                if (stop || ti == tend) {
                    cur = kTraceDone;
                    if (!stop)
                        while (sp > 0) {
                            sp--;
                            const int2 e = peek(sp);
                            if (__int_as_float(e.y) < t) { cur = e.x; break; }
                        }
                    ti = -1;
                }
#if RM_LEAF_REPS > 1
              }
#endif
            }
            live = __ballot_sync(FULL, active && cur != kTraceDone);
        } while (__popc(live) >= need);
    }
}

} // namespace rm
