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

constexpr int kStackDepth = 32;   // tree depth is <= 20 for 5 M triangles

struct TraceCounters { unsigned long long rays, box, tri; };

// A reference to a BVH child: inner node index u >= 1, or a leaf encoded as ~(faceL<<4 | count).
RM_DI int leaf_ref(int faceL, int faceR) { return ~((faceL << 4) | (faceR - faceL)); }

struct RaySetup {
    V3 o, d;
    float inv[3];      // 1.0f / d[i]                    (src/geometry.cpp:48)
    bool par[3];       // |d[i]| < eps_zero              (src/geometry.cpp:42)
};

RM_DI RaySetup setup_ray(V3 o, V3 d) {
    RaySetup r;
    r.o = o; r.d = d;
    r.par[0] = fabsf(d.x) < kEps; r.par[1] = fabsf(d.y) < kEps; r.par[2] = fabsf(d.z) < kEps;
    r.inv[0] = frcp(d.x); r.inv[1] = frcp(d.y); r.inv[2] = frcp(d.z);
    return r;
}

// One axis of rayInBox; `live` carries the early-return state.
RM_DI void slab_axis(bool par, float o, float inv, float b0, float b1, float &tL, float &tR, bool &live) {
    if (!live) return;
    if (par) {
        if (o < b0 || o > b1) { tR = -1.0f; live = false; }
    } else {
        float tn, tf;
        if (inv >= 0.0f) { tn = fmul(fsub(b0, o), inv); tf = fmul(fsub(b1, o), inv); }
        else { tn = fmul(fsub(b1, o), inv); tf = fmul(fsub(b0, o), inv); }
        tL = fmaxf(tL, tn);
        tR = fminf(tR, tf);
        tR = fadd(tR, kEps);
        if (tL > tR) live = false;
    }
}

RM_DI void ray_in_box(const RaySetup &r, float4 a, float4 b, float &tL, float &tR) {
    bool live = true;
    slab_axis(r.par[0], r.o.x, r.inv[0], a.x, a.w, tL, tR, live);
    slab_axis(r.par[1], r.o.y, r.inv[1], a.y, b.x, tL, tR, live);
    slab_axis(r.par[2], r.o.z, r.inv[2], a.z, b.y, tL, tR, live);
}

// returns t or +INF
RM_DI float ray_triangle(const RaySetup &r, float4 q0, float4 q1, float4 q2) {
    V3 v0 = mk3(q0.x, q0.y, q0.z);
    V3 e1 = mk3(q0.w, q1.x, q1.y);
    V3 e2 = mk3(q1.z, q1.w, q2.x);
    V3 h = cross(r.d, e2);
    float a = dot(e1, h);
    if (fdiv(fabsf(a), q2.y) < kEps) return CUDART_INF_F;
    float f = frcp(a);
    V3 s = r.o - v0;
    float u = fmul(f, dot(s, h));
    if (u < 0.0f || u > 1.0f) return CUDART_INF_F;
    V3 q = cross(s, e1);
    float v = fmul(f, dot(r.d, q));
    if (v < 0.0f || fadd(u, v) > 1.0f) return CUDART_INF_F;
    return fmul(f, dot(e2, q));
}

// BVH::rayHit: closest hit in (t_min, t_max).  `stack` is this thread's column of a shared-memory
// array, entries strided by `stride` (conflict-free).  ANYHIT: stop at the first accepted
// triangle - used only where that cannot change the caller's answer (see ray_occluded).
template <bool COUNT, bool ANYHIT>
RM_DI void bvh_ray_hit(const DevScene &S, const RaySetup &r, float t_min, float &t_max, int &face,
                       int2 *stack, int stride, TraceCounters &cnt, float any_limit = 0.0f) {
    cnt.rays++;                      // rays are always counted; COUNT adds box / triangle tests
    int sp = 0;
    int cur;
    if (S.root_is_leaf) {
        float4 rb = __ldg(S.nodes + 3);
        cur = leaf_ref(__float_as_int(rb.z), __float_as_int(rb.w));
    } else cur = 1;
    for (;;) {
        if (cur >= 0) {
            const float4 *n = S.nodes + (size_t(cur) << 2);        // children 2u, 2u+1: one 64-byte block
            float4 a0 = __ldg(n), b0 = __ldg(n + 1), a1 = __ldg(n + 2), b1 = __ldg(n + 3);
            float tL0 = t_min, tR0 = t_max, tL1 = t_min, tR1 = t_max;
            ray_in_box(r, a0, b0, tL0, tR0);
            ray_in_box(r, a1, b1, tL1, tR1);
            if (COUNT) cnt.box += 2;
            int fr0 = __float_as_int(b0.w), fr1 = __float_as_int(b1.w);
            int ref0 = fr0 ? leaf_ref(__float_as_int(b0.z), fr0) : (cur << 1);
            int ref1 = fr1 ? leaf_ref(__float_as_int(b1.z), fr1) : (cur << 1 | 1);
            bool ok0 = tL0 < tR0, ok1 = tL1 < tR1;
            int first, second;
            bool okF, okS;
            float tLS;
            if (tL0 < tL1) { first = ref0; okF = ok0; second = ref1; okS = ok1; tLS = tL1; }
            else { first = ref1; okF = ok1; second = ref0; okS = ok0; tLS = tL0; }
            if (okF) {
                if (okS) { stack[sp * stride] = make_int2(second, __float_as_int(tLS)); sp++; }
                cur = first;
                continue;
            }
            if (okS && tLS < t_max) { cur = second; continue; }
        } else {
            int x = ~cur;
            int f0 = x >> 4, f1 = f0 + (x & 15);
            for (int i = f0; i < f1; i++) {
                const float4 *q = S.tri + size_t(i) * 3;
                float t = ray_triangle(r, __ldg(q), __ldg(q + 1), __ldg(q + 2));
                if (t_min < t && t < t_max) {
                    t_max = t;
                    face = i;
                    if (ANYHIT && t < any_limit) { if (COUNT) cnt.tri += unsigned(i - f0 + 1); return; }
                }
            }
            if (COUNT) cnt.tri += unsigned(f1 - f0);
        }
        // pop: a deferred child is entered only if its tL is still below the current t_max
        for (;;) {
            if (sp == 0) return;
            sp--;
            int2 e = stack[sp * stride];
            if (__int_as_float(e.y) < t_max) { cur = e.x; break; }
        }
    }
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
    float cut = __ldg(&S.tri[size_t(face) * 3 + 2].z);
    if (cut == 0.0f) return false;                       // material without hasFullyTransparentPart
    FaceShade F = load_face(S, face);
    V3 P = r.o + r.d * t;
    V3 bary = barycentric(F.v[0], F.v[1], F.v[2], P);
    V2 uv = interp_uv(F, bary);
    return mat_diffuse_alpha0(S, S.materials[F.material], uv.x, uv.y) < kEps;
}

// Model::rayHit (src/model.cpp:332-341): face = -1 and t = INF on a miss
template <bool COUNT>
RM_DI void ray_hit(const DevScene &S, const RaySetup &r, float &t, int &face, int2 *stack, int stride, TraceCounters &cnt) {
    float t_min = kEps;
    t = CUDART_INF_F;
    face = -1;
    for (int T = 0; T < 8; T++) {
        bvh_ray_hit<COUNT, false>(S, r, t_min, t, face, stack, stride, cnt);
        if (t == CUDART_INF_F) return;
        if (!S.any_cutout || !transparent_test(S, r, t, face)) return;
        t_min = fadd(t, kEps);
        t = CUDART_INF_F;
        face = -1;
    }
}

// Model::rayHit_test (src/model.cpp:343-354): true = blocked before aimDepth.
// The reference runs a full closest-hit search in (t_min, aim+eps) and then asks whether the
// closest hit lies below aimDepth.  When no material has alpha cut-outs, "the closest accepted
// t is < aim" is equivalent to "some accepted t is < aim", and the traversal state is identical
// up to the first accepted triangle, so stopping there (ANYHIT) returns the same boolean.
template <bool COUNT>
RM_DI bool ray_occluded(const DevScene &S, const RaySetup &r, float aim, int2 *stack, int stride, TraceCounters &cnt) {
    float t_min = kEps;
    float t_lim = fadd(aim, kEps);
    if (!S.any_cutout) {
        // stop at the first accepted triangle with t < aim; triangles accepted with t in
        // [aim, aim+eps) only shrink t_max, as in the reference, and the search goes on
        float t = t_lim;
        int face = -1;
        bvh_ray_hit<COUNT, true>(S, r, t_min, t, face, stack, stride, cnt, aim);
        return !(t >= aim);
    }
    float t = t_lim;
    int face = -1;
    for (int T = 0; T < 8; T++) {
        bvh_ray_hit<COUNT, false>(S, r, t_min, t, face, stack, stride, cnt);
        if (t >= aim) return false;
        if (!transparent_test(S, r, t, face)) return true;
        t_min = fadd(t, kEps);
        t = t_lim;
        face = -1;
    }
    return true;
}

} // namespace rm
