// sah_sweep.h — the per-element steps of the device's top-down sweep-SAH builder of the secondary-ray tree (gpu_sah_bvh.cu).
//
// The builder works level by level on ALL nodes of a level at once, and every step is a map or a scan over the n triangles:
// the triangles are kept in three lists, one per axis, each sorted by box centre along its axis and partitioned so that the
// triangles of a node occupy the same range [L, R) in all three.  One level =
//   1. a segmented inclusive scan of box unions along each list, forwards and backwards (6 n items, one scan call)
//      -> for every split position of every node on every axis the surface of what lies left and right of it;
//   2. cost(axis, i) = area_left * count_left + area_right * count_right, minimum per node (64-bit atomicMin of
//      cost | distance from the middle | axis: ties go to the more balanced split);
//   3. per node: split there - or, where the depth cap would otherwise be at risk, at the median of the widest axis -
//      and hand out the children's node indices and next-level slots;
//   4. a stable partition of all three lists by the side each triangle went to (one exclusive scan of 3 n flags + a scatter),
//      which keeps every list sorted inside the children.
// No step depends on how many nodes a level has, so the top of the tree (a few huge nodes) and its bottom (hundreds of
// thousands of tiny ones) run at the same speed, and the result is the exact sweep SAH, not a binned one.
//
// The functions below are the bodies of those maps, `__host__ __device__`, so that tests/tools/sah_sweep_host.cpp can run
// the same code on the CPU with sequential loops in place of the launches and scans (tests/test_cpu_host.py).  There is no
// reference counterpart: the reference's tree (src/bvh.cpp:18-54, object median) is what primary rays traverse.
#pragma once
#include <cmath>
#include <cstdint>

#include <vector_types.h>

#if defined(__CUDACC__)
#define RM_SHD __host__ __device__ __forceinline__
#else
#define RM_SHD inline
#endif

namespace rm_sah {

constexpr unsigned long long kNoSplit = ~0ull;

struct alignas(16) SweepItem {
    float lx, ly, lz;
    int flag;              // 1: a segment (a node's range, in scan direction) starts here
    float hx, hy, hz;
    int pad;
};

// the segmented-scan operator: (f1, b1) + (f2, b2) = (f1 | f2, f2 ? b2 : b1 U b2)
struct SweepUnion {
    RM_SHD SweepItem operator()(const SweepItem &a, const SweepItem &b) const {
        SweepItem r = b;
        if (!b.flag) {
            r.lx = fminf(a.lx, b.lx); r.ly = fminf(a.ly, b.ly); r.lz = fminf(a.lz, b.lz);
            r.hx = fmaxf(a.hx, b.hx); r.hy = fmaxf(a.hy, b.hy); r.hz = fmaxf(a.hz, b.hz);
            r.flag = a.flag;
        }
        return r;
    }
};

RM_SHD float sweep_area(const SweepItem &b) {
    const float dx = b.hx - b.lx, dy = b.hy - b.ly, dz = b.hz - b.lz;
    return dx * dy + dy * dz + dz * dx;
}

struct Level {
    int n;                      // triangles
    const float4 *tlo, *thi;    // triangle boxes, by triangle
    const float4 *tbox;         // the same boxes as {lo, hi} pairs, 32 bytes per triangle: one 256-bit load on the device
    const int *list[3];         // per axis: triangles sorted by box centre inside every node's range
    const int *nodeid;          // per position: the slot of the active node whose range holds it, -1 = finished (a single triangle)
    const int *aL, *aR;         // per slot: the node's range
};

// (selected, not indexed: an indexed member of a kernel parameter is copied to local memory first)
RM_SHD const int *list_of(const Level &V, int a) { return a == 0 ? V.list[0] : a == 1 ? V.list[1] : V.list[2]; }

RM_SHD float centre_key(const float4 &lo, const float4 &hi, int axis) {
    return axis == 0 ? lo.x + hi.x : axis == 1 ? lo.y + hi.y : lo.z + hi.z;
}

// the box of triangle t as one 32-byte record
RM_SHD void load_box(const Level &V, int t, float4 &lo, float4 &hi) {
#ifdef __CUDA_ARCH__
    const float4 *p = V.tbox + 2 * size_t(t);
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
#else
    lo = V.tbox[2 * size_t(t)];
    hi = V.tbox[2 * size_t(t) + 1];
#endif
}

// The scan input is six rows of n items: row = 2 * axis + direction, item i of a forward row is position i of the axis's list, item
// i of a backward row is position n - 1 - i.  `list` = list_of(V, axis).
RM_SHD SweepItem sweep_item_at(const Level &V, const int *list, bool rev, int i) {
    const int pos = rev ? V.n - 1 - i : i;
    const int nd = V.nodeid[pos];
    SweepItem it;
    if (nd < 0) {               // a finished position is a segment of its own whose prefix nobody reads: no need to fetch its box
        it.lx = it.ly = it.lz = it.hx = it.hy = it.hz = 0.0f; it.flag = 1; it.pad = 0;
        return it;
    }
    const int t = list[pos];
    int flag = 1;
    if (i > 0) flag = V.nodeid[rev ? pos + 1 : pos - 1] != nd;
    float4 lo, hi;
    load_box(V, t, lo, hi);
    it.lx = lo.x; it.ly = lo.y; it.lz = lo.z; it.flag = flag;
    it.hx = hi.x; it.hy = hi.y; it.hz = hi.z; it.pad = 0;
    return it;
}

// item `idx` of the scan input as one sequence, idx in [0, 6 n): the rows one after the other
RM_SHD SweepItem sweep_item(const Level &V, int idx) {
    const int row = idx / V.n;
    return sweep_item_at(V, list_of(V, row >> 1), (row & 1) != 0, idx - row * V.n);
}

// candidate c in [0, 3 n): axis a = c / n, split after position i = c % n.  `areas` = sweep_area of the scan's output, same
// indexing as sweep_item.  Returns the node's slot (-1: no candidate here) and the key whose minimum picks the split.
RM_SHD int sweep_candidate(const Level &V, const float *areas, int c, unsigned long long *key) {
    const int n = V.n, a = c / n, i = c - a * n;
    const int nd = V.nodeid[i];
    if (nd < 0) return -1;
    const int L = V.aL[nd], R = V.aR[nd];
    if (i >= R - 1) return -1;
    const int nl = i - L + 1, nr = R - 1 - i, ns = R - L;
    const float fa = areas[size_t(a) * 2 * n + i], ra = areas[size_t(a) * 2 * n + n + (n - 2 - i)];
    float cost = fa * float(nl) + ra * float(nr);
    if (!(cost >= 0.0f)) cost = INFINITY;
    const int d = nl - ns / 2;
    const unsigned zig = d >= 0 ? unsigned(2 * d) : unsigned(-2 * d - 1);
    unsigned bits;
    {
        union { float f; unsigned u; } cv;
        cv.f = cost;
        bits = cv.u;
    }
    *key = (static_cast<unsigned long long>(bits) << 32) | (static_cast<unsigned long long>(zig) << 2) | unsigned(a);
    return nd;
}

RM_SHD int fetch_add(int *p, int v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    const int o = *p;
    *p += v;
    return o;
#endif
}

struct Tree {             // the binary tree: nodes [0, n) are the triangles (left = ~triangle, right = -1, count 1), inner nodes follow
    float4 *lo, *hi;
    int *left, *right, *count;
};

struct Split {            // what a level decides, per slot
    int *axis, *M;        // the list that is cut, and where: [L, M) goes left
    int *childL, *childR; // next-level slots of the children, -1 = a single triangle (finished)
};

struct NextLevel {
    int *aL, *aR, *aB;               // per slot: range, node index
    unsigned long long *best;        // per slot: the minimum of the candidate keys, reset here
    int *next_node, *next_slot;      // counters
};

// smallest k with 3 * 2^k >= ns: the levels a subtree of ns triangles still needs if every split is even (leaves of <= 3)
RM_SHD int levels_needed(int ns) {
    int k = 0;
    long long cap = 3;
    while (cap < ns) { cap *= 2; k++; }
    return k;
}

// slot s of this level (a node of >= 2 triangles): choose the split, create the children
RM_SHD void sweep_decide(const Level &V, const int *aB, int s, unsigned long long key, int level, int depth_cap, const Tree &T, const Split &S, const NextLevel &X) {
    const int L = V.aL[s], R = V.aR[s], ns = R - L, b = aB[s];
    int axis, nl;
    if (key != kNoSplit && level + levels_needed(ns) < depth_cap) {
        axis = int(key & 3);
        const unsigned zig = unsigned(key >> 2) & 0x3fffffffu;
        const int d = (zig & 1) ? -int((zig + 1) >> 1) : int(zig >> 1);
        nl = ns / 2 + d;
    } else {                    // median of the widest axis (by box centres: the lists are sorted by them)
        axis = 0;
        float best = -1.0f;
        for (int a = 0; a < 3; a++) {
            const int t0 = list_of(V, a)[L], t1 = list_of(V, a)[R - 1];
            const float e = centre_key(V.tlo[t1], V.thi[t1], a) - centre_key(V.tlo[t0], V.thi[t0], a);
            if (e > best) { best = e; axis = a; }
        }
        nl = ns / 2;
    }
    if (nl < 1) nl = 1;
    if (nl > ns - 1) nl = ns - 1;
    const int M = L + nl;
    S.axis[s] = axis;
    S.M[s] = M;
    int kid[2], slot[2];
    for (int k = 0; k < 2; k++) {
        const int cl = k ? M : L, cr = k ? R : M;
        if (cr - cl == 1) { kid[k] = list_of(V, axis)[cl]; slot[k] = -1; continue; }
        kid[k] = fetch_add(X.next_node, 1);
        slot[k] = fetch_add(X.next_slot, 1);
        X.aL[slot[k]] = cl;
        X.aR[slot[k]] = cr;
        X.aB[slot[k]] = kid[k];
        X.best[slot[k]] = kNoSplit;
    }
    T.left[b] = kid[0];
    T.right[b] = kid[1];
    T.count[b] = ns;
    S.childL[s] = slot[0];
    S.childR[s] = slot[1];
}

// position i: which side does the triangle at list[axis of its node][i] go to
RM_SHD void sweep_mark(const Level &V, const Split &S, int i, uint8_t *side) {
    const int nd = V.nodeid[i];
    if (nd < 0) return;
    side[list_of(V, S.axis[nd])[i]] = i >= S.M[nd] ? 1 : 0;
}

// c in [0, 3 n): 1 if the triangle at position c % n of list c / n belongs to an active node and goes left
RM_SHD int sweep_goes_left(const Level &V, const uint8_t *side, int c) {
    const int a = c / V.n, i = c - a * V.n;
    return V.nodeid[i] >= 0 && side[list_of(V, a)[i]] == 0;
}

// c in [0, 3 n): move the triangle at position c % n of list c / n to its place in the next level's list; `zeros` = the
// exclusive sum of sweep_goes_left over [0, 3 n)
RM_SHD void sweep_scatter(const Level &V, const Split &S, const uint8_t *side, const int *zeros, int c, int *const out_list[3], int *out_nodeid) {
    const int n = V.n, a = c / n, i = c - a * n;
    const int nd = V.nodeid[i], t = list_of(V, a)[i];
    int *const out = a == 0 ? out_list[0] : a == 1 ? out_list[1] : out_list[2];
    if (nd < 0) {
        out[i] = t;
        if (a == 0) out_nodeid[i] = -1;
        return;
    }
    const int L = V.aL[nd], M = S.M[nd];
    const int zl = zeros[c] - zeros[size_t(a) * n + L];
    int p, child;
    if (side[t] == 0) { p = L + zl; child = S.childL[nd]; }
    else { p = M + (i - L - zl); child = S.childR[nd]; }
    out[p] = t;
    if (a == 0) out_nodeid[p] = child;
}

// inner node b after its children: box = union of theirs
RM_SHD void sweep_refit(const Tree &T, int b) {
    const int l = T.left[b], r = T.right[b];
    const float4 alo = T.lo[l], ahi = T.hi[l], blo = T.lo[r], bhi = T.hi[r];
    float4 lo, hi;
    lo.x = fminf(alo.x, blo.x); lo.y = fminf(alo.y, blo.y); lo.z = fminf(alo.z, blo.z); lo.w = 0.0f;
    hi.x = fmaxf(ahi.x, bhi.x); hi.y = fmaxf(ahi.y, bhi.y); hi.z = fmaxf(ahi.z, bhi.z); hi.w = 0.0f;
    T.lo[b] = lo;
    T.hi[b] = hi;
}

} // namespace rm_sah
