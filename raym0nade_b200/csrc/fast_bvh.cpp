// fast_bvh.cpp — a second BVH over the same triangles for SECONDARY rays (bounce and shadow rays of the estimator).
//
// Primary hits, and the per-ray seam (rm_trace_closest / rm_trace_occluded), always traverse the reference's own tree
// in the reference's own order: their triangle index is a bit-exact contract.  The rays of the estimator are held to
// the north star's statistical bar instead, so they may use a better tree as long as the hit they find is the closest
// accepted triangle under the same box and triangle tests.  This builder makes that tree:
// binned SAH (32 bins per axis), leaves of at most 3 triangles (the reference: object median, 5-10 per leaf), depth capped
// so the traversal stack still fits 8 resident CTAs per SM.  On the bench scene a bounce ray then tests ~85 boxes and ~10
// triangles instead of 98 and 36 (bench.py reports the device counters of both trees).
//
// Output, in the layout the traversal engine already reads (dev_trace.cuh): an array of 32-byte node records in PAIRS -
// block b = records 2b, 2b+1 = the two children of one inner node, block 1 = the children of the root.  A leaf record
// carries [faceL, faceR) into `order` (the builder's own triangle order); an inner record carries faceR = 0 and, in
// faceL, the index of the block that holds its children.  The whole-tree record sits at index 1 (root-is-leaf scenes).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "rm_internal.h"
#include "raym0nade_b200.h"

namespace {

struct Aabb {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    void add(const float *p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void add(const Aabb &b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const { float d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]}; return 2.0f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]); }
};

constexpr int kBins = 32;
inline int bin_of(float c, float lo, float scale) {
    const float f = (c - lo) * scale;
    return f > 0.0f ? (f < float(kBins) ? int(f) : kBins - 1) : 0;      // also maps NaN to bin 0
}

struct Builder {
    const float *pos;                 // [n][9]
    std::vector<Aabb> tb;             // triangle boxes
    std::vector<float> cen;           // [n][3] centroids
    std::vector<int32_t> &order;
    std::vector<RmBvhNode> &nodes;
    std::atomic<int> next_block{2};   // block 0 = {unused, whole-tree record}, block 1 = the root's children
    std::atomic<int> max_depth{0};
    int depth_cap, kLeafMax;
    const std::atomic<bool> *cancel = nullptr;        // set by the owner of a background build: unwind, the tree is not wanted any more

    Builder(const float *p, int n, std::vector<int32_t> &o, std::vector<RmBvhNode> &nd, int cap, int leaf_max)
        : pos(p), order(o), nodes(nd), depth_cap(cap), kLeafMax(leaf_max) {
        tb.resize(n);
        cen.resize(size_t(n) * 3);
        for (int i = 0; i < n; i++) {
            const float *t = pos + size_t(i) * 9;
            Aabb b;
            b.add(t); b.add(t + 3); b.add(t + 6);
            tb[i] = b;
            for (int a = 0; a < 3; a++) {
                const float c = (t[a] + t[3 + a] + t[6 + a]) * (1.0f / 3.0f);
                cen[size_t(i) * 3 + a] = std::isfinite(c) ? c : 0.0f;       // a non-finite triangle must not derail the sort / binning (it can never be hit)
            }
        }
    }

    static void put_box(RmBvhNode &r, const Aabb &b) {
        for (int a = 0; a < 3; a++) { r.v0[a] = b.lo[a]; r.v1[a] = b.hi[a]; }
    }

    // fills `rec` for the triangles order[L, R); its children go into pair block `block` (-1: a freshly allocated one)
    void build(RmBvhNode &rec, int L, int R, int depth, int spawn_levels, int block = -1) {
        if (cancel && cancel->load(std::memory_order_relaxed)) return;
        int seen = max_depth.load();
        while (depth > seen && !max_depth.compare_exchange_weak(seen, depth)) {}
        // big nodes (the top few levels) split their two passes over the triangles across threads
        const int n = R - L;
        const int nt = n > 200000 ? std::min<int>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
        auto chunks = [&](auto &&fn) {                 // fn(thread, begin, end)
            if (nt == 1) { fn(0, L, R); return; }
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back([&, t] { fn(t, L + int(int64_t(n) * t / nt), L + int(int64_t(n) * (t + 1) / nt)); });
            for (auto &x : th) x.join();
        };
        Aabb box, cbox;
        {
            std::vector<Aabb> pb(nt), pc(nt);
            chunks([&](int t, int b, int e) { for (int i = b; i < e; i++) { pb[t].add(tb[order[i]]); pc[t].add(&cen[size_t(order[i]) * 3]); } });
            for (int t = 0; t < nt; t++) { box.add(pb[t]); cbox.add(pc[t]); }
        }
        put_box(rec, box);
        if (n <= kLeafMax) { rec.faceL = L; rec.faceR = R; return; }
        // a subtree of n triangles needs ceil(log2(n / leaf)) more levels if split evenly: SAH is free only while that fits the cap
        const int need = int(std::ceil(std::log2(double(n) / kLeafMax)));
        int M = -1;
        if (depth + need < depth_cap) {
            struct Bins { Aabb bb[3][kBins]; int cnt[3][kBins]; };
            std::vector<Bins> part(nt);
            float lo3[3], scale3[3];
            bool use[3];
            for (int ax = 0; ax < 3; ax++) {
                const float ext = cbox.hi[ax] - cbox.lo[ax];
                use[ax] = ext > 0.0f;
                lo3[ax] = cbox.lo[ax];
                scale3[ax] = use[ax] ? float(kBins) / ext : 0.0f;
            }
            chunks([&](int t, int b, int e) {
                Bins &P = part[t];
                std::memset(P.cnt, 0, sizeof(P.cnt));
                for (int i = b; i < e; i++) {
                    const int tr = order[i];
                    for (int ax = 0; ax < 3; ax++) {
                        if (!use[ax]) continue;
                        const int k = bin_of(cen[size_t(tr) * 3 + ax], lo3[ax], scale3[ax]);
                        P.bb[ax][k].add(tb[tr]);
                        P.cnt[ax][k]++;
                    }
                }
            });
            float best = INFINITY;
            int best_axis = -1, best_bin = -1;
            for (int ax = 0; ax < 3; ax++) {
                if (!use[ax]) continue;
                Aabb bb[kBins];
                int cnt[kBins] = {0};
                for (int t = 0; t < nt; t++)
                    for (int k = 0; k < kBins; k++)
                        if (part[t].cnt[ax][k]) { bb[k].add(part[t].bb[ax][k]); cnt[k] += part[t].cnt[ax][k]; }
                float la[kBins];
                int lc[kBins];
                Aabb acc;
                int c = 0;
                for (int k = 0; k < kBins; k++) { if (cnt[k]) acc.add(bb[k]); c += cnt[k]; la[k] = c ? acc.area() : 0.0f; lc[k] = c; }
                acc = Aabb();
                c = 0;
                for (int k = kBins - 1; k > 0; k--) {
                    if (cnt[k]) acc.add(bb[k]);
                    c += cnt[k];
                    if (!c || !lc[k - 1]) continue;
                    const float cost = la[k - 1] * float(lc[k - 1]) + acc.area() * float(c);
                    if (cost < best) { best = cost; best_axis = ax; best_bin = k - 1; }
                }
            }
            if (best_axis >= 0) {
                const float lo = cbox.lo[best_axis], scale = float(kBins) / (cbox.hi[best_axis] - cbox.lo[best_axis]);
                auto mid = std::partition(order.begin() + L, order.begin() + R, [&](int32_t t) {
                    return bin_of(cen[size_t(t) * 3 + best_axis], lo, scale) <= best_bin;
                });
                M = int(mid - order.begin());
                if (M == L || M == R) M = -1;
            }
        }
        if (M < 0) {        // object median along the widest centroid axis
            int ax = 0;
            float ext[3] = {cbox.hi[0] - cbox.lo[0], cbox.hi[1] - cbox.lo[1], cbox.hi[2] - cbox.lo[2]};
            if (ext[1] > ext[ax]) ax = 1;
            if (ext[2] > ext[ax]) ax = 2;
            M = (L + R) / 2;
            std::nth_element(order.begin() + L, order.begin() + M, order.begin() + R,
                             [&](int32_t a, int32_t b) { return cen[size_t(a) * 3 + ax] < cen[size_t(b) * 3 + ax]; });
        }
        if (block < 0) block = next_block.fetch_add(1);
        rec.faceL = block;
        rec.faceR = 0;
        RmBvhNode &c0 = nodes[size_t(block) * 2], &c1 = nodes[size_t(block) * 2 + 1];
        if (spawn_levels > 0 && n > 50000) {
            std::thread th([&, L, M, depth, spawn_levels] { build(c0, L, M, depth + 1, spawn_levels - 1); });
            build(c1, M, R, depth + 1, spawn_levels - 1);
            th.join();
        } else {
            build(c0, L, M, depth + 1, 0);
            build(c1, M, R, depth + 1, 0);
        }
    }
};

} // namespace

// positions [n][9]; returns the node records (pairs), the triangle order and the tree depth (levels of inner nodes + 1)
// (cancel: set by the owner of a background build to make it unwind; the result is then RM_ERR_STATE and no tree)
int rm_build_fast_bvh_cancellable(const float *positions, int n, int depth_cap, int leaf_max, std::vector<RmBvhNode> &nodes, std::vector<int32_t> &order, int *depth_out,
                                  const std::atomic<bool> *cancel) {
    const int kLeafMax = std::min(std::max(leaf_max, 1), 15);          // a leaf reference holds its count in 4 bits (dev_trace.cuh)
    if (!positions || n <= 0) return rm_fail(RM_ERR_INVALID, "rm_build_fast_bvh: no triangles");
    order.resize(n);
    for (int i = 0; i < n; i++) order[i] = i;
    nodes.assign(size_t(n + 2) * 2, RmBvhNode{});
    const int min_cap = int(std::ceil(std::log2(std::max(1.0, double(n) / kLeafMax)))) + 1;
    Builder B(positions, n, order, nodes, std::max(depth_cap, min_cap), kLeafMax);
    B.cancel = cancel;
    RmBvhNode whole{};
    B.build(whole, 0, n, 0, 4, /*the root's children are block 1, where the engine starts*/ 1);
    nodes[1] = whole;          // only read when the whole scene is one leaf (root_is_leaf)
    nodes.resize(size_t(B.next_block.load()) * 2);
    if (depth_out) *depth_out = B.max_depth.load() + 1;
    if (cancel && cancel->load()) return rm_fail(RM_ERR_STATE, "rm_build_fast_bvh: cancelled");
    return RM_OK;
}

int rm_build_fast_bvh(const float *positions, int n, int depth_cap, int leaf_max, std::vector<RmBvhNode> &nodes, std::vector<int32_t> &order, int *depth_out) {
    return rm_build_fast_bvh_cancellable(positions, n, depth_cap, leaf_max, nodes, order, depth_out, nullptr);
}

// Host-only diagnostic behind the C ABI (no GPU needed): builds the secondary-ray tree for `positions` and checks its
// invariants - every triangle in exactly one leaf, every leaf's box encloses its triangles and every inner box its
// children, links in range, leaves within leaf_max, depth within the cap - before it reports the shape.
extern "C" int rm_secondary_tree_stats(const float *positions, int32_t n, int32_t depth_cap, int32_t leaf_max, int32_t out[4]) {
    std::vector<RmBvhNode> nodes;
    std::vector<int32_t> order;
    int depth = 0;
    int rc = rm_build_fast_bvh(positions, n, depth_cap, leaf_max, nodes, order, &depth);
    if (rc) return rc;
    std::vector<uint8_t> seen(size_t(n), 0);
    int leaves = 0, biggest = 0;
    auto inside = [](const RmBvhNode &outer, const float *lo, const float *hi) {
        for (int a = 0; a < 3; a++) if (lo[a] < outer.v0[a] || hi[a] > outer.v1[a]) return false;
        return true;
    };
    auto check_leaf = [&](const RmBvhNode &c) {
        const int cnt = c.faceR - c.faceL;
        if (cnt < 1 || cnt > std::min(std::max(leaf_max, 1), 15) || c.faceL < 0 || c.faceR > n) return false;
        for (int i = c.faceL; i < c.faceR; i++) {
            const int t = order[i];
            if (t < 0 || t >= n || seen[t]) return false;
            seen[t] = 1;
            for (int v = 0; v < 3; v++) { const float *p = positions + size_t(t) * 9 + v * 3; if (!inside(c, p, p)) return false; }
        }
        leaves++;
        biggest = std::max(biggest, cnt);
        return true;
    };
    bool ok = true;
    if (nodes[1].faceR) ok = check_leaf(nodes[1]);
    else {
        std::vector<std::pair<int, int>> stack = {{1, 1}};           // (record index of the parent, its children's block)
        if (nodes[1].faceL != 1) ok = false;
        while (ok && !stack.empty()) {
            auto [parent, block] = stack.back();
            stack.pop_back();
            for (int k = 0; k < 2 && ok; k++) {
                const size_t ci = size_t(block) * 2 + k;
                if (ci >= nodes.size()) { ok = false; break; }
                const RmBvhNode &c = nodes[ci];
                if (!inside(nodes[parent], c.v0, c.v1)) ok = false;
                else if (c.faceR) ok = check_leaf(c);
                else if (c.faceL < 2 || size_t(c.faceL) * 2 + 1 >= nodes.size()) ok = false;
                else stack.push_back({int(ci), c.faceL});
            }
        }
    }
    for (int i = 0; ok && i < n; i++) ok = seen[i] != 0;
    if (!ok) return rm_fail(RM_ERR_STATE, "rm_secondary_tree_stats: the tree violates an invariant");
    if (out) { out[0] = int32_t(nodes.size() / 2); out[1] = depth; out[2] = leaves; out[3] = biggest; }
    return RM_OK;
}
