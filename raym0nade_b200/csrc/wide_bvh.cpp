// wide_bvh.cpp — the secondary-ray tree in its traversal form: 4-wide nodes with 8-bit quantised child boxes.
//
// fast_bvh.cpp builds a binned-SAH binary tree over the scene's triangles.  The traversal engine of the estimator's
// bounce and shadow rays (dev_trace4.cuh) is bound by instruction issue and by the L1 data pipe - one wavefront per lane
// and load, the rays being incoherent - not by DRAM, so what it needs is FEWER, FATTER steps: this file collapses the
// binary tree into nodes of up to four children and stores each node in ONE 64-byte record (two 256-bit loads per lane and
// step instead of two per binary node, for about half the steps per ray):
//
//   bytes  0..11  o[3]        the node's lower corner (origin of the quantisation grid)
//         12..23  s[3]        grid step per axis: (extent / 255) rounded up
//         24..35  qlo[3][4]   per axis, per child: lower plane in grid units, rounded DOWN
//         36..47  qhi[3][4]   per axis, per child: upper plane in grid units, rounded UP
//         48..51  child_base  index of the node's first inner child (its inner children are consecutive records)
//         52..55  tri_base    first triangle of the node's leaf children (consecutive in the tree's triangle order)
//         56..59  meta[4]     per child: 0 = empty slot, 0x80 | k = the k-th inner child, (offset << 2) | count = a leaf of
//                             `count` (1..3) triangles starting `offset` (< 32) after tri_base
//         60..63  unused
//
// A decoded box o + s * q always ENCLOSES the child's true box (checked here in double precision), so a ray that meets a
// triangle meets the boxes above it: the tree returns the same closest accepted triangle as any other valid hierarchy
// over the same triangles under the same triangle test (the reference's RayTriangleIntersection, src/geometry.cpp:63-87).
// Only the work per ray differs.  There is no reference counterpart: the reference's tree (src/bvh.cpp:18-54) is what
// primary rays and the per-ray seam traverse, in its own order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "rm_internal.h"
#include "raym0nade_b200.h"
#include "wide_bvh.h"

namespace {

struct Box3 { float lo[3], hi[3]; };

inline Box3 box_of(const RmBvhNode &r) { return {{r.v0[0], r.v0[1], r.v0[2]}, {r.v1[0], r.v1[1], r.v1[2]}}; }
inline float half_area(const Box3 &b) {
    const float d[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]};
    return d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
}

struct Collapser {
    const std::vector<RmBvhNode> &bin;      // the binary tree: pair blocks (fast_bvh.cpp)
    const std::vector<int32_t> &order_in;   // its triangle order
    std::vector<RmWideNode> &out;
    std::vector<int32_t> &order_out;
    int max_depth = 0;
    bool ok = true;

    // fills out[self] for the binary inner record `rec` (children in pair block rec.faceL)
    void emit(int self, const RmBvhNode &rec, int depth) {
        max_depth = std::max(max_depth, depth);
        // gather up to four children: open the inner child with the largest surface until four are held or none is inner
        const RmBvhNode *ch[4];
        int n = 0;
        ch[n++] = &bin[size_t(rec.faceL) * 2];
        ch[n++] = &bin[size_t(rec.faceL) * 2 + 1];
        while (n < 4) {
            int best = -1;
            float best_area = -1.0f;
            for (int i = 0; i < n; i++)
                if (ch[i]->faceR == 0) {
                    const float a = half_area(box_of(*ch[i]));
                    if (a > best_area || best < 0) { best_area = a; best = i; }
                }
            if (best < 0) break;
            const RmBvhNode *open = ch[best];
            ch[best] = &bin[size_t(open->faceL) * 2];
            ch[n++] = &bin[size_t(open->faceL) * 2 + 1];
        }
        fill(self, ch, n, depth);
    }

    void fill(int self, const RmBvhNode *const *ch, int n, int depth) {
        RmWideNode w;
        std::memset(&w, 0, sizeof(w));
        float lo[4][3], hi[4][3];
        for (int i = 0; i < n; i++)
            for (int a = 0; a < 3; a++) { lo[i][a] = ch[i]->v0[a]; hi[i][a] = ch[i]->v1[a]; }
        if (!wide_quantise(lo, hi, n, w)) ok = false;
        int n_inner = 0, n_tris = 0;
        for (int i = 0; i < n; i++) (ch[i]->faceR == 0 ? n_inner : n_tris) += ch[i]->faceR == 0 ? 1 : ch[i]->faceR - ch[i]->faceL;
        w.child_base = n_inner ? int32_t(out.size()) : 0;
        w.tri_base = int32_t(order_out.size());
        if (n_inner) out.resize(out.size() + n_inner);         // the inner children's records: consecutive
        int k_inner = 0, tri_off = 0;
        const RmBvhNode *inner[4];
        for (int i = 0; i < n; i++) {
            const RmBvhNode &c = *ch[i];
            if (c.faceR == 0) {
                w.meta[i] = uint8_t(0x80 | k_inner);
                inner[k_inner++] = &c;
            } else {
                const int cnt = c.faceR - c.faceL;
                if (cnt < 1 || cnt > 3 || tri_off > 31) { ok = false; continue; }
                w.meta[i] = uint8_t((tri_off << 2) | cnt);
                for (int t = c.faceL; t < c.faceR; t++) order_out.push_back(order_in[t]);
                tri_off += cnt;
            }
        }
        const int base = w.child_base;
        out[self] = w;
        for (int k = 0; k < k_inner; k++) emit(base + k, *inner[k], depth + 1);
    }
};

} // namespace

int rm_build_wide_bvh(const std::vector<RmBvhNode> &bin, const std::vector<int32_t> &order_in, int n_tris, std::vector<RmWideNode> &out,
                      std::vector<int32_t> &order_out, int *depth_out) {
    if (bin.size() < 2 || n_tris <= 0 || int(order_in.size()) != n_tris) return rm_fail(RM_ERR_INVALID, "rm_build_wide_bvh: no tree");
    out.clear();
    order_out.clear();
    out.reserve(size_t(n_tris));
    order_out.reserve(size_t(n_tris));
    Collapser C{bin, order_in, out, order_out};
    out.resize(1);
    const RmBvhNode &whole = bin[1];
    if (whole.faceR != 0) {                    // the whole scene is one leaf: a root with that single leaf child
        if (whole.faceR - whole.faceL > 3) return rm_fail(RM_ERR_INVALID, "rm_build_wide_bvh: leaves hold at most 3 triangles");
        const RmBvhNode *one[1] = {&whole};
        C.fill(0, one, 1, 0);
    } else C.emit(0, whole, 0);
    if (!C.ok || int(order_out.size()) != n_tris) return rm_fail(RM_ERR_STATE, "rm_build_wide_bvh: the collapse lost a triangle or a box");
    if (depth_out) *depth_out = C.max_depth + 1;
    return RM_OK;
}

// Host-only diagnostic behind the C ABI (no GPU needed): builds the secondary-ray tree for `positions`, collapses it and
// checks the 4-wide form - every triangle in exactly one leaf, every child's DECODED box (o + s * q, evaluated in double)
// encloses every triangle vertex beneath that child, links in range, leaves of 1..3 triangles - before it reports the shape.
int rm_build_fast_bvh(const float *positions, int n, int depth_cap, int leaf_max, std::vector<RmBvhNode> &nodes, std::vector<int32_t> &order, int *depth_out);

extern "C" int rm_wide_tree_stats(const float *positions, int32_t n, int32_t depth_cap, int32_t out[4]) {
    std::vector<RmBvhNode> bin;
    std::vector<int32_t> order, worder;
    std::vector<RmWideNode> w;
    int depth = 0, wdepth = 0;
    int rc = rm_build_fast_bvh(positions, n, depth_cap, 3, bin, order, &depth);
    if (rc) return rc;
    if ((rc = rm_build_wide_bvh(bin, order, n, w, worder, &wdepth))) return rc;
    std::vector<uint8_t> seen(size_t(n), 0);
    struct Item { int node; double lo[3], hi[3]; };        // the decoded box every vertex beneath `node` must lie in
    std::vector<Item> stack;
    Item root{0, {-INFINITY, -INFINITY, -INFINITY}, {INFINITY, INFINITY, INFINITY}};
    stack.push_back(root);
    long long children = 0, leaves = 0;
    bool ok = true;
    while (ok && !stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        if (it.node < 0 || size_t(it.node) >= w.size()) { ok = false; break; }
        const RmWideNode &nd = w[it.node];
        for (int c = 0; c < 4 && ok; c++) {
            const uint8_t m = nd.meta[c];
            if (!m) continue;
            children++;
            Item ch;
            for (int a = 0; a < 3; a++) {
                ch.lo[a] = std::max(it.lo[a], double(nd.o[a]) + double(nd.s[a]) * nd.qlo[a][c]);
                ch.hi[a] = std::min(it.hi[a], double(nd.o[a]) + double(nd.s[a]) * nd.qhi[a][c]);
            }
            if (m & 0x80) {
                ch.node = nd.child_base + (m & 0x7f);
                if (ch.node <= it.node) ok = false;            // children are emitted after their parent
                stack.push_back(ch);
            } else {
                const int cnt = m & 3, first = nd.tri_base + (m >> 2);
                if (cnt < 1 || first < 0 || first + cnt > n) { ok = false; break; }
                leaves++;
                for (int k = first; k < first + cnt && ok; k++) {
                    const int t = worder[k];
                    if (t < 0 || t >= n || seen[t]) { ok = false; break; }
                    seen[t] = 1;
                    for (int v = 0; v < 3 && ok; v++)
                        for (int a = 0; a < 3; a++) {
                            const double x = positions[size_t(t) * 9 + v * 3 + a];
                            if (x == x && (x < ch.lo[a] || x > ch.hi[a])) ok = false;      // (a NaN vertex can never be hit)
                        }
                }
            }
        }
    }
    for (int i = 0; ok && i < n; i++) ok = seen[i] != 0;
    if (!ok) return rm_fail(RM_ERR_STATE, "rm_wide_tree_stats: the 4-wide tree violates an invariant");
    if (out) { out[0] = int32_t(w.size()); out[1] = wdepth; out[2] = int32_t(leaves); out[3] = int32_t(children * 100 / std::max<size_t>(w.size(), 1)); }
    return RM_OK;
}
