// scripts/ploc_proto.cpp - offline quality probe (python scripts/ploc_proto_gen.py glossy 1000000 writes /tmp/ploc/tris.bin, rays.bin;
// g++ -O2 -std=c++17 -Iinclude -Iraym0nade_b200/csrc scripts/ploc_proto.cpp raym0nade_b200/csrc/{fast_bvh,wide_bvh,rm_error}.cpp -lpthread): wide-tree node visits / triangle tests per ray for (a) the host binned-SAH tree + collapse (the product's code),
// (b) a sequential emulation of the device PLOC builder + collapse.  Same wide traversal as dev_trace.cuh (near-first, stack).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "rm_internal.h"
#include "raym0nade_b200.h"
#include "wide_bvh.h"
int rm_build_fast_bvh(const float *positions, int n, int depth_cap, int leaf_max, std::vector<RmBvhNode> &nodes, std::vector<int32_t> &order, int *depth_out);

static std::vector<float> pos, rays;
struct B3 { float lo[3], hi[3]; };
static inline float uarea(const B3 &a, const B3 &b) { float d[3]; for (int k = 0; k < 3; k++) d[k] = std::max(a.hi[k], b.hi[k]) - std::min(a.lo[k], b.lo[k]); return d[0]*d[1] + d[1]*d[2] + d[2]*d[0]; }
static inline float area(const B3 &a) { float d[3]; for (int k = 0; k < 3; k++) d[k] = a.hi[k] - a.lo[k]; return d[0]*d[1] + d[1]*d[2] + d[2]*d[0]; }
static uint64_t spread21(uint32_t v) { uint64_t x = v & 0x1fffffu; x = (x | x << 32) & 0x1f00000000ffffull; x = (x | x << 16) & 0x1f0000ff0000ffull; x = (x | x << 8) & 0x100f00f00f00f00full; x = (x | x << 4) & 0x10c30c30c30c30c3ull; x = (x | x << 2) & 0x1249249249249249ull; return x; }

struct Bin { std::vector<B3> box; std::vector<int> left, right, count; int root; };
static int gLeafMax = 3;
static int METRIC = 0;     // 0 area of union
static int TOPK = 0;       // > 0: clustering stops at <= TOPK clusters, a top-down sweep-SAH tree over the clusters finishes the job
static int TOPW = 0;       // weight of a cluster in the top SAH: 0 triangles beneath, 1 one per cluster, 2 its own SAH cost estimate (area-independent: count^0.5?)
static int top_sah(Bin &T, std::vector<int> &ids, int L, int R, int &next) {
    if (R - L == 1) return ids[L];
    int n = R - L; float best = INFINITY; int bax = -1, bk = -1;
    std::vector<float> la(n); std::vector<double> lw(n);
    auto wt = [&](int c) { return TOPW == 0 ? double(T.count[c]) : TOPW == 1 ? 1.0 : std::sqrt(double(T.count[c])); };
    for (int ax = 0; ax < 3; ax++) {
        std::sort(ids.begin() + L, ids.begin() + R, [&](int a, int b) { return T.box[a].lo[ax] + T.box[a].hi[ax] < T.box[b].lo[ax] + T.box[b].hi[ax]; });
        B3 acc{{1e30f,1e30f,1e30f},{-1e30f,-1e30f,-1e30f}}; double w = 0;
        for (int i = 0; i < n - 1; i++) { const B3 &b = T.box[ids[L + i]]; for (int q = 0; q < 3; q++) { acc.lo[q] = std::min(acc.lo[q], b.lo[q]); acc.hi[q] = std::max(acc.hi[q], b.hi[q]); } w += wt(ids[L + i]); la[i] = area(acc); lw[i] = w; }
        acc = B3{{1e30f,1e30f,1e30f},{-1e30f,-1e30f,-1e30f}}; w = 0;
        for (int i = n - 1; i >= 1; i--) { const B3 &b = T.box[ids[L + i]]; for (int q = 0; q < 3; q++) { acc.lo[q] = std::min(acc.lo[q], b.lo[q]); acc.hi[q] = std::max(acc.hi[q], b.hi[q]); } w += wt(ids[L + i]);
            float cost = float(la[i - 1] * lw[i - 1] + area(acc) * w); if (cost < best) { best = cost; bax = ax; bk = i; } }
    }
    std::sort(ids.begin() + L, ids.begin() + R, [&](int a, int b) { return T.box[a].lo[bax] + T.box[a].hi[bax] < T.box[b].lo[bax] + T.box[b].hi[bax]; });
    int M = L + bk;
    int l = top_sah(T, ids, L, M, next), r = top_sah(T, ids, M, R, next);
    int id = next++;
    for (int q = 0; q < 3; q++) { T.box[id].lo[q] = std::min(T.box[l].lo[q], T.box[r].lo[q]); T.box[id].hi[q] = std::max(T.box[l].hi[q], T.box[r].hi[q]); }
    T.left[id] = l; T.right[id] = r; T.count[id] = T.count[l] + T.count[r];
    return id;
}
static Bin ploc(int n, int R) {
    Bin T; T.box.resize(2 * n); T.left.resize(2 * n); T.right.resize(2 * n); T.count.resize(2 * n);
    B3 sb{{1e30f,1e30f,1e30f},{-1e30f,-1e30f,-1e30f}};
    std::vector<B3> tb(n);
    for (int i = 0; i < n; i++) { B3 b{{1e30f,1e30f,1e30f},{-1e30f,-1e30f,-1e30f}}; for (int v = 0; v < 3; v++) for (int a = 0; a < 3; a++) { float x = pos[size_t(i)*9+v*3+a]; b.lo[a] = std::min(b.lo[a], x); b.hi[a] = std::max(b.hi[a], x); } tb[i] = b; for (int a = 0; a < 3; a++) { sb.lo[a] = std::min(sb.lo[a], b.lo[a]); sb.hi[a] = std::max(sb.hi[a], b.hi[a]); } }
    std::vector<std::pair<uint64_t,int>> keys(n);
    for (int i = 0; i < n; i++) { uint64_t k = 0; for (int a = 0; a < 3; a++) { float c = (pos[size_t(i)*9+a] + pos[size_t(i)*9+3+a] + pos[size_t(i)*9+6+a]) * (1.0f/3.0f); float f = (c - sb.lo[a]) * (2097151.0f / (sb.hi[a] - sb.lo[a])); f = std::min(std::max(f, 0.0f), 2097151.0f); k |= spread21(uint32_t(f)) << a; } keys[i] = {k, i}; }
    std::sort(keys.begin(), keys.end());
    std::vector<int> cl(n), nn(n), out(n);
    for (int k = 0; k < n; k++) { int t = keys[k].second; T.box[k] = tb[t]; T.left[k] = ~t; T.right[k] = -1; T.count[k] = 1; cl[k] = k; }
    int next = n, m = n, rounds = 0;
    while (m > 1 && !(TOPK > 0 && m <= TOPK)) {
        for (int i = 0; i < m; i++) { float best = INFINITY; int bj = -1; for (int j = std::max(i - R, 0); j <= std::min(i + R, m - 1); j++) { if (j == i) continue; float a = uarea(T.box[cl[i]], T.box[cl[j]]); if (METRIC == 1) a = a - area(T.box[cl[i]]) - area(T.box[cl[j]]); if (METRIC == 2) a *= float(T.count[cl[i]] + T.count[cl[j]]); if (METRIC == 3) a *= std::sqrt(float(T.count[cl[i]] + T.count[cl[j]])); if (a < best || bj < 0) { best = a; bj = j; } } nn[i] = bj; }
        int k = 0;
        for (int i = 0; i < m; i++) { int j = nn[i]; if (nn[j] != i) { out[k++] = cl[i]; continue; } if (i > j) continue; int id = next++; const B3 &a = T.box[cl[i]], &b = T.box[cl[j]]; for (int q = 0; q < 3; q++) { T.box[id].lo[q] = std::min(a.lo[q], b.lo[q]); T.box[id].hi[q] = std::max(a.hi[q], b.hi[q]); } T.left[id] = cl[i]; T.right[id] = cl[j]; T.count[id] = T.count[cl[i]] + T.count[cl[j]]; out[k++] = id; }
        m = k; std::swap(cl, out); rounds++;
    }
    if (m > 1) { fprintf(stderr, "  top SAH over %d clusters\n", m); std::vector<int> ids(cl.begin(), cl.begin() + m); auto t0 = std::chrono::steady_clock::now(); T.root = top_sah(T, ids, 0, m, next); fprintf(stderr, "  top build %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); } else
    T.root = cl[0];
    { double ci = 0, cl_ = 0; float ra = area(T.box[T.root]); std::vector<int> stk{T.root}; int maxd = 0; std::vector<int> dep{0};
      while (!stk.empty()) { int b = stk.back(); stk.pop_back(); int dd = dep.back(); dep.pop_back(); maxd = std::max(maxd, dd); if (T.count[b] <= gLeafMax) { cl_ += area(T.box[b]) / ra * T.count[b]; continue; } ci += area(T.box[b]) / ra; stk.push_back(T.left[b]); dep.push_back(dd + 1); stk.push_back(T.right[b]); dep.push_back(dd + 1); }
      fprintf(stderr, "  binary SAH cost: inner %.2f leaf %.2f depth %d\n", ci, cl_, maxd); }
    fprintf(stderr, "ploc R=%d rounds %d\n", R, rounds);
    return T;
}
static int small_sub(const Bin &N, int b, int tri[8]) { if (N.right[b] < 0) { tri[0] = ~N.left[b]; return 1; } int n = 0; n += small_sub(N, N.left[b], tri + n); n += small_sub(N, N.right[b], tri + n); return n; }
static int COLLAPSE_MODE = 0;
static int EXACT_BOXES = 0;          // 1: the traversal tests the children's true boxes instead of the decoded 8-bit ones (what does the grid cost?)
static std::vector<float> gExact;    // [node][child][6]
static void collapse(const Bin &N, std::vector<RmWideNode> &W, std::vector<int> &order, int *levels) {
    struct It { int bin, slot; };
    std::vector<It> cur{{N.root, 0}}, nxt; W.assign(1, RmWideNode{}); order.clear(); int lv = 0;
    while (!cur.empty()) { lv++; nxt.clear();
        for (auto it : cur) { int ch[4], n = 0;
            if (N.count[it.bin] <= gLeafMax) ch[n++] = it.bin; else { ch[n++] = N.left[it.bin]; ch[n++] = N.right[it.bin];
                while (n < 4) { int open = -1; float oa = -1; for (int k = 0; k < n; k++) if (N.count[ch[k]] > gLeafMax) { float a = area(N.box[ch[k]]); if (COLLAPSE_MODE == 1) a *= float(N.count[ch[k]]); if (COLLAPSE_MODE == 2) a = float(N.count[ch[k]]); if (open < 0 || a > oa) { open = k; oa = a; } } if (open < 0) break; int b = ch[open]; ch[open] = N.left[b]; ch[n++] = N.right[b]; } }
            float lo[4][3], hi[4][3]; int ni = 0, nt = 0;
            for (int k = 0; k < n; k++) { for (int a = 0; a < 3; a++) { lo[k][a] = N.box[ch[k]].lo[a]; hi[k][a] = N.box[ch[k]].hi[a]; } if (N.count[ch[k]] > gLeafMax) ni++; else nt += N.count[ch[k]]; }
            RmWideNode w; memset(&w, 0, sizeof(w)); wide_quantise(lo, hi, n, w);
            if (gExact.size() < (size_t(it.slot) + 1) * 24) gExact.resize((size_t(it.slot) + 1) * 24 * 2); for (int k = 0; k < n; k++) for (int a = 0; a < 3; a++) { gExact[size_t(it.slot) * 24 + k * 6 + a] = lo[k][a]; gExact[size_t(it.slot) * 24 + k * 6 + 3 + a] = hi[k][a]; }
            w.child_base = ni ? int(W.size()) : 0; w.tri_base = int(order.size()); if (ni) W.resize(W.size() + ni);
            int ki = 0, toff = 0;
            for (int k = 0; k < n; k++) { if (N.count[ch[k]] > gLeafMax) { w.meta[k] = uint8_t(0x80 | ki); nxt.push_back({ch[k], w.child_base + ki}); ki++; } else { int tri[8]; int c = small_sub(N, ch[k], tri); w.meta[k] = uint8_t((toff << 2) | c); for (int t = 0; t < c; t++) order.push_back(tri[t]); toff += c; } }
            W[it.slot] = w; }
        cur.swap(nxt); }
    *levels = lv;
}
// SAH-optimal collapse (after Ylitie et al. 2017): T[b][k] = cheapest way to hand subtree b to its wide parent as <= k+1 children
static float C_NODE = 1.0f, C_TRI = 0.37f;
static void dp_collapse(const Bin &N, std::vector<RmWideNode> &W, std::vector<int> &order, int *levels) {
    const int total = int(N.box.size());
    std::vector<float> T(size_t(total) * 4, 0.0f); std::vector<uint8_t> J(size_t(total) * 4, 0);     // J[b][k]: children given to the left subtree (0 = b stays whole)
    // post-order
    std::vector<int> post; post.reserve(total); { std::vector<std::pair<int,int>> st{{N.root, 0}}; while (!st.empty()) { auto [b, ph] = st.back(); st.pop_back(); if (N.count[b] <= gLeafMax) { post.push_back(b); continue; } if (ph == 0) { st.push_back({b, 1}); st.push_back({N.left[b], 0}); st.push_back({N.right[b], 0}); } else post.push_back(b); } }
    float ra = area(N.box[N.root]);
    for (int b : post) {
        float *t = &T[size_t(b) * 4]; uint8_t *j = &J[size_t(b) * 4];
        if (N.count[b] <= gLeafMax) { float c = area(N.box[b]) / ra * N.count[b] * C_TRI; for (int k = 0; k < 4; k++) { t[k] = c; j[k] = 0; } continue; }
        const float *tl = &T[size_t(N.left[b]) * 4], *tr = &T[size_t(N.right[b]) * 4];
        // opened into k+1 pieces (k >= 1): left gets a pieces, right k+1-a
        float open[4]; uint8_t oj[4]; open[0] = INFINITY; oj[0] = 0;
        for (int k = 1; k < 4; k++) { open[k] = INFINITY; oj[k] = 0; for (int a = 1; a <= k; a++) { float c = tl[a - 1] + tr[k - a]; if (c < open[k]) { open[k] = c; oj[k] = uint8_t(a); } } }
        const float whole = area(N.box[b]) / ra * C_NODE + open[3];      // b as a wide node of its own: always worth all four slots
        t[0] = whole; j[0] = 0;
        for (int k = 1; k < 4; k++) { t[k] = whole; j[k] = 0; if (open[k] < t[k]) { t[k] = open[k]; j[k] = oj[k]; } if (t[k - 1] < t[k]) { t[k] = t[k - 1]; j[k] = j[k - 1]; /* fewer pieces */ } }
        // remember the opening of b itself in slot 3 of a side table: reuse J[b][.] only for 'as a child'; the node's own split is oj[3]
        J[size_t(b) * 4 + 0] = oj[3];      // (slot 0 as a child is always 'whole', so its J is free)
    }
    struct It { int bin, slot; };
    std::vector<It> cur{{N.root, 0}}, nxt; W.assign(1, RmWideNode{}); order.clear(); int lv = 0;
    while (!cur.empty()) { lv++; nxt.clear();
        for (auto it : cur) { int ch[4], n = 0;
            if (N.count[it.bin] <= gLeafMax) ch[n++] = it.bin; else {
                // expand: (node, pieces-1 budget)
                struct E { int b, k; }; std::vector<E> st; int a = J[size_t(it.bin) * 4 + 0]; st.push_back({N.right[it.bin], 3 - a}); st.push_back({N.left[it.bin], a - 1});
                while (!st.empty()) { E e = st.back(); st.pop_back(); int jj = (N.count[e.b] <= gLeafMax || e.k == 0) ? 0 : J[size_t(e.b) * 4 + e.k];
                    // J[b][k] with k>=1: 0 = whole (or fewer pieces chain) ; we stored fewer-pieces by copying j, so jj is final
                    if (jj == 0) { ch[n++] = e.b; continue; }
                    // find the k' actually used: pieces = largest k' <= e.k with T equal; simpler: recompute
                    int kk = e.k; while (kk > 1 && T[size_t(e.b) * 4 + kk] == T[size_t(e.b) * 4 + kk - 1]) kk--; jj = J[size_t(e.b) * 4 + kk]; if (jj == 0) { ch[n++] = e.b; continue; }
                    st.push_back({N.right[e.b], kk - jj}); st.push_back({N.left[e.b], jj - 1}); } }
            float lo[4][3], hi[4][3]; int ni = 0, nt = 0;
            for (int k = 0; k < n; k++) { for (int a = 0; a < 3; a++) { lo[k][a] = N.box[ch[k]].lo[a]; hi[k][a] = N.box[ch[k]].hi[a]; } if (N.count[ch[k]] > gLeafMax) ni++; else nt += N.count[ch[k]]; }
            RmWideNode w; memset(&w, 0, sizeof(w)); wide_quantise(lo, hi, n, w);
            w.child_base = ni ? int(W.size()) : 0; w.tri_base = int(order.size()); if (ni) W.resize(W.size() + ni);
            int ki = 0, toff = 0;
            for (int k = 0; k < n; k++) { if (N.count[ch[k]] > gLeafMax) { w.meta[k] = uint8_t(0x80 | ki); nxt.push_back({ch[k], w.child_base + ki}); ki++; } else { int tri[8]; int c = small_sub(N, ch[k], tri); w.meta[k] = uint8_t((toff << 2) | c); for (int t = 0; t < c; t++) order.push_back(tri[t]); toff += c; } }
            W[it.slot] = w; }
        cur.swap(nxt); }
    *levels = lv;
}
// traversal of a wide tree: node visits, box tests (non-empty children), tri tests
static void trace(const std::vector<RmWideNode> &W, const std::vector<int> &order, const char *name) {
    size_t m = rays.size() / 6; double visits = 0, tris = 0, boxes = 0, hits = 0, pushes = 0; int maxsp = 0;
    for (size_t r = 0; r < m; r++) {
        const float *R = &rays[r * 6]; float o[3] = {R[0], R[1], R[2]}, d[3] = {R[3], R[4], R[5]}, inv[3];
        for (int a = 0; a < 3; a++) { float dd = std::fabs(d[a]) < 1e-4f ? std::copysign(1e-30f, d[a]) : d[a]; inv[a] = 1.0f / dd; }
        float t = INFINITY, tmin = 1e-4f; struct E { int ref; float tl; }; E st[128]; int sp = 0; int cur = 0;
        for (;;) {
            if (cur >= 0) { const RmWideNode &nd = W[cur]; visits++;
                struct K { float tn; int ref; } ks[4]; int nk = 0;
                for (int c = 0; c < 4; c++) { if (!nd.meta[c]) continue; boxes++; float tn = tmin, tf = t;
                    for (int a = 0; a < 3; a++) { float ax = nd.s[a] * inv[a], bx = (nd.o[a] - o[a]) * inv[a]; float t0 = nd.qlo[a][c] * ax + bx, t1 = nd.qhi[a][c] * ax + bx; if (EXACT_BOXES) { t0 = (gExact[size_t(cur) * 24 + c * 6 + a] - o[a]) * inv[a]; t1 = (gExact[size_t(cur) * 24 + c * 6 + 3 + a] - o[a]) * inv[a]; } if (inv[a] < 0) std::swap(t0, t1); tn = std::max(tn, t0); tf = std::min(tf, t1); }
                    if (tn * 0.999998f <= tf * 1.000002f + 1e-4f) { uint8_t mm = nd.meta[c]; int ref = (mm & 0x80) ? nd.child_base + (mm & 0x7f) : ~(((nd.tri_base + (mm >> 2)) << 4) | (mm & 3)); ks[nk++] = {tn, ref}; } }
                std::sort(ks, ks + nk, [](const K &a, const K &b) { return a.tn < b.tn; });
                for (int c = nk - 1; c >= 1; c--) { st[sp++] = {ks[c].ref, ks[c].tn}; pushes++; } maxsp = std::max(maxsp, sp);
                if (nk) { cur = ks[0].ref; continue; }
            } else { int x = ~cur, ti = x >> 4, cnt = x & 15;
                for (int k = 0; k < cnt; k++) { tris++; const float *p = &pos[size_t(order[ti + k]) * 9];
                    float e1[3], e2[3], h[3], s[3], q[3]; for (int a = 0; a < 3; a++) { e1[a] = p[3+a] - p[a]; e2[a] = p[6+a] - p[a]; }
                    h[0] = d[1]*e2[2] - d[2]*e2[1]; h[1] = d[2]*e2[0] - d[0]*e2[2]; h[2] = d[0]*e2[1] - d[1]*e2[0];
                    float a_ = e1[0]*h[0] + e1[1]*h[1] + e1[2]*h[2]; if (std::fabs(a_) < 1e-12f) continue; float f = 1 / a_;
                    for (int a = 0; a < 3; a++) s[a] = o[a] - p[a]; float u = f * (s[0]*h[0] + s[1]*h[1] + s[2]*h[2]); if (u < 0 || u > 1) continue;
                    q[0] = s[1]*e1[2] - s[2]*e1[1]; q[1] = s[2]*e1[0] - s[0]*e1[2]; q[2] = s[0]*e1[1] - s[1]*e1[0];
                    float v = f * (d[0]*q[0] + d[1]*q[1] + d[2]*q[2]); if (v < 0 || u + v > 1) continue; float tt = f * (e2[0]*q[0] + e2[1]*q[1] + e2[2]*q[2]); if (tt > tmin && tt < t) t = tt; } }
            cur = 0x80000000; while (sp > 0) { sp--; if (st[sp].tl < t) { cur = st[sp].ref; break; } }
            if (cur == int(0x80000000)) break;
        }
        if (t < INFINITY) hits++;
    }
    printf("%-28s nodes %7zu | visits/ray %6.2f boxes/ray %6.2f tris/ray %5.2f pushes/ray %5.2f | est cost %7.0f | hit %.3f maxsp %d\n", name, W.size(), visits / m, boxes / m, tris / m, pushes / m, (visits * 300 + tris * 110) / m, hits / m, maxsp);
}
static std::vector<float> readf(const char *p) { FILE *f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<float> v(n / 4); fread(v.data(), 4, v.size(), f); fclose(f); return v; }
extern "C" int sah_sweep_host(const float *pos, int n, int depth_cap, float *lo_out, float *hi_out, int *left, int *right, int *count, int *root_out);
static Bin sweep_tree(int n, int cap) {          // the host mirror of the device's sweep-SAH builder (tests/tools/sah_sweep_host.cpp)
    Bin T; T.box.resize(2 * n); T.left.resize(2 * n); T.right.resize(2 * n); T.count.resize(2 * n);
    std::vector<float> lo(size_t(8) * n), hi(size_t(8) * n);
    auto t0 = std::chrono::steady_clock::now();
    int lv = sah_sweep_host(pos.data(), n, cap, lo.data(), hi.data(), T.left.data(), T.right.data(), T.count.data(), &T.root);
    fprintf(stderr, "sweep mirror: %d levels, %.0f ms\n", lv, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    for (int b = 0; b < 2 * n - 1; b++) for (int a = 0; a < 3; a++) { T.box[b].lo[a] = lo[size_t(b) * 4 + a]; T.box[b].hi[a] = hi[size_t(b) * 4 + a]; }
    return T;
}
int main(int argc, char **argv) {
    pos = readf("/tmp/ploc/tris.bin"); rays = readf("/tmp/ploc/rays.bin"); int n = int(pos.size() / 9);
    { std::vector<RmBvhNode> bin; std::vector<int32_t> order, worder; std::vector<RmWideNode> w; int d = 0, wd = 0;
      rm_build_fast_bvh(pos.data(), n, 22, 3, bin, order, &d);
      { double ci = 0, cl_ = 0; auto ar = [&](const RmBvhNode &r) { float dx = r.v1[0]-r.v0[0], dy = r.v1[1]-r.v0[1], dz = r.v1[2]-r.v0[2]; return dx*dy+dy*dz+dz*dx; }; float ra = 0; { B3 sb{{1e30f,1e30f,1e30f},{-1e30f,-1e30f,-1e30f}}; for (int k = 0; k < 2; k++) for (int a = 0; a < 3; a++) { sb.lo[a] = std::min(sb.lo[a], bin[2+k].v0[a]); sb.hi[a] = std::max(sb.hi[a], bin[2+k].v1[a]); } ra = area(sb); } ci += 1.0;
        std::vector<int> stk{2, 3}; while (!stk.empty()) { int i = stk.back(); stk.pop_back(); const RmBvhNode &r = bin[i]; if (r.faceR) { cl_ += ar(r) / ra * (r.faceR - r.faceL); continue; } ci += ar(r) / ra; stk.push_back(r.faceL * 2); stk.push_back(r.faceL * 2 + 1); }
        fprintf(stderr, "host binary SAH cost: inner %.2f leaf %.2f depth %d\n", ci, cl_, d); }
      rm_build_wide_bvh(bin, order, n, w, worder, &wd);
      std::vector<int> o2(worder.begin(), worder.end()); char nm[64]; snprintf(nm, 64, "host SAH (levels %d)", wd); trace(w, o2, nm); }
    for (int i = 1; i < argc; i++) { if (!strcmp(argv[i], "exact")) { EXACT_BOXES = 1; continue; } if (!strcmp(argv[i], "quant")) { EXACT_BOXES = 0; continue; }
      if (!strncmp(argv[i], "sweep", 5)) { int cap = 22; sscanf(argv[i], "sweep:%d", &cap); gLeafMax = 3; COLLAPSE_MODE = 0; Bin T = sweep_tree(n, cap); std::vector<RmWideNode> W; std::vector<int> order; int lv; collapse(T, W, order, &lv); char nm[64]; snprintf(nm, 64, "sweep SAH mirror cap%d (lv %d)", cap, lv); trace(W, order, nm);
        for (float ct : {0.2f, 0.37f, 0.6f, 1.0f}) { C_TRI = ct; dp_collapse(T, W, order, &lv); snprintf(nm, 64, "  + DP collapse c_tri %.2f (lv %d)", ct, lv); trace(W, order, nm); } continue; }
      int R = 16, lm = 3, om = 0, mt = 0, tk = 0, tw = 0; sscanf(argv[i], "%d:%d:%d:%d:%d:%d", &R, &lm, &om, &mt, &tk, &tw); TOPK = tk; TOPW = tw; gLeafMax = lm; COLLAPSE_MODE = om; METRIC = mt; Bin T = ploc(n, R); std::vector<RmWideNode> W; std::vector<int> order; int lv; collapse(T, W, order, &lv); char nm[64]; snprintf(nm, 64, "PLOC R=%d l%d o%d m%d top%d w%d (lv %d)", R, lm, om, mt, tk, tw, lv); trace(W, order, nm); }
}
