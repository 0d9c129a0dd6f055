// CPU prototype behind the secondary-ray tree's parameters (fast_bvh.cpp): box / triangle tests and engine steps per ray for
// the reference's median-split BVH (leaf <= 10) and for binned-SAH trees (leaf size, depth cap, bin count), on the bench
// scene's triangles (tris.bin: float32 [n][9]) and incoherent sample rays (rays.bin: float32 [m][6], origins on random
// triangles, directions in the upper hemisphere) - both exported with a few lines of numpy from raym0nade_b200.scenes.
// "est. cost" = inner steps x 135 + leaf steps x 170 instructions (the engine's measured step costs).
//   g++ -O2 -o bvh_proto scripts/bvh_quality_proto.cpp && ./bvh_proto        (in the directory holding tris.bin / rays.bin)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
struct V { float x, y, z; float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } };
static V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Box { V lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
    void add(V p) { lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)}; hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)}; }
    void add(const Box &b) { add(b.lo); add(b.hi); }
    float area() const { V d = hi - lo; return 2 * (d.x * d.y + d.y * d.z + d.z * d.x); } };
struct Node { Box box; int left = -1, right = -1, first = 0, count = 0; };
struct Tri { V v[3]; };
std::vector<Tri> T; std::vector<V> C; std::vector<Box> TB; std::vector<int> order; std::vector<Node> nodes; int maxDepth = 0;
int leafMax = 4, depthCap = 26, gNB = 16; bool useSah = true;
int build(int L, int R, int depth) {
    int id = (int)nodes.size(); nodes.emplace_back(); maxDepth = std::max(maxDepth, depth);
    Box b, cb; for (int i = L; i < R; i++) { b.add(TB[order[i]]); cb.add(C[order[i]]); }
    nodes[id].box = b;
    int n = R - L;
    if (n <= leafMax) { nodes[id].first = L; nodes[id].count = n; return id; }
    int M = -1;
    int need = (int)std::ceil(std::log2(std::max(1.0, double(n) / leafMax)));
    bool forceMedian = !useSah || (depth + need >= depthCap);
    if (!forceMedian) {
        const int NB = gNB; float best = 1e30f; int bestAxis = -1, bestBin = -1;
        for (int ax = 0; ax < 3; ax++) {
            float lo = cb.lo[ax], hi = cb.hi[ax]; if (!(hi > lo)) continue;
            Box bb[64]; int cnt[64] = {0};
            for (int i = L; i < R; i++) { int k = std::min(NB - 1, int((C[order[i]][ax] - lo) / (hi - lo) * NB)); bb[k].add(TB[order[i]]); cnt[k]++; }
            float la[64], ra[64]; int lc[64], rc[64]; Box acc; int c = 0;
            for (int k = 0; k < NB; k++) { if (cnt[k]) acc.add(bb[k]); c += cnt[k]; la[k] = c ? acc.area() : 0; lc[k] = c; }
            acc = Box(); c = 0;
            for (int k = NB - 1; k >= 0; k--) { if (cnt[k]) acc.add(bb[k]); c += cnt[k]; ra[k] = c ? acc.area() : 0; rc[k] = c; }
            for (int k = 0; k + 1 < NB; k++) { if (!lc[k] || !rc[k + 1]) continue; float cost = la[k] * lc[k] + ra[k + 1] * rc[k + 1]; if (cost < best) { best = cost; bestAxis = ax; bestBin = k; } }
        }
        if (bestAxis >= 0) {
            float lo = cb.lo[bestAxis], hi = cb.hi[bestAxis];
            auto mid = std::partition(order.begin() + L, order.begin() + R, [&](int t) { return std::min(gNB - 1, int((C[t][bestAxis] - lo) / (hi - lo) * gNB)) <= bestBin; });
            M = int(mid - order.begin());
            if (M == L || M == R) M = -1;
        }
    }
    if (M < 0) {   // median split on the max-variance axis (the reference's rule)
        double em[3] = {0, 0, 0}, em2[3] = {0, 0, 0};
        for (int i = L; i < R; i++) for (int a = 0; a < 3; a++) { double c = C[order[i]][a]; em[a] += c; em2[a] += c * c; }
        double D[3]; for (int a = 0; a < 3; a++) D[a] = em2[a] - em[a] * em[a] / n;
        int ax = 0; if (D[1] > D[0]) ax = 1; if (D[2] > D[0] && D[2] > D[1]) ax = 2;
        M = (L + R) / 2;
        std::nth_element(order.begin() + L, order.begin() + M, order.begin() + R, [&](int a, int b) { return C[a][ax] < C[b][ax]; });
    }
    int l = build(L, M, depth + 1), r = build(M, R, depth + 1);
    nodes[id].left = l; nodes[id].right = r; return id;
}
static bool slab(const Box &b, V o, V inv, float tmin, float tmax, float &tn) {
    float t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; a++) { float x0 = (b.lo[a] - o[a]) * inv[a], x1 = (b.hi[a] - o[a]) * inv[a]; if (x0 > x1) std::swap(x0, x1); t0 = std::max(t0, x0); t1 = std::min(t1, x1) + 1e-4f; if (t0 > t1) return false; }
    tn = t0; return t0 < t1;
}
static float tri(const Tri &t, V o, V d) {
    V e1 = t.v[1] - t.v[0], e2 = t.v[2] - t.v[0], h = cross(d, e2); float a = dot(e1, h); if (std::fabs(a) < 1e-12f) return 1e30f;
    float f = 1 / a; V s = o - t.v[0]; float u = f * dot(s, h); if (u < 0 || u > 1) return 1e30f; V q = cross(s, e1); float v = f * dot(d, q); if (v < 0 || u + v > 1) return 1e30f; return f * dot(e2, q);
}
int main(int argc, char **argv) {
    FILE *f = fopen("tris.bin", "rb"); fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET); int n = sz / 36; T.resize(n); fread(T.data(), 36, n, f); fclose(f);
    f = fopen("rays.bin", "rb"); fseek(f, 0, SEEK_END); sz = ftell(f); fseek(f, 0, SEEK_SET); int nr = sz / 24; std::vector<float> rays(nr * 6); fread(rays.data(), 24, nr, f); fclose(f);
    C.resize(n); TB.resize(n);
    for (int i = 0; i < n; i++) { Box b; for (int k = 0; k < 3; k++) b.add(T[i].v[k]); TB[i] = b; C[i] = {(T[i].v[0].x + T[i].v[1].x + T[i].v[2].x) / 3, (T[i].v[0].y + T[i].v[1].y + T[i].v[2].y) / 3, (T[i].v[0].z + T[i].v[1].z + T[i].v[2].z) / 3}; }
    struct Cfg { const char *name; bool sah; int leaf; int cap; int nb; } cfgs[] = {{"SAH leaf3 cap22 16 bins", true, 3, 22, 16}, {"SAH leaf3 cap22 32 bins", true, 3, 22, 32}, {"SAH leaf3 cap22 8 bins", true, 3, 22, 8}, {"SAH leaf3 cap23 32 bins", true, 3, 23, 32}};
    for (auto &cf : cfgs) {
        useSah = cf.sah; leafMax = cf.leaf; depthCap = cf.cap; gNB = cf.nb; nodes.clear(); nodes.reserve(2 * n); order.resize(n); for (int i = 0; i < n; i++) order[i] = i; maxDepth = 0;
        build(0, n, 0);
        double inner = 0, box = 0, tris = 0, leaves = 0, lsteps = 0, hits = 0, maxsp = 0;
        for (int r = 0; r < nr; r++) {
            V o{rays[r * 6], rays[r * 6 + 1], rays[r * 6 + 2]}, d{rays[r * 6 + 3], rays[r * 6 + 4], rays[r * 6 + 5]}, inv{1 / d.x, 1 / d.y, 1 / d.z};
            float t = 1e30f; int stack[128]; float stl[128]; int sp = 0; int cur = 0; float dummy;
            if (!slab(nodes[0].box, o, inv, 1e-4f, t, dummy)) continue;
            while (true) {
                const Node &N = nodes[cur];
                if (N.count) { leaves++; lsteps += (N.count + 1) / 2; for (int i = 0; i < N.count; i++) { tris++; float tt = tri(T[order[N.first + i]], o, d); if (tt > 1e-4f && tt < t) t = tt; } cur = -1; }
                else {
                    inner++; box += 2; float tl, tr; bool hl = slab(nodes[N.left].box, o, inv, 1e-4f, t, tl), hr = slab(nodes[N.right].box, o, inv, 1e-4f, t, tr);
                    if (hl && hr) { int first = tl < tr ? N.left : N.right, second = tl < tr ? N.right : N.left; stack[sp] = second; stl[sp] = std::max(tl, tr); sp++; if (sp > maxsp) maxsp = sp; cur = first; }
                    else if (hl) cur = N.left; else if (hr) cur = N.right; else cur = -1;
                }
                if (cur < 0) { while (sp > 0) { sp--; if (stl[sp] < t) { cur = stack[sp]; break; } } if (cur < 0) break; }
            }
            if (t < 1e30f) hits++;
        }
        double cost = inner / nr * 135 + lsteps / nr * 170;
        printf("%-32s nodes %8zu depth %2d maxstack %2.0f | inner %.1f box %.1f tri %.1f leaves %.1f leafsteps %.1f hit %.2f | est. cost %.0f\n", cf.name, nodes.size(), maxDepth, maxsp, inner / nr, box / nr, tris / nr, leaves / nr, lsteps / nr, hits / nr, cost);
    }
}
