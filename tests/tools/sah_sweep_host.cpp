// tests/tools/sah_sweep_host.cpp — test infrastructure: the level loop of the device's sweep-SAH builder (gpu_sah_bvh.cu) with
// sequential loops in place of the launches and scans, around the very same per-element functions (csrc/sah_sweep.h).
// Lets tests/test_cpu_host.py check the builder's logic - ranges, scan indexing, partition, depth rule - without a GPU:
//     g++ -O2 -std=c++17 -shared -fPIC -I/usr/local/cuda/include -Iraym0nade_b200/csrc tests/tools/sah_sweep_host.cpp -o tests/tools/_sah_sweep_host.so
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "sah_sweep.h"

using namespace rm_sah;

// pos [n][9] -> binary tree arrays of 2 n entries each (lo / hi as [2n][4] floats); returns the number of levels, < 0 on failure
extern "C" int sah_sweep_host(const float *pos, int n, int depth_cap, float *lo_out, float *hi_out, int *left, int *right, int *count, int *root_out) {
    if (n < 1) return -1;
    std::vector<float4> tlo(size_t(2) * n), thi(size_t(2) * n);
    Tree T{tlo.data(), thi.data(), left, right, count};
    for (int t = 0; t < n; t++) {
        const float *p = pos + size_t(t) * 9;
        float lo[3], hi[3];
        bool finite = true;
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(std::min(p[a], p[3 + a]), p[6 + a]);
            hi[a] = std::max(std::max(p[a], p[3 + a]), p[6 + a]);
            finite = finite && std::isfinite(lo[a]) && std::isfinite(hi[a]);
        }
        if (!finite) for (int a = 0; a < 3; a++) lo[a] = hi[a] = 0.0f;
        tlo[t] = float4{lo[0], lo[1], lo[2], 0.0f};
        thi[t] = float4{hi[0], hi[1], hi[2], 0.0f};
        left[t] = ~t; right[t] = -1; count[t] = 1;
    }
    std::vector<float4> tbox(size_t(2) * n);
    for (int t = 0; t < n; t++) { tbox[2 * size_t(t)] = tlo[t]; tbox[2 * size_t(t) + 1] = thi[t]; }
    std::vector<int> list[2][3], nodeid[2];
    for (int k = 0; k < 2; k++) { for (int a = 0; a < 3; a++) list[k][a].resize(n); nodeid[k].assign(n, 0); }
    for (int a = 0; a < 3; a++) {
        std::iota(list[0][a].begin(), list[0][a].end(), 0);
        std::stable_sort(list[0][a].begin(), list[0][a].end(), [&](int x, int y) { return centre_key(tlo[x], thi[x], a) < centre_key(tlo[y], thi[y], a); });
    }
    int root = 0, levels = 0;
    if (n >= 2) {
        std::vector<int> aL[2], aR[2], aB[2];
        std::vector<unsigned long long> best[2];
        for (int k = 0; k < 2; k++) { aL[k].resize(n); aR[k].resize(n); aB[k].resize(n); best[k].resize(n); }
        std::vector<int> s_axis(n), s_M(n), s_cl(n), s_cr(n), zeros(size_t(3) * n);
        std::vector<uint8_t> side(n);
        std::vector<SweepItem> items(size_t(6) * n);
        std::vector<float> areas(size_t(6) * n);
        int next_node = n + 1, n_slots = 1, cur = 0;
        root = n;
        aL[0][0] = 0; aR[0][0] = n; aB[0][0] = n; best[0][0] = kNoSplit;
        std::vector<std::pair<int, int>> ranges;          // inner nodes created per level
        ranges.push_back({n, n + 1});
        for (int level = 0; n_slots > 0; level++, cur ^= 1) {
            if (level > 200) return -2;
            levels = level + 1;
            Level V{n, tlo.data(), thi.data(), tbox.data(), {list[cur][0].data(), list[cur][1].data(), list[cur][2].data()}, nodeid[cur].data(), aL[cur].data(), aR[cur].data()};
            for (int idx = 0; idx < 6 * n; idx++) items[idx] = sweep_item(V, idx);
            {
                SweepUnion op;
                SweepItem acc = items[0];
                areas[0] = sweep_area(acc);
                for (int idx = 1; idx < 6 * n; idx++) { acc = op(acc, items[idx]); areas[idx] = sweep_area(acc); }
            }
            for (int c = 0; c < 3 * n; c++) {
                unsigned long long key;
                const int nd = sweep_candidate(V, areas.data(), c, &key);
                if (nd >= 0 && key < best[cur][nd]) best[cur][nd] = key;
            }
            Split S{s_axis.data(), s_M.data(), s_cl.data(), s_cr.data()};
            int next_slot = 0;
            const int node0 = next_node;
            NextLevel X{aL[cur ^ 1].data(), aR[cur ^ 1].data(), aB[cur ^ 1].data(), best[cur ^ 1].data(), &next_node, &next_slot};
            for (int s = 0; s < n_slots; s++) sweep_decide(V, aB[cur].data(), s, best[cur][s], level, depth_cap, T, S, X);
            ranges.push_back({node0, next_node});
            for (int i = 0; i < n; i++) sweep_mark(V, S, i, side.data());
            {
                int acc = 0;
                for (int c = 0; c < 3 * n; c++) { zeros[c] = acc; acc += sweep_goes_left(V, side.data(), c); }
            }
            int *out_list[3] = {list[cur ^ 1][0].data(), list[cur ^ 1][1].data(), list[cur ^ 1][2].data()};
            for (int c = 0; c < 3 * n; c++) sweep_scatter(V, S, side.data(), zeros.data(), c, out_list, nodeid[cur ^ 1].data());
            n_slots = next_slot;
        }
        if (next_node != 2 * n - 1) return -3;
        for (int l = int(ranges.size()) - 1; l >= 0; l--)
            for (int b = ranges[l].first; b < ranges[l].second; b++) sweep_refit(T, b);
    }
    for (int b = 0; b < 2 * n - 1; b++) {
        lo_out[size_t(b) * 4 + 0] = tlo[b].x; lo_out[size_t(b) * 4 + 1] = tlo[b].y; lo_out[size_t(b) * 4 + 2] = tlo[b].z; lo_out[size_t(b) * 4 + 3] = 0.0f;
        hi_out[size_t(b) * 4 + 0] = thi[b].x; hi_out[size_t(b) * 4 + 1] = thi[b].y; hi_out[size_t(b) * 4 + 2] = thi[b].z; hi_out[size_t(b) * 4 + 3] = 0.0f;
    }
    *root_out = root;
    return levels;
}
