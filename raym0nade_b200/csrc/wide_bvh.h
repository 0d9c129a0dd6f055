// wide_bvh.h — record layout of the 4-wide, 8-bit quantised secondary-ray tree (wide_bvh.cpp builds it, dev_trace4.cuh walks it)
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "rm_types.h"

struct RmWideNode {
    float o[3];             // lower corner of the node = origin of its quantisation grid
    float s[3];             // grid step per axis
    uint8_t qlo[3][4];      // [axis][child] lower plane, grid units (rounded down)
    uint8_t qhi[3][4];      // [axis][child] upper plane, grid units (rounded up)
    int32_t child_base;     // record index of the first inner child
    int32_t tri_base;       // first triangle of the leaf children
    uint8_t meta[4];        // 0 empty | 0x80 + k: k-th inner child | offset << 2 + count: leaf
    int32_t _pad;
};
static_assert(sizeof(RmWideNode) == 64, "one wide node is one 64-byte record");

#ifdef __CUDACC__
#define RM_HD __host__ __device__
#else
#define RM_HD
#endif

// The quantisation grid of a node and the 8-bit boxes of its n (1..4) children on it: o = the lower corner of their union,
// s = extent / 255 rounded up until o + 255 s reaches the upper corner, lower planes rounded down and upper planes up, then
// corrected so that - evaluated in double precision - every decoded box o + s * q encloses the child's box.  Fills o, s, qlo,
// qhi (slots >= n zeroed); shared by the host builder (wide_bvh.cpp) and the device builder (gpu_bvh.cu).  Returns false
// when a box could not be enclosed (non-finite input).
RM_HD inline bool wide_quantise(const float lo[4][3], const float hi[4][3], int n, RmWideNode &w) {
    bool ok = true;
    for (int a = 0; a < 3; a++) {
        float ulo = lo[0][a], uhi = hi[0][a];
        for (int i = 1; i < n; i++) { ulo = lo[i][a] < ulo ? lo[i][a] : ulo; uhi = hi[i][a] > uhi ? hi[i][a] : uhi; }
        w.o[a] = ulo;
        float s = float((double(uhi) - double(ulo)) / 255.0);
        if (!(s > 0.0f)) s = 0.0f;
        for (int guard = 0; guard < 64 && double(ulo) + 255.0 * double(s) < double(uhi); guard++) s = nextafterf(s, INFINITY);
        w.s[a] = s;
        for (int i = 0; i < 4; i++) {
            int l = 0, h = 0;
            if (i < n) {
                const double o = ulo, sd = s;
                if (sd > 0.0) {
                    l = int(floor((double(lo[i][a]) - o) / sd));
                    h = int(ceil((double(hi[i][a]) - o) / sd));
                    l = l < 0 ? 0 : (l > 255 ? 255 : l);
                    h = h < 0 ? 0 : (h > 255 ? 255 : h);
                    while (l > 0 && o + sd * l > double(lo[i][a])) l--;
                    while (h < 255 && o + sd * h < double(hi[i][a])) h++;
                    if (o + sd * l > double(lo[i][a]) || o + sd * h < double(hi[i][a])) ok = false;
                } else if (double(lo[i][a]) < o || double(hi[i][a]) > o) ok = false;      // flat node along this axis: every plane is o itself
            }
            w.qlo[a][i] = uint8_t(l);
            w.qhi[a][i] = uint8_t(h);
        }
    }
    return ok;
}

// bin / order_in: the binary tree of rm_build_fast_bvh (leaves of at most 3 triangles); out[0] is the root
int rm_build_wide_bvh(const std::vector<RmBvhNode> &bin, const std::vector<int32_t> &order_in, int n_tris, std::vector<RmWideNode> &out,
                      std::vector<int32_t> &order_out, int *depth_out);
