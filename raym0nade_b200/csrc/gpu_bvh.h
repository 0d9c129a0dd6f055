// gpu_bvh.h — device builds of the 4-wide secondary-ray tree (gpu_bvh.cu: Morton sort + PLOC; gpu_sah_bvh.cu: top-down sweep SAH)
#pragma once
#include <vector_types.h>
struct RmContext;
// a binary tree on the device: nodes [0, n) are the triangles, inner nodes follow
struct RmBinTree {
    float4 *lo, *hi;       // box; lo.w / hi.w unused
    int *left, *right;     // children; a triangle node has left = ~triangle, right = -1
    int *count;            // triangles beneath
};
// d_pos: device positions [n][9]; scene bounds from the reference tree's root box.  Fills ctx->b_nodes_wide (RmWideNode records,
// record 0 = the root) and ctx->b_facemap_wide (the tree's triangle order -> face index) on the device.
int rm_gpu_build_wide(RmContext *ctx, const float *d_pos, int n, const float scene_lo[3], const float scene_hi[3], int *levels_out, int *nodes_out);
// the same product from the sweep-SAH builder; depth_cap as for the host builder (fast_bvh.cpp); fallback_point: any point
// inside the scene bounds (where triangles with non-finite vertices are parked: they can never be hit)
int rm_gpu_build_wide_sah(RmContext *ctx, const float *d_pos, int n, int depth_cap, const float fallback_point[3], int *levels_out, int *nodes_out);
// collapse of a device binary tree into the 4-wide form (subtrees of <= 3 triangles become leaves)
int rm_gpu_collapse_wide(RmContext *ctx, const RmBinTree &N, int root, int n, int *levels_out, int *nodes_out);
