// gpu_bvh.h — device build of the 4-wide secondary-ray tree (gpu_bvh.cu)
#pragma once
struct RmContext;
// d_pos: device positions [n][9]; scene bounds from the reference tree's root box.  Fills ctx->b_nodes_wide (RmWideNode records,
// record 0 = the root) and ctx->b_facemap_wide (the tree's triangle order -> face index) on the device.
int rm_gpu_build_wide(RmContext *ctx, const float *d_pos, int n, const float scene_lo[3], const float scene_hi[3], int *levels_out, int *nodes_out);
