// wide_bvh.h — record layout of the 4-wide, 8-bit quantised secondary-ray tree (wide_bvh.cpp builds it, dev_trace4.cuh walks it)
#pragma once
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

// bin / order_in: the binary tree of rm_build_fast_bvh (leaves of at most 3 triangles); out[0] is the root
int rm_build_wide_bvh(const std::vector<RmBvhNode> &bin, const std::vector<int32_t> &order_in, int n_tris, std::vector<RmWideNode> &out,
                      std::vector<int32_t> &order_out, int *depth_out);
