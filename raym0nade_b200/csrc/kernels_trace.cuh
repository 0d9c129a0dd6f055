// kernels_trace.cuh — ray-casting kernels: primary rays (K1) and the batched per-ray seam.
//
// All of them are persistent kernels around trace_engine (dev_trace.cuh): a fixed grid of
// kTraceCtasPerSm CTAs per SM, each warp pulling chunks of rays from a device-side cursor.
#pragma once
#include <algorithm>
#include "dev_trace.cuh"

namespace rm {

#ifndef RM_TRACE_BLOCK
#define RM_TRACE_BLOCK 128
#endif
constexpr int kTraceBlock = RM_TRACE_BLOCK;
constexpr int kTraceCtasPerSm = 1024 / RM_TRACE_BLOCK;      // 1024 threads x 64 registers = the SM's register file

// Work counters in device memory: [0] rays, [1] box tests, [2] triangle tests.
RM_DI void flush_counters(const TraceCounters &c, unsigned long long *g, bool count_tests) {
    unsigned long long r = c.rays, b = c.box, t = c.tri;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o);
        if (count_tests) { b += __shfl_xor_sync(0xffffffffu, b, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
    }
    if ((threadIdx.x & 31) == 0) {
        if (r) atomicAdd(g, r);
        if (count_tests) { atomicAdd(g + 1, b); atomicAdd(g + 2, t); }
    }
}

// Primary ray of pixel (x, y) exactly as renderPixel forms it (src/render.cpp:466-477):
//   d = view + accuracy * ((x - w/2) * right + (y - h/2) * up);  Dir = normalize(d)
RM_DI V3 primary_d(const DevArgs &A, int x, int y) {
    float rayX = fsub(float(x), fdiv(float(A.width), 2.0f));
    float rayY = fsub(float(y), fdiv(float(A.height), 2.0f));
    return A.direction + A.accuracy * (rayX * A.right + rayY * A.up);
}

// K1: ray index -> pixel through 8x4 tiles, so the 32 rays a warp fetches together are one tile
// and consecutive chunks are neighbouring tiles (coherent node fetches).
struct PrimaryJob {
    static constexpr bool kOcclusion = false;
    DevArgs A;
    int tiles_x;
    int *tri_idx;
    float *t_out;
    RM_DI bool pixel_of(int i, int &x, int &y) const {
        int tile = i >> 5, l = i & 31;
        x = (tile % tiles_x) * 8 + (l & 7);
        y = (tile / tiles_x) * 4 + (l >> 3);
        return x < A.width && y < A.height;
    }
    RM_DI bool load(int i, V3 &o, V3 &d, float &aim) const {
        int x, y;
        if (!pixel_of(i, x, y)) return false;
        o = A.position;
        d = normalize(primary_d(A, x, y));
        aim = CUDART_INF_F;
        return true;
    }
    RM_DI void hit(int i, float t, int face) const {
        int x, y;
        pixel_of(i, x, y);
        tri_idx[y * A.width + x] = face;
        t_out[y * A.width + x] = t;
    }
    RM_DI void visibility(int, bool) const {}
};

// Batched Model::rayHit / rayHit_test over caller-supplied rays (org/dir [n][3]).
struct RayListJob {
    const float *org, *dir, *aim_in;
    RM_DI bool load(int i, V3 &o, V3 &d, float &aim) const {
        const size_t k = size_t(i) * 3;
        o = mk3(org[k], org[k + 1], org[k + 2]);
        d = mk3(dir[k], dir[k + 1], dir[k + 2]);
        aim = aim_in ? aim_in[i] : CUDART_INF_F;
        return true;
    }
};
struct ClosestJob : RayListJob {
    static constexpr bool kOcclusion = false;
    int *tri_idx;
    float *t_out;
    RM_DI void hit(int i, float t, int face) const { tri_idx[i] = face; t_out[i] = t; }
    RM_DI void visibility(int, bool) const {}
};
struct OccludedJob : RayListJob {
    static constexpr bool kOcclusion = true;
    unsigned char *out;
    RM_DI void hit(int, float, int) const {}
    RM_DI void visibility(int i, bool occluded) const { out[i] = occluded ? 1 : 0; }
};

template <class Job, bool COUNT, bool WIDE = false>
__global__ void __launch_bounds__(kTraceBlock, kTraceCtasPerSm) k_trace(DevScene S, Job job, int n_host, const int *__restrict__ n_dev,
                                                                                int *cursor, unsigned long long *counters, TraceTune tune) {
    // one deferred child per tree level and thread: [levels][kTraceBlock] int2, sized by the launcher from the scene's
    // tree depth - the shallower the stack, the more of the SM's 256 KB stays L1 for node and triangle lines
    extern __shared__ int2 stack[];
    TraceCounters cnt = {0, 0, 0};
    const int n = n_dev ? min(*n_dev, n_host) : n_host;          // queue kernels read their length on the device
    trace_engine<Job, COUNT, WIDE>(S, job, n, cursor, stack + threadIdx.x, kTraceBlock, cnt, tune);
    flush_counters(cnt, counters, COUNT);
}

#ifdef __CUDACC__       // (the launch syntax below is nvcc's; tests/tools/cuda_on_host.h compiles this header with g++)
// Launch geometry of every trace kernel: kTraceCtasPerSm CTAs per SM, stack sized by the tree depth.
template <class Job>
inline void launch_trace(const DevScene &S, int stack_levels, bool count, int grid, cudaStream_t st, const Job &job, int n_host, const int *n_dev,
                         int *cursor, unsigned long long *counters, const TraceTune &tune) {
    TraceTune t = tune;
    t.smem_levels = std::min(stack_levels, tune.smem_levels > 0 ? tune.smem_levels : stack_levels);
    const size_t smem = size_t(t.smem_levels) * kTraceBlock * sizeof(int2);
    if (S.wide) {          // the 4-wide secondary-ray tree (wide_bvh.cpp)
        if (count) k_trace<Job, true, true><<<grid, kTraceBlock, smem, st>>>(S, job, n_host, n_dev, cursor, counters, t);
        else k_trace<Job, false, true><<<grid, kTraceBlock, smem, st>>>(S, job, n_host, n_dev, cursor, counters, t);
        return;
    }
    if (count) k_trace<Job, true><<<grid, kTraceBlock, smem, st>>>(S, job, n_host, n_dev, cursor, counters, t);
    else k_trace<Job, false><<<grid, kTraceBlock, smem, st>>>(S, job, n_host, n_dev, cursor, counters, t);
}
#endif

} // namespace rm
