// kernels_trace.cuh — ray-casting kernels: primary rays (K1) and the batched per-ray seam.
#pragma once
#include "dev_trace.cuh"

namespace rm {

constexpr int kTraceBlock = 128;

// Work counters in device memory: [0] rays, [1] box tests, [2] triangle tests.
RM_DI void flush_counters(const TraceCounters &c, unsigned long long *g, bool count_tests) {
    unsigned long long r = c.rays, b = c.box, t = c.tri;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o);
        if (count_tests) { b += __shfl_xor_sync(0xffffffffu, b, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(g, r);
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

// K1: one thread per pixel; a warp covers an 8x4 pixel tile so its rays stay coherent.
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock) k_trace_primary(DevScene S, DevArgs A, int *__restrict__ tri_idx,
                                                               float *__restrict__ t_out, unsigned long long *counters) {
    __shared__ int2 stack[kStackDepth * kTraceBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    TraceCounters cnt = {0, 0, 0};
    if (x < A.width && y < A.height) {
        RaySetup r = setup_ray(A.position, normalize(primary_d(A, x, y)));
        float t;
        int face;
        ray_hit<COUNT>(S, r, t, face, stack + threadIdx.x, kTraceBlock, cnt);
        tri_idx[y * A.width + x] = face;
        t_out[y * A.width + x] = t;
    }
    flush_counters(cnt, counters, COUNT);
}

// Batched Model::rayHit over caller-supplied rays (org/dir [n][3]).
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock) k_trace_closest(DevScene S, long long n, const float *__restrict__ org,
                                                               const float *__restrict__ dir, int *__restrict__ tri_idx,
                                                               float *__restrict__ t_out, unsigned long long *counters) {
    __shared__ int2 stack[kStackDepth * kTraceBlock];
    long long i = (long long)blockIdx.x * kTraceBlock + threadIdx.x;
    TraceCounters cnt = {0, 0, 0};
    if (i < n) {
        RaySetup r = setup_ray(mk3(org[i * 3], org[i * 3 + 1], org[i * 3 + 2]), mk3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]));
        float t;
        int face;
        ray_hit<COUNT>(S, r, t, face, stack + threadIdx.x, kTraceBlock, cnt);
        tri_idx[i] = face;
        t_out[i] = t;
    }
    flush_counters(cnt, counters, COUNT);
}

// Batched Model::rayHit_test.
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock) k_trace_occluded(DevScene S, long long n, const float *__restrict__ org,
                                                                const float *__restrict__ dir, const float *__restrict__ aim,
                                                                unsigned char *__restrict__ out, unsigned long long *counters) {
    __shared__ int2 stack[kStackDepth * kTraceBlock];
    long long i = (long long)blockIdx.x * kTraceBlock + threadIdx.x;
    TraceCounters cnt = {0, 0, 0};
    if (i < n) {
        RaySetup r = setup_ray(mk3(org[i * 3], org[i * 3 + 1], org[i * 3 + 2]), mk3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]));
        out[i] = ray_occluded<COUNT>(S, r, aim[i], stack + threadIdx.x, kTraceBlock, cnt) ? 1 : 0;
    }
    flush_counters(cnt, counters, COUNT);
}

} // namespace rm
