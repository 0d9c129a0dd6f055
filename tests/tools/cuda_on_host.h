// cuda_on_host.h — TEST-ONLY shim that lets g++ compile the device headers of raym0nade_b200/csrc (dev_math / dev_trace /
// dev_surface / dev_bsdf / dev_texture .cuh) as ordinary host code, so the CPU suite can run the very source the kernels are
// built from against the reference's golden vectors (tests/tools/device_on_host.cpp).  The arithmetic intrinsics map to the
// IEEE operations they denote (the translation unit is compiled with -ffp-contract=off and without fast-math, so `a * b`
// IS __fmul_rn); warp-level and atomic primitives are declared so that the traversal engine parses, and trap if called -
// the engine itself only runs on the GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include <cuda_runtime.h>            // float2 / float4 / int2 / dim3 / make_float4 ... (host-includable)

// The launch geometry a kernel body reads.  On the host a kernel without barriers, shared memory or warp primitives is an
// ordinary function of (blockIdx, threadIdx): rm_host_launch below calls it once per thread of the grid, in order.
static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;
#define __launch_bounds__(...)

// Kernels that stage a tile in shared memory and meet at __syncthreads(): `__shared__` becomes one static array per
// process and rm_host_launch_blocks runs the blocks one after the other, each as blockDim real threads that meet at a
// barrier (a thread that leaves the kernel early drops out of it, as on the device).
#ifdef __shared__
#undef __shared__
#endif
#define __shared__ static
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
struct RmBlockBarrier {
    std::mutex m;
    std::condition_variable cv;
    int expected = 0, waiting = 0;
    unsigned generation = 0;
    void arrive(bool leave) {
        std::unique_lock<std::mutex> l(m);
        if (leave) expected--; else waiting++;
        if (waiting >= expected) { waiting = 0; generation++; cv.notify_all(); return; }
        if (leave) return;
        const unsigned g = generation;
        cv.wait(l, [&] { return g != generation; });
    }
};
static RmBlockBarrier *rm_block_barrier = nullptr;

template <class Kernel, class... Args>
static void rm_host_launch(Kernel kernel, dim3 grid, dim3 block, Args... args) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++)
        for (unsigned tz = 0; tz < block.z; tz++) for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) {
            blockIdx = {bx, by, bz}; threadIdx = {tx, ty, tz};
            kernel(args...);
        }
}

#ifdef __noinline__
#undef __noinline__
#endif
#define __noinline__

using std::isfinite;
using std::isnan;
using std::max;
using std::min;

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }

// GPU-only primitives: present for the parser, never executed on the host
[[noreturn]] static inline void rm_gpu_only() { std::abort(); }
// One warp = 32 real threads in lock step wherever they vote or shuffle.  Full-mask collectives only (all the engine uses):
// every collective call is one barrier; results travel through three rotating slots so that a fast lane's next call can
// never overwrite what a slow lane has not read yet.  rm_host_launch_warp sets it up.
// meeting point of the lanes named by a partial mask (a shuffle among the lanes that took a branch): `n` lanes per meeting
struct RmGroupBarrier {
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    unsigned generation = 0;
    void arrive(int n) {
        std::unique_lock<std::mutex> l(m);
        if (++waiting >= n) { waiting = 0; generation++; cv.notify_all(); return; }
        const unsigned g = generation;
        cv.wait(l, [&] { return g != generation; });
    }
};
struct RmWarp {
    RmBlockBarrier barrier;
    RmGroupBarrier group;
    uint64_t group_slots[32];
    std::atomic<unsigned> votes[3];
    uint64_t slots[3][32];
};
static thread_local RmWarp *rm_warp = nullptr;          // the warp this host thread is a lane of
static thread_local unsigned rm_warp_votes = 0, rm_warp_shuffles = 0;      // per-lane call counts: the rotation index of each kind
static inline unsigned rm_lane() { return (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31u; }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    if (!rm_warp || mask != 0xffffffffu) rm_gpu_only();
    const unsigned k = rm_warp_votes++ % 3, lane = rm_lane();
    if (pred) rm_warp->votes[k].fetch_or(1u << lane);
    rm_warp->barrier.arrive(false);
    const unsigned v = rm_warp->votes[k].load();
    if (lane == 0) rm_warp->votes[(k + 2) % 3].store(0);      // the slot of the vote after next: nobody is in it yet, or still
    return v;
}
template <class T> static inline T __shfl_sync(unsigned mask, T value, int src, int = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of a 32- or 64-bit value");
    if (!rm_warp || mask == 0u) rm_gpu_only();
    uint64_t bits = 0;
    std::memcpy(&bits, &value, sizeof(T));
    if (mask != 0xffffffffu) {                                 // only the lanes of `mask` are here (alloc_slot): they meet among themselves,
        const int n = __builtin_popcount(mask);                // once to publish and once more before anyone may publish again
        rm_warp->group_slots[rm_lane()] = bits;
        rm_warp->group.arrive(n);
        T got;
        std::memcpy(&got, &rm_warp->group_slots[src & 31], sizeof(T));
        rm_warp->group.arrive(n);
        return got;
    }
    const unsigned k = rm_warp_shuffles++ % 3, lane = rm_lane();
    rm_warp->slots[k][lane] = bits;
    rm_warp->barrier.arrive(false);
    T out;
    std::memcpy(&out, &rm_warp->slots[k][src & 31], sizeof(T));
    return out;
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T value, unsigned delta, int = 32) {
    const unsigned lane = rm_lane();
    const T from = __shfl_sync(mask, value, int(lane >= delta ? lane - delta : lane));
    return lane >= delta ? from : value;
}
// lanes holding the same value (full mask only): 32 broadcasts
template <class T> static inline unsigned __match_any_sync(unsigned mask, T value) {
    if (mask != 0xffffffffu) rm_gpu_only();
    unsigned same = 0;
    for (int src = 0; src < 32; src++)
        if (__shfl_sync(mask, value, src) == value) same |= 1u << src;
    return same;
}
static inline unsigned __activemask() { rm_gpu_only(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { rm_gpu_only(); }

// `body(lane)` on the 32 lanes of one emulated warp
template <class Body>
static void rm_host_launch_warp(Body body) {
    RmWarp warp;
    for (auto &v : warp.votes) v.store(0);
    warp.barrier.expected = 32;
    std::vector<std::thread> lanes;
    for (unsigned lane = 0; lane < 32; lane++)
        lanes.emplace_back([&, lane] {
            rm_warp = &warp;
            gridDim = dim3(1); blockDim = dim3(32);
            blockIdx = {0, 0, 0}; threadIdx = {lane, 0, 0};
            rm_warp_votes = rm_warp_shuffles = 0;
            body(int(lane));
            warp.barrier.arrive(true);
        });
    for (auto &t : lanes) t.join();
}
static inline void __syncthreads() { if (rm_block_barrier) rm_block_barrier->arrive(false); else rm_gpu_only(); }

template <class Kernel, class... Args>
static void rm_host_launch_blocks(Kernel kernel, dim3 grid, dim3 block, Args... args) {
    RmBlockBarrier barrier;
    rm_block_barrier = &barrier;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        const unsigned nthreads = block.x * block.y * block.z;
        barrier.expected = int(nthreads);
        barrier.waiting = 0;
        std::vector<std::unique_ptr<RmWarp>> warps;             // the block's warps: lanes vote and shuffle within their own
        for (unsigned w = 0; w * 32 < nthreads; w++) {
            warps.emplace_back(new RmWarp());
            for (auto &v : warps.back()->votes) v.store(0);
            warps.back()->barrier.expected = int(std::min(32u, nthreads - w * 32));
        }
        std::vector<std::thread> threads;
        for (unsigned tz = 0; tz < block.z; tz++) for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++)
            threads.emplace_back([=, &barrier, &warps] {
                gridDim = grid; blockDim = block;
                blockIdx = {bx, by, bz}; threadIdx = {tx, ty, tz};
                RmWarp *mine = warps[(tx + block.x * (ty + block.y * tz)) / 32].get();
                rm_warp = mine;
                rm_warp_votes = rm_warp_shuffles = 0;
                kernel(args...);
                mine->barrier.arrive(true);
                barrier.arrive(true);
            });
        for (auto &t : threads) t.join();
    }
    rm_block_barrier = nullptr;
}
template <class T> static inline T __shfl_down_sync(unsigned, T, unsigned, int = 32) { rm_gpu_only(); }
template <class T> static inline T __shfl_xor_sync(unsigned, T, int, int = 32) { rm_gpu_only(); }
template <class T, class U> static inline T atomicAdd(T *p, U v) {
    if constexpr (std::is_floating_point<T>::value) {          // fp32 atomic add: compare-and-swap on the bits
        T seen;
        __atomic_load(p, &seen, __ATOMIC_SEQ_CST);
        for (;;) { T want = seen + T(v); if (__atomic_compare_exchange(p, &seen, &want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return seen; }
    } else return __atomic_fetch_add(p, T(v), __ATOMIC_SEQ_CST);
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <class T, class U> static inline T atomicMax(T *, U) { rm_gpu_only(); }
template <class T, class U> static inline T atomicExch(T *p, U v) { T want = T(v), old; __atomic_exchange(p, &want, &old, __ATOMIC_SEQ_CST); return old; }
template <class T, class U, class W> static inline T atomicCAS(T *p, U compare, W value) {
    T expected = T(compare), want = T(value);
    __atomic_compare_exchange(p, &expected, &want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;                                           // the value seen: `compare` when the swap happened
}
