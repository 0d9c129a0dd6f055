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

#include <cuda_runtime.h>            // float2 / float4 / int2 / make_float4 ... (host-includable)
#include <device_launch_parameters.h>

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
static inline unsigned __ballot_sync(unsigned, int) { rm_gpu_only(); }
static inline unsigned __activemask() { rm_gpu_only(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { rm_gpu_only(); }
static inline void __syncthreads() { rm_gpu_only(); }
template <class T> static inline T __shfl_sync(unsigned, T, int, int = 32) { rm_gpu_only(); }
template <class T> static inline T __shfl_down_sync(unsigned, T, unsigned, int = 32) { rm_gpu_only(); }
template <class T> static inline T __shfl_xor_sync(unsigned, T, int, int = 32) { rm_gpu_only(); }
template <class T, class U> static inline T atomicAdd(T *, U) { rm_gpu_only(); }
template <class T, class U> static inline T atomicMax(T *, U) { rm_gpu_only(); }
template <class T, class U> static inline T atomicExch(T *, U) { rm_gpu_only(); }
template <class T, class U, class W> static inline T atomicCAS(T *, U, W) { rm_gpu_only(); }
