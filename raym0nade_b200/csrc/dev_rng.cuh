// dev_rng.cuh — counter-based RNG replacing the reference's per-thread std::mt19937.
//
// The reference draws every uniform through Generator::operator() (src/component.cpp:5-10):
// std::uniform_real_distribution<float>(1e-6f, 1-1e-6f) over mt19937, i.e. one 32-bit draw
// x -> generate_canonical = float(x) / 2^32 (clipped below 1) -> u*(b-a)+a.  Several branches
// rely on the open range, so the mapping from 32-bit draw to float is
// reproduced exactly; only the source of the 32-bit words changes: Philox4x32-10 keyed by the
// render seed, with the counter naming (pixel, sample, stream, block).  Every (pixel, sample)
// owns an independent stream, so results do not depend on scheduling, wave size or GPU count.
#pragma once
#include <cstdint>
#include "dev_math.cuh"

namespace rm {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Generator::operator() on one raw 32-bit draw
__host__ __device__ inline float uniform_from_u32(uint32_t x) {
    float u = (float)x * 2.3283064365386963e-10f;          // float(x) / 2^32 (power of two: exact scaling)
    if (u >= 1.0f) u = 0.99999994f;                         // nextafter(1, 0)
    const float a = 1e-6f, b = 1.0f - 1e-6f;
#ifdef __CUDA_ARCH__
    return __fadd_rn(__fmul_rn(u, __fsub_rn(b, a)), a);
#else
    float s = b - a;
    float m = u * s;
    return m + a;
#endif
}

enum : uint32_t { kStreamIndirect = 0u, kStreamDirect = 1u };

#ifdef __CUDACC__
// one out-of-line copy per kernel: every gen() call site would otherwise inline the ten rounds
static __device__ __noinline__ Philox4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t k0, uint32_t k1) {
    return philox4x32_10(c0, c1, c2, 0u, k0, k1);
}
#endif

struct Rng {
    uint32_t pixel, sample_stream;      // counter words 0,1: pixel id; sample index | stream << 31
    uint32_t k0, k1;                    // key = render seed
    uint32_t drawn;                     // draws consumed so far on this stream
    Philox4 block;
    uint32_t block_id;

    RM_DI void init(uint64_t seed, uint32_t pixel_, uint32_t sample, uint32_t stream, uint32_t drawn_ = 0) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        pixel = pixel_; sample_stream = sample | (stream << 31);
        drawn = drawn_;
        block_id = 0xffffffffu;
    }
    RM_DI uint32_t next_u32() {
        uint32_t b = drawn >> 2;
        if (b != block_id) { block = philox_block(pixel, sample_stream, b, k0, k1); block_id = b; }
        uint32_t i = drawn & 3u;
        drawn++;
        return i == 0 ? block.x : (i == 1 ? block.y : (i == 2 ? block.z : block.w));
    }
    RM_DI float operator()() { return uniform_from_u32(next_u32()); }
};

} // namespace rm
