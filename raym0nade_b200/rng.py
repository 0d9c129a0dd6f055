"""Host-side (numpy) statement of the device RNG streams (csrc/dev_rng.cuh).

Every (pixel, sample, stream) owns one Philox4x32-10 stream: key = the 64-bit render seed,
counter = (pixel, sample | stream << 31, block, 0); draw i of the stream is word i % 4 of
block i // 4.  The tests use this to replay a GPU sample through the reference's own code
with identical random numbers.
"""
import numpy as np

STREAM_INDIRECT, STREAM_DIRECT = 0, 1
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all arguments broadcastable uint32 arrays -> 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0, p1 = np.uint64(_M0) * c0, np.uint64(_M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0, k1 = (k0 + np.uint64(_W0)) & _MASK, (k1 + np.uint64(_W1)) & _MASK
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def stream_u32(seed, pixel, sample, stream, n_draws):
    """First n_draws 32-bit draws of the stream of (pixel, sample, stream)."""
    nb = (n_draws + 3) // 4
    blocks = np.arange(nb, dtype=np.uint32)
    w = philox4x32_10(np.uint32(pixel), np.uint32(sample | (stream << 31)), blocks, np.uint32(0),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, 1).reshape(-1)[:n_draws]


def uniform_from_u32(x):
    """Generator::operator() of the reference (src/component.cpp:5-10) on raw 32-bit draws:
    libstdc++ generate_canonical<float,24> over one 32-bit word, then u*(b-a)+a with
    a = 1e-6f, b = 1-1e-6f, every step in fp32."""
    f32 = np.float32
    u = np.asarray(x, np.uint32).astype(f32) * f32(2.0 ** -32)
    u = np.where(u >= f32(1.0), np.nextafter(f32(1.0), f32(0.0)), u).astype(f32)
    a, b = f32(1e-6), f32(1.0) - f32(1e-6)
    return (u * (b - a) + a).astype(f32)
