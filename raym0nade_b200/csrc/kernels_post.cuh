// kernels_post.cuh — post pass: shade + tone/gamma (the step between resolve and FXAA) and FXAA.
//
//   k_fxaa          Photo::FXAA             src/image.cpp:358-452
//   k_shade_gamma   Photo::shade            src/image.cpp:215-246
//                   Photo::gammaCorrection  src/image.cpp:454-468
// FXAA is a tiled stencil: each CTA stages the luminance of its 32x8 tile plus a one-pixel
// halo in shared memory (the 3x3 neighbourhood every pixel needs); the <= 12 gather taps
// along the gradient can reach 24 pixels away and are read through L1/L2.
#pragma once
#include "dev_math.cuh"
#include "rm_types.h"

namespace rm {

constexpr int kFxTileW = 32, kFxTileH = 8;

RM_DI float min4(float a, float b, float c, float d) { float m = a; if (b < m) m = b; if (c < m) m = c; if (d < m) m = d; return m; }
RM_DI float max4(float a, float b, float c, float d) { float m = a; if (m < b) m = b; if (m < c) m = c; if (m < d) m = d; return m; }

// Data movement: the tile's rgb (12 B / pixel) comes in and goes out through shared memory as 16-byte vectors - a tile row is
// 384 contiguous bytes, 24 float4 - instead of three stride-12 scalar accesses per pixel; the two halo columns are scalar.
// Shared row layout: [32 interior pixels][right halo][left halo], so the interior starts 16-byte aligned.
constexpr int kFxRow = kFxTileW * 3 + 8;          // floats per shared row (104: a multiple of 4)
RM_DI int fx_col(int lx) { return lx >= 0 ? lx * 3 : kFxTileW * 3 + 3; }      // lx in [-1, 32]

// Two passes.  Only a few percent of a frame's pixels sit on an edge (range >= threshold), but they carry the expensive part -
// five divisions and the 12 gather taps - and in one kernel a warp runs that part with the two or three lanes that need it
// (ncu, round 2: the tap loop held 45 % of k_fxaa's instructions at 4.7 of 32 lanes).  So:
//   k_fxaa        every pixel: luminance tile, 4-neighbour range test; writes the centre through unchanged (what a pixel below the
//                 threshold gets) and appends the pixels at or above it to a list
//   k_fxaa_edges  one thread per listed pixel, dense warps: 3x3 luminances, edge direction, the 12 taps, sub-pixel blend
// Same operations on the same values in the same order as Photo::FXAA, so the result is bit-equal.
__global__ void __launch_bounds__(kFxTileW * kFxTileH) k_fxaa(const float *__restrict__ in, float *__restrict__ out, int width, int height,
                                                               int *__restrict__ edge_list, int *edge_count) {
    __shared__ __align__(16) float rgb[kFxTileH + 2][kFxRow];
    __shared__ float luma[kFxTileH + 2][kFxTileW + 2];
    const int x0 = blockIdx.x * kFxTileW, y0 = blockIdx.y * kFxTileH;
    const int tid = threadIdx.y * kFxTileW + threadIdx.x;
    const bool vec = (width & 3) == 0 && x0 + kFxTileW <= width;       // every row segment of the tile is 16-byte aligned and inside the image
    if (vec) {
        for (int i = tid; i < (kFxTileH + 2) * (kFxTileW * 3 / 4); i += kFxTileW * kFxTileH) {
            const int row = i / (kFxTileW * 3 / 4), q = i % (kFxTileW * 3 / 4), gy = y0 + row - 1;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (gy >= 0 && gy < height) {
                v = __ldg(reinterpret_cast<const float4 *>(in + (size_t(gy) * width + x0) * 3) + q);
                // the interior rows go straight out again: a pixel below the threshold keeps its colour, the others are rewritten by k_fxaa_edges
                if (row >= 1 && row <= kFxTileH) reinterpret_cast<float4 *>(out + (size_t(gy) * width + x0) * 3)[q] = v;
            }
            reinterpret_cast<float4 *>(rgb[row])[q] = v;
        }
        for (int i = tid; i < (kFxTileH + 2) * 6; i += kFxTileW * kFxTileH) {
            const int row = i / 6, k = i % 6, gy = y0 + row - 1, gx = k < 3 ? x0 + kFxTileW : x0 - 1;
            float v = 0.0f;
            if (gy >= 0 && gy < height && gx >= 0 && gx < width) v = __ldg(in + (size_t(gy) * width + gx) * 3 + k % 3);
            rgb[row][kFxTileW * 3 + k] = v;
        }
    } else {
        for (int i = tid; i < (kFxTileH + 2) * (kFxTileW + 2) * 3; i += kFxTileW * kFxTileH) {
            const int row = i / ((kFxTileW + 2) * 3), r = i % ((kFxTileW + 2) * 3), lx = r / 3 - 1, ch = r % 3;
            const int gx = x0 + lx, gy = y0 + row - 1;
            float v = 0.0f;
            if (gx >= 0 && gx < width && gy >= 0 && gy < height) {
                v = __ldg(in + (size_t(gy) * width + gx) * 3 + ch);
                if (row >= 1 && row <= kFxTileH && lx >= 0 && lx < kFxTileW) out[(size_t(gy) * width + gx) * 3 + ch] = v;
            }
            rgb[row][fx_col(lx) + ch] = v;
        }
    }
    __syncthreads();
    for (int i = tid; i < (kFxTileH + 2) * (kFxTileW + 2); i += kFxTileW * kFxTileH) {
        const int row = i / (kFxTileW + 2), lx = i % (kFxTileW + 2) - 1;
        const float *p = rgb[row] + fx_col(lx);
        luma[row][lx + 1] = lum(mk3(p[0], p[1], p[2]));
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const int lx = threadIdx.x + 1, ly = threadIdx.y + 1;
    bool edge = false;
    if (x < width && y < height) {
        const float M = luma[ly][lx];
        const float N = y > 0 ? luma[ly - 1][lx] : M, Sl = y < height - 1 ? luma[ly + 1][lx] : M;
        const float E = x < width - 1 ? luma[ly][lx + 1] : M, Wl = x > 0 ? luma[ly][lx - 1] : M;
        const float rangeMax = max4(N, Sl, E, Wl);
        const float range = fsub(rangeMax, min4(N, Sl, E, Wl));
        float thr = fmul(rangeMax, 0.125f);                 // EDGE_THRESHOLD_MAX
        thr = (0.0312f < thr) ? thr : 0.0312f;              // std::max(EDGE_THRESHOLD_MIN, ...)
        edge = !(range < thr);
    }
    // append the edge pixels: one atomic per warp
    const unsigned m = __ballot_sync(0xffffffffu, edge);
    if (m) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(edge_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (edge) edge_list[base + __popc(m & ((1u << lane) - 1u))] = y * width + x;
    }
}

RM_DI float fx_luma_at(const float *__restrict__ in, int width, int x, int y) {
    const float *p = in + (size_t(y) * width + x) * 3;
    return lum(mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2)));
}

// The expensive part of Photo::FXAA for one pixel at or above the threshold (src/image.cpp:384-446): 3x3 luminances, edge
// direction, the 12 taps, sub-pixel blend.  Shared by both forms of the pass below.
RM_DI void fx_edge_pixel(const float *__restrict__ in, float *__restrict__ out, int width, int height, int x, int y) {
    const size_t pix = size_t(y) * width + x;
    const float *pc = in + pix * 3;
    const V3 center = mk3(__ldg(pc), __ldg(pc + 1), __ldg(pc + 2));
    const float M = lum(center);
    const bool hasN = y > 0, hasS = y < height - 1, hasE = x < width - 1, hasW = x > 0;
    const float N = hasN ? fx_luma_at(in, width, x, y - 1) : M, Sl = hasS ? fx_luma_at(in, width, x, y + 1) : M;
    const float E = hasE ? fx_luma_at(in, width, x + 1, y) : M, Wl = hasW ? fx_luma_at(in, width, x - 1, y) : M;
    const float range = fsub(max4(N, Sl, E, Wl), min4(N, Sl, E, Wl));
    const float NW = (hasN && hasW) ? fx_luma_at(in, width, x - 1, y - 1) : M, NE = (hasN && hasE) ? fx_luma_at(in, width, x + 1, y - 1) : M;
    const float SW = (hasS && hasW) ? fx_luma_at(in, width, x - 1, y + 1) : M, SE = (hasS && hasE) ? fx_luma_at(in, width, x + 1, y + 1) : M;
    const float third = fdiv(1.0f, 3.0f);
    const float edgeHorz = fmul(fabsf(fsub(fadd(fadd(NW, Wl), SW), fadd(fadd(NE, E), SE))), third);
    const float edgeVert = fmul(fabsf(fsub(fadd(fadd(NW, N), NE), fadd(fadd(SW, Sl), SE))), third);
    const bool isH = edgeHorz >= edgeVert;
    const float stepLength = isH ? fdiv(1.0f, float(width)) : fdiv(1.0f, float(height));
    float g = fdiv(isH ? edgeHorz : edgeVert, range);
    g = (g < -2.0f) ? -2.0f : ((2.0f < g) ? 2.0f : g);  // std::clamp
    const float u = fdiv(float(x), float(width)), v = fdiv(float(y), float(height));
    V3 finalColor = center;
    float bestDelta = 0.0f;
    const float gs = fmul(g, stepLength);
    // The 12 taps step along ONE axis: sampleUv = uv + (0, off) on a horizontal edge, uv + (off, 0) on a vertical one
    // (src/image.cpp:412-418; the other offset is 0.0f and uv + 0.0f = uv, both being >= 0).  So one coordinate - texel, validity,
    // address - is formed once per pixel and only the other per tap.  Every fetch is issued up front (coordinates are clamped, so
    // a tap outside [0, 1] may be read and is then ignored, as the reference's `continue` ignores it); the running maximum is
    // taken in tap order.
    const float fw = float(width), fh = float(height);
    const float c_fix = isH ? u : v, c_var = isH ? v : u, n_fix = isH ? fw : fh, n_var = isH ? fh : fw;
    const int m_fix = (isH ? width : height) - 1, m_var = (isH ? height : width) - 1;
    const bool fix_ok = !(c_fix < 0.0f || c_fix > 1.0f);
    int i_fix = int(fmul(c_fix, n_fix));
    i_fix = i_fix < 0 ? 0 : (m_fix < i_fix ? m_fix : i_fix);
    const float *tap_base = in + (isH ? size_t(i_fix) * 3 : size_t(i_fix) * width * 3);
    const size_t tap_stride = isH ? size_t(width) * 3 : size_t(3);
    V3 tapc[12];
    bool tapv[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {                      // QUALITY
        const float sc = fadd(c_var, fmul(gs, float(i + 1)));
        tapv[i] = fix_ok && !(sc < 0.0f || sc > 1.0f);
        int si = int(fmul(sc, n_var));
        si = si < 0 ? 0 : (m_var < si ? m_var : si);
        const float *ps = tap_base + size_t(si) * tap_stride;
        tapc[i] = mk3(__ldg(ps), __ldg(ps + 1), __ldg(ps + 2));
    }
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const float delta = fabsf(fsub(lum(tapc[i]), M));
        if (tapv[i] && delta > bestDelta) { bestDelta = delta; finalColor = tapc[i]; }
    }
    float sub = fmul(fadd(fmul(fabsf(fsub(fadd(N, Sl), fmul(2.0f, M))), 2.0f), fabsf(fsub(fadd(E, Wl), fmul(2.0f, M)))), 0.25f);
    sub = (1.0f < sub) ? 1.0f : sub;                    // std::min(..., 1.0f)
    const float a = fmul(sub, 0.75f);                   // SUBPIXEL_QUALITY
    const V3 r = center * fsub(1.0f, a) + finalColor * a;        // glm::mix
    float *po = out + pix * 3;
    po[0] = r.x; po[1] = r.y; po[2] = r.z;
}

__global__ void __launch_bounds__(256) k_fxaa_edges(const float *__restrict__ in, float *__restrict__ out, int width, int height,
                                                    const int *__restrict__ edge_list, const int *__restrict__ edge_count) {
    const int n = *edge_count;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int pix = edge_list[e];
        fx_edge_pixel(in, out, width, height, pix % width, pix / width);
    }
}

// ---- FXAA in one launch, for frames whose width is a multiple of four (every row is then a whole number of 16-byte vectors).
// The tiled kernel above spends ~200 warp instructions per 32 pixels on index arithmetic around 12-byte pixels and is bound by
// instruction issue at a quarter of the HBM roofline.  Here a warp owns a strip of 128 x ROWS pixels and walks it top to bottom:
//   * rows travel by the bulk-copy engine (TMA, cp.async.bulk): lane 0 asks for a row of the strip - its 128 pixels plus four
//     on either side, 1632 contiguous bytes - to be dropped into one of the warp's STAGES shared-memory buffers and signalled on
//     an mbarrier, up to kFxStages rows ahead of the row being worked on; the copy-through every pixel below the threshold gets is
//     one bulk store from that same buffer back to the output frame.  No LDG / STG, no register staging, the loads of several
//     rows in flight per warp;
//   * a lane owns FOUR consecutive pixels of a row = 48 bytes = three 16-byte shared-memory loads -> four luminances; the rows
//     above / at / below the current row stay in registers (a rolling window), left / right neighbours come across lanes by
//     shuffle, the two pixels beside the warp's span from the buffer's margins;
//   * the 4-neighbour range test appends the strip's edge pixels (ballot + popc) to a per-warp list in shared memory as 16-bit
//     strip coordinates; when the strip is done - or the list may overflow - the warp runs the expensive part, fx_edge_pixel,
//     for them in dense lanes, while the strip's pixels are still in L1 / L2.  No global list, no atomics, no second launch.
// The range test uses FMNMX; std::min / std::max chains (what the reference evaluates) agree with it unless a luminance is NaN,
// which a per-row check routes to the exact chains.  Same operations on the same values in the same order as Photo::FXAA:
// bit-equal.
#ifdef __CUDACC__
constexpr int kFxStripWarps = 1, kFxStages = 3;       // one warp per CTA: the block scheduler hands out strips as warps finish
constexpr int kFxRowBytes = 136 * 12;               // 4 + 128 + 4 pixels
constexpr int kFxRowStride = 1664;                  // ... padded to a multiple of 128 bytes
constexpr int kFxListCap = 512;                     // edge pixels a warp gathers before it works them off
constexpr int kFxWarpBytes = kFxStages * kFxRowStride + kFxListCap * 2 + 128;      // row buffers, edge list, mbarriers: a multiple of 128
static_assert(kFxWarpBytes % 128 == 0 && kFxRowStride >= kFxRowBytes, "shared-memory layout of k_fxaa_strip");

RM_DI unsigned fx_smem(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
RM_DI void fx_bar_init(unsigned bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory"); }
RM_DI void fx_bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
RM_DI void fx_bar_wait(unsigned bar, unsigned parity) {
    unsigned ok, spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();          // a copy that never lands is an error, not a hang
    } while (!ok);
}
RM_DI void fx_bulk_store(void *dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(32 * kFxStripWarps, 32 / kFxStripWarps) k_fxaa_strip(const float *__restrict__ in, float *__restrict__ out, int width, int height, int ROWS) {
    extern __shared__ __align__(128) unsigned char fx_dyn[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *mine = fx_dyn + size_t(wid) * kFxWarpBytes;
    float *rows = reinterpret_cast<float *>(mine);                                                   // [kFxStages][kFxRowStride / 4]
    unsigned short *s_edges = reinterpret_cast<unsigned short *>(mine + kFxStages * kFxRowStride);    // [kFxListCap]
    const unsigned bar0 = fx_smem(mine + kFxStages * kFxRowStride + kFxListCap * 2);                  // kFxStages mbarriers
    const int spans = (width + 127) >> 7, bands = (height + ROWS - 1) / ROWS;
    const int strip = blockIdx.x * kFxStripWarps + wid;
    if (strip >= spans * bands) return;                 // whole warps leave; the kernel has no CTA-wide barrier
    const int span_x = (strip % spans) << 7, y0 = (strip / spans) * ROWS;
    const int y1 = min(y0 + ROWS, height);
    const int x0 = span_x + lane * 4;
    const bool active = x0 < width;                     // width % 4 == 0: a lane's four pixels are all inside or all outside
    const bool first = x0 == 0, last = x0 + 4 == width;
    // the rows this strip reads: [ylo, yhi]; the columns a row copy covers: [xb, xe)
    const int ylo = max(y0 - 1, 0), yhi = min(y1, height - 1), nrows = yhi - ylo + 1;
    const int xb = max(span_x - 4, 0), xe = min(span_x + 132, width);
    const unsigned row_bytes = unsigned(xe - xb) * 12u, own_bytes = unsigned(min(128, width - span_x)) * 12u;
    const unsigned dst_off = unsigned(xb - (span_x - 4)) * 12u;           // pixel span_x always lands 48 bytes into the buffer

    const size_t row_floats = size_t(width) * 3;
    const float *src_row = in + (size_t(ylo) * width + xb) * 3;          // lane 0: the next row to ask for ...
    float *dst_row = out + (size_t(y0) * width + span_x) * 3;             // ... and the next owned row to write through
    const unsigned rows_s = fx_smem(rows) + dst_off;
    auto issue = [&](int k) {                           // lane 0: ask for row k of the strip's sequence (called with k = 0, 1, 2, ...)
        const int st = k % kFxStages;
        fx_bulk_load(rows_s + st * kFxRowStride, src_row, row_bytes, bar0 + st * 8);
        src_row += row_floats;
    };
    if (lane == 0) {
        for (int st = 0; st < kFxStages; st++) fx_bar_init(bar0 + st * 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int k = 0; k < kFxStages && k < nrows; k++) issue(k);
    }
    __syncwarp();

    int n_list = 0;                                     // warp-uniform
    auto flush = [&]() {
        // every copy-through this warp has issued must have landed before an edge pixel is rewritten
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        for (int e = lane; e < n_list; e += 32) {
            const int c = s_edges[e];
            fx_edge_pixel(in, out, width, height, span_x + (c & 127), y0 + (c >> 7));
        }
        __syncwarp();
        n_list = 0;
    };

    float prv[6], cur[6], nxt[6];
    bool nan_prv = false, nan_cur = false, nan_nxt = false;
#pragma unroll
    for (int k = 0; k < 6; k++) prv[k] = cur[k] = nxt[k] = 0.0f;
#pragma unroll 1
    for (int k = 0; k <= nrows; k++) {
        if (k < nrows) {                                // warp-uniform
            const int st = k % kFxStages, y = ylo + k;
            fx_bar_wait(bar0 + st * 8, unsigned(k / kFxStages) & 1u);
            const float *buf = rows + st * (kFxRowStride / 4);
            if (lane == 0 && y >= y0 && y < y1) { fx_bulk_store(dst_row, fx_smem(buf) + 48, own_bytes); dst_row += row_floats; }
            const float4 a = reinterpret_cast<const float4 *>(buf + 12)[lane * 3], b = reinterpret_cast<const float4 *>(buf + 12)[lane * 3 + 1],
                         c = reinterpret_cast<const float4 *>(buf + 12)[lane * 3 + 2];
            const float *hp = buf + (lane == 0 ? 9 : 12 + 128 * 3);      // the pixel left of the span / right of it
            const float h = lum(mk3(hp[0], hp[1], hp[2]));
            nxt[1] = lum(mk3(a.x, a.y, a.z));
            nxt[2] = lum(mk3(a.w, b.x, b.y));
            nxt[3] = lum(mk3(b.z, b.w, c.x));
            nxt[4] = lum(mk3(c.y, c.z, c.w));
            const float up = __shfl_up_sync(0xffffffffu, nxt[4], 1), dn = __shfl_down_sync(0xffffffffu, nxt[1], 1);
            nxt[0] = first ? nxt[1] : (lane == 0 ? h : up);              // x == 0: lumaW = lumaM
            nxt[5] = last ? nxt[4] : (lane == 31 ? h : dn);              // x == width - 1: lumaE = lumaM
            const float chk = fadd(fadd(fadd(nxt[0], nxt[1]), fadd(nxt[2], nxt[3])), fadd(nxt[4], nxt[5]));
            nan_nxt = __any_sync(0xffffffffu, active && !(fabsf(chk) <= 3.0e38f));      // a NaN (or an infinity) somewhere in this row
            __syncwarp();                               // every lane has its row in registers: the buffer before this one is free
            // refill the buffer of row k - 1 with row k - 1 + STAGES once the copy-through issued from it has read it
            if (lane == 0 && k >= 1 && k - 1 + kFxStages < nrows) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                issue(k - 1 + kFxStages);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 6; j++) nxt[j] = cur[j];                // y == height - 1: lumaS = lumaM
            nan_nxt = nan_cur;
        }
        const int y = ylo + k - 1;                      // the row under test: cur, between prv and nxt
        if (k >= 1 && y >= y0 && y < y1) {
            if (y == 0) {                               // lumaN = lumaM
#pragma unroll
                for (int j = 0; j < 6; j++) prv[j] = cur[j];
                nan_prv = nan_cur;
            }
            bool edge[4];
            if (!(nan_prv || nan_cur || nan_nxt)) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float N = prv[1 + j], Sl = nxt[1 + j], E = cur[2 + j], Wl = cur[j];
                    const float rangeMax = fmaxf(fmaxf(N, Sl), fmaxf(E, Wl));
                    const float range = fsub(rangeMax, fminf(fminf(N, Sl), fminf(E, Wl)));
                    const float thr = fmaxf(fmul(rangeMax, 0.125f), 0.0312f);
                    edge[j] = active && !(range < thr);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float N = prv[1 + j], Sl = nxt[1 + j], E = cur[2 + j], Wl = cur[j];
                    const float rangeMax = max4(N, Sl, E, Wl);
                    const float range = fsub(rangeMax, min4(N, Sl, E, Wl));
                    float thr = fmul(rangeMax, 0.125f);                 // EDGE_THRESHOLD_MAX
                    thr = (0.0312f < thr) ? thr : 0.0312f;              // std::max(EDGE_THRESHOLD_MIN, ...)
                    edge[j] = active && !(range < thr);
                }
            }
            if (n_list + 128 > kFxListCap) flush();
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned m = __ballot_sync(0xffffffffu, edge[j]);
                if (edge[j]) s_edges[n_list + __popc(m & lt)] = static_cast<unsigned short>(((y - y0) << 7) | (lane * 4 + j));
                n_list += __popc(m);
            }
        }
#pragma unroll
        for (int j = 0; j < 6; j++) { prv[j] = cur[j]; cur[j] = nxt[j]; }
        nan_prv = nan_cur; nan_cur = nan_nxt;
    }
    if (n_list) flush();
    // a CTA's shared memory must outlive the bulk stores that read it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#endif  // __CUDACC__

// Photo::ShadeOption bits (include/image.h:17-36)
enum { kBaseColor = 1, kEmission = 2, kDirect = 4, kIndirect = 8, kDiffuse = 16, kSpecular = 32, kShapeNormal = 64, kSurfaceNormal = 128 };

// Photo::gammaCorrection for one pixel (src/image.cpp:454-468)
RM_DI V3 gamma_pixel(V3 pix) {
    pix.x = (pix.x < 0.0f) ? 0.0f : pix.x; pix.y = (pix.y < 0.0f) ? 0.0f : pix.y; pix.z = (pix.z < 0.0f) ? 0.0f : pix.z;
    float C = lum(pix);
    if (C > 0.75f) {
        float bound = fadd(fdiv(tanhf(fmul(3.0f, fsub(C, 0.75f))), 3.0f), 0.75f);
        pix = div_recip(pix, C) * bound;
    }
    pix.x = (1.0f < pix.x) ? 1.0f : pix.x; pix.y = (1.0f < pix.y) ? 1.0f : pix.y; pix.z = (1.0f < pix.z) ? 1.0f : pix.z;
    const float ig = fdiv(1.0f, 2.2f);
    return mk3(powf(pix.x, ig), powf(pix.y, ig), powf(pix.z, ig));
}

__global__ void k_shade_gamma(const RmHitInfo *__restrict__ G, const RmRadiance *__restrict__ Dd, const RmRadiance *__restrict__ Ds,
                              const RmRadiance *__restrict__ Id, const RmRadiance *__restrict__ Is, int npix, float exposure, int options,
                              bool do_gamma, float *__restrict__ rgb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float *g = reinterpret_cast<const float *>(G + i);
    V3 pix;
    if (options & kShapeNormal) pix = div_recip(mk3(g[0], g[1], g[2]) + splat3(1.0f), 2.0f);
    else if (options & kSurfaceNormal) pix = div_recip(mk3(g[3], g[4], g[5]) + splat3(1.0f), 2.0f);
    else {
        V3 rd = splat3(0.0f), rs = splat3(0.0f);
        if (options & kDirect) {
            if (options & kDiffuse) rd = rd + mk3(Dd[i].radiance[0], Dd[i].radiance[1], Dd[i].radiance[2]);
            if (options & kSpecular) rs = rs + mk3(Ds[i].radiance[0], Ds[i].radiance[1], Ds[i].radiance[2]);
        }
        if (options & kIndirect) {
            if (options & kDiffuse) rd = rd + mk3(Id[i].radiance[0], Id[i].radiance[1], Id[i].radiance[2]);
            if (options & kSpecular) rs = rs + mk3(Is[i].radiance[0], Is[i].radiance[1], Is[i].radiance[2]);
        }
        if (!(options & (kDirect | kIndirect))) rd = splat3(1.0f);
        V3 dc = (options & kBaseColor) ? mk3(g[9], g[10], g[11]) : splat3(1.0f);
        pix = dc * rd + rs;
        if (options & kEmission) pix = pix + mk3(g[6], g[7], g[8]) * exposure;
    }
    if (do_gamma) pix = gamma_pixel(pix);
    rgb[3 * i] = pix.x; rgb[3 * i + 1] = pix.y; rgb[3 * i + 2] = pix.z;
}

// Photo::gammaCorrection alone, in place (used when bloom sits between shade and gamma)
__global__ void k_gamma(float *__restrict__ rgb, int npix) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    V3 pix = gamma_pixel(mk3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
    rgb[3 * i] = pix.x; rgb[3 * i + 1] = pix.y; rgb[3 * i + 2] = pix.z;
}

// ------------------------------------------------------------------ image-space passes over the radiance planes
// (SURVEY.md section 8f).  The four planes Dd, Ds, Id, Is are handled by one launch each pass; every pass reads one
// set of planes and writes another (the reference filters into temporaries and copies back).
struct Planes4 { RmRadiance *p[4]; };

RM_DI float4 ld_rad(const RmRadiance *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
RM_DI void st_rad(RmRadiance *p, V3 r, float var) { *reinterpret_cast<float4 *>(p) = make_float4(r.x, r.y, r.z, var); }

// clamp / Photo::spatialClamp (src/image.cpp:30-82): luminance, 7-tap blur along x then y (taps outside the image
// skipped - adding their +0 products instead is the same sum), and a rescale of pixels that outshine 36x the
// blurred luminance of their neighbourhood.  One CTA = a 32x8 tile with a 3-pixel apron in shared memory;
// blockIdx.z = plane.  Summation order is the reference's, so the result is bit-equal.
constexpr int kClW = 32, kClH = 8, kClR = 3;
__constant__ float c_tap7[7] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};

__global__ void __launch_bounds__(kClW * kClH) k_spatial_clamp(Planes4 in, Planes4 out, int width, int height) {
    __shared__ float s_lum[kClH + 2 * kClR][kClW + 2 * kClR];
    __shared__ float s_bx[kClH + 2 * kClR][kClW];
    const RmRadiance *src = in.p[blockIdx.z];
    const int x0 = blockIdx.x * kClW, y0 = blockIdx.y * kClH;
    const int tid = threadIdx.y * kClW + threadIdx.x;
    for (int k = tid; k < (kClH + 2 * kClR) * (kClW + 2 * kClR); k += kClW * kClH) {
        const int ty = k / (kClW + 2 * kClR), tx = k % (kClW + 2 * kClR);
        const int gx = x0 + tx - kClR, gy = y0 + ty - kClR;
        float l = 0.0f;
        if (gx >= 0 && gx < width && gy >= 0 && gy < height) {
            const float4 r = ld_rad(src + size_t(gy) * width + gx);
            l = lum(mk3(r.x, r.y, r.z));
        }
        s_lum[ty][tx] = l;
    }
    __syncthreads();
    for (int k = tid; k < (kClH + 2 * kClR) * kClW; k += kClW * kClH) {
        const int ty = k / kClW, tx = k % kClW;
        float acc = 0.0f;
#pragma unroll
        for (int d = 0; d < 7; d++) acc = fadd(acc, fmul(s_lum[ty][tx + d], c_tap7[d]));
        // rows outside the image must contribute nothing to the vertical pass
        const int gy = y0 + ty - kClR;
        s_bx[ty][tx] = (gy >= 0 && gy < height) ? acc : 0.0f;
    }
    __syncthreads();
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx >= width || gy >= height) return;
    float acc = 0.0f;
#pragma unroll
    for (int d = 0; d < 7; d++) acc = fadd(acc, fmul(s_bx[threadIdx.y + d][threadIdx.x], c_tap7[d]));
    const float centre = fmul(c_tap7[3], c_tap7[3]);
    const float l = s_lum[threadIdx.y + kClR][threadIdx.x + kClR];
    const float others = fsub(acc, fmul(centre, l));
    const size_t p = size_t(gy) * width + gx;
    float4 r = ld_rad(src + p);
    if (l > fmul(36.0f, others)) {
        const float sc = fdiv(fdiv(others, fadd(l, kEps)), fsub(1.0f, centre));
        r.x = fmul(r.x, sc); r.y = fmul(r.y, sc); r.z = fmul(r.z, sc);
    }
    *reinterpret_cast<float4 *>(out.p[blockIdx.z] + p) = r;
}

// filterVar (src/image.cpp:84-107): 3x3 binomial estimate of the variance; radiance passes through
__global__ void k_filter_var(Planes4 in, Planes4 out, int width, int height) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    const RmRadiance *src = in.p[blockIdx.z];
    const float tap[3] = {0.25f, 0.5f, 0.25f};
    V3 E = splat3(0.0f);
    float E2 = 0.0f, Vs = 0.0f;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            const int nx = x + dx, ny = y + dy;
            if (nx < 0 || nx >= width || ny < 0 || ny >= height) continue;
            const float4 q = ld_rad(src + size_t(ny) * width + nx);
            const V3 r = mk3(q.x, q.y, q.z);
            const float w = fmul(tap[dx + 1], tap[dy + 1]);
            E = E + r * w;
            E2 = fadd(E2, fmul(dot(r, r), w));
            Vs = fadd(Vs, fmul(q.w, w));
        }
    const size_t p = size_t(y) * width + x;
    const float4 c = ld_rad(src + p);
    st_rad(out.p[blockIdx.z] + p, mk3(c.x, c.y, c.z), fsub(fadd(Vs, E2), dot(E, E)));
}

// The G-buffer fields the edge-stopping weight reads at a NEIGHBOUR (src/image.cpp:109-158), packed from the 88-byte
// HitInfo records into two float4 per pixel: {position, metallic} {surfaceNormal, specular} + opacity.
struct FilterG { float4 *pm, *ns; float *opacity; };

__global__ void k_filter_pack(const RmHitInfo *__restrict__ G, FilterG F, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float *g = reinterpret_cast<const float *>(G + i);
    F.pm[i] = make_float4(g[12], g[13], g[14], g[17]);
    F.ns[i] = make_float4(g[3], g[4], g[5], g[15]);
    F.opacity[i] = g[18];
}

// One a-trous pass of filterRadiance (src/image.cpp:160-191) for all four planes: the geometric factors of getWeight
// (normal power, grazing-angle term, material distance) are shared by the planes, the radiance-distance term and
// exp() are per plane.  The weight is exp() of distances times the 1024th power of a normal dot product: it is evaluated with
// the SFU's exp2 (pow1024_near_one, exp2_fast: ~1e-6 relative) and the sums are contracted to FMAs, so results agree with the
// reference to a few 1e-6 relative, not bit for bit - the tolerance is stated in tests/test_gpu_post.py.
// The edge-stopping weight is an exp() of distances: it tolerates approximate square roots and quotients (MUFU.SQRT / MUFU.RCP,
// ~2 ulp) where the rest of the library insists on the correctly rounded ones - the pass already differs from the reference by
// the ulps of powf / expf, and the weight's relative error stays below 2e-6 (|k| <= 7.5), far inside the test's 2e-4.  The
// IEEE sequences were a third of the kernel's instructions (ncu, profiles/r02k_ncu_source_k_atrous.txt).
RM_DI float sqrt_fast(float x) {
#ifdef __CUDA_ARCH__
    float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return sqrtf(x);
#endif
}
RM_DI float div_fast(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
RM_DI float exp2_fast(float x) {
#ifdef __CUDA_ARCH__
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return exp2f(x);
#endif
}
// contracted dot product / length: for the DISTANCE terms of the weight only (sensitivity ~1); the normal dot product that is
// raised to the 1024th power keeps the reference's own rounding sequence (one ulp of it is 6e-5 of the weight)
RM_DI float dot_fma(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
RM_DI float length_fast(V3 v) { return sqrt_fast(dot_fma(v, v)); }
// d^1024 for d in (0.98, 1 + a few ulps]: exp2(1024 * log2(d)) with ln(d) = -(e + e^2/2 + ... + e^6/6), e = 1 - d (exact by
// Sterbenz; the next term, e^7/7, is below 1e-11 of the sum) - six FMAs and one MUFU.EX2 instead of powf's ~45 instructions,
// relative error ~1e-6 (the rounding of an exponent of magnitude <= 30, and ex2.approx's 2^-22)
RM_DI float pow1024_near_one(float d) {
    const float e = 1.0f - d;
    const float p = e * fmaf(e, fmaf(e, fmaf(e, fmaf(e, fmaf(e, 1.0f / 6.0f, 0.2f), 0.25f), 1.0f / 3.0f), 0.5f), 1.0f);
    return exp2_fast(-1477.3196798378912f * p);            // 1024 / ln 2
}

__constant__ float c_tap5[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};

__global__ void __launch_bounds__(128, 5) k_atrous(const RmHitInfo *__restrict__ G, FilterG F, Planes4 in, Planes4 out, int width, int height, int step) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    const size_t p = size_t(y) * width + x;
    const float4 Ap = __ldg(F.pm + p);
    const V3 posp = mk3(Ap.x, Ap.y, Ap.z);
    if (!isfinite_any(posp)) {                  // background: the reference leaves its zero-initialised buffers
#pragma unroll
        for (int j = 0; j < 4; j++) st_rad(out.p[j] + p, splat3(0.0f), 0.0f);
        return;
    }
    const float4 Bp = __ldg(F.ns + p);
    const V3 snp = mk3(Bp.x, Bp.y, Bp.z);
    const float opp = __ldg(F.opacity + p);
    const float *gp = reinterpret_cast<const float *>(G + p);
    const V3 shp = mk3(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2));
    const float rough = __ldg(gp + 16);
    const float spec_scale = (rough < 4e-2f) ? 4e-2f : rough;          // std::max(Gp.roughness, eps_r)
    float4 Lp[4];
    float sig[4], rsig[4], wsum[4], var[4];
    V3 acc[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        Lp[j] = ld_rad(in.p[j] + p);
        sig[j] = fadd(fsqrt(Lp[j].w), 1e-2f);                          // sigma_l * sqrt(Lp.Var) + eps
        rsig[j] = div_fast(1.0f, sig[j]);
        wsum[j] = 0.0f; var[j] = 0.0f; acc[j] = splat3(0.0f);
    }
#pragma unroll 1
    for (int dy = -2; dy <= 2; dy++) {
#pragma unroll 1
        for (int dx = -2; dx <= 2; dx++) {
            const int nx = x + dx * step, ny = y + dy * step;
            if (nx < 0 || nx >= width || ny < 0 || ny >= height) continue;
            const size_t q = size_t(ny) * width + nx;
            const float base = fmul(c_tap5[dx + 2], c_tap5[dy + 2]);
            float4 Lq[4];
            float w[4] = {base, base, base, base};
            if (dx != 0 || dy != 0) {
                const float4 Aq = __ldg(F.pm + q);
                const V3 posq = mk3(Aq.x, Aq.y, Aq.z);
                float wn = 0.0f, kg = 0.0f;
                if (isfinite_any(posq)) {
                    const float4 Bq = __ldg(F.ns + q);
                    const float d = dot(snp, mk3(Bq.x, Bq.y, Bq.z));
                    // the reference drops a neighbour whose normal weight d^1024 is below 1e-6, i.e. d < 0.98660: for d <= 0.98
                    // (d^1024 <= 1.1e-9) that is known without the power
                    if (d > 0.98f) wn = pow1024_near_one(d);
                    if (wn < 1e-6f) wn = 0.0f;
                    else {
#pragma unroll
                        for (int j = 0; j < 4; j++) Lq[j] = ld_rad(in.p[j] + q);
                        const V3 dpos = posq - posp;
                        const float sn = fabsf(div_fast(dot_fma(shp, dpos), length_fast(dpos)));          // |dot(shapeNormal, normalize(dpos))|
                        const float tanT = div_fast(sn, sqrt_fast(fmaf(-sn, sn, 1.0f)) + kEps);
                        const float dm = length_fast(mk3(fsub(Ap.w, Aq.w), fsub(Bp.w, Bq.w), fsub(opp, __ldg(F.opacity + q))));
                        // k = ((0 - tan/sigma_z) - radianceDiff) - materialDiff/sigma_m; the middle term is per plane
                        kg = -tanT;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float dr = fmul(length_fast(mk3(fsub(Lp[j].x, Lq[j].x), fsub(Lp[j].y, Lq[j].y), fsub(Lp[j].z, Lq[j].z))), rsig[j]);
                            const float k = (kg - dr) - dm;
                            float wj = 0.0f;
                            if (!(k < -7.5f)) {
                                wj = wn * exp2_fast(k * 1.4426950408889634f);
                                if (j & 1) wj *= spec_scale;
                                if (!isfinite(wj)) wj = 0.0f;
                            }
                            w[j] = base * wj;
                        }
                    }
                }
                // a dropped neighbour enters every sum with weight base * 0 = +0: with finite planes (accumulateInwardRadiance admits
                // nothing else) it changes no sum, so its radiance is not even fetched
                if (wn == 0.0f) continue;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) Lq[j] = Lp[j];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                wsum[j] += w[j];
                acc[j] = mk3(fmaf(Lq[j].x, w[j], acc[j].x), fmaf(Lq[j].y, w[j], acc[j].y), fmaf(Lq[j].z, w[j], acc[j].z));
                var[j] = fmaf(Lq[j].w * w[j], w[j], var[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) st_rad(out.p[j] + p, div_true(acc[j], wsum[j]), fdiv(var[j], fmul(wsum[j], wsum[j])));
}

// Photo::bloom (src/image.cpp:248-283): bright pass, then five dilated 5x5 binomial blurs, each added at 1/6
__global__ void k_bloom_bright(const float *__restrict__ rgb, float *__restrict__ glow, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const V3 c = mk3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    const float L = lum(c);
    V3 g = splat3(0.0f);
    if (!(L < 1.0f)) {
        const V3 b = div_recip(c, powf(L, 0.65f));
        g = mk3(fmaxf(fsub(b.x, 1.0f), 0.0f), fmaxf(fsub(b.y, 1.0f), 0.0f), fmaxf(fsub(b.z, 1.0f), 0.0f));
    }
    glow[3 * i] = g.x; glow[3 * i + 1] = g.y; glow[3 * i + 2] = g.z;
}

__global__ void k_bloom_pass(const float *__restrict__ prev, float *__restrict__ glow, float *__restrict__ rgb, int width, int height, int step) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    const float tap[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    V3 acc = splat3(0.0f);
#pragma unroll
    for (int ky = -2; ky <= 2; ky++)
#pragma unroll
        for (int kx = -2; kx <= 2; kx++) {
            const int nx = x + kx * step, ny = y + ky * step;
            if (nx < 0 || nx >= width || ny < 0 || ny >= height) continue;
            const float *q = prev + (size_t(ny) * width + nx) * 3;
            acc = acc + (mk3(__ldg(q), __ldg(q + 1), __ldg(q + 2)) * tap[ky + 2]) * tap[kx + 2];
        }
    const size_t i = size_t(y) * width + x;
    glow[3 * i] = acc.x; glow[3 * i + 1] = acc.y; glow[3 * i + 2] = acc.z;
    const V3 add = div_recip(acc, 6.0f);
    rgb[3 * i] = fadd(rgb[3 * i], add.x); rgb[3 * i + 1] = fadd(rgb[3 * i + 1], add.y); rgb[3 * i + 2] = fadd(rgb[3 * i + 2], add.z);
}

// ------------------------------------------------------------------ depth of field
// Photo::depthFeildBlur (src/image.cpp:285-356).  The reference visits the pixels nearest first and lets each scatter its
// colour over a disc; a destination stops accepting once it has gathered 0.99, so the result depends on the visiting
// order.  Here every DESTINATION pixel replays, in that same order, the sources whose disc can reach it: the order
// (a stable sort by camera distance) is made on the host with the reference's own comparator, k_dof_tile_lists filters
// it per 32x32 tile (order kept), and k_dof_gather walks a tile's list once per destination - the same fp32 operations
// in the same sequence per destination as the reference's scatter, without atomics.
constexpr int kDofTile = 32;
constexpr float kDofMaxCoC = 96.0f;

// [0] number of NaN distances, [1] the largest circle of confusion in whole pixels (how far a source reaches)
__global__ void k_dof_stats(const float *__restrict__ depth, const float2 *__restrict__ src, int npix, int *stat) {
    int nans = 0, reach = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        if (depth[p] != depth[p]) nans++;
        const float c0 = src[p].x;
        if (c0 == c0) reach = max(reach, int(c0));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { nans += __shfl_xor_sync(0xffffffffu, nans, o); reach = max(reach, __shfl_xor_sync(0xffffffffu, reach, o)); }
    if ((threadIdx.x & 31) == 0) { if (nans) atomicAdd(stat, nans); if (reach) atomicMax(stat + 1, reach); }
}

__global__ void k_iota(int *p, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
}

// per source pixel: depth, CoC0 and the disc's total weight (the first pair of loops of the reference)
__global__ void k_dof_prepare(const RmHitInfo *__restrict__ G, V3 cam, float focus, float CoC, int npix, float *__restrict__ depth,
                              float2 *__restrict__ src) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float *g = reinterpret_cast<const float *>(G + p);
    const float d = length(mk3(g[12], g[13], g[14]) - cam);
    depth[p] = d;
    const float spread = fadd(fmul(CoC, fabsf(fsub(1.0f, fdiv(focus, d)))), kEps);
    const float c0 = (kDofMaxCoC < spread) ? kDofMaxCoC : spread;          // std::min(spread, MaxCoC); NaN stays NaN
    float total = 0.0f;
    if (c0 == c0) {
        const int radius = int(c0);
        for (int dy = -radius; dy <= radius; dy++) {
            const int xlen = int(fsqrt(fsub(fmul(c0, c0), float(dy * dy))));
            int xl, xr;
            for (xl = -xlen; xl <= 0; xl++) {
                float w = fsub(c0, fsqrt(float(xl * xl + dy * dy)));
                w = (1.0f < w) ? 1.0f : w;
                if (w == 1.0f) break;
                total = fadd(total, w);
            }
            for (xr = xlen; xr > 0; xr--) {
                float w = fsub(c0, fsqrt(float(xr * xr + dy * dy)));
                w = (1.0f < w) ? 1.0f : w;
                if (w == 1.0f) break;
                total = fadd(total, w);
            }
            total = fadd(total, float(xr - xl + 1));
        }
    }
    src[p] = make_float2(c0, total);
}

// For one tile: the sources of the depth-sorted order that lie within `reach` pixels of the tile, order kept.
__global__ void __launch_bounds__(256) k_dof_tile_lists(const int *__restrict__ sorted, int npix, int width, int height, int reach, int tiles_x,
                                                        int cap, int *__restrict__ lists, int *__restrict__ counts) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int tile = blockIdx.x;
    const int x0 = (tile % tiles_x) * kDofTile - reach, x1 = (tile % tiles_x) * kDofTile + kDofTile - 1 + reach;
    const int y0 = (tile / tiles_x) * kDofTile - reach, y1 = (tile / tiles_x) * kDofTile + kDofTile - 1 + reach;
    int *out = lists + size_t(tile) * cap;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = 0; b < npix; b += blockDim.x) {
        const int i = b + threadIdx.x;
        int q = -1;
        bool in = false;
        if (i < npix) {
            q = __ldg(sorted + i);
            const int qx = q % width, qy = q / width;
            in = qx >= x0 && qx <= x1 && qy >= y0 && qy <= y1;
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; w++) before += s_warp[w];
        if (in) out[before + __popc(m & ((1u << lane) - 1u))] = q;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[tile] = s_base;
}

// One thread per destination pixel: replay the tile's sources in order (the second pair of loops of the reference, seen
// from the destination), then pixel = blurred / gained.
__global__ void __launch_bounds__(kDofTile *kDofTile) k_dof_gather(const float *__restrict__ rgb_in, const float2 *__restrict__ src,
                                                                     const int *__restrict__ lists, const int *__restrict__ counts, int cap,
                                                                     int width, int height, int tiles_x, float *__restrict__ rgb_out) {
    const int tile = blockIdx.x;
    const int x = (tile % tiles_x) * kDofTile + threadIdx.x, y = (tile / tiles_x) * kDofTile + threadIdx.y;
    const bool live = x < width && y < height;
    const int *list = lists + size_t(tile) * cap;
    const int n = counts[tile];
    V3 acc = splat3(0.0f);
    float gained = 0.0f;
    for (int i = 0; i < n; i++) {
        const int q = __ldg(list + i);                       // the same source for the whole CTA
        const float2 s = __ldg(src + q);
        const float c0 = s.x;
        if (!(c0 == c0) || !live) continue;                  // a source without a finite depth scatters nowhere
        const int dx = x - q % width, dy = y - q / width;
        const int radius = int(c0);
        if (dy < -radius || dy > radius) continue;
        const int xlen = int(fsqrt(fsub(fmul(c0, c0), float(dy * dy))));
        if (dx < -xlen || dx > xlen) continue;
        float w = fsub(c0, fsqrt(float(dx * dx + dy * dy)));
        w = fdiv((1.0f < w) ? 1.0f : w, s.y);
        if (w < kEps) continue;
        if (fadd(gained, w) > 0.99f) w = fsub(0.99f, gained);
        if (w < kEps) continue;
        const float *c = rgb_in + size_t(q) * 3;
        acc = acc + mk3(__ldg(c), __ldg(c + 1), __ldg(c + 2)) * w;
        gained = fadd(gained, w);
    }
    if (!live) return;
    const V3 r = div_recip(acc, gained);
    float *o = rgb_out + (size_t(y) * width + x) * 3;
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}

} // namespace rm
