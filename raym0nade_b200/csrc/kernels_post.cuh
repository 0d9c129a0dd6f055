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

__global__ void __launch_bounds__(kFxTileW * kFxTileH) k_fxaa(const float *__restrict__ in, float *__restrict__ out, int width, int height) {
    __shared__ float luma[kFxTileH + 2][kFxTileW + 2];
    const int x0 = blockIdx.x * kFxTileW, y0 = blockIdx.y * kFxTileH;
    const int tid = threadIdx.y * kFxTileW + threadIdx.x;
    for (int i = tid; i < (kFxTileH + 2) * (kFxTileW + 2); i += kFxTileW * kFxTileH) {
        int ly = i / (kFxTileW + 2), lx = i % (kFxTileW + 2);
        int gx = x0 + lx - 1, gy = y0 + ly - 1;
        float l = 0.0f;
        if (gx >= 0 && gx < width && gy >= 0 && gy < height) {
            const float *p = in + (size_t(gy) * width + gx) * 3;
            l = lum(mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2)));
        }
        luma[ly][lx] = l;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= width || y >= height) return;
    const int lx = threadIdx.x + 1, ly = threadIdx.y + 1;
    const float *pc = in + (size_t(y) * width + x) * 3;
    const V3 center = mk3(__ldg(pc), __ldg(pc + 1), __ldg(pc + 2));
    float *po = out + (size_t(y) * width + x) * 3;
    const float M = luma[ly][lx];
    const bool hasN = y > 0, hasS = y < height - 1, hasE = x < width - 1, hasW = x > 0;
    const float N = hasN ? luma[ly - 1][lx] : M, Sl = hasS ? luma[ly + 1][lx] : M;
    const float E = hasE ? luma[ly][lx + 1] : M, Wl = hasW ? luma[ly][lx - 1] : M;
    const float rangeMin = min4(N, Sl, E, Wl), rangeMax = max4(N, Sl, E, Wl);
    const float range = fsub(rangeMax, rangeMin);
    float thr = fmul(rangeMax, 0.125f);                 // EDGE_THRESHOLD_MAX
    thr = (0.0312f < thr) ? thr : 0.0312f;              // std::max(EDGE_THRESHOLD_MIN, ...)
    if (range < thr) { po[0] = center.x; po[1] = center.y; po[2] = center.z; return; }
    const float NW = (hasN && hasW) ? luma[ly - 1][lx - 1] : M, NE = (hasN && hasE) ? luma[ly - 1][lx + 1] : M;
    const float SW = (hasS && hasW) ? luma[ly + 1][lx - 1] : M, SE = (hasS && hasE) ? luma[ly + 1][lx + 1] : M;
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
    for (int i = 0; i < 12; i++) {                      // QUALITY
        float off = fmul(gs, float(i + 1));
        float su = fadd(u, isH ? 0.0f : off), sv = fadd(v, isH ? off : 0.0f);
        if (su < 0.0f || su > 1.0f || sv < 0.0f || sv > 1.0f) continue;
        int sx = int(fmul(su, float(width))), sy = int(fmul(sv, float(height)));
        sx = sx < 0 ? 0 : (width - 1 < sx ? width - 1 : sx);
        sy = sy < 0 ? 0 : (height - 1 < sy ? height - 1 : sy);
        const float *ps = in + (size_t(sy) * width + sx) * 3;
        V3 sc = mk3(__ldg(ps), __ldg(ps + 1), __ldg(ps + 2));
        float delta = fabsf(fsub(lum(sc), M));
        if (delta > bestDelta) { bestDelta = delta; finalColor = sc; }
    }
    float sub = fmul(fadd(fmul(fabsf(fsub(fadd(N, Sl), fmul(2.0f, M))), 2.0f), fabsf(fsub(fadd(E, Wl), fmul(2.0f, M)))), 0.25f);
    sub = (1.0f < sub) ? 1.0f : sub;                    // std::min(..., 1.0f)
    const float a = fmul(sub, 0.75f);                   // SUBPIXEL_QUALITY
    V3 r = center * fsub(1.0f, a) + finalColor * a;     // glm::mix
    po[0] = r.x; po[1] = r.y; po[2] = r.z;
}

// Photo::ShadeOption bits (include/image.h:17-36)
enum { kBaseColor = 1, kEmission = 2, kDirect = 4, kIndirect = 8, kDiffuse = 16, kSpecular = 32, kShapeNormal = 64, kSurfaceNormal = 128 };

__global__ void k_shade_gamma(const RmHitInfo *__restrict__ G, const RmRadiance *__restrict__ Dd, const RmRadiance *__restrict__ Ds,
                              const RmRadiance *__restrict__ Id, const RmRadiance *__restrict__ Is, int npix, float exposure, int options,
                              float *__restrict__ rgb) {
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
    // gammaCorrection
    pix.x = (pix.x < 0.0f) ? 0.0f : pix.x; pix.y = (pix.y < 0.0f) ? 0.0f : pix.y; pix.z = (pix.z < 0.0f) ? 0.0f : pix.z;
    float C = lum(pix);
    if (C > 0.75f) {
        float bound = fadd(fdiv(tanhf(fmul(3.0f, fsub(C, 0.75f))), 3.0f), 0.75f);
        pix = div_recip(pix, C) * bound;
    }
    pix.x = (1.0f < pix.x) ? 1.0f : pix.x; pix.y = (1.0f < pix.y) ? 1.0f : pix.y; pix.z = (1.0f < pix.z) ? 1.0f : pix.z;
    const float ig = fdiv(1.0f, 2.2f);
    rgb[3 * i] = powf(pix.x, ig); rgb[3 * i + 1] = powf(pix.y, ig); rgb[3 * i + 2] = powf(pix.z, ig);
}

} // namespace rm
