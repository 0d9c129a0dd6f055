// port_post.cpp — CPU restatement of the reference's image-space passes that consume the four radiance
// planes and the G-buffer:
//   clamp / Photo::spatialClamp        src/image.cpp:30-82     7x7 separable luminance blur, outlier rescale
//   filterVar                          src/image.cpp:84-107    3x3 variance estimate
//   getWeight / filterRadiance         src/image.cpp:109-200   edge-stopping a-trous pass (5x5, dilated)
//   Photo::filter                      src/image.cpp:193-213   variance pass + steps 1,2,4,8,16 per plane
//   Photo::bloom                       src/image.cpp:248-283   bright-pass + 5 dilated 5x5 blurs
//   Photo::depthFeildBlur              src/image.cpp:285-356   depth-ordered disc scatter with a 0.99 gain cap
//
// TEST INFRASTRUCTURE (see port_math.h).  Pinned against the compiled reference (oracle/_ref) by
// tests/test_cpu_oracle.py.  Which libm flavour each unqualified call resolves to in the reference's
// image.cpp was established against that build and is noted at each call.
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "port_math.h"
#include "../../include/rm_types.h"

namespace port {
namespace {

inline vec3 rad(const RmRadiance &r) { return v3(r.radiance); }

// ---- spatial clamp ---------------------------------------------------------------------------------
const float kTap7[7] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};

// 1-D 7-tap pass along x (stride 1) or y (stride width); taps outside the image are skipped, not renormalised
void blur7(const std::vector<float> &src, std::vector<float> &dst, int width, int height, bool along_x) {
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            float acc = 0.0f;
            for (int k = -3; k <= 3; k++) {
                int nx = along_x ? x + k : x, ny = along_x ? y : y + k;
                if (nx < 0 || nx >= width || ny < 0 || ny >= height) continue;
                acc += src[size_t(ny) * width + nx] * kTap7[k + 3];
            }
            dst[size_t(y) * width + x] = acc;
        }
}

void spatialClampPlane(RmRadiance *plane, int width, int height) {
    const size_t n = size_t(width) * height;
    std::vector<float> lum(n), a(n), b(n);
    for (size_t i = 0; i < n; i++) lum[i] = dot(rad(plane[i]), RGB_Weight);
    blur7(lum, a, width, height, true);
    blur7(a, b, width, height, false);
    const float centre = kTap7[3] * kTap7[3];             // the pixel's own share of the 7x7 kernel
    for (size_t i = 0; i < n; i++) {
        float others = b[i] - centre * lum[i];
        if (lum[i] > 36.0f * others) {
            float s = others / (lum[i] + eps_zero) / (1.0f - centre);
            for (int c = 0; c < 3; c++) plane[i].radiance[c] *= s;
        }
    }
}

// ---- variance pass ---------------------------------------------------------------------------------
void variancePass(RmRadiance *plane, int width, int height) {
    const float tap[3] = {0.25f, 0.5f, 0.25f};
    std::vector<float> out(size_t(width) * height);
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            vec3 E = v3(0.0f);
            float E2 = 0.0f, V = 0.0f;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int nx = x + dx, ny = y + dy;
                    if (nx < 0 || nx >= width || ny < 0 || ny >= height) continue;
                    const RmRadiance &q = plane[size_t(ny) * width + nx];
                    float w = tap[dx + 1] * tap[dy + 1];
                    E = E + rad(q) * w;
                    E2 += dot(rad(q), rad(q)) * w;
                    V += q.Var * w;
                }
            out[size_t(y) * width + x] = V + E2 - dot(E, E);
        }
    for (size_t i = 0; i < out.size(); i++) plane[i].Var = out[i];
}

// ---- edge-stopping weight ----------------------------------------------------------------------------
// image.cpp sees <cmath> through <vector>/<iostream> and glm through image.h, with `using namespace glm`:
// pow / sqrt / exp / abs on floats bind to the float overloads (glm's, which forward to std::) - verified
// bit-for-bit against oracle/_ref.
float edgeWeight(const RmHitInfo &Gp, const RmHitInfo &Gq, const RmRadiance &Lp, const RmRadiance &Lq, bool specular) {
    if (!finite_any(v3(Gq.position))) return 0.0f;
    float w = std::pow(std::max(0.0f, dot(v3(Gp.surfaceNormal), v3(Gq.surfaceNormal))), 1024.0f);
    if (w < 1e-6f) return 0.0f;
    float k = 0.0f;
    vec3 dir = normalize(v3(Gq.position) - v3(Gp.position));
    float s = std::fabs(dot(v3(Gp.shapeNormal), dir));
    float tanTheta = s / (std::sqrt(1.0f - s * s) + eps_zero);
    k += -tanTheta / 1.0f;
    float dRad = length(rad(Lp) - rad(Lq)) / (1.0f * std::sqrt(Lp.Var) + 1e-2f);
    k += -dRad;
    float dMat = length(v3(Gp.metallic - Gq.metallic, Gp.specular - Gq.specular, Gp.opacity - Gq.opacity));
    k += -dMat / 1.0f;
    if (k < -7.5f) return 0.0f;
    w *= std::exp(k);
    if (specular) w *= std::max(Gp.roughness, 4e-2f);
    if (!std::isfinite(w)) return 0.0f;
    return w;
}

void atrousPass(RmRadiance *plane, const RmHitInfo *G, int width, int height, bool specular, int step) {
    const float tap[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    const size_t n = size_t(width) * height;
    std::vector<RmRadiance> out(n);
    std::memset(static_cast<void *>(out.data()), 0, n * sizeof(RmRadiance));
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            const size_t p = size_t(y) * width + x;
            if (!finite_any(v3(G[p].position))) continue;          // background pixels come out black
            float wsum = 0.0f, var = 0.0f;
            vec3 acc = v3(0.0f);
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++) {
                    int nx = x + dx * step, ny = y + dy * step;
                    if (nx < 0 || nx >= width || ny < 0 || ny >= height) continue;
                    const size_t q = size_t(ny) * width + nx;
                    float w = tap[dx + 2] * tap[dy + 2];
                    if (dx != 0 || dy != 0) w *= edgeWeight(G[p], G[q], plane[p], plane[q], specular);
                    wsum += w;
                    acc = acc + rad(plane[q]) * w;
                    var += plane[q].Var * w * w;
                }
            acc = div_assign(acc, wsum);
            var /= wsum * wsum;
            out[p].radiance[0] = acc.x; out[p].radiance[1] = acc.y; out[p].radiance[2] = acc.z;
            out[p].Var = var;
        }
    std::memcpy(static_cast<void *>(plane), out.data(), n * sizeof(RmRadiance));
}

void filterPlane(RmRadiance *plane, const RmHitInfo *G, int width, int height, bool specular) {
    variancePass(plane, width, height);
    for (int step = 1; step <= 16; step *= 2) atrousPass(plane, G, width, height, specular, step);
}

// ---- bloom -------------------------------------------------------------------------------------------
void bloomImage(vec3 *img, int width, int height) {
    const float tap[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    const size_t n = size_t(width) * height;
    std::vector<vec3> glow(n, v3(0.0f)), prev;
    for (size_t i = 0; i < n; i++) {
        float L = dot(img[i], RGB_Weight);
        if (L < 1.0f) continue;
        vec3 c = div_scalar(img[i], std::pow(L, 0.65f));
        glow[i] = v3(std::max(c.x - 1.0f, 0.0f), std::max(c.y - 1.0f, 0.0f), std::max(c.z - 1.0f, 0.0f));
    }
    for (int step = 1; step <= 16; step *= 2) {
        prev = glow;
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) {
                vec3 acc = v3(0.0f);
                for (int ky = -2; ky <= 2; ky++)
                    for (int kx = -2; kx <= 2; kx++) {
                        int nx = x + kx * step, ny = y + ky * step;
                        if (nx < 0 || nx >= width || ny < 0 || ny >= height) continue;
                        acc = acc + prev[size_t(ny) * width + nx] * tap[ky + 2] * tap[kx + 2];
                    }
                glow[size_t(y) * width + x] = acc;
            }
        for (size_t i = 0; i < n; i++) img[i] = img[i] + div_scalar(glow[i], 6.0f);
    }
}

// ---- depth of field ---------------------------------------------------------------------------------------
// Photo::depthFeildBlur (src/image.cpp:285-356): pixels are visited nearest first (stable sort by camera distance); each
// scatters its colour over a disc of radius CoC0 = min(CoC |1 - focus/depth| + eps, 96) with weights that fall off over
// the last pixel of the radius, normalised by the disc's total weight; a destination stops accepting once it has
// gathered 0.99.  The outcome depends on the visiting order, which is why this pass stays sequential.
void depthOfField(const RmHitInfo *G, vec3 *img, int width, int height, vec3 cam, float focus, float CoC) {
    const size_t n = size_t(width) * height;
    struct Px { int x, y; float depth; };
    std::vector<Px> px;
    px.reserve(n);
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) px.push_back({x, y, length(v3(G[size_t(y) * width + x].position) - cam)});
    std::stable_sort(px.begin(), px.end(), [](const Px &a, const Px &b) { return a.depth < b.depth; });
    std::vector<vec3> out(n, v3(0.0f));
    std::vector<float> gained(n, 0.0f);
    for (size_t i = 0; i < n; i++) {
        const Px &p = px[i];
        const float spread = CoC * std::fabs(1.0f - focus / p.depth) + eps_zero;
        const float c0 = (96.0f < spread) ? 96.0f : spread;                 // std::min(spread, MaxCoC)
        const int radius = int(c0);
        const vec3 colour = img[size_t(p.y) * width + p.x];
        float total = 0.0f;
        for (int dy = -radius; dy <= radius; dy++) {
            const int xlen = int(std::sqrt(c0 * c0 - float(dy * dy)));
            int xl, xr;
            for (xl = -xlen; xl <= 0; xl++) {                                 // soft rim on the left ...
                float w = c0 - sqrtf(float(xl * xl + dy * dy));
                w = (1.0f < w) ? 1.0f : w;
                if (w == 1.0f) break;
                total += w;
            }
            for (xr = xlen; xr > 0; xr--) {                                   // ... and on the right
                float w = c0 - sqrtf(float(xr * xr + dy * dy));
                w = (1.0f < w) ? 1.0f : w;
                if (w == 1.0f) break;
                total += w;
            }
            total += float(xr - xl + 1);                                      // the full-weight span between them
        }
        for (int dy = -radius; dy <= radius; dy++) {
            const int xlen = int(std::sqrt(c0 * c0 - float(dy * dy)));
            for (int dx = -xlen; dx <= xlen; dx++) {
                const int nx = p.x + dx, ny = p.y + dy;
                if (nx < 0 || nx >= width || ny < 0 || ny >= height) continue;
                float w = c0 - sqrtf(float(dx * dx + dy * dy));
                w = ((1.0f < w) ? 1.0f : w) / total;
                if (w < eps_zero) continue;
                const size_t id = size_t(ny) * width + nx;
                if (gained[id] + w > 0.99f) w = 0.99f - gained[id];
                if (w < eps_zero) continue;
                out[id] = out[id] + colour * w;
                gained[id] += w;
            }
        }
    }
    for (size_t i = 0; i < n; i++) img[i] = div_scalar(out[i], gained[i]);
}

} // namespace
} // namespace port

extern "C" {

// Photo::depthFeildBlur on an rgb frame, in place
void port_depth_field_blur(const RmHitInfo *g, float *rgb, int width, int height, const float *camera_position, float focus, float CoC) {
    port::depthOfField(g, reinterpret_cast<port::vec3 *>(rgb), width, height, port::v3(camera_position), focus, CoC);
}


// stages: bit 0 = Photo::spatialClamp, bit 1 = Photo::filter; planes are updated in place
void port_denoise(const RmHitInfo *g, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is, int width, int height, int stages) {
    RmRadiance *pl[4] = {Dd, Ds, Id, Is};
    if (stages & 1)
        for (int k = 0; k < 4; k++) port::spatialClampPlane(pl[k], width, height);
    if (stages & 2) {
        std::vector<std::thread> th;
        for (int k = 0; k < 4; k++) th.emplace_back([=] { port::filterPlane(pl[k], g, width, height, (k & 1) != 0); });
        for (auto &t : th) t.join();
    }
}

// Photo::bloom on an rgb frame, in place
void port_bloom(float *rgb, int width, int height) { port::bloomImage(reinterpret_cast<port::vec3 *>(rgb), width, height); }

} // extern "C"
