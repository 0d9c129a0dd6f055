// port.cpp — CPU restatement of the deterministic stages of Raym0nade's hot path.
//
// TEST INFRASTRUCTURE (see port_math.h).  Plain C++, recursive like the reference, no glm.
// Every function cites the reference file:line it follows (paths relative to the
// lemonchu/Raym0nade tree).  Pinned against oracle/_ref (the reference compiled unmodified)
// through the golden vectors in tests/golden/ and directly in tests/test_cpu_oracle.py.
//
// Covered: geometry kernels, BVH build + traversal (with box/triangle test counters = the B and
// T of the bytes-per-ray roofline figure), Model::rayHit / rayHit_test with alpha cut-outs,
// texture / material / sky lookups, getHitInfo (G-buffer), BSDF evaluation, radiance split,
// RNG float mapping, shade + gamma, FXAA.
// Not restated here: the stochastic estimator (sampleRay & co.) - it is checked against
// oracle/_ref by per-sample replay (tests/test_gpu_render.py).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "port_math.h"
#include "../../include/rm_types.h"

namespace port {

// ----------------------------------------------------------------- geometry (src/geometry.cpp)
struct Ray { vec3 origin, direction; };
struct Box { vec3 v0, v1; };

// rayInBox, src/geometry.cpp:40-61
void rayInBox(const Ray &ray, const Box &box, float &tL, float &tR) {
    for (int i = 0; i < 3; ++i) {
        if (std::abs(ray.direction[i]) < eps_zero) {
            if (ray.origin[i] < box.v0[i] || ray.origin[i] > box.v1[i]) { tR = -1.0; return; }
        } else {
            float invD = 1.0f / ray.direction[i];
            if (invD >= 0) {
                tL = std::fmax(tL, (box.v0[i] - ray.origin[i]) * invD);
                tR = std::fmin(tR, (box.v1[i] - ray.origin[i]) * invD);
            } else {
                tL = std::fmax(tL, (box.v1[i] - ray.origin[i]) * invD);
                tR = std::fmin(tR, (box.v0[i] - ray.origin[i]) * invD);
            }
            tR += eps_zero;
            if (tL > tR) return;
        }
    }
}

// RayTriangleIntersection, src/geometry.cpp:63-87
float rayTriangle(const Ray &ray, vec3 v0, vec3 v1, vec3 v2) {
    vec3 edge1 = v1 - v0, edge2 = v2 - v0;
    vec3 h = cross(ray.direction, edge2);
    float a = dot(edge1, h);
    if (std::abs(a) / length(edge1) < eps_zero) return INFINITY;
    float f = 1.0f / a;
    vec3 s = ray.origin - v0;
    float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return INFINITY;
    vec3 q = cross(s, edge1);
    float v = f * dot(ray.direction, q);
    if (v < 0.0f || u + v > 1.0f) return INFINITY;
    return f * dot(edge2, q);
}

// barycentric, src/geometry.cpp:89-103
vec3 barycentric(vec3 v0, vec3 v1, vec3 v2, vec3 P) {
    vec3 v0v1 = v1 - v0, v0v2 = v2 - v0;
    vec3 n = cross(v0v1, v0v2);
    float denom = dot(n, n);
    vec3 v0P = P - v0;
    float alpha = dot(cross(v0P, v0v2), n) / denom;
    float beta = dot(cross(v0v1, v0P), n) / denom;
    float gamma = 1.0f - alpha - beta;
    return {gamma, alpha, beta};
}

// ----------------------------------------------------------------- scene
struct Texture {
    int width = 0, height = 0, channels = 0, map_depth = 0;
    std::vector<uint8_t> data[8];
    bool empty() const { return data[0].empty(); }
};

struct Material {
    Texture tex[4];       // diffuse, specular, emissive, normals
    int id = 0;
    bool cutout = false;
    float opacity = 1.0f, ior = 1.0f, roughness = 0.8f;
    vec3 transmitting{0, 0, 0};
};

struct Face {
    vec3 v[3];
    vec2 uv[3];
    vec3 n[3];
    int material;
    int original;
    vec3 center() const { return div_scalar(v[0] + v[1] + v[2], 3.0f); }      // src/component.cpp:37-39
};

struct Node { Box box; int faceL = 0, faceR = 0; };

struct Counters { uint64_t rays = 0, box = 0, tri = 0; };
static thread_local Counters tl;

struct Scene {
    std::vector<Material> materials;
    std::vector<Face> faces;
    std::vector<Node> nodes;
    int sky_w = 0, sky_h = 0;
    std::vector<vec3> sky;       // premultiplied
    std::vector<float> sky_cdf;
};

// ImageData::generateMipmaps, src/material.cpp:113-148
void generateMipmaps(Texture &t) {
    t.map_depth = 8;
    for (int level = 1; level < 8; ++level) {
        int pw = t.width >> (level - 1), ph = t.height >> (level - 1), cw = t.width >> level, ch = t.height >> level;
        if (cw == 0 || ch == 0) { t.map_depth = level; break; }
        t.data[level].resize(size_t(cw) * ch * t.channels);
        for (int y = 0; y < ch; ++y)
            for (int x = 0; x < cw; ++x)
                for (int c = 0; c < t.channels; ++c) {
                    int sum = 0;
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx)
                            sum += t.data[level - 1][(size_t((y * 2 + dy) % ph) * pw + (x * 2 + dx) % pw) * t.channels + c];
                    t.data[level][(size_t(y) * cw + x) * t.channels + c] = uint8_t(sum / 4);
                }
    }
}

// BVH::dfs_build / nodeCount, src/bvh.cpp:18-46
int nodeCount(int u, int n) { return (n <= 10) ? u : nodeCount(u << 1 | 1, (n + 1) >> 1); }

void dfs_build(Scene &S, int u, int L, int R) {
    Node &nd = S.nodes[u];
    if (R - L <= 10) {
        nd.faceL = L; nd.faceR = R;
        nd.box = {v3(INFINITY), v3(-INFINITY)};
        for (int i = L; i < R; i++) {
            const Face &f = S.faces[i];
            for (int k = 0; k < 3; k++) {
                float mn = std::min(f.v[0][k], std::min(f.v[1][k], f.v[2][k])), mx = std::max(f.v[0][k], std::max(f.v[1][k], f.v[2][k]));
                float *lo = &nd.box.v0.x + k, *hi = &nd.box.v1.x + k;
                *lo = std::fmin(*lo, mn);
                *hi = std::fmax(*hi, mx);
            }
        }
        return;
    }
    vec3 Em = v3(0.0f), Em2 = v3(0.0f);
    for (int i = L; i < R; i++) { vec3 m = S.faces[i].center(); Em = Em + m; Em2 = Em2 + m * m; }
    vec3 D = Em2 - div_scalar(Em * Em, float(R - L));
    int axis = 0;
    if (D[1] > D[0]) axis = 1;
    if (D[2] > D[0] && D[2] > D[1]) axis = 2;
    int M = (R + L) / 2;
    std::nth_element(S.faces.begin() + L, S.faces.begin() + M, S.faces.begin() + R,
                     [axis](const Face &a, const Face &b) { return a.center()[axis] < b.center()[axis]; });
    dfs_build(S, u << 1, L, M);
    dfs_build(S, u << 1 | 1, M, R);
    const Box &a = S.nodes[u << 1].box, &b = S.nodes[u << 1 | 1].box;
    S.nodes[u].box = {v3(std::fmin(a.v0.x, b.v0.x), std::fmin(a.v0.y, b.v0.y), std::fmin(a.v0.z, b.v0.z)),
                      v3(std::fmax(a.v1.x, b.v1.x), std::fmax(a.v1.y, b.v1.y), std::fmax(a.v1.z, b.v1.z))};
}

// ----------------------------------------------------------------- traversal (src/bvh.cpp:56-92)
struct HitRecord { float t_min, t_max; int face; };

void dfs_rayHit(const Scene &S, int u, const Ray &ray, HitRecord &hit) {
    const Node &nd = S.nodes[u];
    if (nd.faceR) {
        for (int i = nd.faceL; i < nd.faceR; i++) {
            const Face &f = S.faces[i];
            tl.tri++;
            float t = rayTriangle(ray, f.v[0], f.v[1], f.v[2]);
            if (hit.t_min < t && t < hit.t_max) { hit.t_max = t; hit.face = i; }
        }
        return;
    }
    float tL0 = hit.t_min, tR0 = hit.t_max, tL1 = hit.t_min, tR1 = hit.t_max;
    tl.box += 2;
    rayInBox(ray, S.nodes[u << 1].box, tL0, tR0);
    rayInBox(ray, S.nodes[u << 1 | 1].box, tL1, tR1);
    if (tL0 < tL1) {
        if (tL0 < tR0) dfs_rayHit(S, u << 1, ray, hit);
        if (tL1 < tR1 && tL1 < hit.t_max) dfs_rayHit(S, u << 1 | 1, ray, hit);
    } else {
        if (tL1 < tR1) dfs_rayHit(S, u << 1 | 1, ray, hit);
        if (tL0 < tR0 && tL0 < hit.t_max) dfs_rayHit(S, u << 1, ray, hit);
    }
}

// ----------------------------------------------------------------- textures (src/material.cpp:26-94, 337-383)
void wrap(int &x, int m) { if (x < 0 || x >= m) { x %= m; if (x < 0) x += m; } }

vec4 bilinear4(const std::vector<uint8_t> &d, int w, int h, float u, float v) {
    float x = u * float(w), y = v * float(h);
    int x0 = int(floorf(x)), y0 = int(floorf(y));
    float dx = x - float(x0), dy = y - float(y0);
    wrap(x0, w); wrap(y0, h);
    int x1 = (x0 + 1) % w, y1 = (y0 + 1) % h;
    float out[4];
    for (int c = 0; c < 4; c++) {       // vec4 / 255.0f is a true division in glm
        float c00 = float(d[(size_t(y0) * w + x0) * 4 + c]) / 255.0f, c01 = float(d[(size_t(y0) * w + x1) * 4 + c]) / 255.0f;
        float c10 = float(d[(size_t(y1) * w + x0) * 4 + c]) / 255.0f, c11 = float(d[(size_t(y1) * w + x1) * 4 + c]) / 255.0f;
        float c0 = c00 * (1.0f - dx) + c01 * dx, c1 = c10 * (1.0f - dx) + c11 * dx;
        out[c] = c0 * (1.0f - dy) + c1 * dy;
    }
    return {out[0], out[1], out[2], out[3]};
}

vec3 bilinear3(const std::vector<uint8_t> &d, int w, int h, float u, float v) {
    float x = u * float(w), y = v * float(h);
    int x0 = int(floorf(x)), y0 = int(floorf(y));
    float dx = x - float(x0), dy = y - float(y0);
    wrap(x0, w); wrap(y0, h);
    int x1 = (x0 + 1) % w, y1 = (y0 + 1) % h;
    const float r = 1.0f / 255.0f;      // vec3 / 255.0f is v * (1/255) in glm
    float out[3];
    for (int c = 0; c < 3; c++) {
        float c00 = float(d[(size_t(y0) * w + x0) * 3 + c]) * r, c01 = float(d[(size_t(y0) * w + x1) * 3 + c]) * r;
        float c10 = float(d[(size_t(y1) * w + x0) * 3 + c]) * r, c11 = float(d[(size_t(y1) * w + x1) * 3 + c]) * r;
        float c0 = c00 * (1.0f - dx) + c01 * dx, c1 = c10 * (1.0f - dx) + c11 * dx;
        out[c] = c0 * (1.0f - dy) + c1 * dy;
    }
    return {out[0], out[1], out[2]};
}

void mipSelect(const Texture &t, float &depth, int &level, int &next, float &blend) {
    depth = std::max(std::min(depth, float(t.map_depth - 1)), 0.0f);
    level = int(depth);
    next = std::min(level + 1, t.map_depth - 1);
    blend = depth - float(level);
}

vec4 get4(const Texture &t, float u, float v, float depth) {       // ImageData::get<vec4>
    v = 1.0f - v;
    int level, next; float blend;
    mipSelect(t, depth, level, next, blend);
    vec4 a = bilinear4(t.data[level], t.width >> level, t.height >> level, u, v);
    vec4 b = bilinear4(t.data[next], t.width >> next, t.height >> next, u, v);
    return {a.x * (1.0f - blend) + b.x * blend, a.y * (1.0f - blend) + b.y * blend, a.z * (1.0f - blend) + b.z * blend, a.w * (1.0f - blend) + b.w * blend};
}

vec3 get3(const Texture &t, float u, float v, float depth) {       // ImageData::get<vec3>
    v = 1.0f - v;
    int level, next; float blend;
    mipSelect(t, depth, level, next, blend);
    vec3 a = bilinear3(t.data[level], t.width >> level, t.height >> level, u, v);
    vec3 b = bilinear3(t.data[next], t.width >> next, t.height >> next, u, v);
    return a * (1.0f - blend) + b * blend;
}

float lod(const Texture &t, float duv) { return std::isnan(duv) ? 0.0f : log2f(duv * float(t.width)); }

vec4 getDiffuseColor(const Material &m, float u, float v, float duv) {
    if (m.tex[0].empty()) return {1, 1, 1, 1};
    vec4 c = get4(m.tex[0], u, v, lod(m.tex[0], duv));
    return {powf(c.x, 2.2f), powf(c.y, 2.2f), powf(c.z, 2.2f), c.w};
}
vec3 getEmissiveColor(const Material &m, float u, float v, float duv) {
    if (m.tex[2].empty()) return v3(0.0f);
    vec4 c = get4(m.tex[2], u, v, lod(m.tex[2], duv));
    return {powf(c.x, 2.2f), powf(c.y, 2.2f), powf(c.z, 2.2f)};
}
vec3 getNormal(const Material &m, float u, float v, float duv) {
    if (m.tex[3].empty()) return v3(0.0f);
    vec3 c = get3(m.tex[3], u, v, lod(m.tex[3], duv));
    return {c.x * 2.0f - 1.0f, c.y * 2.0f - 1.0f, c.z * 2.0f - 1.0f};
}
void getSurfaceData(const Material &m, float u, float v, float &roughness, float &metallic) {
    if (m.tex[1].empty()) { metallic = 0.0f; roughness = m.roughness; return; }
    vec4 c = get4(m.tex[1], u, v, 0);
    metallic = std::min(c.z, 0.99f);
    roughness = std::max(c.y, 1e-3f);
}

// SkyBox::get, src/component.cpp:120-140
vec3 skyGet(const Scene &S, vec3 dir) {
    if (S.sky.empty()) return v3(0.0f);
    float theta = atan2f(-dir.x, dir.z), phi = acosf(dir.y);
    if (theta < 0.0f) theta += 2.0f * PI;
    int u = int(theta / (2.0f * PI) * float(S.sky_w)), v = int(phi / PI * float(S.sky_h));
    u = std::max(0, std::min(u, S.sky_w - 1));
    v = std::max(0, std::min(v, S.sky_h - 1));
    phi = PI * (float(v) + 0.5f) / float(S.sky_h);
    float area = sinf(phi) * 2.0f * PI / float(S.sky_w * S.sky_h);
    return div_scalar(S.sky[size_t(v) * S.sky_w + u], area);
}

// ----------------------------------------------------------------- Model::rayHit & co. (src/model.cpp:217-354)
bool transparentTest(const Scene &S, const Ray &ray, const HitRecord &hit) {
    const Face &f = S.faces[hit.face];
    const Material &m = S.materials[f.material];
    if (!m.cutout) return false;
    vec3 P = ray.origin + ray.direction * hit.t_max;
    vec3 b = barycentric(f.v[0], f.v[1], f.v[2], P);
    vec2 uv = b.x * f.uv[0] + b.y * f.uv[1] + b.z * f.uv[2];
    return getDiffuseColor(m, uv.x, uv.y, NAN).w < eps_zero;
}

HitRecord modelRayHit(const Scene &S, const Ray &ray) {
    HitRecord hit{eps_zero, INFINITY, -1};
    for (int T = 0; T < 8; T++) {
        tl.rays++;
        dfs_rayHit(S, 1, ray, hit);
        if (hit.t_max == INFINITY || !transparentTest(S, ray, hit)) return hit;
        hit = {hit.t_max + eps_zero, INFINITY, -1};
    }
    return hit;
}

bool modelRayHitTest(const Scene &S, const Ray &ray, float aim) {
    HitRecord hit{eps_zero, aim + eps_zero, -1};
    for (int T = 0; T < 8; T++) {
        tl.rays++;
        dfs_rayHit(S, 1, ray, hit);
        if (hit.t_max >= aim) return false;
        if (!transparentTest(S, ray, hit)) return true;
        hit = {hit.t_max + eps_zero, aim + eps_zero, -1};
    }
    return true;
}

// ----------------------------------------------------------------- surface record (src/model.cpp:232-328, src/render.cpp:44-79)
struct HitInfo {
    vec3 shapeNormal{NAN, NAN, NAN}, surfaceNormal{NAN, NAN, NAN}, emission{0, 0, 0}, baseColor{0, 0, 0}, position{NAN, NAN, NAN};
    float specular = 0.04f, roughness = 0.8f, metallic = 0.0f, opacity = 1.0f, eta = 1.0f;
    int id = 0;
    bool entering = true;
};
struct RayDifferential { vec3 dPdx{0, 0, 0}, dPdy{0, 0, 0}, dDdx{0, 0, 0}, dDdy{0, 0, 0}; };

bool reverseFix(vec3 &v, vec3 Dir) { if (dot(v, Dir) < 0.0f) { v = v * -1.0f; return false; } return true; }

void getHitNormals(const Face &f, vec3 inDir, vec3 bary, vec3 &shapeNormal, vec3 &raw, bool &entering) {
    vec3 crossV0 = cross(f.v[1] - f.v[0], f.v[2] - f.v[0]);
    shapeNormal = normalize(crossV0);
    entering = reverseFix(shapeNormal, -inDir);
    raw = shapeNormal;
    float area = length(crossV0) / 2.0f;
    if (area > 1e-2f) return;
    vec3 n0 = f.n[0], n1 = f.n[1], n2 = f.n[2];
    reverseFix(n0, shapeNormal); reverseFix(n1, shapeNormal); reverseFix(n2, shapeNormal);
    raw = normalize(bary.x * (dot(n0, shapeNormal) > 0.85f ? n0 : shapeNormal) + bary.y * (dot(n1, shapeNormal) > 0.85f ? n1 : shapeNormal) +
                    bary.z * (dot(n2, shapeNormal) > 0.85f ? n2 : shapeNormal));
    if (!finite_any(raw)) raw = shapeNormal;
}

void calc_dPdxy(const Ray &ray, float t, vec3 normal, const RayDifferential &bd, vec3 &dPdx, vec3 &dPdy) {
    float dtdx = -dot(bd.dPdx + t * bd.dDdx, normal) / dot(ray.direction, normal);
    float dtdy = -dot(bd.dPdy + t * bd.dDdy, normal) / dot(ray.direction, normal);
    dPdx = bd.dPdx + dtdx * ray.direction + t * bd.dDdx;
    dPdy = bd.dPdy + dtdy * ray.direction + t * bd.dDdy;
}

vec2 getDuv(const Face &f, vec3 dP) {
    vec3 b = barycentric(f.v[0], f.v[1], f.v[2], f.v[0] + dP);
    return (b.x - 1.0f) * f.uv[0] + b.y * f.uv[1] + b.z * f.uv[2];
}

void calcSurfaceNormal(const Face &f, vec3 nm, vec3 shapeNormal, vec3 &surfaceNormal) {
    vec3 e1 = f.v[1] - f.v[0], e2 = f.v[2] - f.v[0];
    vec2 d1 = f.uv[1] - f.uv[0], d2 = f.uv[2] - f.uv[0];
    float k = 1.0f / (d1.x * d2.y - d2.x * d1.y);
    vec3 tbU = k * (d2.y * e1 - d1.y * e2), tbV = k * (-d2.x * e1 + d1.x * e2);
    vec3 tangent = normalize(tbU - shapeNormal * dot(shapeNormal, tbU));
    vec3 bitangent = normalize(tbV - shapeNormal * dot(shapeNormal, tbV) - tangent * dot(tangent, tbV));
    vec3 sav = surfaceNormal;
    surfaceNormal = normalize(tangent * nm.x + bitangent * nm.y + surfaceNormal);
    if (!finite_any(surfaceNormal)) surfaceNormal = sav;
}

void getHitInfo(const Scene &S, const HitRecord &hit, const Ray &ray, const RayDifferential &bd, HitInfo &h) {
    const Face &f = S.faces[hit.face];
    vec3 bary = barycentric(f.v[0], f.v[1], f.v[2], h.position);
    getHitNormals(f, ray.direction, bary, h.shapeNormal, h.surfaceNormal, h.entering);
    vec3 raw = h.surfaceNormal, dPdx, dPdy;
    calc_dPdxy(ray, hit.t_max, h.shapeNormal, bd, dPdx, dPdy);
    vec2 uv = bary.x * f.uv[0] + bary.y * f.uv[1] + bary.z * f.uv[2];
    const Material &m = S.materials[f.material];
    h.id = m.id;
    getSurfaceData(m, uv.x, uv.y, h.roughness, h.metallic);
    h.opacity = m.opacity;
    h.eta = m.ior;
    if (h.opacity > 1.0f - eps_zero) h.entering = true;
    vec2 dUVdx = getDuv(f, dPdx), dUVdy = getDuv(f, dPdy);
    float duv = finite_any(dUVdx) ? (length(dUVdx) + length(dUVdy)) / 2.0f : NAN;
    if (h.opacity < eps_zero) h.baseColor = m.transmitting;
    else { vec4 c = getDiffuseColor(m, uv.x, uv.y, duv); h.baseColor = {c.x, c.y, c.z}; }
    h.emission = getEmissiveColor(m, uv.x, uv.y, duv);
    calcSurfaceNormal(f, getNormal(m, uv.x, uv.y, duv), h.shapeNormal, h.surfaceNormal);
    if (dot(h.surfaceNormal, ray.direction) >= 0.0f) h.surfaceNormal = raw;
    if (dot(h.surfaceNormal, ray.direction) >= 0.0f) h.surfaceNormal = h.shapeNormal;
}

// initRayDiff, src/render.cpp:439-446
void initRayDiff(vec3 d, const RmRenderArgs &a, RayDifferential &bd) {
    vec3 dddx = a.accuracy * v3(a.right), dddy = a.accuracy * v3(a.up);
    float dd = dot(d, d), ddx = dot(d, dddx), ddy = dot(d, dddy);
    bd.dDdx = div_scalar(dd * dddx - d * ddx, sqrtf(dd) * dd);
    bd.dDdy = div_scalar(dd * dddy - d * ddy, sqrtf(dd) * dd);
}

vec3 primaryD(const RmRenderArgs &a, int x, int y) {      // src/render.cpp:466-470
    float rayX = float(x) - float(a.width) / 2.0f, rayY = float(y) - float(a.height) / 2.0f;
    return v3(a.direction) + a.accuracy * (rayX * v3(a.right) + rayY * v3(a.up));
}

// ----------------------------------------------------------------- BSDF evaluation (src/sampling.cpp:10-199)
float sqr(float x) { return x * x; }
float clampf(float x, float a, float b) { return x < a ? a : (x > b ? b : x); }
float mixf(float a, float b, float t) { return a * (1 - t) + b * t; }
float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }
float SchlickFresnel(float u) { return pow5(clampf(1 - u, 0, 1)); }
float GTR1(float NdotH, float a) {
    if (a >= 1) return 1 / PI;
    float a2 = a * a, t = 1 + (a2 - 1) * NdotH * NdotH;
    return (a2 - 1) / (PI * logf(a2) * t);
}
float GTR2(float NdotH, float a) { float a2 = a * a, t = 1 + (a2 - 1) * NdotH * NdotH; return a2 / (PI * t * t); }
float smithG_GGX(float NdotV, float alphaG) { float a = alphaG * alphaG, b = NdotV * NdotV; return 1 / (NdotV + sqrtf(a + b - a * b)); }
float sqrt_s(float x) { return x <= 0.0f ? 0.0f : (float)std::sqrt((double)x); }     // geometry.cpp: sqrt resolves to the double version

vec3 getBRDF(const HitInfo &s, vec3 V, vec3 L) {
    vec3 N = s.surfaceNormal;
    float NdotL = dot(N, L), NdotV = dot(N, V);
    if (NdotL <= 0.0f || NdotV <= 0.0f) return v3(0.0f);
    vec3 H = normalize(L + V);
    float NdotH = dot(N, H), LdotH = dot(L, H);
    vec3 Cdlin = s.baseColor;
    float Cdlum = dot(Cdlin, RGB_Weight);
    const float subsurface = 0, specularTint = 0, sheen = 0, sheenTint = 0, clearcoat = 1.5f, clearcoatGloss = 0.2f, clearcoatTint = 0;
    vec3 Ctint = Cdlum > 0.0f ? div_scalar(Cdlin, Cdlum) : v3(1.0f);
    vec3 Cspec0 = mix(s.specular * mix(v3(1.0f), Ctint, specularTint), Cdlin, s.metallic);
    vec3 Csheen = mix(v3(1.0f), Ctint, sheenTint);
    float FL = SchlickFresnel(NdotL), FV = SchlickFresnel(NdotV);
    float Fd90 = 0.5f + 2.0f * LdotH * LdotH * s.roughness;
    float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
    float Fss90 = LdotH * LdotH * s.roughness;
    float Fss = mixf(1.0f, Fss90, FL) * mixf(1.0f, Fss90, FV);
    float ss = 1.25f * (Fss * (1.0f / (NdotL + NdotV) - .5f) + .5f);
    float Ds = GTR2(NdotH, s.roughness), FH;
    if (s.eta <= 1.0f) FH = SchlickFresnel(LdotH);
    else {
        float cosI = fabsf(LdotH), sinI = sqrt_s(1.0f - cosI * cosI), sinT = s.eta * sinI;
        if (sinT >= 1.0f) FH = 1.0f;
        else {
            float cosT = sqrt_s(1.0f - sinT * sinT), R0 = (s.eta - 1.0f) / (s.eta + 1.0f);
            R0 = R0 * R0;
            FH = mixf(R0, 1.0f, SchlickFresnel(cosT));
        }
    }
    vec3 Fs = mix(Cspec0, v3(1.0f), FH);
    float Gs = smithG_GGX(NdotL, s.roughness);
    Gs *= smithG_GGX(NdotV, s.roughness);
    vec3 Fsheen = FH * sheen * Csheen;
    float Dr = GTR1(NdotH, mixf(.1f, .001f, clearcoatGloss)), Fr = mixf(.04f, 1.0f, FH);
    float Gr = smithG_GGX(NdotL, .25f) * smithG_GGX(NdotV, .25f);
    vec3 ret = ((1.0f / PI) * mixf(Fd, ss, subsurface)) * Cdlin + Fsheen;
    ret = ret * (1.0f - s.metallic);
    ret = ret + (0.25f * clearcoat * Gr * Fr * Dr) * mix(v3(1.0f), Ctint, clearcoatTint);
    ret = ret * s.opacity;
    ret = ret + Fs * Ds * Gs;
    return ret * NdotL;
}

vec3 getBTDF(const HitInfo &s, vec3 V, vec3 L) {
    vec3 N = s.surfaceNormal;
    float NdotL = dot(N, L);
    if (NdotL >= 0.0f) return v3(0.0f);
    vec3 H = normalize(L + s.eta * V);
    float D = GTR2(dot(N, H), s.roughness);
    float btdf = D * (-NdotL);
    float LdotH = dot(L, H), NdotV = dot(N, V), HdotV = dot(H, V);
    btdf *= std::abs(LdotH * HdotV) / (std::abs(NdotL * NdotV) + eps_zero);
    float k = s.eta * HdotV + LdotH;
    btdf /= k * k;
    return v3(btdf);
}

vec3 getBSDF(const HitInfo &s, vec3 V, vec3 L) {
    if (s.opacity > 1.0f - eps_zero || s.entering) return getBRDF(s, V, L);
    return getBTDF(s, V, L);
}

// ----------------------------------------------------------------- accumulateInwardRadiance (src/image.cpp:615-659)
struct Radiance { vec3 radiance{0, 0, 0}; float Var = 0; };

void accumBasic(Radiance &r, vec3 in, float w) {
    if (!finite_any(in) || !std::isfinite(w)) return;
    r.radiance = r.radiance + in * w;
    r.Var += dot(in, in) * w;
}
void accumulateInwardRadiance(vec3 baseColor, vec3 bsdfPdf, vec3 light, float weight, Radiance &rd, Radiance &rs) {
    if (length(light) < eps_zero) return;
    vec3 base0 = normalize(baseColor);
    if (length(baseColor) < eps_zero) { accumBasic(rs, light * bsdfPdf, weight); return; }
    const vec3 White = normalize(v3(1.0f));
    float XdotY = dot(base0, White);
    if (XdotY > 0.99f) { accumBasic(rd, div_vec(light * bsdfPdf, baseColor), weight); return; }
    vec3 perp = normalize(cross(base0, White));
    vec3 bp = bsdfPdf - perp * dot(perp, bsdfPdf);
    float d1 = dot(bp, White), d2 = dot(bp, base0);
    float AplusB = (d1 + d2) / (1 + XdotY), AminusB = (d1 - d2) / (1 - XdotY), B = (AplusB - AminusB) / 2.0f;
    accumBasic(rd, div_scalar(light * B, length(baseColor)), weight);
    accumBasic(rs, light * (bsdfPdf - B * base0), weight);
}

// ----------------------------------------------------------------- FXAA (src/image.cpp:358-452)
void fxaa(const vec3 *data, vec3 *output, int width, int height) {
    auto lumaOf = [](vec3 c) { return dot(c, RGB_Weight); };
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            vec3 C = data[y * width + x];
            float M = lumaOf(C);
            float N = y > 0 ? lumaOf(data[(y - 1) * width + x]) : M, Sx = y < height - 1 ? lumaOf(data[(y + 1) * width + x]) : M;
            float E = x < width - 1 ? lumaOf(data[y * width + x + 1]) : M, W = x > 0 ? lumaOf(data[y * width + x - 1]) : M;
            float rangeMin = std::min({N, Sx, E, W}), rangeMax = std::max({N, Sx, E, W}), range = rangeMax - rangeMin;
            if (range < std::max(0.0312f, rangeMax * 0.125f)) { output[y * width + x] = C; continue; }
            float NW = (y > 0 && x > 0) ? lumaOf(data[(y - 1) * width + x - 1]) : M, NE = (y > 0 && x < width - 1) ? lumaOf(data[(y - 1) * width + x + 1]) : M;
            float SW = (y < height - 1 && x > 0) ? lumaOf(data[(y + 1) * width + x - 1]) : M, SE = (y < height - 1 && x < width - 1) ? lumaOf(data[(y + 1) * width + x + 1]) : M;
            float edgeHorz = std::abs((NW + W + SW) - (NE + E + SE)) * (1.0f / 3.0f), edgeVert = std::abs((NW + N + NE) - (SW + Sx + SE)) * (1.0f / 3.0f);
            bool isH = edgeHorz >= edgeVert;
            float stepLength = isH ? 1.0f / float(width) : 1.0f / float(height);
            float g = std::clamp((isH ? edgeHorz : edgeVert) / range, -2.0f, 2.0f);
            float u = float(x) / float(width), v = float(y) / float(height);
            vec3 finalColor = C;
            float bestDelta = 0.0f;
            for (int i = 0; i < 12; i++) {
                float ox = isH ? 0.0f : g * stepLength * float(i + 1), oy = isH ? g * stepLength * float(i + 1) : 0.0f;
                float su = u + ox, sv = v + oy;
                if (su < 0.0f || su > 1.0f || sv < 0.0f || sv > 1.0f) continue;
                int sx = std::clamp(int(su * float(width)), 0, width - 1), sy = std::clamp(int(sv * float(height)), 0, height - 1);
                vec3 sc = data[sy * width + sx];
                float delta = std::abs(lumaOf(sc) - M);
                if (delta > bestDelta) { bestDelta = delta; finalColor = sc; }
            }
            float sub = std::min((std::abs(N + Sx - 2.0f * M) * 2.0f + std::abs(E + W - 2.0f * M)) * 0.25f, 1.0f);
            output[y * width + x] = mix(C, finalColor, sub * 0.75f);
        }
}

// Photo::shade (src/image.cpp:215-246)
vec3 gammaPixel(vec3 pix);
vec3 shadePixel(const RmHitInfo &G, const RmRadiance *pl[4], size_t i, float exposure, int options) {
    vec3 pix;
    if (options & 64) pix = div_scalar(v3(G.shapeNormal) + v3(1.0f), 2.0f);
    else if (options & 128) pix = div_scalar(v3(G.surfaceNormal) + v3(1.0f), 2.0f);
    else {
        vec3 rd = v3(0.0f), rs = v3(0.0f);
        if (options & 4) { if (options & 16) rd = rd + v3(pl[0][i].radiance); if (options & 32) rs = rs + v3(pl[1][i].radiance); }
        if (options & 8) { if (options & 16) rd = rd + v3(pl[2][i].radiance); if (options & 32) rs = rs + v3(pl[3][i].radiance); }
        if (!(options & (4 | 8))) rd = v3(1.0f);
        vec3 dc = (options & 1) ? v3(G.baseColor) : v3(1.0f);
        pix = dc * rd + rs;
        if (options & 2) pix = pix + v3(G.emission) * exposure;
    }
    return pix;
}

// Photo::gammaCorrection (src/image.cpp:454-468)
vec3 gammaPixel(vec3 pix) {
    auto mx = [](float x, float y) { return (x < y) ? y : x; };
    auto mn = [](float x, float y) { return (y < x) ? y : x; };
    pix = {mx(pix.x, 0.0f), mx(pix.y, 0.0f), mx(pix.z, 0.0f)};
    float C = dot(pix, RGB_Weight);
    if (C > 0.75f) pix = div_scalar(pix, C) * (tanhf(3.0f * (C - 0.75f)) / 3.0f + 0.75f);
    pix = {mn(pix.x, 1.0f), mn(pix.y, 1.0f), mn(pix.z, 1.0f)};
    const float ig = 1.0f / 2.2f;
    return {powf(pix.x, ig), powf(pix.y, ig), powf(pix.z, ig)};
}

template <typename F>
void parallelRows(int height, int threads, F fn) {
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int i = 0; i < std::max(1, threads); i++)
        pool.emplace_back([&]() { for (;;) { int y = next.fetch_add(1); if (y >= height) break; fn(y); } });
    for (auto &t : pool) t.join();
}

HitInfo fromRm(const RmHitInfo &g) {
    HitInfo h;
    h.shapeNormal = v3(g.shapeNormal); h.surfaceNormal = v3(g.surfaceNormal); h.emission = v3(g.emission);
    h.baseColor = v3(g.baseColor); h.position = v3(g.position);
    h.specular = g.specular; h.roughness = g.roughness; h.metallic = g.metallic; h.opacity = g.opacity; h.eta = g.eta;
    h.id = g.id; h.entering = g.entering != 0;
    return h;
}
void toRm(const HitInfo &h, RmHitInfo &g) {
    std::memset(&g, 0, sizeof(g));
    auto put = [](float *d, vec3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
    put(g.shapeNormal, h.shapeNormal); put(g.surfaceNormal, h.surfaceNormal); put(g.emission, h.emission);
    put(g.baseColor, h.baseColor); put(g.position, h.position);
    g.specular = h.specular; g.roughness = h.roughness; g.metallic = h.metallic; g.opacity = h.opacity; g.eta = h.eta;
    g.id = h.id; g.entering = h.entering ? 1 : 0;
}

} // namespace port

using namespace port;

extern "C" void port_bloom(float *rgb, int width, int height);

extern "C" {

void *port_scene_create(const RmRawScene *raw) {
    auto *S = new Scene();
    S->materials.resize(raw->n_materials);
    for (int i = 0; i < raw->n_materials; i++) {
        const RmRawMaterial &r = raw->materials[i];
        Material &m = S->materials[i];
        const int t[4] = {r.tex_diffuse, r.tex_specular, r.tex_emissive, r.tex_normals};
        for (int k = 0; k < 4; k++) {
            if (t[k] < 0) continue;
            const RmRawTexture &rt = raw->textures[t[k]];
            Texture &tx = m.tex[k];
            tx.width = rt.width; tx.height = rt.height; tx.channels = rt.channels;
            tx.data[0].assign(rt.pixels, rt.pixels + size_t(rt.width) * rt.height * rt.channels);
            generateMipmaps(tx);
        }
        m.id = i; m.opacity = r.opacity; m.ior = r.ior; m.roughness = r.roughness; m.transmitting = v3(r.transmitting_color);
        for (size_t b = 3; b < m.tex[0].data[0].size(); b += 4)       // hasTransparentPart, src/material.cpp:102-107
            if (m.tex[0].data[0][b] < 255) { m.cutout = true; break; }
    }
    S->faces.resize(raw->n_faces);
    for (int k = 0; k < raw->n_meshes; k++)
        for (int f = raw->meshes[k].face_begin; f < raw->meshes[k].face_end; f++) {
            Face &F = S->faces[f];
            for (int c = 0; c < 3; c++) {
                F.v[c] = v3(raw->positions + (size_t(f) * 3 + c) * 3);
                F.uv[c] = {raw->uvs[(size_t(f) * 3 + c) * 2], raw->uvs[(size_t(f) * 3 + c) * 2 + 1]};
                F.n[c] = v3(raw->normals + (size_t(f) * 3 + c) * 3);
            }
            F.material = raw->meshes[k].material;
            F.original = f;
        }
    if (raw->sky_rgb && raw->sky_width > 0) {      // SkyBox::Init, src/component.cpp:54-67
        S->sky_w = raw->sky_width; S->sky_h = raw->sky_height;
        size_t n = size_t(S->sky_w) * S->sky_h;
        S->sky.resize(n); S->sky_cdf.resize(n);
        for (int v = 0; v < S->sky_h; v++)
            for (int u = 0; u < S->sky_w; u++) {
                float phi = PI * (float(v) + 0.5f) / float(S->sky_h);
                float area = sinf(phi) * 2.0f * PI / float(S->sky_w * S->sky_h);
                size_t id = size_t(v) * S->sky_w + u;
                S->sky[id] = v3(raw->sky_rgb + id * 3) * area;
                float C = dot(S->sky[id], RGB_Weight);
                S->sky_cdf[id] = id == 0 ? C : S->sky_cdf[id - 1] + C;
            }
    }
    S->nodes.assign(size_t(nodeCount(1, raw->n_faces)) + 1, Node{});
    dfs_build(*S, 1, 0, raw->n_faces);
    return S;
}
void port_scene_destroy(void *h) { delete static_cast<Scene *>(h); }
int port_node_count(void *h) { return int(static_cast<Scene *>(h)->nodes.size()); }

void port_bvh_export(void *h, RmBvhNode *nodes, int32_t *perm) {
    auto *S = static_cast<Scene *>(h);
    std::memset(nodes, 0, sizeof(RmBvhNode) * S->nodes.size());
    std::vector<int> st{1};
    while (!st.empty()) {
        int u = st.back(); st.pop_back();
        const Node &n = S->nodes[u];
        RmBvhNode &o = nodes[u];
        o.v0[0] = n.box.v0.x; o.v0[1] = n.box.v0.y; o.v0[2] = n.box.v0.z;
        o.v1[0] = n.box.v1.x; o.v1[1] = n.box.v1.y; o.v1[2] = n.box.v1.z;
        o.faceL = n.faceL; o.faceR = n.faceR;
        if (!n.faceR) { st.push_back(u << 1); st.push_back(u << 1 | 1); }
    }
    for (size_t i = 0; i < S->faces.size(); i++) perm[i] = S->faces[i].original;
}

void port_trace_primary(void *h, const RmRenderArgs *a, int threads, int32_t *tri_idx, float *t, uint64_t *counters3) {
    const Scene &S = *static_cast<Scene *>(h);
    std::atomic<uint64_t> c0(0), c1(0), c2(0);
    parallelRows(a->height, threads, [&](int y) {
        Counters before = tl;
        for (int x = 0; x < a->width; x++) {
            Ray ray = {v3(a->position), normalize(primaryD(*a, x, y))};
            HitRecord hit = modelRayHit(S, ray);
            tri_idx[y * a->width + x] = hit.t_max == INFINITY ? -1 : hit.face;
            t[y * a->width + x] = hit.t_max;
        }
        c0 += tl.rays - before.rays; c1 += tl.box - before.box; c2 += tl.tri - before.tri;
    });
    if (counters3) { counters3[0] = c0; counters3[1] = c1; counters3[2] = c2; }
}

void port_trace_closest(void *h, int64_t n, const float *org, const float *dir, int32_t *tri_idx, float *t, uint64_t *counters3) {
    const Scene &S = *static_cast<Scene *>(h);
    Counters before = tl;
    for (int64_t i = 0; i < n; i++) {
        HitRecord hit = modelRayHit(S, {v3(org + i * 3), v3(dir + i * 3)});
        tri_idx[i] = hit.t_max == INFINITY ? -1 : hit.face;
        t[i] = hit.t_max;
    }
    if (counters3) { counters3[0] = tl.rays - before.rays; counters3[1] = tl.box - before.box; counters3[2] = tl.tri - before.tri; }
}

void port_trace_occluded(void *h, int64_t n, const float *org, const float *dir, const float *aim, uint8_t *out) {
    const Scene &S = *static_cast<Scene *>(h);
    for (int64_t i = 0; i < n; i++) out[i] = modelRayHitTest(S, {v3(org + i * 3), v3(dir + i * 3)}, aim[i]) ? 1 : 0;
}

// renderPixel with spp = 0 (src/render.cpp:466-495): G-buffer incl. the red nudge (kept: the early return at 529-530)
void port_gbuffer(void *h, const RmRenderArgs *a, int threads, RmHitInfo *out) {
    const Scene &S = *static_cast<Scene *>(h);
    parallelRows(a->height, threads, [&](int y) {
        for (int x = 0; x < a->width; x++) {
            vec3 d = primaryD(*a, x, y);
            RayDifferential bd;
            initRayDiff(d, *a, bd);
            vec3 Dir = normalize(d);
            Ray ray = {v3(a->position), Dir};
            HitInfo G;
            HitRecord hit = modelRayHit(S, ray);
            if (hit.t_max == INFINITY) {
                G.position = v3(NAN);
                if (!S.sky.empty()) G.emission = skyGet(S, Dir);
            } else {
                G.position = ray.origin + hit.t_max * ray.direction;
                getHitInfo(S, hit, ray, bd, G);
                float C0 = dot(G.baseColor, RGB_Weight);
                if (C0 < 2e-2 || (C0 < 0.8f && length(div_scalar(G.baseColor, C0) - v3(1.0f)) < 2e-2)) G.baseColor.x += 4e-2;
            }
            toRm(G, out[y * a->width + x]);
        }
    });
}

void port_fxaa(const float *in, float *out, int width, int height) {
    fxaa(reinterpret_cast<const vec3 *>(in), reinterpret_cast<vec3 *>(out), width, height);
}

void port_postprocess(const RmHitInfo *g, const RmRadiance *Dd, const RmRadiance *Ds, const RmRadiance *Id, const RmRadiance *Is,
                      int width, int height, float exposure, int options, float *rgb_out) {
    const RmRadiance *pl[4] = {Dd, Ds, Id, Is};
    size_t n = size_t(width) * height;
    std::vector<vec3> img(n), tmp;
    for (size_t i = 0; i < n; i++) img[i] = shadePixel(g[i], pl, i, exposure, options);
    if (options & 256) port_bloom(reinterpret_cast<float *>(img.data()), width, height);      // DoBloom (port_post.cpp)
    for (size_t i = 0; i < n; i++) img[i] = gammaPixel(img[i]);
    if (options & 512) { tmp.resize(n); fxaa(img.data(), tmp.data(), width, height); img.swap(tmp); }
    std::memcpy(rgb_out, img.data(), n * sizeof(vec3));
}

void port_kat_ray_in_box(int64_t n, const float *rays, const float *boxes, float *tlr) {
    for (int64_t i = 0; i < n; i++) rayInBox({v3(rays + i * 6), v3(rays + i * 6 + 3)}, {v3(boxes + i * 6), v3(boxes + i * 6 + 3)}, tlr[i * 2], tlr[i * 2 + 1]);
}
void port_kat_ray_triangle(int64_t n, const float *rays, const float *tris, float *t) {
    for (int64_t i = 0; i < n; i++) t[i] = rayTriangle({v3(rays + i * 6), v3(rays + i * 6 + 3)}, v3(tris + i * 9), v3(tris + i * 9 + 3), v3(tris + i * 9 + 6));
}
void port_kat_barycentric(int64_t n, const float *tris, const float *p, float *out) {
    for (int64_t i = 0; i < n; i++) { vec3 b = barycentric(v3(tris + i * 9), v3(tris + i * 9 + 3), v3(tris + i * 9 + 6), v3(p + i * 3)); out[i * 3] = b.x; out[i * 3 + 1] = b.y; out[i * 3 + 2] = b.z; }
}
void port_kat_bsdf(int which, int64_t n, const RmHitInfo *surf, const float *in_dirs, const float *out_dirs, float *out) {
    for (int64_t i = 0; i < n; i++) {
        HitInfo s = fromRm(surf[i]);
        vec3 V = v3(in_dirs + i * 3), L = v3(out_dirs + i * 3);
        vec3 c = which == 0 ? getBSDF(s, V, L) : (which == 1 ? getBRDF(s, V, L) : getBTDF(s, V, L));
        out[i * 3] = c.x; out[i * 3 + 1] = c.y; out[i * 3 + 2] = c.z;
    }
}
void port_kat_accumulate(int64_t n, const float *base, const float *s7, float *out) {
    for (int64_t i = 0; i < n; i++) {
        Radiance d, s;
        accumulateInwardRadiance(v3(base + i * 3), v3(s7 + i * 7), v3(s7 + i * 7 + 3), s7[i * 7 + 6], d, s);
        float *o = out + i * 8;
        o[0] = d.radiance.x; o[1] = d.radiance.y; o[2] = d.radiance.z; o[3] = d.Var;
        o[4] = s.radiance.x; o[5] = s.radiance.y; o[6] = s.radiance.z; o[7] = s.Var;
    }
}
void port_kat_material_fetch(void *h, int material, int which, int64_t n, const float *uvd, float *out) {
    const Material &m = static_cast<Scene *>(h)->materials[material];
    for (int64_t i = 0; i < n; i++) {
        float u = uvd[i * 3], v = uvd[i * 3 + 1], d = uvd[i * 3 + 2], *o = out + i * 4;
        o[0] = o[1] = o[2] = o[3] = 0.0f;
        if (which == 0) { vec4 c = getDiffuseColor(m, u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w; }
        else if (which == 1) { vec3 c = getEmissiveColor(m, u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
        else if (which == 2) { vec3 c = getNormal(m, u, v, d); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
        else getSurfaceData(m, u, v, o[0], o[1]);
    }
}
void port_kat_sky_get(void *h, int64_t n, const float *dirs, float *out) {
    const Scene &S = *static_cast<Scene *>(h);
    for (int64_t i = 0; i < n; i++) { vec3 c = skyGet(S, v3(dirs + i * 3)); out[i * 3] = c.x; out[i * 3 + 1] = c.y; out[i * 3 + 2] = c.z; }
}
// Generator::operator() on raw 32-bit draws (src/component.cpp:5-10 over libstdc++ generate_canonical)
void port_uniform_from_u32(const uint32_t *u32, int n, float *out) {
    const float a = 1e-6f, b = 1.0f - 1e-6f;
    for (int i = 0; i < n; i++) {
        float u = float(u32[i]) / 4294967296.0f;
        if (u >= 1.0f) u = std::nextafter(1.0f, 0.0f);
        out[i] = u * (b - a) + a;
    }
}

} // extern "C"
