// dev_texture.cuh — texture, material and environment lookups on the device.
//
// Restates (not translates) the reference's fetch rules:
//   get_basic / get_bilinear / ImageData::get        src/material.cpp:26-94
//   Material::getDiffuseColor / getNormal / getEmissiveColor / getSurfaceData, gammaPow   :337-383
//   SkyBox::get                                       src/component.cpp:120-140
// Addressing is wrap, v flipped, no half-texel offset; RGBA8 fetches divide by 255
// (glm vec4 / scalar is a true division) while RGB8 fetches multiply by 1/255 (glm
// vec3 / scalar), see dev_math.cuh.
#pragma once
#include "dev_scene.cuh"

namespace rm {

struct V4 { float x, y, z, w; };

RM_DI int wrap_index(int x, int m) {           // the `_mod` lambda, src/material.cpp:51-56
    if ((m & (m - 1)) == 0) return x & (m - 1);            // power-of-two size: the two's-complement AND is the mathematical mod
    if (x < 0 || x >= m) { x %= m; if (x < 0) x += m; }
    return x;
}
RM_DI int wrap_next(int x, int m) { return x + 1 == m ? 0 : x + 1; }          // (x + 1) % m for 0 <= x < m

RM_DI float lerp2(float a, float b, float t) { return fadd(fmul(a, fsub(1.0f, t)), fmul(b, t)); }

// bilinear RGBA8: all four channels.  `lut` holds k / 255.0f (glm vec4 / scalar is a true division; the
// table entries are that quotient, correctly rounded, so the decode is bit-identical without dividing)
RM_DI V4 bilinear_rgba(const uint8_t *data, const float *__restrict__ lut, int w, int h, float u, float v) {
    float x = fmul(u, float(w)), y = fmul(v, float(h));
    int x0 = int(floorf(x)), y0 = int(floorf(y));
    float dx = fsub(x, float(x0)), dy = fsub(y, float(y0));
    x0 = wrap_index(x0, w);
    y0 = wrap_index(y0, h);
    int x1 = wrap_next(x0, w), y1 = wrap_next(y0, h);
    const uchar4 *p = reinterpret_cast<const uchar4 *>(data);
    uchar4 t00 = __ldg(p + (y0 * w + x0)), t01 = __ldg(p + (y0 * w + x1));
    uchar4 t10 = __ldg(p + (y1 * w + x0)), t11 = __ldg(p + (y1 * w + x1));
#define RM_CH(c) lerp2(lerp2(__ldg(lut + t00.c), __ldg(lut + t01.c), dx), lerp2(__ldg(lut + t10.c), __ldg(lut + t11.c), dx), dy)
    V4 r;
    r.x = RM_CH(x); r.y = RM_CH(y); r.z = RM_CH(z); r.w = RM_CH(w);
#undef RM_CH
    return r;
}

// bilinear RGB8 (stride 3)
RM_DI V3 bilinear_rgb(const uint8_t *data, int w, int h, float u, float v) {
    float x = fmul(u, float(w)), y = fmul(v, float(h));
    int x0 = int(floorf(x)), y0 = int(floorf(y));
    float dx = fsub(x, float(x0)), dy = fsub(y, float(y0));
    x0 = wrap_index(x0, w);
    y0 = wrap_index(y0, h);
    int x1 = wrap_next(x0, w), y1 = wrap_next(y0, h);
    const uint8_t *p00 = data + (y0 * w + x0) * 3, *p01 = data + (y0 * w + x1) * 3;
    const uint8_t *p10 = data + (y1 * w + x0) * 3, *p11 = data + (y1 * w + x1) * 3;
    const float r255 = frcp(255.0f);
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
        c[k] = lerp2(lerp2(fmul(float(__ldg(p00 + k)), r255), fmul(float(__ldg(p01 + k)), r255), dx),
                     lerp2(fmul(float(__ldg(p10 + k)), r255), fmul(float(__ldg(p11 + k)), r255), dx), dy);
    return mk3(c[0], c[1], c[2]);
}

// ImageData::get level selection (src/material.cpp:82-88): returns level, nextLevel, blend
RM_DI void mip_select(const DevTexture &t, float depth, int &level, int &next, float &blend) {
    float top = float(t.map_depth - 1);
    depth = (top < depth) ? top : depth;          // std::min(depth, top)
    depth = (depth < 0.0f) ? 0.0f : depth;        // std::max(depth, 0.0f)
    level = int(depth);
    next = min(level + 1, t.map_depth - 1);
    blend = fsub(depth, float(level));
}

// ImageData::get<vec4> (src/material.cpp:81-94): trilinear = two bilinear fetches blended by the LOD fraction
RM_NI V4 texture_rgba_impl(const DevTexture *__restrict__ textures, const uint8_t *__restrict__ texels, const float *__restrict__ lut,
                           int tex, float u, float v, float depth) {
    const DevTexture t = textures[tex];
    v = fsub(1.0f, v);
    int level, next;
    float blend;
    mip_select(t, depth, level, next, blend);
    // a zero LOD fraction (magnified or past the last level: most fetches) makes the blend a * 1 + b * 0 = a exactly
    // (texels are finite), so the second level is not fetched
    V4 lv[2];
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        if (k == 1 && blend == 0.0f) { lv[1] = lv[0]; break; }
        const int l = k ? next : level;
        lv[k] = bilinear_rgba(texels + t.offset[l], lut, t.width >> l, t.height >> l, u, v);
    }
    if (blend == 0.0f) return lv[0];
    V4 r;
    r.x = lerp2(lv[0].x, lv[1].x, blend); r.y = lerp2(lv[0].y, lv[1].y, blend);
    r.z = lerp2(lv[0].z, lv[1].z, blend); r.w = lerp2(lv[0].w, lv[1].w, blend);
    return r;
}
RM_DI V4 texture_rgba(const DevScene &S, int tex, float u, float v, float depth) {
    return texture_rgba_impl(S.textures, S.texels, S.div255, tex, u, v, depth);
}

RM_DI V3 texture_rgb(const DevScene &S, int tex, float u, float v, float depth) {   // one call site (normal maps)
    const DevTexture t = S.textures[tex];
    v = fsub(1.0f, v);
    int level, next;
    float blend;
    mip_select(t, depth, level, next, blend);
    V3 a = bilinear_rgb(S.texels + t.offset[level], t.width >> level, t.height >> level, u, v);
    if (blend == 0.0f) return a;               // a * 1 + b * 0 = a exactly
    V3 b = bilinear_rgb(S.texels + t.offset[next], t.width >> next, t.height >> next, u, v);
    return mk3(lerp2(a.x, b.x, blend), lerp2(a.y, b.y, blend), lerp2(a.z, b.z, blend));
}

// duv = NaN means "no mip-mapping" (src/material.cpp:348)
RM_DI float lod_of(const DevScene &S, int tex, float duv) {
    return isnan(duv) ? 0.0f : log2f(fmul(duv, float(S.textures[tex].width)));
}

// gammaPow (src/material.cpp:337-346)
RM_NI float gamma_pow(float c) { return powf(c, 2.2f); }

RM_DI V4 mat_diffuse(const DevScene &S, const DevMaterial &m, float u, float v, float duv) {
    V4 c;
    if (m.tex[0] < 0) { c.x = c.y = c.z = c.w = 1.0f; return c; }
    c = texture_rgba(S, m.tex[0], u, v, lod_of(S, m.tex[0], duv));
    c.x = gamma_pow(c.x); c.y = gamma_pow(c.y); c.z = gamma_pow(c.z);
    return c;
}

// alpha of the diffuse texture at LOD 0: all TransparentTest needs (src/model.cpp:228-229)
RM_DI float mat_diffuse_alpha0(const DevScene &S, const DevMaterial &m, float u, float v) {
    if (m.tex[0] < 0) return 1.0f;
    return texture_rgba(S, m.tex[0], u, v, 0.0f).w;
}

RM_DI V3 mat_emissive(const DevScene &S, const DevMaterial &m, float u, float v, float duv) {
    if (m.tex[2] < 0) return splat3(0.0f);
    V4 c = texture_rgba(S, m.tex[2], u, v, lod_of(S, m.tex[2], duv));
    return mk3(gamma_pow(c.x), gamma_pow(c.y), gamma_pow(c.z));
}

RM_DI V3 mat_normal(const DevScene &S, const DevMaterial &m, float u, float v, float duv) {
    if (m.tex[3] < 0) return splat3(0.0f);
    V3 c = texture_rgb(S, m.tex[3], u, v, lod_of(S, m.tex[3], duv));
    return mk3(fsub(fmul(c.x, 2.0f), 1.0f), fsub(fmul(c.y, 2.0f), 1.0f), fsub(fmul(c.z, 2.0f), 1.0f));
}

RM_DI void mat_surface(const DevScene &S, const DevMaterial &m, float u, float v, float &roughness, float &metallic) {
    if (m.tex[1] < 0) { metallic = 0.0f; roughness = m.roughness; return; }
    V4 c = texture_rgba(S, m.tex[1], u, v, 0.0f);
    metallic = (0.99f < c.z) ? 0.99f : c.z;        // std::min(surfaceData[2], 0.99f)
    roughness = (c.y < 1e-3f) ? 1e-3f : c.y;       // std::max(surfaceData[1], 1e-3f)
}

// SkyBox::get: nearest texel of the premultiplied map, divided by the texel's solid angle again
RM_DI V3 sky_get(const DevScene &S, V3 dir) {
    if (S.sky_width == 0) return splat3(0.0f);
    float theta = atan2f(-dir.x, dir.z);
    float phi = acosf(dir.y);
    if (theta < 0.0f) theta = fadd(theta, fmul(2.0f, kPi));
    int u = int(fmul(fdiv(theta, fmul(2.0f, kPi)), float(S.sky_width)));
    int v = int(fmul(fdiv(phi, kPi), float(S.sky_height)));
    u = u < 0 ? 0 : (u >= S.sky_width ? S.sky_width - 1 : u);
    v = v < 0 ? 0 : (v >= S.sky_height ? S.sky_height - 1 : v);
    phi = fdiv(fmul(kPi, fadd(float(v), 0.5f)), float(S.sky_height));
    float area = fdiv(fmul(fmul(sinf(phi), 2.0f), kPi), float(S.sky_width * S.sky_height));
    const float *d = S.sky_data + (size_t(v) * S.sky_width + u) * 3;
    return div_recip(mk3(__ldg(d), __ldg(d + 1), __ldg(d + 2)), area);
}

} // namespace rm
