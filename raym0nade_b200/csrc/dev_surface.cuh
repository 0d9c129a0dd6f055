// dev_surface.cuh — hit record -> surface record (the reference's HitInfo).
//
// Restates, in the reference's operation order:
//   getHitNormals / reverseFix / getDuv / calcSurfaceNormal / getHitMaterial   src/model.cpp:232-328
//   calc_dPdxy / calc_dDdxy / getHitInfo / initRayDiff                          src/render.cpp:44-79,439-446
#pragma once
#include "dev_trace.cuh"

namespace rm {

// HitInfo (include/model.h:12-18) in registers.
struct Surface {
    V3 shapeNormal, surfaceNormal, emission, baseColor, position;
    float specular, roughness, metallic, opacity, eta;
    int id;
    bool entering;
};

RM_DI Surface default_surface() {          // HitInfo::HitInfo(), src/model.cpp:4-7
    Surface s;
    s.position = s.shapeNormal = s.surfaceNormal = splat3(CUDART_NAN_F);
    s.opacity = 1.0f; s.specular = 0.04f; s.roughness = 0.8f; s.metallic = 0.0f; s.eta = 1.0f;
    s.emission = splat3(0.0f); s.baseColor = splat3(0.0f);
    s.entering = true; s.id = 0;
    return s;
}

struct RayDiff { V3 dPdx, dPdy, dDdx, dDdy; };

RM_DI bool reverse_fix(V3 &v, V3 dir) {
    if (dot(v, dir) < 0.0f) { v = v * -1.0f; return false; }
    return true;
}

// getHitNormals (src/model.cpp:240-270)
RM_DI void hit_normals(const V3 v[3], const V3 n[3], V3 inDir, V3 bary, V3 &shapeNormal, V3 &surfaceNormal_raw, bool &entering) {
    V3 crossV0 = cross(v[1] - v[0], v[2] - v[0]);
    shapeNormal = normalize(crossV0);
    entering = reverse_fix(shapeNormal, -inDir);
    surfaceNormal_raw = shapeNormal;
    float area = fdiv(length(crossV0), 2.0f);
    if (area > 1e-2f) return;
    V3 n0 = n[0], n1 = n[1], n2 = n[2];
    reverse_fix(n0, shapeNormal);
    reverse_fix(n1, shapeNormal);
    reverse_fix(n2, shapeNormal);
    V3 a = bary.x * (dot(n0, shapeNormal) > 0.85f ? n0 : shapeNormal);
    V3 b = bary.y * (dot(n1, shapeNormal) > 0.85f ? n1 : shapeNormal);
    V3 c = bary.z * (dot(n2, shapeNormal) > 0.85f ? n2 : shapeNormal);
    surfaceNormal_raw = normalize((a + b) + c);
    if (!isfinite_any(surfaceNormal_raw)) surfaceNormal_raw = shapeNormal;
}

// calc_dPdxy (src/render.cpp:44-50)
RM_DI void calc_dPdxy(V3 dir, float hit_t, V3 normal, const RayDiff &bd, V3 &dPdx, V3 &dPdy) {
    float dn = dot(dir, normal);
    float dtdx = fdiv(-dot(bd.dPdx + hit_t * bd.dDdx, normal), dn);
    float dtdy = fdiv(-dot(bd.dPdy + hit_t * bd.dDdy, normal), dn);
    dPdx = (bd.dPdx + dtdx * dir) + hit_t * bd.dDdx;
    dPdy = (bd.dPdy + dtdy * dir) + hit_t * bd.dDdy;
}

// calc_dDdxy (src/render.cpp:52-60); hit_dNdx = hit_dNdy = 0
RM_DI void calc_dDdxy(V3 dir, V3 normal, const RayDiff &bd, V3 &dDdx, V3 &dDdy) {
    const V3 z = splat3(0.0f);
    float dDNdx = fadd(dot(z, dir), dot(bd.dDdx, normal));
    float dDNdy = fadd(dot(z, dir), dot(bd.dDdy, normal));
    float dn = dot(dir, normal);
    dDdx = bd.dDdx - 2.0f * (dn * z + dDNdx * normal);
    dDdy = bd.dDdy - 2.0f * (dn * z + dDNdy * normal);
}

// initRayDiff (src/render.cpp:439-446); d is the UN-normalised primary direction
RM_DI RayDiff init_ray_diff(V3 d, const DevArgs &A) {
    RayDiff r;
    r.dPdx = splat3(0.0f); r.dPdy = splat3(0.0f);
    V3 dddx = A.accuracy * A.right, dddy = A.accuracy * A.up;
    float dd = dot(d, d), ddx = dot(d, dddx), ddy = dot(d, dddy);
    float den = fmul(fsqrt(dd), dd);
    r.dDdx = div_recip(dd * dddx - d * ddx, den);
    r.dDdy = div_recip(dd * dddy - d * ddy, den);
    return r;
}

// getDuv (src/model.cpp:272-277)
RM_DI V2 get_duv(const FaceShade &F, V3 dP) {
    V3 b = barycentric(F.v[0], F.v[1], F.v[2], F.v[0] + dP);
    return (fsub(b.x, 1.0f) * F.uv[0] + b.y * F.uv[1]) + b.z * F.uv[2];
}

// calcSurfaceNormal (src/model.cpp:279-298)
RM_DI void calc_surface_normal(const FaceShade &F, V3 normalMap, V3 shapeNormal, V3 &surfaceNormal) {
    V3 edge1 = F.v[1] - F.v[0], edge2 = F.v[2] - F.v[0];
    V2 d1 = F.uv[1] - F.uv[0], d2 = F.uv[2] - F.uv[0];
    float f = frcp(fsub(fmul(d1.x, d2.y), fmul(d2.x, d1.y)));
    V3 tbU = f * (d2.y * edge1 - d1.y * edge2);
    V3 tbV = f * ((-d2.x) * edge1 + d1.x * edge2);
    V3 tangent = normalize(tbU - shapeNormal * dot(shapeNormal, tbU));
    V3 bitangent = normalize((tbV - shapeNormal * dot(shapeNormal, tbV)) - tangent * dot(tangent, tbV));
    V3 sav = surfaceNormal;
    surfaceNormal = normalize((tangent * normalMap.x + bitangent * normalMap.y) + surfaceNormal);
    if (!isfinite_any(surfaceNormal)) surfaceNormal = sav;
}

// getHitInfo (src/render.cpp:62-79).  s.position must already hold the hit point.
RM_DI void get_hit_info(const DevScene &S, int face, float hit_t, V3 dir, const RayDiff &bd, V3 &dPdx, V3 &dPdy, Surface &s) {
    const FaceShade F = load_face(S, face);
    const V3 bary = barycentric(F.v[0], F.v[1], F.v[2], s.position);
    hit_normals(F.v, F.n, dir, bary, s.shapeNormal, s.surfaceNormal, s.entering);
    const V3 raw = s.surfaceNormal;
    calc_dPdxy(dir, hit_t, s.shapeNormal, bd, dPdx, dPdy);
    // getHitMaterial (src/model.cpp:300-328)
    const V2 uv = interp_uv(F, bary);
    const DevMaterial m = S.materials[F.material];
    s.id = F.material;
    mat_surface(S, m, uv.x, uv.y, s.roughness, s.metallic);
    s.opacity = m.opacity;
    s.eta = m.ior;
    if (s.opacity > fsub(1.0f, kEps)) s.entering = true;
    V2 dUVdx = get_duv(F, dPdx), dUVdy = get_duv(F, dPdy);
    float duv = isfinite_any(dUVdx) ? fdiv(fadd(length(dUVdx), length(dUVdy)), 2.0f) : CUDART_NAN_F;
    if (s.opacity < kEps) s.baseColor = mk3(m.tc[0], m.tc[1], m.tc[2]);
    else { V4 c = mat_diffuse(S, m, uv.x, uv.y, duv); s.baseColor = mk3(c.x, c.y, c.z); }
    s.emission = mat_emissive(S, m, uv.x, uv.y, duv);
    V3 nm = mat_normal(S, m, uv.x, uv.y, duv);
    calc_surface_normal(F, nm, s.shapeNormal, s.surfaceNormal);
    if (dot(s.surfaceNormal, dir) >= 0.0f) s.surfaceNormal = raw;
    if (dot(s.surfaceNormal, dir) >= 0.0f) s.surfaceNormal = s.shapeNormal;
}

// ---- the reference's 88-byte HitInfo image in global memory (RmHitInfo) ----
RM_DI void store_hitinfo(RmHitInfo *g, const Surface &s) {
    float *f = reinterpret_cast<float *>(g);
    f[0] = s.shapeNormal.x; f[1] = s.shapeNormal.y; f[2] = s.shapeNormal.z;
    f[3] = s.surfaceNormal.x; f[4] = s.surfaceNormal.y; f[5] = s.surfaceNormal.z;
    f[6] = s.emission.x; f[7] = s.emission.y; f[8] = s.emission.z;
    f[9] = s.baseColor.x; f[10] = s.baseColor.y; f[11] = s.baseColor.z;
    f[12] = s.position.x; f[13] = s.position.y; f[14] = s.position.z;
    f[15] = s.specular; f[16] = s.roughness; f[17] = s.metallic; f[18] = s.opacity; f[19] = s.eta;
    reinterpret_cast<int *>(g)[20] = s.id;
    reinterpret_cast<int *>(g)[21] = s.entering ? 1 : 0;
}

RM_DI Surface load_hitinfo(const RmHitInfo *g) {
    const float *f = reinterpret_cast<const float *>(g);
    Surface s;
    s.shapeNormal = mk3(f[0], f[1], f[2]);
    s.surfaceNormal = mk3(f[3], f[4], f[5]);
    s.emission = mk3(f[6], f[7], f[8]);
    s.baseColor = mk3(f[9], f[10], f[11]);
    s.position = mk3(f[12], f[13], f[14]);
    s.specular = f[15]; s.roughness = f[16]; s.metallic = f[17]; s.opacity = f[18]; s.eta = f[19];
    s.id = reinterpret_cast<const int *>(g)[20];
    s.entering = (reinterpret_cast<const int *>(g)[21] & 0xff) != 0;
    return s;
}

} // namespace rm
