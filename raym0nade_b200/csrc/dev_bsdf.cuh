// dev_bsdf.cuh — BSDF evaluation and importance sampling, light / environment sampling.
//
// Restates the reference's sampling.cpp in its own operation order (fp32, IEEE NaN
// semantics kept: several comparisons rely on NaN being unordered):
//   helpers sqr/clamp/mix/pow5/SchlickFresnel/GTR1/GTR2/smithG_GGX      src/sampling.cpp:10-47
//   BSDF::getBRDF / getBTDF / getBSDF                                    :49-199
//   sampleBRDF / clamp(vec3) / sampleCos / preciseRefraction             :203-308
//   sampleReflection / refract / sampleBTDF                              :310-391
//   sample / getLightObjectWeight / generateRandomPointInLightFace /
//   sampleLightFace / sampleSkyBox                                       :396-465
//   getTangentSpace / getTangentSpaceWithInDir                           src/geometry.cpp:105-124
//   RandomDistribution::operator() / pdf                                 src/component.cpp:20-31
#pragma once
#include "dev_surface.cuh"
#include "dev_rng.cuh"

namespace rm {

RM_DI float sqr(float x) { return fmul(x, x); }
RM_DI float clampf(float x, float a, float b) { return x < a ? a : (x > b ? b : x); }
RM_DI float mixf(float a, float b, float t) { return fadd(fmul(a, fsub(1.0f, t)), fmul(b, t)); }
RM_DI V3 mix3(V3 x, V3 y, float a) { return x * fsub(1.0f, a) + y * a; }       // glm::mix(vec3, vec3, float)
RM_DI float pow5(float x) { float x2 = fmul(x, x); return fmul(fmul(x2, x2), x); }
RM_DI float schlick(float u) { return pow5(clampf(fsub(1.0f, u), 0.0f, 1.0f)); }
RM_DI float sqrt_s(float x) { return x <= 0.0f ? 0.0f : fsqrt(x); }            // src/geometry.cpp:12-20

RM_DI float GTR1(float NdotH, float a) {
    if (a >= 1.0f) return fdiv(1.0f, kPi);
    float a2 = fmul(a, a);
    float t = fadd(1.0f, fmul(fmul(fsub(a2, 1.0f), NdotH), NdotH));
    return fdiv(fsub(a2, 1.0f), fmul(fmul(kPi, logf(a2)), t));
}
RM_DI float GTR2(float NdotH, float a) {
    float a2 = fmul(a, a);
    float t = fadd(1.0f, fmul(fmul(fsub(a2, 1.0f), NdotH), NdotH));
    return fdiv(a2, fmul(fmul(kPi, t), t));
}
RM_DI float smithG_GGX(float NdotV, float alphaG) {
    float a = fmul(alphaG, alphaG), b = fmul(NdotV, NdotV);
    return fdiv(1.0f, fadd(NdotV, fsqrt(fsub(fadd(a, b), fmul(a, b)))));
}

struct Bsdf {
    V3 inDir;          // points away from the surface (BSDF::inDir)
    Surface s;
};

RM_DI V3 get_brdf(const Bsdf &B, V3 L) {
    const V3 V = B.inDir, N = B.s.surfaceNormal;
    float NdotL = dot(N, L), NdotV = dot(N, V);
    if (NdotL <= 0.0f || NdotV <= 0.0f) return splat3(0.0f);
    V3 H = normalize(L + V);
    float NdotH = dot(N, H), LdotH = dot(L, H);
    V3 Cdlin = B.s.baseColor;
    float Cdlum = lum(Cdlin);
    const float subsurface = 0.0f, specularTint = 0.0f, sheen = 0.0f, sheenTint = 0.0f, clearcoatGloss = 0.2f, clearcoatTint = 0.0f;
    V3 Ctint = Cdlum > 0.0f ? div_recip(Cdlin, Cdlum) : splat3(1.0f);
    V3 Cspec0 = mix3(B.s.specular * mix3(splat3(1.0f), Ctint, specularTint), Cdlin, B.s.metallic);
    V3 Csheen = mix3(splat3(1.0f), Ctint, sheenTint);
    float FL = schlick(NdotL), FV = schlick(NdotV);
    float Fd90 = fadd(0.5f, fmul(fmul(fmul(2.0f, LdotH), LdotH), B.s.roughness));
    float Fd = fmul(mixf(1.0f, Fd90, FL), mixf(1.0f, Fd90, FV));
    float Fss90 = fmul(fmul(LdotH, LdotH), B.s.roughness);
    float Fss = fmul(mixf(1.0f, Fss90, FL), mixf(1.0f, Fss90, FV));
    float ss = fmul(1.25f, fadd(fmul(Fss, fsub(fdiv(1.0f, fadd(NdotL, NdotV)), 0.5f)), 0.5f));
    float Ds = GTR2(NdotH, B.s.roughness);
    float FH;
    if (B.s.eta <= 1.0f) FH = schlick(LdotH);
    else {
        float cosI = fabsf(LdotH);
        float sinI = sqrt_s(fsub(1.0f, fmul(cosI, cosI)));
        float sinT = fmul(B.s.eta, sinI);
        if (sinT >= 1.0f) FH = 1.0f;
        else {
            float cosT = sqrt_s(fsub(1.0f, fmul(sinT, sinT)));
            float R0 = fdiv(fsub(B.s.eta, 1.0f), fadd(B.s.eta, 1.0f));
            R0 = fmul(R0, R0);
            FH = mixf(R0, 1.0f, schlick(cosT));
        }
    }
    V3 Fs = mix3(Cspec0, splat3(1.0f), FH);
    float Gs = fmul(smithG_GGX(NdotL, B.s.roughness), smithG_GGX(NdotV, B.s.roughness));
    V3 Fsheen = fmul(FH, sheen) * Csheen;
    float Dr = GTR1(NdotH, mixf(0.1f, 0.001f, clearcoatGloss));
    float Fr = mixf(0.04f, 1.0f, FH);
    float Gr = fmul(smithG_GGX(NdotL, 0.25f), smithG_GGX(NdotV, 0.25f));
    V3 ret = fmul(fdiv(1.0f, kPi), mixf(Fd, ss, subsurface)) * Cdlin + Fsheen;
    ret = ret * fsub(1.0f, B.s.metallic);
    ret = ret + fmul(fmul(fmul(0.375f, Gr), Fr), Dr) * mix3(splat3(1.0f), Ctint, clearcoatTint);
    ret = ret * B.s.opacity;
    ret = ret + (Fs * Ds) * Gs;
    return ret * NdotL;
}

RM_DI V3 get_btdf(const Bsdf &B, V3 L) {
    const V3 N = B.s.surfaceNormal, V = B.inDir;
    float NdotL = dot(N, L);
    if (NdotL >= 0.0f) return splat3(0.0f);
    V3 H = normalize(L + B.s.eta * V);
    float D = GTR2(dot(N, H), B.s.roughness);
    float btdf = fmul(D, -NdotL);
    float LdotH = dot(L, H), NdotV = dot(N, V), HdotV = dot(H, V);
    btdf = fmul(btdf, fdiv(fabsf(fmul(LdotH, HdotV)), fadd(fabsf(fmul(NdotL, NdotV)), kEps)));
    float k = fadd(fmul(B.s.eta, HdotV), LdotH);
    btdf = fdiv(btdf, fmul(k, k));
    return splat3(btdf);
}

RM_DI V3 get_bsdf(const Bsdf &B, V3 outDir) {
    if (B.s.opacity > fsub(1.0f, kEps) || B.s.entering) return get_brdf(B, outDir);
    return get_btdf(B, outDir);
}

RM_DI void tangent_space(V3 normal, V3 &tangent, V3 &bitangent) {
    V3 v0 = fabsf(normal.x) < 0.8f ? mk3(1.0f, 0.0f, 0.0f) : mk3(0.0f, 1.0f, 0.0f);
    tangent = normalize(cross(v0, normal));
    bitangent = cross(normal, tangent);
}
RM_DI void tangent_space_in(V3 normal, V3 inDir, V3 &tangent, V3 &bitangent) {
    float c = dot(normal, inDir);
    if (c < fadd(-1.0f, 1e-3f)) { tangent_space(normal, tangent, bitangent); return; }
    bitangent = normalize(cross(normal, inDir));
    tangent = cross(bitangent, normal);
}

constexpr int kMaxTrys = 16;

RM_DI void clamp_lum(V3 &pdf) {                 // clamp(vec3&), src/sampling.cpp:238-243
    float C = lum(pdf);
    if (C > 64.0f) pdf = div_true(pdf, fdiv(C, 64.0f));
}

// GTR2 half-vector in the local frame (shared by sampleBRDF and sampleBTDF)
RM_DI V3 sample_gtr2_H(const Bsdf &B, Rng &gen, V3 tangent, V3 bitangent, float &cosTheta) {
    float u = gen();
    float phi = fmul(fmul(gen(), 2.0f), kPi);
    cosTheta = fsqrt(fdiv(fsub(1.0f, u), fadd(1.0f, fmul(fsub(sqr(B.s.roughness), 1.0f), u))));
    float sinTheta = fsqrt(fsub(1.0f, fmul(cosTheta, cosTheta)));
    float sp, cp;
    sincosf(phi, &sp, &cp);
    return (fmul(sinTheta, cp) * tangent + fmul(sinTheta, sp) * bitangent) + cosTheta * B.s.surfaceNormal;
}

// sampleCos (src/sampling.cpp:245-269) and sampleBRDF (203-236) share one rejection loop here: `ggx`
// selects the proposal (cosine-weighted direction or GTR2 half-vector), so the BRDF evaluation exists
// once in the instruction stream.  Draw order and arithmetic per try are the reference's.
RM_DI void sample_lobe(const Bsdf &B, Rng &gen, bool ggx, V3 tangent, V3 bitangent, V3 &outDir, V3 &brdfPdf, float &pdf, int &fails) {
#pragma unroll 1
    for (int T = 1; T <= kMaxTrys; T++) {
        if (ggx) {
            float cosTheta;
            V3 H = sample_gtr2_H(B, gen, tangent, bitangent, cosTheta);
            outDir = fmul(2.0f, dot(B.inDir, H)) * H - B.inDir;
            float LdotH = dot(outDir, H), LdotN = dot(outDir, B.s.surfaceNormal);
            if (LdotH <= 0.0f || LdotN <= 0.0f) pdf = 0.0f;
            else pdf = fdiv(GTR2(cosTheta, B.s.roughness), fmul(4.0f, LdotH));
        } else {
            float u = gen();
            float phi = fmul(fmul(gen(), 2.0f), kPi);
            float d = fsqrt(u);
            float z = fsqrt(fsub(1.0f, fmul(d, d)));
            float sp, cp;
            sincosf(phi, &sp, &cp);
            float x = fmul(d, cp), y = fmul(d, sp);
            outDir = (x * tangent + y * bitangent) + z * B.s.surfaceNormal;
            pdf = fdiv(dot(outDir, B.s.surfaceNormal), kPi);
        }
        if (dot(outDir, B.s.shapeNormal) > 0.0f && pdf > 0.0f) {
            brdfPdf = div_recip(get_brdf(B, outDir), pdf);
            if (!ggx) clamp_lum(brdfPdf);
            return;
        }
        fails++;
    }
    pdf = 0.0f;
    brdfPdf = splat3(0.0f);
    outDir = splat3(CUDART_NAN_F);
}

RM_DI void precise_refraction(const Bsdf &B, V3 &outDir, float &F) {
    const V3 V = B.inDir, N = B.s.surfaceNormal;
    float eta = B.s.eta;
    float NdotV = dot(N, V);
    if (NdotV < 0.0f) { F = 1.0f; outDir = splat3(CUDART_NAN_F); return; }
    float delta = fsub(1.0f, fmul(fmul(eta, eta), fsub(1.0f, fmul(NdotV, NdotV))));
    if (delta < 0.0f) { F = 1.0f; outDir = splat3(CUDART_NAN_F); return; }
    float k = fsub(fmul(eta, NdotV), fsqrt(delta));
    outDir = k * N - eta * V;
    float LdotN = -dot(outDir, N);
    float R0 = fdiv(fsub(1.0f, eta), fadd(1.0f, eta));
    R0 = fmul(R0, R0);
    F = fadd(R0, fmul(fsub(1.0f, R0), schlick(LdotN)));
}

// one-sample lobe pick (src/sampling.cpp:310-340): both proposals are always drawn (cosine first, then
// GGX - the order of the random stream), the pick uses a further draw
RM_DI void sample_reflection(const Bsdf &B, Rng &gen, V3 &Dir, V3 &brdfPdf, int &fails) {
    V3 tangent, bitangent;
    tangent_space_in(B.s.surfaceNormal, B.inDir, tangent, bitangent);
    V3 dirs[2], bs[2];
    float pdfs[2];
    int fl[2] = {0, 0};
#pragma unroll 1
    for (int k = 0; k < 2; k++) sample_lobe(B, gen, k == 1, tangent, bitangent, dirs[k], bs[k], pdfs[k], fl[k]);
    float p1 = fdiv(pdfs[0], fadd(pdfs[0], pdfs[1]));
    const int pick = gen() < p1 ? 0 : 1;
    Dir = dirs[pick]; brdfPdf = bs[pick]; fails += fl[pick];
}

RM_DI V3 refract_dir(V3 V, V3 N, float eta) {
    float NdotV = dot(N, V);
    float delta = fsub(1.0f, fmul(fmul(eta, eta), fsub(1.0f, fmul(NdotV, NdotV))));
    if (delta < 0.0f) return splat3(CUDART_NAN_F);
    float k = fsub(fmul(eta, NdotV), fsqrt(delta));
    return k * N - eta * V;
}

RM_DI void sample_btdf(const Bsdf &B, Rng &gen, V3 &outDir, V3 &btdfPdf, int &fails) {
    if (fabsf(fsub(B.s.eta, 1.0f)) < kEps) { outDir = -B.inDir; btdfPdf = splat3(1.0f); return; }
    V3 tangent, bitangent;
    tangent_space_in(B.s.surfaceNormal, B.inDir, tangent, bitangent);
    const V3 V = B.inDir;
    for (int T = 1; T <= kMaxTrys; T++) {
        float cosTheta, weight;
        V3 H = sample_gtr2_H(B, gen, tangent, bitangent, cosTheta);
        outDir = refract_dir(V, H, B.s.eta);
        if (!isfinite_any(outDir)) weight = 0.0f;
        else {
            float VdotH = dot(V, H), VdotN = dot(V, B.s.surfaceNormal), HdotN = dot(H, B.s.surfaceNormal);
            weight = fdiv(fabsf(VdotH), fabsf(fmul(VdotN, HdotN)));
        }
        if (weight > 0.0f && dot(outDir, B.s.shapeNormal) < 0.0f) { btdfPdf = splat3(weight); return; }
        fails++;
    }
    btdfPdf = splat3(0.0f);
    outDir = splat3(CUDART_NAN_F);
}

// ---- discrete distributions ----
// RandomDistribution::operator(): lower_bound(prefix, total * u)
RM_DI int cdf_sample(const float *cdf, int n, float u) {
    float x = fmul(__ldg(cdf + n - 1), u);
    int lo = 0, len = n;
    while (len > 0) {                       // std::lower_bound
        int half = len >> 1;
        if (__ldg(cdf + lo + half) < x) { lo = lo + half + 1; len = len - half - 1; }
        else len = half;
    }
    return lo;
}
// The same lower_bound, started from the bracket a guide table gives: guide[j] = lower_bound(cdf, total * j / G).  Rounding
// is monotone, so j/G <= u < (j+1)/G puts total*u between the two thresholds and the answer inside [guide[j], guide[j+1]]:
// identical index, ~12 fewer dependent loads on a 2 M-entry sky CDF.
RM_DI int cdf_sample_guided(const float *cdf, int n, float u, const int32_t *guide) {
    float x = fmul(__ldg(cdf + n - 1), u);
    int j = int(u * float(kSkyGuide));
    j = j < 0 ? 0 : (j > kSkyGuide - 1 ? kSkyGuide - 1 : j);
    int lo = __ldg(guide + j), len = __ldg(guide + j + 1) - lo;
    while (len > 0) {
        int half = len >> 1;
        if (__ldg(cdf + lo + half) < x) { lo = lo + half + 1; len = len - half - 1; }
        else len = half;
    }
    return lo;
}
RM_DI float cdf_pdf(const float *cdf, int n, int i) {
    float now = __ldg(cdf + i);
    if (i > 0) now = fsub(now, __ldg(cdf + i - 1));
    return fdiv(now, __ldg(cdf + n - 1));
}

// ---- next-event estimation ----
constexpr int kMaxLights = 32;              // light objects per scene the NEE weight table holds

// getLightObjectWeight (src/sampling.cpp:406-417)
RM_DI float light_weights(const DevScene &S, const Bsdf &B, float *w) {
    float total = 0.0f;
    const V3 pos = B.s.position;
    for (int i = 0; i < S.n_lights; i++) {
        const DevLight &L = S.lights[i];
        V3 c = mk3(L.center[0], L.center[1], L.center[2]);
        V3 lightDir = normalize(c - pos);
        float distance = length(c - pos);
        float C = lum(get_bsdf(B, lightDir));
        w[i] = fdiv(fmul(C, L.power), fadd(fmul(distance, distance), 1e-3f));
        total = fadd(total, w[i]);
    }
    return total;
}

// sample() (src/sampling.cpp:396-404)
RM_DI int pick_light(const float *w, int n, float r) {
    float c = 0.0f;
    for (int i = 0; i < n; i++) {
        c = fadd(c, w[i]);
        if (r <= c) return i;
    }
    return -1;
}

// sampleLightFace + generateRandomPointInLightFace (src/sampling.cpp:419-448)
RM_DI void sample_light_face(const DevScene &S, const DevLight &L, V3 pos, Rng &gen, V3 &lightPos, int &fails) {
    const float *cdf = S.light_cdf + L.face_offset;
    for (int T = 1; T <= kMaxTrys; T++) {
        int fi = cdf_sample(cdf, L.n_faces, gen());
        const float *p = S.light_pos + size_t(L.face_offset + fi) * 9;
        const float *nn = S.light_nrm + size_t(L.face_offset + fi) * 9;
        V3 v[3] = {mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8])};
        V3 n[3] = {mk3(nn[0], nn[1], nn[2]), mk3(nn[3], nn[4], nn[5]), mk3(nn[6], nn[7], nn[8])};
        float a = gen(), b = gen();
        if (fadd(a, b) > 1.0f) { a = fsub(1.0f, a); b = fsub(1.0f, b); }
        float c = fsub(fsub(1.0f, a), b);
        lightPos = (a * v[0] + b * v[1]) + c * v[2];
        V3 lightDir = normalize(lightPos - pos);
        V3 shapeN, surfN;
        bool entering;
        hit_normals(v, n, lightDir, mk3(a, b, c), shapeN, surfN, entering);
        float cosPhi = dot(surfN, -lightDir);
        if (cosPhi > 0.0f && gen() < cosPhi) return;
        fails++;
    }
    lightPos = splat3(CUDART_NAN_F);
}

// sampleSkyBox (src/sampling.cpp:450-465)
RM_DI void sample_sky(const DevScene &S, V3 shapeNormal, Rng &gen, V3 &Dir, V3 &light) {
    const int n = S.sky_width * S.sky_height;
    for (int T = 1; T <= kMaxTrys; T++) {
        int idx = cdf_sample_guided(S.sky_cdf, n, gen(), S.sky_guide);
        int u = idx % S.sky_width, v = idx / S.sky_width;
        float phi = fdiv(fmul(kPi, fadd(float(v), 0.5f)), float(S.sky_height));
        float theta = fdiv(fmul(fmul(2.0f, kPi), fadd(float(u), 0.5f)), float(S.sky_width));
        float sph, cph, sth, cth;
        sincosf(phi, &sph, &cph);
        sincosf(theta, &sth, &cth);
        Dir = mk3(fmul(-sph, sth), cph, fmul(sph, cth));
        if (dot(Dir, shapeNormal) > 0.0f) {
            const float *d = S.sky_data + size_t(idx) * 3;
            light = div_recip(mk3(__ldg(d), __ldg(d + 1), __ldg(d + 2)), cdf_pdf(S.sky_cdf, n, idx));
            return;
        }
    }
    Dir = splat3(CUDART_NAN_F);
}

} // namespace rm
