// dev_math.cuh — fp32 vector arithmetic for the device path.
//
// The reference computes in plain IEEE-754 single precision with no FMA contraction
// (x86-64 baseline), through glm 1.0.0.  Bit-exact stages (primary
// hits, G-buffer, FXAA) therefore use the round-to-nearest intrinsics below, which nvcc
// never fuses, and follow glm's operation order:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z        lib/glm/glm/detail/func_geometric.inl:52-53
//   cross(x,y)    = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)      :79-82
//   length(v)     = sqrt(dot(v,v))                                                    :12
//   normalize(v)  = v * (1 / sqrt(dot(v,v)))                                          :104
//   vec3 / scalar = v * (1 / scalar)   (!)               lib/glm/glm/detail/type_vec3.inl:582-585
//   vec3 /= scalar, vec3 / vec3, vec4 / scalar, vec2 / scalar = true division         :290-293
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace rm {

struct V2 { float x, y; };
struct V3 { float x, y, z; };

#define RM_DI __device__ __forceinline__
// out-of-line device function: one copy per kernel instead of one per call site (the shading kernels are
// instruction-cache bound); arguments are scalars so they travel in registers
#define RM_NI static __device__ __noinline__

RM_DI float fmul(float a, float b) { return __fmul_rn(a, b); }
RM_DI float fadd(float a, float b) { return __fadd_rn(a, b); }
RM_DI float fsub(float a, float b) { return __fsub_rn(a, b); }
RM_DI float fdiv(float a, float b) { return __fdiv_rn(a, b); }
RM_DI float fsqrt(float a) { return __fsqrt_rn(a); }
RM_DI float frcp(float a) { return __frcp_rn(a); }          // correctly rounded, i.e. the same bits as 1.0f / a

RM_DI V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RM_DI V3 splat3(float s) { return mk3(s, s, s); }
RM_DI V2 mk2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }

RM_DI V3 operator+(V3 a, V3 b) { return mk3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
RM_DI V3 operator-(V3 a, V3 b) { return mk3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
RM_DI V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
RM_DI V3 operator*(V3 a, V3 b) { return mk3(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
RM_DI V3 operator*(V3 a, float s) { return mk3(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
RM_DI V3 operator*(float s, V3 a) { return mk3(fmul(s, a.x), fmul(s, a.y), fmul(s, a.z)); }
RM_DI V3 div_recip(V3 a, float s) { float r = frcp(s); return a * r; }            // glm `vec3 / scalar`
RM_DI V3 div_true(V3 a, float s) { return mk3(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }   // glm `vec3 /= scalar`
RM_DI V3 div_true(V3 a, V3 b) { return mk3(fdiv(a.x, b.x), fdiv(a.y, b.y), fdiv(a.z, b.z)); } // glm `vec3 / vec3`
RM_DI float dot(V3 a, V3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
RM_DI V3 cross(V3 x, V3 y) {
    return mk3(fsub(fmul(x.y, y.z), fmul(y.y, x.z)), fsub(fmul(x.z, y.x), fmul(y.z, x.x)),
               fsub(fmul(x.x, y.y), fmul(y.x, x.y)));
}
RM_DI float length(V3 v) { return fsqrt(dot(v, v)); }
RM_DI V3 normalize(V3 v) { return v * frcp(fsqrt(dot(v, v))); }
RM_DI float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

RM_DI V2 operator+(V2 a, V2 b) { return mk2(fadd(a.x, b.x), fadd(a.y, b.y)); }
RM_DI V2 operator-(V2 a, V2 b) { return mk2(fsub(a.x, b.x), fsub(a.y, b.y)); }
RM_DI V2 operator*(float s, V2 a) { return mk2(fmul(s, a.x), fmul(s, a.y)); }
RM_DI float length(V2 v) { return fsqrt(fadd(fmul(v.x, v.x), fmul(v.y, v.y))); }

// isfinite(vec3) in the reference is an OR over components (src/geometry.cpp:8-10)
RM_DI bool isfinite_any(V3 v) { return isfinite(v.x) || isfinite(v.y) || isfinite(v.z); }
RM_DI bool isfinite_any(V2 v) { return isfinite(v.x) || isfinite(v.y); }

constexpr float kEps = 1e-4f;                       // eps_zero            include/geometry.h:13
constexpr float kPi = 3.14159265358979323846f;      // PI                  include/geometry.h:14
#define RM_LUM rm::mk3(0.3f, 0.6f, 0.1f)            // RGB_Weight          include/geometry.h:16

RM_DI float lum(V3 c) { return dot(c, RM_LUM); }

} // namespace rm
