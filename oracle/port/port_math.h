// port_math.h — fp32 vector helpers for the CPU restatement (oracle/port).
//
// TEST INFRASTRUCTURE.  oracle/port is an independent, glm-free restatement of the reference's
// hot-path algorithms in plain C++; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg may use it.  It shares no code with raym0nade_b200/ (the product) - only the
// data-format header include/rm_types.h.
//
// glm 1.0.0 conventions the reference relies on (lib/glm/glm/detail):
//   dot = (x*x' + y*y') + z*z'            func_geometric.inl:52-53
//   cross as written                       func_geometric.inl:79-82
//   normalize(v) = v * (1/sqrt(dot(v,v)))  func_geometric.inl:104, func_exponential.inl:138
//   vec3 / scalar = v * (1/scalar)         type_vec3.inl:582-585   (!)
//   vec3 /= scalar, vec3/vec3, vec4/scalar, vec2/scalar = true division
// Compile with -ffp-contract=off: the reference's x86-64 baseline build has no FMA.
#pragma once
#include <cmath>
#include <cstdint>

namespace port {

struct vec2 { float x, y; };
struct vec3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 { float x, y, z, w; };

inline vec3 v3(float x, float y, float z) { return {x, y, z}; }
inline vec3 v3(float s) { return {s, s, s}; }
inline vec3 v3(const float *p) { return {p[0], p[1], p[2]}; }
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 div_scalar(vec3 a, float s) { float r = 1.0f / s; return {a.x * r, a.y * r, a.z * r}; }   // glm: vec3 / scalar
inline vec3 div_assign(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }                       // glm: vec3 /= scalar
inline vec3 div_vec(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }                     // glm: vec3 / vec3
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 x, vec3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline vec3 mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }

inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator*(float s, vec2 a) { return {s * a.x, s * a.y}; }
inline float length(vec2 v) { return std::sqrt(v.x * v.x + v.y * v.y); }

inline bool finite_any(vec3 v) { return std::isfinite(v.x) || std::isfinite(v.y) || std::isfinite(v.z); }   // src/geometry.cpp:8-10
inline bool finite_any(vec2 v) { return std::isfinite(v.x) || std::isfinite(v.y); }

constexpr float eps_zero = 1e-4f;                       // include/geometry.h:13
const float PI = 3.14159265358979323846f;               // include/geometry.h:14
const vec3 RGB_Weight = {0.3f, 0.6f, 0.1f};             // include/geometry.h:16

} // namespace port
