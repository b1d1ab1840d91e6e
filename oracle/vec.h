// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product
// path (moonshine_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs may use it.
//
// Minimal HLSL-flavoured FP32 vector types for the CPU restatement of
// /root/reference/shaders/**. Arithmetic conventions (the reference leaves them to DXC and
// the Vulkan driver, so they are *defined here* and mirrored by the CUDA path):
//   * everything is IEEE binary32, compiled with -ffp-contract=off (no FMA contraction);
//   * dot(a,b)   = ((a.x*b.x + a.y*b.y) + a.z*b.z)           (left to right)
//   * cross(a,b) = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
//   * normalize(v) = v * (1.0f / sqrt(dot(v,v)))
//   * lerp(a,b,t) = a + t*(b-a)                                
//   * mul(M3x4, (p,1)) row r = ((m[r][0]*x + m[r][1]*y) + m[r][2]*z) + m[r][3]
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint2  { uint32_t x, y; };
struct uint3  { uint32_t x, y, z; };

inline float2 f2(float x, float y) { return {x, y}; }
inline float3 f3(float x, float y, float z) { return {x, y, z}; }
inline float3 f3(float s) { return {s, s, s}; }

inline float2 operator+(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float2 operator-(float2 a, float2 b) { return {a.x - b.x, a.y - b.y}; }
inline float2 operator*(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
inline float2 operator*(float a, float2 b) { return {a * b.x, a * b.y}; }
inline float2 operator*(float2 a, float b) { return {a.x * b, a.y * b}; }
inline float2 operator/(float2 a, float2 b) { return {a.x / b.x, a.y / b.y}; }
inline float2 operator/(float2 a, float b) { return {a.x / b, a.y / b}; }
inline float2 operator-(float2 a) { return {-a.x, -a.y}; }

inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(float a, float3 b) { return {a * b.x, a * b.y, a * b.z}; }
inline float3 operator*(float3 a, float b) { return {a.x * b, a.y * b, a.z * b}; }
inline float3 operator/(float3 a, float3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator/(float3 a, float b) { return {a.x / b, a.y / b, a.z / b}; }
inline float3 operator-(float3 a) { return {-a.x, -a.y, -a.z}; }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
inline float3& operator/=(float3& a, float b) { a = a / b; return a; }

inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float3 cross(float3 a, float3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(float3 a) { return std::sqrt(dot(a, a)); }
inline float distance(float3 a, float3 b) { return length(a - b); }
inline float3 normalize(float3 a) { float inv = 1.0f / std::sqrt(dot(a, a)); return a * inv; }
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float3 lerp(float3 a, float3 b, float t) { return a + t * (b - a); }
inline float3 lerp(float3 a, float3 b, float3 t) { return a + t * (b - a); }
inline float clampf(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }
inline float saturate(float v) { return clampf(v, 0.0f, 1.0f); }
inline float signf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
inline float3 absv(float3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline float2 absv(float2 a) { return {std::fabs(a.x), std::fabs(a.y)}; }
inline bool all_zero(float3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

inline int32_t as_int(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline float as_float(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }

// row-major 3x4 affine matrix (Mat3x4 of hydra/moonshine.h:33-35; world.hlsl:7)
struct Mat3x4 { float m[3][4]; };

inline float3 mul_point(const Mat3x4& M, float3 p) {
    return {((M.m[0][0] * p.x + M.m[0][1] * p.y) + M.m[0][2] * p.z) + M.m[0][3],
            ((M.m[1][0] * p.x + M.m[1][1] * p.y) + M.m[1][2] * p.z) + M.m[1][3],
            ((M.m[2][0] * p.x + M.m[2][1] * p.y) + M.m[2][2] * p.z) + M.m[2][3]};
}
inline float3 mul_vec(const Mat3x4& M, float3 v) {
    return {(M.m[0][0] * v.x + M.m[0][1] * v.y) + M.m[0][2] * v.z,
            (M.m[1][0] * v.x + M.m[1][1] * v.y) + M.m[1][2] * v.z,
            (M.m[2][0] * v.x + M.m[2][1] * v.y) + M.m[2][2] * v.z};
}
// mul(transpose(M), v) with M's 3x3 part: column c of the 3x3 dotted with v
// (reflection_frame.hlsl:23-29 via world.hlsl:172-173: Frame::inSpace(transpose(toMesh)))
inline float3 mul_transpose3(const Mat3x4& M, float3 v) {
    return {(M.m[0][0] * v.x + M.m[1][0] * v.y) + M.m[2][0] * v.z,
            (M.m[0][1] * v.x + M.m[1][1] * v.y) + M.m[2][1] * v.z,
            (M.m[0][2] * v.x + M.m[1][2] * v.y) + M.m[2][2] * v.z};
}

}  // namespace orc
