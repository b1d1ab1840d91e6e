// Device-side FP32 helpers for the sm_100a kernels.
//
// This translation unit is compiled with -fmad=false: the compiler never contracts a*b+c, so every
// expression below evaluates exactly as written (IEEE binary32, round-to-nearest, -prec-div/-prec-sqrt
// defaults). FMA is used only where it is spelled out (`fmaf`), i.e. in the conservative BVH slab test.
// Conventions (DESIGN.md §"Arithmetic"): dot = ((x+y)+z), normalize = v * (1/sqrt(dot)),
// lerp(a,b,t) = a + t*(b-a); pow(x,2)/pow(x,5) of the HLSL are explicit products.
//
// Each function cites the HLSL it restates (paths relative to /root/reference/shaders/).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "dtypes.h"

namespace msne {

#define DI __device__ __forceinline__

DI float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
DI float3 mk3(float s) { return make_float3(s, s, s); }
DI float2 mk2(float x, float y) { return make_float2(x, y); }
DI float3 xyz(float4 v) { return make_float3(v.x, v.y, v.z); }

DI float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
DI float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
DI float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
DI float3 operator*(float a, float3 b) { return mk3(a * b.x, a * b.y, a * b.z); }
DI float3 operator*(float3 a, float b) { return mk3(a.x * b, a.y * b, a.z * b); }
DI float3 operator/(float3 a, float b) { return mk3(a.x / b, a.y / b, a.z / b); }
DI float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
DI float2 operator+(float2 a, float2 b) { return mk2(a.x + b.x, a.y + b.y); }
DI float2 operator-(float2 a, float2 b) { return mk2(a.x - b.x, a.y - b.y); }
DI float2 operator*(float a, float2 b) { return mk2(a * b.x, a * b.y); }
DI float2 operator*(float2 a, float2 b) { return mk2(a.x * b.x, a.y * b.y); }
DI float2 operator/(float2 a, float2 b) { return mk2(a.x / b.x, a.y / b.y); }
DI float2 operator/(float2 a, float b) { return mk2(a.x / b, a.y / b); }

DI float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
DI float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
DI float3 cross(float3 a, float3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DI float3 normalize(float3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
DI float length(float3 a) { return sqrtf(dot(a, a)); }
DI float lerp1(float a, float b, float t) { return a + t * (b - a); }
DI float3 lerp3(float3 a, float3 b, float t) { return a + t * (b - a); }
DI float3 lerp3(float3 a, float3 b, float3 t) { return a + t * (b - a); }
DI float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
DI float signf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
DI bool allZero(float3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

// math.hlsl:3-7
#define MS_PI 3.14159265f
#define MS_INFINITY 1000000000000.0f
#define MS_AIR_IOR 1.000277f

// 3x4 row-major affine transforms stored as 12 floats
DI float3 mulPoint(const float* m, float3 p) {
    return mk3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3],
               ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
               ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
DI float3 mulVec(const float* m, float3 v) {
    return mk3((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[4] * v.x + m[5] * v.y) + m[6] * v.z, (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
}
// mul(transpose(M3x3), v)   (reflection_frame.hlsl:23-29 through world.hlsl:172-173)
DI float3 mulTranspose3(const float* m, float3 v) {
    return mk3((m[0] * v.x + m[4] * v.y) + m[8] * v.z, (m[1] * v.x + m[5] * v.y) + m[9] * v.z, (m[2] * v.x + m[6] * v.y) + m[10] * v.z);
}

// ---- utils/random.hlsl:7-46 (integer exact) ---------------------------------------------------
DI uint32_t hashLcg(uint32_t a) { return a * 747796405u + 2891336453u; }
DI uint32_t hashRxsMXs(uint32_t a) { const uint32_t b = ((a >> ((a >> 28u) + 4u)) ^ a) * 277803737u; return (b >> 22u) ^ b; }
DI uint32_t hashPcg(uint32_t a) { return hashRxsMXs(hashLcg(a)); }
struct Rng {
    uint32_t state;
    DI static Rng fromSeed(uint32_t sx, uint32_t sy, uint32_t sz) { Rng r; r.state = hashPcg(sx + hashPcg(sy + hashPcg(sz))); return r; }
    DI float getFloat() { state = hashLcg(state); return float(hashRxsMXs(state) >> 8) * 0x1p-24f; }
    DI float2 getFloat2() { float x = getFloat(); float y = getFloat(); return mk2(x, y); }  // x first
};

// ---- utils/math.hlsl ---------------------------------------------------------------------------
DI float luminance(float3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }   // :17-21
DI float3 faceForward(float3 n, float3 d) { return dot(n, d) > 0 ? n : -n; }              // :23-25
DI float offsetAlongNormal1(float p, float n) {                                          // :32-42
    const float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    int of_i = (int)(n * int_scale);
    float p_i = __int_as_float(__float_as_int(p) + (p < 0.f ? -of_i : of_i));
    return fabsf(p) < origin ? p + n * float_scale : p_i;
}
DI float3 offsetAlongNormal(float3 p, float3 n) { return mk3(offsetAlongNormal1(p.x, n.x), offsetAlongNormal1(p.y, n.y), offsetAlongNormal1(p.z, n.z)); }
DI void coordinateSystem(float3 v1, float3& v2, float3& v3) {                             // :55-63
    if (fabsf(v1.x) > fabsf(v1.y)) v2 = mk3(-v1.z, 0.0f, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
    else v2 = mk3(0.0f, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
    v3 = cross(v2, v1);
}

// ---- defined binary32 transcendentals -------------------------------------------------------------
// HLSL's sin / cos / log / atan2 / acos have implementation-defined precision. Here they are DEFINED as
// fixed sequences of correctly rounded IEEE binary32 operations - Cody-Waite reduction + the Cephes single-precision minimax
// polynomials - instead of CUDA's libdevice, so that any conforming implementation of the same sequences (the CPU oracle states
// them independently) returns the same bits: with them the whole renderer is reproducible bit for bit across host and device.
// Accuracy against float64: sin/cos <= 1.5 ulp, log <= 1 ulp, atan2 <= 3.1 ulp, acos <= 4.1 ulp (tests/test_oracle_kat.py).
DI void sinCosDef(float x, float& s, float& c) {
    const float q = rintf(x * 0.636619772367581343f);                   // round to nearest, ties to even
    float r = x - q * 1.5703125f;                                        // pi/2 in three parts: the products are exact for |x| < 1e5
    r = r - q * 4.837512969970703125e-4f;
    r = r - q * 7.54978995489188216e-8f;
    const float z = r * r;
    const float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z;
    cp = cp - 0.5f * z;
    cp = cp + 1.0f;
    const int n = (int)q & 3;
    const float ss = (n & 1) ? cp : sp, cc = (n & 1) ? sp : cp;
    s = (n & 2) ? -ss : ss;
    c = (n == 1 || n == 2) ? -cc : cc;
}
DI float sinDef(float x) { float s, c; sinCosDef(x, s, c); return s; }
DI float cosDef(float x) { float s, c; sinCosDef(x, s, c); return c; }
DI float logDef(float x) {                                               // x positive and normal
    const int bits = __float_as_int(x);
    int e = ((bits >> 23) & 0xFF) - 126;
    float m = __int_as_float((bits & 0x007FFFFF) | 0x3F000000);          // [0.5, 1)
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    const float z = m * m;
    float y = ((((((((7.0376836292e-2f * m - 1.1514610310e-1f) * m + 1.1676998740e-1f) * m - 1.2420140846e-1f) * m + 1.4249322787e-1f) * m
                 - 1.6668057665e-1f) * m + 2.0000714765e-1f) * m - 2.4999993993e-1f) * m + 3.3333331174e-1f) * m * z;
    const float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y - 0.5f * z;
    float r = m + y;
    r = r + 0.693359375f * fe;
    return r;
}
DI float atanPosDef(float x) {                                           // x >= 0
    float y0 = 0.0f;
    if (x > 2.414213562373095f) { y0 = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y0 = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    const float z = x * x;
    const float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x;
    return y0 + p;
}
DI float atan2Def(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax == 0.0f) a = ay == 0.0f ? 0.0f : 1.5707963267948966f;
    else a = atanPosDef(ay / ax);
    if (x < 0.0f) a = 3.14159265358979323846f - a;
    return y < 0.0f ? -a : a;
}
DI float acosDef(float z) { return 2.0f * atan2Def(sqrtf(1.0f - z), sqrtf(1.0f + z)); }

// ---- utils/mappings.hlsl -----------------------------------------------------------------------
DI float2 squareToTriangle(float2 sq) { float a = 1 - sqrtf(1 - sq.x); float b = sq.y * sqrtf(1 - sq.x); return mk2(a, b); }  // :5-9
DI float2 squareToGaussian(float2 sq) {                                                   // :11-17
    const float u1 = 1.0f - sq.x, u2 = sq.y;
    const float r = sqrtf(-2.0f * logDef(u1));
    const float theta = 2 * MS_PI * u2;
    float sn, cs; sinCosDef(theta, sn, cs);
    return r * mk2(cs, sn);
}
DI float2 squareToUniformDiskConcentric(float2 sq) {                                      // :19-37
    float2 uO = 2.0f * sq - mk2(1.0f, 1.0f);
    if (uO.x == 0.0f && uO.y == 0.0f) return mk2(0.0f, 0.0f);
    float theta, r;
    if (fabsf(uO.x) > fabsf(uO.y)) { r = uO.x; theta = (MS_PI / 4) * (uO.y / uO.x); }
    else { r = uO.y; theta = (MS_PI / 2) - (MS_PI / 4) * (uO.x / uO.y); }
    float sn, cs; sinCosDef(theta, sn, cs);
    return r * mk2(cs, sn);
}
DI float3 squareToCosineHemisphere(float2 sq) {                                           // :39-44
    float2 d = squareToUniformDiskConcentric(sq);
    float z = sqrtf(fmaxf(0.0f, 1.0f - dot(d, d)));
    return mk3(d.x, d.y, z);
}
DI float3 sphericalToCartesian(float sinTheta, float cosTheta, float phi) { float sn, cs; sinCosDef(phi, sn, cs); return mk3(sinTheta * cs, sinTheta * sn, cosTheta); }  // :53-55
DI float2 cartesianToSpherical(float3 v) {                                                // :59-64
    float p = atan2Def(v.y, v.x);
    float phi = (p < 0) ? (p + 2 * MS_PI) : p;
    float theta = acosDef(v.z);
    return mk2(phi, theta);
}
DI float3 squareToEqualAreaSphere(float2 sq) {                                            // :67-83
    const float2 uv = 2.0f * sq - mk2(1.0f, 1.0f);
    const float2 uvp = mk2(fabsf(uv.x), fabsf(uv.y));
    const float signedDistance = 1.0f - (uvp.x + uvp.y);
    const float d = fabsf(signedDistance);
    const float r = 1.0f - d;
    const float phi = (r == 0.0f ? 1.0f : (uvp.y - uvp.x) / r + 1.0f) * MS_PI / 4.0f;
    const float3 signs = mk3(signf(uv.x), signf(uv.y), signf(signedDistance));
    float sn, cs; sinCosDef(phi, sn, cs);
    return signs * mk3(cs * r * sqrtf(2.0f - r * r), sn * r * sqrtf(2.0f - r * r), 1.0f - r * r);
}
DI float2 squareToEqualAreaSphereInverse(float3 dir) {                                    // :85-99
    const float3 a = mk3(fabsf(dir.x), fabsf(dir.y), fabsf(dir.z));
    const float r = sqrtf(1.0f - a.z);
    float phi = (a.x == 0.0f && a.y == 0.0f) ? 0.0f : atan2Def(fminf(a.x, a.y), fmaxf(a.x, a.y)) * 2.0f / MS_PI;
    if (a.x < a.y) phi = 1.0f - phi;
    float2 uv = mk2(r - phi * r, phi * r);
    if (dir.z < 0) uv = mk2(1.0f, 1.0f) - mk2(uv.y, uv.x);
    uv = uv * mk2(signf(dir.x), signf(dir.y));
    return (uv + mk2(1.0f, 1.0f)) / 2.0f;
}
// Same two quotients as the HLSL's branches (rand/p or (rand-p)/(1-p)), selected per lane and divided ONCE: in a warp
// both branches are always taken by somebody, so the branchy form executes two IEEE divisions per call (20 calls per
// env-map sample).
DI bool coinFlipRemap(float p, float& rand) {                                             // :103-111
    const bool heads = rand < p;
    const float num = heads ? rand : rand - p;
    const float den = heads ? p : 1.0f - p;
    rand = num / den;
    return heads;
}
// float -> uint conversion with defined behaviour for negative / NaN / huge inputs (-> 0 / 0 / max)
DI uint32_t floatToUint(float f) { return f >= 0.0f ? (f < 4294967040.0f ? (uint32_t)f : 0xFFFFFFFFu) : 0u; }

}  // namespace msne
