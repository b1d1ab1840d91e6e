// ORACLE — TEST INFRASTRUCTURE ONLY (see vec.h).
// CPU restatement of /root/reference/shaders/utils/{random,math,mappings}.hlsl.
// Every function cites the HLSL lines it follows. pow(x,2)/pow(x,5) are restated as explicit
// multiplications. sin/cos/log/atan2/acos: the HLSL intrinsics have implementation-defined
// precision, so they are DEFINED here (sinDef ... acosDef below) as fixed sequences of IEEE binary32
// operations; the CUDA path states the same sequences independently (device_math.cuh) and must agree bit for bit.
#pragma once
#include "vec.h"

namespace orc {

// math.hlsl:3-7
static const float PI = 3.14159265f;
static const float INFINITY_F = 1000000000000.0f;  // "pranked" 1e12
static const uint32_t MAX_UINT = 0xFFFFFFFFu;
static const float AIR_IOR = 1.000277f;

// ---- random.hlsl:7-46 -------------------------------------------------------------------
namespace Hash {
inline uint32_t lcg(uint32_t a) { return a * 747796405u + 2891336453u; }          // :8-12
inline uint32_t rxs_m_xs(uint32_t a) {                                             // :15-18
    const uint32_t b = ((a >> ((a >> 28u) + 4u)) ^ a) * 277803737u;
    return (b >> 22u) ^ b;
}
inline uint32_t pcg(uint32_t a) { return rxs_m_xs(lcg(a)); }                       // :20-22
}  // namespace Hash

struct Rng {
    uint32_t state;
    static Rng fromSeed(uint32_t sx, uint32_t sy, uint32_t sz) {                   // :28-32
        Rng r;
        r.state = Hash::pcg(sx + Hash::pcg(sy + Hash::pcg(sz)));
        return r;
    }
    void stepState() { state = Hash::lcg(state); }                                 // :34-36
    float getFloat() {                                                             // :38-45
        stepState();
        uint32_t h = Hash::rxs_m_xs(state);
        return float(h >> 8) * 0x1p-24f;
    }
    // float2(rng.getFloat(), rng.getFloat()): evaluation order fixed x first, then y
    // (the reference leaves it to DXC).
    float2 getFloat2() { float x = getFloat(); float y = getFloat(); return {x, y}; }
};

// ---- math.hlsl --------------------------------------------------------------------------
inline float luminance(float3 c) {                                                 // :17-21
    return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z;
}
inline float3 faceForward(float3 n, float3 d) { return dot(n, d) > 0 ? n : -n; }   // :23-25

// math.hlsl:32-42 — Waechter & Binder integer-ULP ray-origin offset.
inline float offsetAlongNormal1(float p, float n) {
    const float origin = 1.0f / 32.0f;
    const float float_scale = 1.0f / 65536.0f;
    const float int_scale = 256.0f;
    int32_t of_i = (int32_t)(n * int_scale);  // truncates toward zero
    float p_i = as_float(as_int(p) + (p < 0.f ? -of_i : of_i));
    return std::fabs(p) < origin ? p + n * float_scale : p_i;
}
inline float3 offsetAlongNormal(float3 p, float3 n) {
    return {offsetAlongNormal1(p.x, n.x), offsetAlongNormal1(p.y, n.y), offsetAlongNormal1(p.z, n.z)};
}

inline void coordinateSystem(float3 v1, float3& v2, float3& v3) {                  // :55-63
    if (std::fabs(v1.x) > std::fabs(v1.y)) {
        v2 = f3(-v1.z, 0.0f, v1.x) / std::sqrt(v1.x * v1.x + v1.z * v1.z);
    } else {
        v2 = f3(0.0f, v1.z, -v1.y) / std::sqrt(v1.y * v1.y + v1.z * v1.z);
    }
    v3 = cross(v2, v1);
}

// ---- defined binary32 transcendentals ------------------------------------------------------
// Every operation is a correctly rounded IEEE binary32 +, -, *, / or sqrt in the order written (no FMA contraction), so any
// conforming implementation returns the same bits. Accuracy (tests/test_oracle_kat.py, against float64): sin/cos <= 2 ulp on
// [-2 pi, 4 pi], log <= 2 ulp on [2^-24, 1], atan2 <= 3 ulp, acos <= 4 ulp away from +-1 - the class of a GPU's own intrinsics.
// sin/cos: Cody-Waite reduction by multiples of pi/2 (three-part constant; exact products for |x| < 1e5), then the Cephes
// single-precision minimax polynomials on [-pi/4, pi/4].
inline void sinCosDef(float x, float& s, float& c) {
    const float q = std::nearbyint(x * 0.636619772367581343f);          // round to nearest, ties to even
    float r = x - q * 1.5703125f;
    r = r - q * 4.837512969970703125e-4f;
    r = r - q * 7.54978995489188216e-8f;
    const float z = r * r;
    const float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z;
    cp = cp - 0.5f * z;
    cp = cp + 1.0f;
    switch ((int32_t)q & 3) {
        case 0: s = sp; c = cp; break;
        case 1: s = cp; c = -sp; break;
        case 2: s = -sp; c = -cp; break;
        default: s = -cp; c = sp; break;
    }
}
inline float sinDef(float x) { float s, c; sinCosDef(x, s, c); return s; }
inline float cosDef(float x) { float s, c; sinCosDef(x, s, c); return c; }
// log of a positive normal number: x = m 2^e with m in [sqrt(1/2), sqrt(2)), Cephes logf polynomial in m - 1
inline float logDef(float x) {
    const int32_t bits = as_int(x);
    int32_t e = ((bits >> 23) & 0xFF) - 126;
    float m = as_float((bits & 0x007FFFFF) | 0x3F000000);                // [0.5, 1)
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
// atan of x >= 0 (Cephes atanf: two argument reductions, degree-4 polynomial in x^2)
inline float atanPosDef(float x) {
    float y0 = 0.0f;
    if (x > 2.414213562373095f) { y0 = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y0 = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    const float z = x * x;
    const float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x;
    return y0 + p;
}
inline float atan2Def(float y, float x) {
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a;
    if (ax == 0.0f) a = ay == 0.0f ? 0.0f : 1.5707963267948966f;
    else a = atanPosDef(ay / ax);
    if (x < 0.0f) a = 3.14159265358979323846f - a;
    return y < 0.0f ? -a : a;
}
inline float acosDef(float z) { return 2.0f * atan2Def(std::sqrt(1.0f - z), std::sqrt(1.0f + z)); }

// ---- mappings.hlsl ----------------------------------------------------------------------
inline float2 squareToTriangle(float2 sq) {                                        // :5-9
    float a = 1 - std::sqrt(1 - sq.x);
    float b = sq.y * std::sqrt(1 - sq.x);
    return {a, b};
}
inline float2 squareToGaussian(float2 sq) {                                        // :11-17
    const float u1 = 1.0f - sq.x;
    const float u2 = sq.y;
    const float r = std::sqrt(-2.0f * logDef(u1));
    const float theta = 2 * PI * u2;
    return r * f2(cosDef(theta), sinDef(theta));
}
inline float2 squareToUniformDiskConcentric(float2 sq) {                           // :19-37
    float2 uOffset = 2.0f * sq - f2(1.0f, 1.0f);
    if (uOffset.x == 0.0f && uOffset.y == 0.0f) return {0.0f, 0.0f};
    float theta, r;
    if (std::fabs(uOffset.x) > std::fabs(uOffset.y)) {
        r = uOffset.x;
        theta = (PI / 4) * (uOffset.y / uOffset.x);
    } else {
        r = uOffset.y;
        theta = (PI / 2) - (PI / 4) * (uOffset.x / uOffset.y);
    }
    return r * f2(cosDef(theta), sinDef(theta));
}
inline float3 squareToCosineHemisphere(float2 sq) {                                // :39-44
    float2 d = squareToUniformDiskConcentric(sq);
    float z = std::sqrt(std::fmax(0.0f, 1.0f - dot(d, d)));
    return {d.x, d.y, z};
}
inline float3 sphericalToCartesian(float sinTheta, float cosTheta, float phi) {    // :53-55
    return {sinTheta * cosDef(phi), sinTheta * sinDef(phi), cosTheta};
}
inline float2 cartesianToSpherical(float3 v) {                                     // :59-64
    float p = atan2Def(v.y, v.x);
    float phi = (p < 0) ? (p + 2 * PI) : p;
    float theta = acosDef(v.z);
    return {phi, theta};
}
inline float3 squareToEqualAreaSphere(float2 sq) {                                 // :67-83
    const float2 uv = 2.0f * sq - f2(1.0f, 1.0f);
    const float2 uvp = absv(uv);
    const float signedDistance = 1.0f - (uvp.x + uvp.y);
    const float d = std::fabs(signedDistance);
    const float r = 1.0f - d;
    const float phi = (r == 0.0f ? 1.0f : (uvp.y - uvp.x) / r + 1.0f) * PI / 4.0f;
    const float3 signs = {signf(uv.x), signf(uv.y), signf(signedDistance)};
    return signs * f3(cosDef(phi) * r * std::sqrt(2.0f - r * r),
                      sinDef(phi) * r * std::sqrt(2.0f - r * r),
                      1.0f - r * r);
}
inline float2 squareToEqualAreaSphereInverse(float3 dir) {                         // :85-99
    const float3 xyz = absv(dir);
    const float r = std::sqrt(1.0f - xyz.z);
    float phi = (xyz.x == 0.0f && xyz.y == 0.0f)
                    ? 0.0f
                    : atan2Def(std::fmin(xyz.x, xyz.y), std::fmax(xyz.x, xyz.y)) * 2.0f / PI;
    if (xyz.x < xyz.y) phi = 1.0f - phi;
    float2 uv = {r - phi * r, phi * r};
    if (dir.z < 0) uv = f2(1.0f, 1.0f) - f2(uv.y, uv.x);
    uv = uv * f2(signf(dir.x), signf(dir.y));
    return (uv + f2(1.0f, 1.0f)) / 2.0f;
}
inline bool coinFlipRemap(float p, float& rand) {                                  // :103-111
    if (rand < p) {
        rand /= p;
        return true;
    } else {
        rand = (rand - p) / (1.0f - p);
        return false;
    }
}

}  // namespace orc
