// Device restatement of the reference's shading code for the wavefront kernels:
//   shaders/hrtsystem/reflection_frame.hlsl (Frame), material.hlsl (GGX, Fresnel, Lambert, StandardPBR,
//   PerfectMirror, Glass, MaterialVariant, getTextureFrame, getEmissive), world.hlsl (MeshAttributes),
//   light.hlsl (EnvMap, MeshLights without their embedded shadow TraceRay — shadow rays are queued and
//   traced by the next trace pass, see kernels.cu).
// Texture sampling restates the reference's samplers in software so that filtering is exact FP32
// (MaterialManager.zig:425-444 repeat/linear; BackgroundManager.zig:78-95 mirrored-repeat/linear):
//   u = s*W - 0.5, i0 = floor(u), a = u - i0, texel = lerp(lerp(t00,t10,a), lerp(t01,t11,a), b).
#pragma once
#include <cuda_fp16.h>
#include "device_math.cuh"

#ifndef MSNE_POS_FROM_TRIREC
#define MSNE_POS_FROM_TRIREC 1
#endif
#ifndef MSNE_MATERIAL_CONSTANTS
#define MSNE_MATERIAL_CONSTANTS 1   // 1x1 material textures are read as constants of the material record (sampleMaterialSlot)
#endif
#ifndef MSNE_EMITTER_RECORDS
#define MSNE_EMITTER_RECORDS 1   // mesh-light samples read the 64-byte per-entry record (emitterPoint) instead of the attribute chain
#endif
#ifndef MSNE_ENV_QUADS
#define MSNE_ENV_QUADS 1   // EnvMap::sample reads the quad-major pyramid (one LDG.128 per level instead of four scalar loads)
#endif

namespace msne {

// ---- textures ------------------------------------------------------------------------------------
// `luts`: 512 floats, [0,256) = sRGB -> linear (IEC formula in double, rounded once), [256,512) = c / 255.0f (UNORM), both
// tabulated on the host with the same IEEE operations the per-texel arithmetic would do, so values are identical and
// the kernels do a cached load instead of a pow / an IEEE division per channel.
DI float unorm8(const float* __restrict__ luts, uint32_t c) { return __ldg(luts + 256 + c); }
DI float srgb8(const float* __restrict__ luts, uint32_t c) { return __ldg(luts + c); }
DI float4 fetchTexel(const float* __restrict__ luts, const DTexture& t, uint32_t x, uint32_t y) {
    const size_t i = (size_t)y * t.w + x;
    switch (t.fmt) {
        case TEX_R32F: return make_float4(__ldg((const float*)t.data + i), 0.f, 0.f, 1.f);
        case TEX_RG32F: { float2 v = __ldg((const float2*)t.data + i); return make_float4(v.x, v.y, 0.f, 1.f); }
        case TEX_RGBA32F: return __ldg((const float4*)t.data + i);
        case TEX_R8_UNORM: return make_float4(unorm8(luts, __ldg((const uint8_t*)t.data + i)), 0.f, 0.f, 1.f);
        case TEX_RG8_UNORM: { uchar2 v = __ldg((const uchar2*)t.data + i); return make_float4(unorm8(luts, v.x), unorm8(luts, v.y), 0.f, 1.f); }
        case TEX_RGBA8_SRGB: {
            uchar4 v = __ldg((const uchar4*)t.data + i);
            return make_float4(srgb8(luts, v.x), srgb8(luts, v.y), srgb8(luts, v.z), unorm8(luts, v.w));
        }
        default: {  // TEX_RGBA16F
            uint2 raw = __ldg((const uint2*)t.data + i);
            __half2 a = *reinterpret_cast<__half2*>(&raw.x), b = *reinterpret_cast<__half2*>(&raw.y);
            return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
        }
    }
}
// REPEAT addressing; power-of-two sizes (the common case) take the mask instead of the ~20-instruction integer modulo
DI int wrapRepeat(int i, int n) {
    if ((n & (n - 1)) == 0) return i & (n - 1);
    int m = i % n; return m < 0 ? m + n : m;
}
DI int wrapMirror(int i, int n) {
    int m = i % (2 * n); if (m < 0) m += 2 * n;
    const int k = m - n;
    const int mir = k >= 0 ? k : -(1 + k);
    return (n - 1) - mir;
}
DI float4 lerp4(float4 a, float4 b, float t) { return make_float4(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z), a.w + t * (b.w - a.w)); }
DI float bilerp1(float t00, float t10, float t01, float t11, float a, float b) {
    const float top = t00 + a * (t10 - t00), bot = t01 + a * (t11 - t01);
    return top + b * (bot - top);
}

// dTextures[idx].SampleLevel(dTextureSampler, uv, 0): exact FP32 bilinear lerp(lerp(t00,t10,a),lerp(t01,t11,a),b).
// The format switch runs once per sample (not per texel) and each case filters only the channels the format has; the
// constant channels (0, 0, 1) are what filtering them would give for finite weights (0 + a*(0-0), 1 + a*(1-1)).
// Deliberately NOT inlined: it is used ~6x per hit; inlined, k_shade was 189 KB of SASS and stalled on instruction
// fetch (profiles/r01b). Arguments/return are scalars in registers, so the call costs no local-memory traffic.
#ifndef MSNE_INLINE_TEX
__device__ __noinline__
#else
DI
#endif
float4 sampleTextureImpl(const DTexture* __restrict__ textures, const float* __restrict__ luts, uint32_t idx, float u_, float v_) {
    const DTexture t = textures[idx];
    if (t.w == 1 && t.h == 1) return fetchTexel(luts, t, 0, 0);
    const float u = u_ * (float)t.w - 0.5f, v = v_ * (float)t.h - 0.5f;
    const float fu = floorf(u), fv = floorf(v);
    const float a = u - fu, b = v - fv;
    const int i0 = wrapRepeat((int)fu, t.w), i1 = wrapRepeat((int)fu + 1, t.w), j0 = wrapRepeat((int)fv, t.h), j1 = wrapRepeat((int)fv + 1, t.h);
    const size_t o00 = (size_t)j0 * t.w + i0, o10 = (size_t)j0 * t.w + i1, o01 = (size_t)j1 * t.w + i0, o11 = (size_t)j1 * t.w + i1;
    switch (t.fmt) {
        case TEX_R8_UNORM: {
            const uint8_t* d = (const uint8_t*)t.data;
            const uint32_t c00 = __ldg(d + o00), c10 = __ldg(d + o10), c01 = __ldg(d + o01), c11 = __ldg(d + o11);
            return make_float4(bilerp1(unorm8(luts, c00), unorm8(luts, c10), unorm8(luts, c01), unorm8(luts, c11), a, b), 0.f, 0.f, 1.f);
        }
        case TEX_RG8_UNORM: {
            const uchar2* d = (const uchar2*)t.data;
            const uchar2 c00 = __ldg(d + o00), c10 = __ldg(d + o10), c01 = __ldg(d + o01), c11 = __ldg(d + o11);
            return make_float4(bilerp1(unorm8(luts, c00.x), unorm8(luts, c10.x), unorm8(luts, c01.x), unorm8(luts, c11.x), a, b),
                               bilerp1(unorm8(luts, c00.y), unorm8(luts, c10.y), unorm8(luts, c01.y), unorm8(luts, c11.y), a, b), 0.f, 1.f);
        }
        case TEX_RGBA8_SRGB: {
            const uchar4* d = (const uchar4*)t.data;
            const uchar4 c00 = __ldg(d + o00), c10 = __ldg(d + o10), c01 = __ldg(d + o01), c11 = __ldg(d + o11);
            return make_float4(bilerp1(srgb8(luts, c00.x), srgb8(luts, c10.x), srgb8(luts, c01.x), srgb8(luts, c11.x), a, b),
                               bilerp1(srgb8(luts, c00.y), srgb8(luts, c10.y), srgb8(luts, c01.y), srgb8(luts, c11.y), a, b),
                               bilerp1(srgb8(luts, c00.z), srgb8(luts, c10.z), srgb8(luts, c01.z), srgb8(luts, c11.z), a, b),
                               bilerp1(unorm8(luts, c00.w), unorm8(luts, c10.w), unorm8(luts, c01.w), unorm8(luts, c11.w), a, b));
        }
        case TEX_R32F: {
            const float* d = (const float*)t.data;
            return make_float4(bilerp1(__ldg(d + o00), __ldg(d + o10), __ldg(d + o01), __ldg(d + o11), a, b), 0.f, 0.f, 1.f);
        }
        case TEX_RG32F: {
            const float2* d = (const float2*)t.data;
            const float2 c00 = __ldg(d + o00), c10 = __ldg(d + o10), c01 = __ldg(d + o01), c11 = __ldg(d + o11);
            return make_float4(bilerp1(c00.x, c10.x, c01.x, c11.x, a, b), bilerp1(c00.y, c10.y, c01.y, c11.y, a, b), 0.f, 1.f);
        }
        default: {  // TEX_RGBA32F, TEX_RGBA16F
            const float4 t00 = fetchTexel(luts, t, i0, j0), t10 = fetchTexel(luts, t, i1, j0), t01 = fetchTexel(luts, t, i0, j1), t11 = fetchTexel(luts, t, i1, j1);
            return lerp4(lerp4(t00, t10, a), lerp4(t01, t11, a), b);
        }
    }
}
DI float4 sampleTexture(const DeviceScene& sc, uint32_t idx, float2 uv) { return sampleTextureImpl(sc.textures, sc.texLuts, idx, uv.x, uv.y); }
// texture slot `slot` of a material: the host-resolved constant when the image is 1x1 (same bits as the sampler's result)
DI float4 sampleMaterialSlot(const DeviceScene& sc, const DMaterial& dm, uint32_t slot, uint32_t idx, float2 uv) {
    if (MSNE_MATERIAL_CONSTANTS && (__ldg(&dm.constMask) >> slot) & 1u) return __ldg(reinterpret_cast<const float4*>(dm.constant[slot]));
    return sampleTexture(sc, idx, uv);
}

// ---- reflection_frame.hlsl:3-85 ------------------------------------------------------------------
struct Frame {
    float3 n, s, t;
    DI Frame inSpaceTranspose(const float* toMesh) const {                              // :23-29
        Frame f;
        f.n = normalize(mulTranspose3(toMesh, n));
        f.s = normalize(mulTranspose3(toMesh, s));
        f.t = normalize(mulTranspose3(toMesh, t));
        return f;
    }
    DI void reorthogonalize() { s = normalize(s - n * dot(n, s)); t = normalize(cross(n, s)); }   // :31-35
    DI float3 worldToFrame(float3 v) const { return mk3(dot(s, v), dot(t, v), dot(n, v)); }       // :37-40
    DI float3 frameToWorld(float3 v) const {                                              // :42-45
        return mk3((s.x * v.x + t.x * v.y) + n.x * v.z, (s.y * v.x + t.y * v.y) + n.y * v.z, (s.z * v.x + t.z * v.y) + n.z * v.z);
    }
};
DI float cosTheta(float3 v) { return v.z; }
DI float cos2Theta(float3 v) { return v.z * v.z; }
DI float sin2Theta(float3 v) { return fmaxf(0.0f, 1.0f - cos2Theta(v)); }
DI float tan2Theta(float3 v) { return sin2Theta(v) / cos2Theta(v); }
DI bool sameHemisphere(float3 a, float3 b) { return a.z * b.z > 0.0f; }

// ---- material.hlsl -------------------------------------------------------------------------------
DI float ggxD(float a, float3 m) {                                                        // :31-35
    const float a2 = a * a, c = cosTheta(m);
    const float inner = (c * c) * (a2 - 1) + 1;
    const float denom = MS_PI * (inner * inner);
    return a2 / denom;
}
DI float ggxLambda(float a, float3 v) {                                                   // :37-41
    const float t2 = tan2Theta(v);
    if (isinf(t2)) return 0.0f;
    return (sqrtf(1.0f + (a * a) * t2) - 1.0f) / 2.0f;
}
DI float ggxG(float a, float3 w_i, float3 w_o) { return 1.0f / (1.0f + ggxLambda(a, w_i) + ggxLambda(a, w_o)); }   // :44-46
DI float3 ggxSample(float a, float3 w_o, float2 sq) {                                     // :50-62
    const float tanThetaSquared = a * a * sq.x / (1 - sq.x);
    const float cosThetaSquared = 1 / (1 + tanThetaSquared);
    const float sinT = sqrtf(fmaxf(0.0f, 1 - cosThetaSquared));
    const float cosT = sqrtf(cosThetaSquared);
    const float phi = 2 * MS_PI * sq.y;
    float3 h = sphericalToCartesian(sinT, cosT, phi);
    if (!sameHemisphere(w_o, h)) h = -h;
    return h;
}
DI float ggxPdf(float a, float3 m) { return ggxD(a, m) * fabsf(cosTheta(m)); }          // :64-66

DI float schlickWeight(float c) { const float x = 1 - c; const float x2 = x * x; return x2 * x2 * x; }   // :76-78
DI float3 schlick3(float c, float3 R0) { const float w = schlickWeight(c); return lerp3(mk3(w), mk3(1.0f), R0); }   // :84-86
DI float fresnelDielectric(float cosThetaI, float etaI, float etaT) {                     // :96-122
    cosThetaI = clampf(cosThetaI, -1, 1);
    if (!(cosThetaI > 0)) { const float tmp = etaI; etaI = etaT; etaT = tmp; cosThetaI = fabsf(cosThetaI); }
    const float sinThetaI = sqrtf(fmaxf(0.0f, 1 - cosThetaI * cosThetaI));
    const float sinThetaT = etaI / etaT * sinThetaI;
    if (sinThetaT >= 1) return 1;
    const float cosThetaT = sqrtf(fmaxf(0.0f, 1 - sinThetaT * sinThetaT));
    const float r_parl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
    const float r_perp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
    return (r_parl * r_parl + r_perp * r_perp) / 2;
}

struct MatSample { float3 dirFs; float pdf; };

// the four variants share one parameter block, loaded ONCE per hit (the reference re-runs ::load and its
// texture fetches inside each of pdf/eval/sample, material.hlsl:411,432,453 — the fetches are pure)
struct MaterialParams {
    uint32_t type;
    float3 color;     // StandardPBR.color / Lambert.r
    float metalness;
    float alpha;      // GGX α = max(roughness², 0.001)
    float ior;        // StandardPBR.ior / Glass.intIOR
    // per-hit invariants of pdf/eval/sample, evaluated once by prepareMaterial with the operations the HLSL repeats in
    // every call (same operands -> same bits): color/π (:264, :153), p_spec (:218, :241), Λ(w_o) (:37-41 inside G)
    float3 diffuse;
    float pSpec;
    float lambdaO;
};

DI float lambertPdf(float3 w_i, float3 w_o) { return sameHemisphere(w_i, w_o) ? fabsf(cosTheta(w_i)) / MS_PI : 0.0f; }   // :157-159
DI MatSample lambertSample(float3 w_o, float2 sq) {                                       // :165-173
    float3 w_i = squareToCosineHemisphere(sq);
    if (w_o.z < 0.0f) w_i.z *= -1;
    MatSample s; s.dirFs = w_i; s.pdf = lambertPdf(w_i, w_o);
    return s;
}
DI float microfacetPdf(const MaterialParams& m, float3 w_i, float3 w_o) {                // :201-205
    if (!sameHemisphere(w_o, w_i)) return 0.0f;
    const float3 h = normalize(w_i + w_o);
    return ggxPdf(m.alpha, h) / (4.0f * dot(w_o, h));
}
DI float pbrPdf(const MaterialParams& m, float3 w_i, float3 w_o) {                        // :240-249
    return lerp1(lambertPdf(w_i, w_o), microfacetPdf(m, w_i, w_o), m.pSpec);
}
#ifndef MSNE_PBR_EVAL_NOINLINE
#define MSNE_PBR_EVAL_NOINLINE 0   // measurement variant: StandardPBR::eval (three call sites per hit) out of line
#endif
// scalar arguments / return in registers (no struct through local memory)
#if MSNE_PBR_EVAL_NOINLINE
__device__ __noinline__
#else
DI
#endif
float3 pbrEvalImpl(float cr, float cg, float cb, float metalness, float alpha, float ior, float dr, float dg, float db, float lambdaO,
                   float wix, float wiy, float wiz, float wox, float woy, float woz);
DI float3 pbrEval(const MaterialParams& m, float3 w_i, float3 w_o) {                      // :251-266
    return pbrEvalImpl(m.color.x, m.color.y, m.color.z, m.metalness, m.alpha, m.ior, m.diffuse.x, m.diffuse.y, m.diffuse.z, m.lambdaO,
                       w_i.x, w_i.y, w_i.z, w_o.x, w_o.y, w_o.z);
}
#if MSNE_PBR_EVAL_NOINLINE
__device__ __noinline__
#else
DI
#endif
float3 pbrEvalImpl(float cr, float cg, float cb, float metalness, float alpha, float ior, float dr, float dg, float db, float lambdaO,
                   float wix, float wiy, float wiz, float wox, float woy, float woz) {
    MaterialParams m;
    m.color = mk3(cr, cg, cb); m.metalness = metalness; m.alpha = alpha; m.ior = ior; m.diffuse = mk3(dr, dg, db); m.lambdaO = lambdaO;
    const float3 w_i = mk3(wix, wiy, wiz), w_o = mk3(wox, woy, woz);
    const float3 h = normalize(w_i + w_o);
    const float fD = fresnelDielectric(dot(w_i, h), MS_AIR_IOR, m.ior);
    const float3 fMetallic = schlick3(dot(w_i, h), m.color);
    const float3 F = lerp3(mk3(fD), fMetallic, m.metalness);
    const float G = 1.0f / (1.0f + ggxLambda(m.alpha, w_i) + m.lambdaO);   // ggxG (:44-46) with the cached Λ(w_o)
    const float D = ggxD(m.alpha, h);
    const float3 specular = sameHemisphere(w_o, w_i) ? (F * G * D) / (4.0f * fabsf(cosTheta(w_i)) * fabsf(cosTheta(w_o))) : mk3(0.0f);
    return specular + (1.0f - m.metalness) * m.diffuse;
}
DI MatSample pbrSample(const MaterialParams& m, float3 w_o, float2 sq) {                  // :217-238 (+ :207-215)
    const float pSpec = m.pSpec;
    MatSample s;
    if (coinFlipRemap(pSpec, sq.x)) {
        const float3 h = ggxSample(m.alpha, w_o, sq);
        const float3 refl = w_o - (2.0f * h) * dot(w_o, h);   // reflect(w_o, h)
        const float3 w_i = -refl;
        const float microPdf = sameHemisphere(w_o, w_i) ? ggxPdf(m.alpha, h) / (4.0f * dot(w_o, h)) : 0.0f;
        const float pdf2 = lambertPdf(w_i, w_o);
        s.pdf = lerp1(pdf2, microPdf, pSpec);
        s.dirFs = w_i;
    } else {
        const MatSample ds = lambertSample(w_o, sq);
        const float pdf2 = microfacetPdf(m, ds.dirFs, w_o);
        s.pdf = lerp1(ds.pdf, pdf2, pSpec);
        s.dirFs = ds.dirFs;
    }
    return s;
}
DI float3 refractDir(float3 wi, float3 n, float eta) {                                    // :334-343
    const float cosThetaI = dot(n, wi);
    const float sin2ThetaI = fmaxf(0.0f, 1 - cosThetaI * cosThetaI);
    const float sin2ThetaT = eta * eta * sin2ThetaI;
    if (sin2ThetaT >= 1) return mk3(0.0f);
    const float cosThetaT = sqrtf(1 - sin2ThetaT);
    return eta * -wi + (eta * cosThetaI - cosThetaT) * n;
}
DI MatSample glassSample(float ior, float3 w_o, float2 sq) {                              // :359-381
    const float fresnel = fresnelDielectric(cosTheta(w_o), MS_AIR_IOR, ior);
    MatSample s;
    if (sq.x < fresnel) {
        s.pdf = fresnel;
        s.dirFs = mk3(-w_o.x, -w_o.y, w_o.z);
    } else {
        float etaI, etaT;
        if (cosTheta(w_o) > 0) { etaI = MS_AIR_IOR; etaT = ior; } else { etaT = MS_AIR_IOR; etaI = ior; }
        s.dirFs = refractDir(w_o, faceForward(mk3(0.0f, 0.0f, 1.0f), w_o), etaI / etaT);
        s.pdf = allZero(s.dirFs) ? 0.0f : 1.0f - fresnel;
    }
    return s;
}
DI float3 glassEval(float ior, float3 w_i, float3 w_o) {                                  // :387-394
    const float fresnel = fresnelDielectric(cosTheta(w_o), MS_AIR_IOR, ior);
    if (sameHemisphere(w_i, w_o)) return mk3(fresnel / fabsf(cosTheta(w_i)));
    return mk3((1.0f - fresnel) / fabsf(cosTheta(w_i)));
}

// MaterialVariant (material.hlsl:395-487)
DI bool matIsDelta(uint32_t type) { return type == MAT_PERFECT_MIRROR || type == MAT_GLASS; }
DI float matPdf(const MaterialParams& m, float3 w_i, float3 w_o) {
    if (m.type == MAT_STANDARD_PBR) return pbrPdf(m, w_i, w_o);
    if (m.type == MAT_LAMBERT) return lambertPdf(w_i, w_o);
    return 0.0f;
}
DI float3 matEval(const MaterialParams& m, float3 w_i, float3 w_o) {
    if (m.type == MAT_STANDARD_PBR) return pbrEval(m, w_i, w_o);
    if (m.type == MAT_LAMBERT) return m.diffuse;
    if (m.type == MAT_PERFECT_MIRROR) return mk3(1.0f / fabsf(cosTheta(w_i)));
    return glassEval(m.ior, w_i, w_o);
}
DI MatSample matSample(const MaterialParams& m, float3 w_o, float2 sq) {
    if (m.type == MAT_STANDARD_PBR) return pbrSample(m, w_o, sq);
    if (m.type == MAT_LAMBERT) return lambertSample(w_o, sq);
    if (m.type == MAT_PERFECT_MIRROR) { MatSample s; s.pdf = 1.0f; s.dirFs = mk3(-w_o.x, -w_o.y, w_o.z); return s; }
    return glassSample(m.ior, w_o, sq);
}
DI MaterialParams loadMaterial(const DeviceScene& sc, const DMaterial& dm, float2 uv) {
    MaterialParams m;
    m.type = dm.type; m.color = mk3(0.0f); m.metalness = 0.0f; m.alpha = 0.001f; m.ior = dm.ior;
    if (dm.type == MAT_STANDARD_PBR) {                                                    // :192-199
        m.color = xyz(sampleMaterialSlot(sc, dm, MAT_SLOT_COLOR, dm.color, uv));
        m.metalness = sampleMaterialSlot(sc, dm, MAT_SLOT_METALNESS, dm.metalness, uv).x;
        const float roughness = sampleMaterialSlot(sc, dm, MAT_SLOT_ROUGHNESS, dm.roughness, uv).x;
        m.alpha = fmaxf(roughness * roughness, 0.001f);
    } else if (dm.type == MAT_LAMBERT) {                                                  // :149-155
        m.color = xyz(sampleMaterialSlot(sc, dm, MAT_SLOT_COLOR, dm.color, uv));
    }
    return m;
}
// w_o = the outgoing direction in the shading frame, the same for every pdf/eval/sample call of this hit
DI void prepareMaterial(MaterialParams& m, float3 w_o) {
    m.diffuse = m.color / MS_PI;
    m.pSpec = 1.0f / (1.0f + (1 - m.metalness));
    m.lambdaO = m.type == MAT_STANDARD_PBR ? ggxLambda(m.alpha, w_o) : 0.0f;
}
DI float3 decodeNormal(float2 rg) {                                                       // :489-492
    rg = 2.0f * rg - mk2(1.0f, 1.0f);
    return mk3(rg.x, rg.y, sqrtf(1.0f - clampf(dot(rg, rg), 0.0f, 1.0f)));
}
DI Frame getTextureFrame(const DeviceScene& sc, const DConstants& k, const DMaterial& dm, float2 uv, const Frame& tangentFrame) {   // :506-517
    const float4 s = sampleMaterialSlot(sc, dm, MAT_SLOT_NORMAL, dm.normal, uv);
    const float3 nTs = k.two_component_normal ? decodeNormal(mk2(s.x, s.y)) : mk3(s.x, s.y, s.z);
    Frame f = tangentFrame;
    f.n = normalize(tangentFrame.frameToWorld(nTs));
    f.reorthogonalize();
    return f;
}

// ---- world.hlsl:86-176 -----------------------------------------------------------------------------

DI float3 ld3(const float* p, uint32_t i) { return mk3(__ldg(p + 3 * (size_t)i), __ldg(p + 3 * (size_t)i + 1), __ldg(p + 3 * (size_t)i + 2)); }
DI float2 ld2(const float* p, uint32_t i) { return mk2(__ldg(p + 2 * (size_t)i), __ldg(p + 2 * (size_t)i + 1)); }
DI float3 interp3(float3 b, float3 v1, float3 v2, float3 v3) { return (b.x * v1 + b.y * v2) + b.z * v3; }
DI float2 interp2(float3 b, float2 v1, float2 v2, float2 v3) { return (b.x * v1 + b.y * v2) + b.z * v3; }

DI void getTangentBitangent(float3 p0, float3 p1, float3 p2, float2 t0, float2 t1, float2 t2, float3& tangent, float3& bitangent) {   // :86-100
    const float2 dT02 = t0 - t2, dT12 = t1 - t2;
    const float3 dP02 = p0 - p2, dP12 = p1 - p2;
    const float det = dT02.x * dT12.y - dT02.y * dT12.x;
    if (det == 0.0f) {
        coordinateSystem(normalize(cross(p2 - p0, p1 - p0)), tangent, bitangent);
    } else {
        tangent = normalize((dT12.y * dP02 - dT02.y * dP12) / det);
        bitangent = normalize((-dT12.x * dP02 + dT02.x * dP12) / det);
    }
}

// Object-space triangle frame (world.hlsl:137-147): tangent / bitangent from the texcoords, geometric normal, reorthogonalised.
DI Frame objectTriangleFrame(float3 p0, float3 p1, float3 p2, float2 t0, float2 t1, float2 t2) {
    Frame f;
    getTangentBitangent(p0, p1, p2, t0, t1, t2, f.s, f.t);
    f.n = normalize(cross(p0 - p2, p1 - p2));
    f.reorthogonalize();
    return f;
}
// attribute indices of a triangle's corners (world.hlsl:130): `cond ? uint3 : float3` promotes both arms to float3 before
// converting back to uint3
DI void attributeIndices(const DConstants& k, uint32_t primitiveIndex, uint32_t ix, uint32_t iy, uint32_t iz, uint32_t& ax, uint32_t& ay, uint32_t& az) {
    if (k.indexed_attributes) { ax = (uint32_t)(float)ix; ay = (uint32_t)(float)iy; az = (uint32_t)(float)iz; }
    else { ax = (uint32_t)(float)(primitiveIndex * 3 + 0); ay = (uint32_t)(float)(primitiveIndex * 3 + 1); az = (uint32_t)(float)(primitiveIndex * 3 + 2); }
}
DI void cornerTexcoords(const DMesh& mesh, uint32_t ax, uint32_t ay, uint32_t az, float2& t0, float2& t1, float2& t2) {
    if (mesh.uv != nullptr) { t0 = ld2(mesh.uv, ax); t1 = ld2(mesh.uv, ay); t2 = ld2(mesh.uv, az); }
    else { t0 = mk2(0, 0); t1 = mk2(1, 0); t2 = mk2(1, 1); }
}
DI void loadRows(const float* rows, float m[12]) {
#pragma unroll
    for (int i = 0; i < 3; i++) { const float4 r = __ldg(reinterpret_cast<const float4*>(rows) + i); m[4 * i] = r.x; m[4 * i + 1] = r.y; m[4 * i + 2] = r.z; m[4 * i + 3] = r.w; }
}

// Per-pooled-triangle frame records (MSNE_TRI_FRAMES, kernels.cu k_tri_frames; 64 B each, DeviceScene::triFrames): everything of
// lookupAndInterpolate(...).inWorld(...) that depends on the TRIANGLE and not on the hit point - the tangent frame from positions and
// texcoords, its reorthogonalisation, and (for members of the merged static BLAS, whose instance is known) its transform to world
// space - is evaluated once per scene build by the very same device functions, so a hit reads the bits it would have computed:
//   f[0] = (object-space s, w = 1 if f[1..3] are in world space, else 0)   f[1] = n   f[2] = s   f[3] = t
// f[1..3]: world space for static members, object space for triangles of shared BLASes (their instance is only known at the hit).
// A hit on a mesh with normals needs f[0], f[1] (one 32-byte sector); s and t of the triangle frame are only read by meshes
// without normals and by the last fallback of the shading-frame selection (integrator.hlsl:93-104).
#ifndef MSNE_TRI_FRAMES
#define MSNE_TRI_FRAMES 1
#endif
struct MeshAttributes {
    float3 position; float2 texcoord; Frame triangleFrame, frame;
    bool triangleFrameComplete;   // false: triangleFrame.s / .t are still in the frame record (completeTriangleFrame)
};

// triangleFrame.s / .t of a hit whose lookup left them in the frame record
DI void completeTriangleFrame(const DeviceScene& sc, MeshAttributes& a, uint32_t instanceIndex, uint32_t triIndex) {
    if (!MSNE_TRI_FRAMES || a.triangleFrameComplete) return;
    const float4* fr = sc.triFrames + 4 * (size_t)triIndex;
    const float4 f0 = __ldg(fr), f2 = __ldg(fr + 2), f3 = __ldg(fr + 3);
    a.triangleFrame.s = mk3(f2.x, f2.y, f2.z); a.triangleFrame.t = mk3(f3.x, f3.y, f3.z);
    if (f0.w == 0.0f) {
        float m[12];
        loadRows(sc.instances[instanceIndex].w2o, m);
        a.triangleFrame.s = normalize(mulTranspose3(m, a.triangleFrame.s));   // Frame::inSpaceTranspose, component by component
        a.triangleFrame.t = normalize(mulTranspose3(m, a.triangleFrame.t));
    }
    a.triangleFrameComplete = true;
}

// MeshAttributes::lookupAndInterpolate(...).inWorld(...)  (world.hlsl:114-176)
// HIT: `tri` is the pooled triangle record of a traced hit, `triIndex` its index (light samples: HIT = false). Its vertices are the mesh's
// object-space positions copied verbatim (runtime.cpp meshTriangles), so the hit path takes them from the 48-byte record the trace
// kernel just touched instead of chasing index -> position (same bits, one dependent hop and nine scalar loads less).
template <bool HIT = false>
DI MeshAttributes lookupAttributes(const DeviceScene& sc, const DConstants& k, uint32_t instanceIndex, uint32_t meshIdx, uint32_t primitiveIndex, float2 attribs,
                                   const TriRec* tri = nullptr, uint32_t triIndex = 0) {
    const DMesh mesh = sc.meshes[meshIdx];
    const float3 bary = mk3(1.0f - attribs.x - attribs.y, attribs.x, attribs.y);
    MeshAttributes a;
    const uint32_t ix = __ldg(mesh.idx + 3 * (size_t)primitiveIndex), iy = __ldg(mesh.idx + 3 * (size_t)primitiveIndex + 1), iz = __ldg(mesh.idx + 3 * (size_t)primitiveIndex + 2);
    float3 p0, p1, p2;
    if (MSNE_POS_FROM_TRIREC && HIT) {
        const float4* tp = reinterpret_cast<const float4*>(tri);
        const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
        p0 = mk3(t0.x, t0.y, t0.z); p1 = mk3(t1.x, t1.y, t1.z); p2 = mk3(t2.x, t2.y, t2.z);
    } else {
        p0 = ld3(mesh.pos, ix); p1 = ld3(mesh.pos, iy); p2 = ld3(mesh.pos, iz);
    }
    a.position = interp3(bary, p0, p1, p2);
    uint32_t ax, ay, az;
    attributeIndices(k, primitiveIndex, ix, iy, iz, ax, ay, az);
    float2 t0, t1, t2;
    cornerTexcoords(mesh, ax, ay, az, t0, t1, t2);
    a.texcoord = interp2(bary, t0, t1, t2);
    const DInstance* in = sc.instances + instanceIndex;
    float m[12];
    loadRows(in->o2w, m);
    a.position = mulPoint(m, a.position);
    loadRows(in->w2o, m);
    if (MSNE_TRI_FRAMES && HIT) {
        const float4* fr = sc.triFrames + 4 * (size_t)triIndex;
        const float4 f0 = __ldg(fr), f1 = __ldg(fr + 1);
        const bool inWorld = f0.w != 0.0f;
        const float3 recN = mk3(f1.x, f1.y, f1.z);
        a.triangleFrame.n = inWorld ? recN : normalize(mulTranspose3(m, recN));
        a.triangleFrameComplete = false;
        if (mesh.nrm != nullptr) {
            const float3 n0 = ld3(mesh.nrm, ax), n1 = ld3(mesh.nrm, ay), n2 = ld3(mesh.nrm, az);
            a.frame.s = mk3(f0.x, f0.y, f0.z);   // object-space s of the triangle frame; t is rebuilt by reorthogonalize()
            a.frame.t = mk3(0.0f);
            a.frame.n = normalize(interp3(bary, n0, n1, n2));
            a.frame.reorthogonalize();
            a.frame = a.frame.inSpaceTranspose(m);
        } else {
            completeTriangleFrame(sc, a, instanceIndex, triIndex);
            a.frame = a.triangleFrame;            // the same three vectors through the same transform
        }
        return a;
    }
    a.triangleFrame = objectTriangleFrame(p0, p1, p2, t0, t1, t2);
    a.frame = a.triangleFrame;
    if (mesh.nrm != nullptr) {
        const float3 n0 = ld3(mesh.nrm, ax), n1 = ld3(mesh.nrm, ay), n2 = ld3(mesh.nrm, az);
        a.frame.n = normalize(interp3(bary, n0, n1, n2));
        a.frame.reorthogonalize();
    }
    // inWorld
    a.triangleFrame = a.triangleFrame.inSpaceTranspose(m);
    a.frame = a.frame.inSpaceTranspose(m);
    a.triangleFrameComplete = true;
    return a;
}

// The part of lookupAndInterpolate(...).inWorld(...) that MeshLights::sample uses (light.hlsl:130-158 reads position,
// texcoord and triangleFrame.n only), fed from the per-entry emitter record (dtypes.h emitterTris) instead of the
// instance -> geometry -> mesh -> index -> vertex chain: same operands, same operations, four dependent hops less.
struct LightPoint { float3 position, n; float2 texcoord; };
DI LightPoint emitterPoint(const DeviceScene& sc, uint32_t entry, uint32_t instanceIndex, float2 attribs) {
    const float4* r = sc.emitterTris + 4 * (size_t)entry;
    const float4 ra = __ldg(r), rb = __ldg(r + 1), rc = __ldg(r + 2), rd = __ldg(r + 3);
    const float3 p0 = mk3(ra.x, ra.y, ra.z), p1 = mk3(rb.x, rb.y, rb.z), p2 = mk3(rc.x, rc.y, rc.z);
    const float2 t0 = mk2(ra.w, rb.w), t1 = mk2(rc.w, rd.x), t2 = mk2(rd.y, rd.z);
    const float3 bary = mk3(1.0f - attribs.x - attribs.y, attribs.x, attribs.y);
    LightPoint lp;
    lp.position = interp3(bary, p0, p1, p2);
    lp.texcoord = interp2(bary, t0, t1, t2);
    lp.n = normalize(cross(p0 - p2, p1 - p2));   // Frame::reorthogonalize (world.hlsl:147) leaves n as is
    const DInstance* in = sc.instances + instanceIndex;
    float m[12];
#pragma unroll
    for (int i = 0; i < 3; i++) { const float4 q = __ldg(reinterpret_cast<const float4*>(in->o2w) + i); m[4 * i] = q.x; m[4 * i + 1] = q.y; m[4 * i + 2] = q.z; m[4 * i + 3] = q.w; }
    lp.position = mulPoint(m, lp.position);
#pragma unroll
    for (int i = 0; i < 3; i++) { const float4 q = __ldg(reinterpret_cast<const float4*>(in->w2o) + i); m[4 * i] = q.x; m[4 * i + 1] = q.y; m[4 * i + 2] = q.z; m[4 * i + 3] = q.w; }
    lp.n = normalize(mulTranspose3(m, lp.n));     // Frame::inSpaceTranspose (reflection_frame.hlsl:23-29)
    return lp;
}

// ---- light.hlsl ------------------------------------------------------------------------------------
DI float areaMeasureToSolidAngleMeasure(float3 pos1, float3 pos2, float3 dir1, float3 dir2) {    // :105-110
    const float3 dlt = pos1 - pos2;
    const float r2 = dot(dlt, dlt);
    const float lightCos = dot(-dir1, dir2);
    return lightCos > 0.0f ? r2 / lightCos : 0.0f;
}
DI float powerHeuristic(uint32_t numf, float fPdf, uint32_t numg, float gPdf) {                  // integrator.hlsl:10-16
    const float f = numf * fPdf, g = numg * gPdf;
    const float f2 = f * f;
    return f2 / (f2 + g * g);
}

// luminanceTexture.Load(uint3(x,y,level)) with out-of-bounds -> 0 (SURVEY §8 a17 hazard)
DI float envLumLoad(const DeviceScene& sc, uint32_t x, uint32_t y, uint32_t level) {
    const uint32_t sz = sc.envSize >> level;
    if (x >= sz || y >= sz) return 0.0f;
    return __ldg(sc.envLum + sc.envLevelOffset[level] + (size_t)y * sz + x);
}
DI float3 envRgbLoad(const DeviceScene& sc, uint32_t x, uint32_t y) {
    if (x >= sc.envSize || y >= sc.envSize) return mk3(0.0f);
    return xyz(__ldg(reinterpret_cast<const float4*>(sc.envRgb) + (size_t)y * sc.envSize + x));
}
struct LightSample { float3 dirWs, radiance; float pdf; };
// EnvMap::sample without the shadow ray (light.hlsl:47-73).
// Two value-preserving shortcuts relative to the literal loop (bit-exact, checked against the oracle's literal one):
//  * the first iteration runs at the 1x1 mip where the (1,0),(0,1),(1,1) loads are out of bounds and return 0
//    (SURVEY 8a hazard): p = 0/(L+0) = 0, coinFlipRemap(0, r) leaves idx at 0 and maps r -> (r-0)/(1-0) = r exactly,
//    so it is skipped (when L == 0 it yields NaN just like the literal loop: 0/0 one level further down);
//  * probs_y are two of the four texels already loaded for probs_x, so they are not fetched again.
// Below the top level idx+1 is always in range, so no bounds checks are needed.
DI LightSample envSample(const DeviceScene& sc, float2 rand) {
    const uint32_t size = sc.envSize, mipCount = sc.envLevels;
    uint32_t ix = 0, iy = 0;
    for (uint32_t level = mipCount - 1; level-- > 0;) {
        ix *= 2; iy *= 2;
#if MSNE_ENV_QUADS
        // quad-major copy of the mip (k_env_fold): the four texels under the parent texel are one 16-byte load
        const float4 qd = __ldg(sc.envQuads + sc.envQuadOffset[level] + (iy >> 1) * (size >> (level + 1)) + (ix >> 1));
        const float l00 = qd.x, l10 = qd.y, l01 = qd.z, l11 = qd.w;
#else
        const uint32_t sz = size >> level;
        const float* mip = sc.envLum + sc.envLevelOffset[level];
        const float l00 = __ldg(mip + (size_t)iy * sz + ix), l10 = __ldg(mip + (size_t)iy * sz + ix + 1);
        const float l01 = __ldg(mip + (size_t)(iy + 1) * sz + ix), l11 = __ldg(mip + (size_t)(iy + 1) * sz + ix + 1);
#endif
        const float px0 = l00 + l01;
        const float px1 = l10 + l11;
        const bool right = coinFlipRemap(px1 / (px0 + px1), rand.x);
        ix += right ? 1u : 0u;
        const float py0 = right ? l10 : l00;
        const float py1 = right ? l11 : l01;
        iy += coinFlipRemap(py1 / (py0 + py1), rand.y) ? 1u : 0u;
    }
    const float integral = envLumLoad(sc, 0, 0, mipCount - 1);
    const float discretePdf = envLumLoad(sc, ix, iy, 0) * float(size * size) / integral;
    const float2 uv = (mk2((float)ix, (float)iy) + rand) / mk2((float)size, (float)size);
    LightSample ls;
    ls.pdf = discretePdf / (4.0f * MS_PI);
    ls.dirWs = squareToEqualAreaSphere(uv);
    ls.radiance = envRgbLoad(sc, ix, iy);
    return ls;
}
// EnvMap::eval (light.hlsl:83-97)
DI void envEval(const DeviceScene& sc, float3 dirWs, float3& radiance, float& pdf) {
    const uint32_t size = sc.envSize, mipCount = sc.envLevels;
    const float2 uv = squareToEqualAreaSphereInverse(dirWs);
    const float integral = envLumLoad(sc, 0, 0, mipCount - 1);
    const uint32_t ix = min(floatToUint(uv.x * (float)size), size), iy = min(floatToUint(uv.y * (float)size), size);   // upper clamp is `size` (sic)
    const float discretePdf = envLumLoad(sc, ix, iy, 0) * float(size * size) / integral;
    pdf = discretePdf / (4.0f * MS_PI);
    radiance = envRgbLoad(sc, ix, iy);
}
// EnvMap::incomingRadiance (light.hlsl:99-102): bilinear, mirrored repeat
DI float3 envIncomingRadiance(const DeviceScene& sc, float3 dirWs) {
    const float2 uv = squareToEqualAreaSphereInverse(dirWs);
    const int n = (int)sc.envSize;
    const float4* rgb = reinterpret_cast<const float4*>(sc.envRgb);
    if (n == 1) return xyz(__ldg(rgb));
    const float u = uv.x * (float)n - 0.5f, v = uv.y * (float)n - 0.5f;
    const float fu = floorf(u), fv = floorf(v);
    const float a = u - fu, b = v - fv;
    const int i0 = wrapMirror((int)fu, n), i1 = wrapMirror((int)fu + 1, n), j0 = wrapMirror((int)fv, n), j1 = wrapMirror((int)fv + 1, n);
    const float4 r = lerp4(lerp4(__ldg(rgb + (size_t)j0 * n + i0), __ldg(rgb + (size_t)j0 * n + i1), a),
                           lerp4(__ldg(rgb + (size_t)j1 * n + i0), __ldg(rgb + (size_t)j1 * n + i1), a), b);
    return xyz(r);
}

}  // namespace msne
