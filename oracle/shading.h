// ORACLE — TEST INFRASTRUCTURE ONLY (see vec.h).
// CPU restatement of the reference's device program, one struct/function per HLSL twin:
//   reflection_frame.hlsl -> Frame;  material.hlsl -> GGX/Fresnel/Lambert/StandardPBR/
//   PerfectMirror/Glass/MaterialVariant/getTextureFrame/getEmissive;  world.hlsl ->
//   MeshAttributes;  light.hlsl -> EnvMap/MeshLights;  camera.hlsl -> generateRay;
//   integrator.hlsl -> PathTracingIntegrator::incomingRadiance;  main.hlsl -> raygen/storeColor.
// All paths relative to /root/reference/shaders/hrtsystem/.
#pragma once
#include <cstdio>
#include "bvh.h"

namespace orc {

// ---- reflection_frame.hlsl:3-85 ---------------------------------------------------------
struct Frame {
    float3 n, s, t;
    Frame inSpaceTranspose(const Mat3x4& toMesh) const {                           // :23-29 + world.hlsl:172-173
        Frame f;
        f.n = normalize(mul_transpose3(toMesh, n));
        f.s = normalize(mul_transpose3(toMesh, s));
        f.t = normalize(mul_transpose3(toMesh, t));
        return f;
    }
    void reorthogonalize() {                                                       // :31-35
        s = normalize(s - n * dot(n, s));
        t = normalize(cross(n, s));
    }
    float3 worldToFrame(float3 v) const { return {dot(s, v), dot(t, v), dot(n, v)}; }            // :37-40
    float3 frameToWorld(float3 v) const {                                          // :42-45: mul(transpose({s,t,n}), v)
        return {(s.x * v.x + t.x * v.y) + n.x * v.z, (s.y * v.x + t.y * v.y) + n.y * v.z, (s.z * v.x + t.z * v.y) + n.z * v.z};
    }
    static float cosTheta(float3 v) { return v.z; }
    static float cos2Theta(float3 v) { return v.z * v.z; }
    static float sin2Theta(float3 v) { return std::fmax(0.0f, 1.0f - cos2Theta(v)); }
    static float tan2Theta(float3 v) { return sin2Theta(v) / cos2Theta(v); }
    static bool sameHemisphere(float3 a, float3 b) { return a.z * b.z > 0.0f; }
};

// ---- material.hlsl ----------------------------------------------------------------------
struct GGX {                                                                       // :20-67
    float a;  // α
    float D(float3 m) const {                                                      // :31-35
        float a2 = a * a;
        float c = Frame::cosTheta(m);
        float inner = (c * c) * (a2 - 1) + 1;
        float denom = PI * (inner * inner);
        return a2 / denom;
    }
    float Lambda(float3 v) const {                                                 // :37-41
        float t2 = Frame::tan2Theta(v);
        if (std::isinf(t2)) return 0.0f;
        return (std::sqrt(1.0f + (a * a) * t2) - 1.0f) / 2.0f;
    }
    float G(float3 w_i, float3 w_o) const { return 1.0f / (1.0f + Lambda(w_i) + Lambda(w_o)); }  // :44-46
    float3 sample(float3 w_o, float2 sq) const {                                   // :50-62
        float tanThetaSquared = a * a * sq.x / (1 - sq.x);
        float cosThetaSquared = 1 / (1 + tanThetaSquared);
        float sinTheta = std::sqrt(std::fmax(0.0f, 1 - cosThetaSquared));
        float cosTheta = std::sqrt(cosThetaSquared);
        float phi = 2 * PI * sq.y;
        float3 h = sphericalToCartesian(sinTheta, cosTheta, phi);
        if (!Frame::sameHemisphere(w_o, h)) h = -h;
        return h;
    }
    float pdf(float3 /*w_o*/, float3 m) const { return D(m) * std::fabs(Frame::cosTheta(m)); }   // :64-66
};

namespace Fresnel {                                                                // :71-123
inline float schlickWeight(float c) { float x = 1 - c; float x2 = x * x; return x2 * x2 * x; }   // :76-78 pow(1-c,5)
inline float3 schlick(float c, float3 R0) { float w = schlickWeight(c); return lerp(f3(w), f3(1.0f), R0); }  // :84-86
inline float dielectric(float cosThetaI, float etaI, float etaT) {                 // :96-122
    cosThetaI = clampf(cosThetaI, -1, 1);
    bool entering = cosThetaI > 0;
    if (!entering) { float tmp = etaI; etaI = etaT; etaT = tmp; cosThetaI = std::fabs(cosThetaI); }
    float sinThetaI = std::sqrt(std::fmax(0.0f, 1 - cosThetaI * cosThetaI));
    float sinThetaT = etaI / etaT * sinThetaI;
    if (sinThetaT >= 1) return 1;
    float cosThetaT = std::sqrt(std::fmax(0.0f, 1 - sinThetaT * sinThetaT));
    float r_parl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
    float r_perp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
    return (r_parl * r_parl + r_perp * r_perp) / 2;
}
}  // namespace Fresnel

struct MaterialSample { float3 dirFs; float pdf; };                                // :125-128

struct Lambert {                                                                   // :137-175
    float3 r;
    float pdf(float3 w_i, float3 w_o) const { return Frame::sameHemisphere(w_i, w_o) ? std::fabs(Frame::cosTheta(w_i)) / PI : 0.0f; }
    float3 eval(float3, float3) const { return r / PI; }
    MaterialSample sample(float3 w_o, float2 sq) const {
        float3 w_i = squareToCosineHemisphere(sq);
        if (w_o.z < 0.0f) w_i.z *= -1;
        return {w_i, pdf(w_i, w_o)};
    }
};

struct StandardPBR {                                                               // :179-270
    GGX distr; float3 color; float metalness; float ior;
    float microfacetPdf(float3 w_i, float3 w_o) const {                            // :201-205
        if (!Frame::sameHemisphere(w_o, w_i)) return 0.0f;
        float3 h = normalize(w_i + w_o);
        return distr.pdf(w_o, h) / (4.0f * dot(w_o, h));
    }
    MaterialSample microfacetSample(float3 w_o, float2 sq) const {                 // :207-215
        float3 h = distr.sample(w_o, sq);
        // -reflect(w_o, h) = -(w_o - 2*h*dot(w_o,h))
        float3 refl = w_o - (2.0f * h) * dot(w_o, h);
        float3 w_i = -refl;
        MaterialSample s;
        s.pdf = Frame::sameHemisphere(w_o, w_i) ? distr.pdf(w_o, h) / (4.0f * dot(w_o, h)) : 0.0f;
        s.dirFs = w_i;
        return s;
    }
    MaterialSample sample(float3 w_o, float2 sq) const {                           // :217-238
        float specularWeight = 1, diffuseWeight = 1 - metalness;
        float pSpec = specularWeight / (specularWeight + diffuseWeight);
        MaterialSample s;
        Lambert diffuse{color};
        if (coinFlipRemap(pSpec, sq.x)) {
            MaterialSample ms = microfacetSample(w_o, sq);
            float pdf2 = diffuse.pdf(ms.dirFs, w_o);
            s.pdf = lerp(pdf2, ms.pdf, pSpec);
            s.dirFs = ms.dirFs;
        } else {
            MaterialSample ds = diffuse.sample(w_o, sq);
            float pdf2 = microfacetPdf(ds.dirFs, w_o);
            s.pdf = lerp(ds.pdf, pdf2, pSpec);
            s.dirFs = ds.dirFs;
        }
        return s;
    }
    float pdf(float3 w_i, float3 w_o) const {                                      // :240-249
        float specularWeight = 1, diffuseWeight = 1 - metalness;
        float pSpec = specularWeight / (specularWeight + diffuseWeight);
        float lambert_pdf = Lambert{color}.pdf(w_i, w_o);
        float micro_pdf = microfacetPdf(w_i, w_o);
        return lerp(lambert_pdf, micro_pdf, pSpec);
    }
    float3 eval(float3 w_i, float3 w_o) const {                                    // :251-266
        float3 h = normalize(w_i + w_o);
        float fD = Fresnel::dielectric(dot(w_i, h), AIR_IOR, ior);
        float3 fMetallic = Fresnel::schlick(dot(w_i, h), color);
        float3 F = lerp(f3(fD), fMetallic, metalness);
        float G = distr.G(w_i, w_o);
        float D = distr.D(h);
        float3 specular = Frame::sameHemisphere(w_o, w_i)
                              ? (F * G * D) / (4.0f * std::fabs(Frame::cosTheta(w_i)) * std::fabs(Frame::cosTheta(w_o)))
                              : f3(0.0f);
        float3 diffuse = Lambert{color}.eval(w_i, w_o);
        return specular + (1.0f - metalness) * diffuse;
    }
};

struct PerfectMirror {                                                             // :313-332
    MaterialSample sample(float3 w_o, float2) const { return {f3(-w_o.x, -w_o.y, w_o.z), 1.0f}; }
    float pdf(float3, float3) const { return 0.0f; }
    float3 eval(float3 w_i, float3) const { return f3(1.0f / std::fabs(Frame::cosTheta(w_i))); }
};

inline float3 refractDir(float3 wi, float3 n, float eta) {                         // :334-343
    float cosThetaI = dot(n, wi);
    float sin2ThetaI = std::fmax(0.0f, 1 - cosThetaI * cosThetaI);
    float sin2ThetaT = eta * eta * sin2ThetaI;
    if (sin2ThetaT >= 1) return f3(0.0f);
    float cosThetaT = std::sqrt(1 - sin2ThetaT);
    return eta * -wi + (eta * cosThetaI - cosThetaT) * n;
}

struct Glass {                                                                     // :345-393
    float intIOR;
    MaterialSample sample(float3 w_o, float2 sq) const {
        float fresnel = Fresnel::dielectric(Frame::cosTheta(w_o), AIR_IOR, intIOR);
        MaterialSample s;
        if (sq.x < fresnel) {
            s.pdf = fresnel;
            s.dirFs = f3(-w_o.x, -w_o.y, w_o.z);
        } else {
            float etaI, etaT;
            if (Frame::cosTheta(w_o) > 0) { etaI = AIR_IOR; etaT = intIOR; } else { etaT = AIR_IOR; etaI = intIOR; }
            s.dirFs = refractDir(w_o, faceForward(f3(0.0f, 0.0f, 1.0f), w_o), etaI / etaT);
            s.pdf = all_zero(s.dirFs) ? 0.0f : 1.0f - fresnel;
        }
        return s;
    }
    float pdf(float3, float3) const { return 0.0f; }
    float3 eval(float3 w_i, float3 w_o) const {
        float fresnel = Fresnel::dielectric(Frame::cosTheta(w_o), AIR_IOR, intIOR);
        if (Frame::sameHemisphere(w_i, w_o)) return f3(fresnel / std::fabs(Frame::cosTheta(w_i)));
        return f3((1.0f - fresnel) / std::fabs(Frame::cosTheta(w_i)));
    }
};

struct SceneData;  // below

// MaterialVariant (material.hlsl:395-487): the reference re-runs ::load (texture fetches) inside each of
// pdf/eval/sample; the fetches are pure so loading once is value-identical.
struct MaterialVariant {
    uint32_t type;
    StandardPBR pbr; Lambert lambert; Glass glass; PerfectMirror mirror;
    static MaterialVariant load(const SceneData& sc, const MaterialRec& m, float2 uv);
    float pdf(float3 w_i, float3 w_o) const {
        switch (type) { case MAT_STANDARD_PBR: return pbr.pdf(w_i, w_o); case MAT_LAMBERT: return lambert.pdf(w_i, w_o);
                        case MAT_PERFECT_MIRROR: return mirror.pdf(w_i, w_o); default: return glass.pdf(w_i, w_o); }
    }
    float3 eval(float3 w_i, float3 w_o) const {
        switch (type) { case MAT_STANDARD_PBR: return pbr.eval(w_i, w_o); case MAT_LAMBERT: return lambert.eval(w_i, w_o);
                        case MAT_PERFECT_MIRROR: return mirror.eval(w_i, w_o); default: return glass.eval(w_i, w_o); }
    }
    MaterialSample sample(float3 w_o, float2 sq) const {
        switch (type) { case MAT_STANDARD_PBR: return pbr.sample(w_o, sq); case MAT_LAMBERT: return lambert.sample(w_o, sq);
                        case MAT_PERFECT_MIRROR: return mirror.sample(w_o, sq); default: return glass.sample(w_o, sq); }
    }
    bool isDelta() const { return type == MAT_PERFECT_MIRROR || type == MAT_GLASS; }
};

struct SceneData {
    std::vector<Texture> textures;
    std::vector<MaterialRec> materials;
    std::vector<MeshRec> meshes;
    std::vector<GeometryRec> geometries;  // flat jagged array (Accel.zig:73-76)
    std::vector<InstanceRec> instances;
    std::vector<AliasEntry> aliasTable;   // [0] = header (Accel.zig:531-533)
    EnvMapData env;
    Accel accel;
    Constants consts;
    bool brute = false;
    mutable uint64_t raysClosest = 0, raysShadow = 0;  // only touched via per-thread counters in render()
};

inline MaterialVariant MaterialVariant::load(const SceneData& sc, const MaterialRec& m, float2 uv) {
    MaterialVariant v; v.type = m.type;
    v.pbr = {}; v.lambert = {}; v.glass = {};
    if (m.type == MAT_STANDARD_PBR) {                                              // :192-199
        float4 c = sample_bilinear(sc.textures[m.color], uv, false);
        v.pbr.color = {c.x, c.y, c.z};
        v.pbr.metalness = sample_bilinear(sc.textures[m.metalness], uv, false).x;
        float roughness = sample_bilinear(sc.textures[m.roughness], uv, false).x;
        v.pbr.distr.a = std::fmax(roughness * roughness, 0.001f);
        v.pbr.ior = m.ior;
    } else if (m.type == MAT_LAMBERT) {                                            // :149-155
        float4 c = sample_bilinear(sc.textures[m.color], uv, false);
        v.lambert.r = {c.x, c.y, c.z};
    } else if (m.type == MAT_GLASS) {                                              // :353-357
        v.glass.intIOR = m.ior;
    }
    return v;
}

// material.hlsl:489-522
inline float3 decodeNormal(float2 rg) {
    rg = rg * 2.0f - f2(1.0f, 1.0f);
    return {rg.x, rg.y, std::sqrt(1.0f - saturate(dot(rg, rg)))};
}
inline Frame getTextureFrame(const SceneData& sc, uint32_t materialIndex, float2 uv, const Frame& tangentFrame) {
    const MaterialRec& m = sc.materials[materialIndex];
    float4 s = sample_bilinear(sc.textures[m.normal], uv, false);
    float3 nTs = sc.consts.two_component_normal_texture ? decodeNormal({s.x, s.y}) : f3(s.x, s.y, s.z);
    float3 nWs = normalize(tangentFrame.frameToWorld(nTs));
    Frame f = tangentFrame; f.n = nWs; f.reorthogonalize();
    return f;
}
inline float3 getEmissive(const SceneData& sc, uint32_t materialIndex, float2 uv) {
    float4 e = sample_bilinear(sc.textures[sc.materials[materialIndex].emissive], uv, false);
    return {e.x, e.y, e.z};
}

// ---- world.hlsl:86-176 ------------------------------------------------------------------
inline void getTangentBitangent(float3 p0, float3 p1, float3 p2, float2 t0, float2 t1, float2 t2, float3& tangent, float3& bitangent) {
    float2 dT02 = t0 - t2, dT12 = t1 - t2;
    float3 dP02 = p0 - p2, dP12 = p1 - p2;
    float det = dT02.x * dT12.y - dT02.y * dT12.x;
    if (det == 0.0f) {
        coordinateSystem(normalize(cross(p2 - p0, p1 - p0)), tangent, bitangent);
    } else {
        tangent = normalize((dT12.y * dP02 - dT02.y * dP12) / det);
        bitangent = normalize((-dT12.x * dP02 + dT02.x * dP12) / det);
    }
}
inline float3 interp(float3 b, float3 v1, float3 v2, float3 v3) { return b.x * v1 + b.y * v2 + b.z * v3; }
inline float2 interp(float3 b, float2 v1, float2 v2, float2 v3) { return b.x * v1 + b.y * v2 + b.z * v3; }

struct MeshAttributes {
    float3 position; float2 texcoord; Frame triangleFrame, frame;
    static MeshAttributes lookupAndInterpolate(const SceneData& sc, uint32_t instanceIndex, uint32_t geometryIndex, uint32_t primitiveIndex, float2 attribs) {
        uint32_t instanceID = sc.instances[instanceIndex].customIndex;
        const MeshRec& mesh = sc.meshes[sc.geometries[instanceID + geometryIndex].mesh];
        float3 bary = {1.0f - attribs.x - attribs.y, attribs.x, attribs.y};
        MeshAttributes a;
        uint3 ind = mesh.indices[primitiveIndex];
        float3 p0 = mesh.positions[ind.x], p1 = mesh.positions[ind.y], p2 = mesh.positions[ind.z];
        a.position = interp(bary, p0, p1, p2);
        // :130 — the non-indexed path goes through float3, i.e. exact only below 2^24
        // (the HLSL ternary `cond ? uint3 : float3` promotes BOTH arms to float3 before the uint3 conversion)
        uint3 ai = sc.consts.indexed_attributes
                       ? uint3{(uint32_t)(float)ind.x, (uint32_t)(float)ind.y, (uint32_t)(float)ind.z}
                       : uint3{(uint32_t)(float)(primitiveIndex * 3 + 0), (uint32_t)(float)(primitiveIndex * 3 + 1), (uint32_t)(float)(primitiveIndex * 3 + 2)};
        float2 t0, t1, t2;
        if (!mesh.texcoords.empty()) { t0 = mesh.texcoords[ai.x]; t1 = mesh.texcoords[ai.y]; t2 = mesh.texcoords[ai.z]; }
        else { t0 = {0, 0}; t1 = {1, 0}; t2 = {1, 1}; }
        a.texcoord = interp(bary, t0, t1, t2);
        getTangentBitangent(p0, p1, p2, t0, t1, t2, a.triangleFrame.s, a.triangleFrame.t);
        a.triangleFrame.n = normalize(cross(p0 - p2, p1 - p2));
        a.triangleFrame.reorthogonalize();
        if (!mesh.normals.empty()) {
            float3 n0 = mesh.normals[ai.x], n1 = mesh.normals[ai.y], n2 = mesh.normals[ai.z];
            a.frame = a.triangleFrame;
            a.frame.n = normalize(interp(bary, n0, n1, n2));
            a.frame.reorthogonalize();
        } else {
            a.frame = a.triangleFrame;
        }
        return a;
    }
    MeshAttributes inWorld(const SceneData& sc, uint32_t instanceIndex) const {    // :166-176
        const InstanceRec& in = sc.instances[instanceIndex];
        MeshAttributes r = *this;
        r.position = mul_point(in.transform, position);
        r.triangleFrame = triangleFrame.inSpaceTranspose(in.worldToInstance);
        r.frame = frame.inSpaceTranspose(in.worldToInstance);
        return r;
    }
};

// ---- light.hlsl -------------------------------------------------------------------------
struct LightSample { float3 dirWs, radiance; float pdf; };
struct LightEval { float3 radiance; float pdf; };

// minAlpha: diagnostic only (never feeds a result) - the smallest GGX alpha of any StandardPBR surface the paths counted here
// touched; tests use it to tell pixels whose value went through a near-singular D(m) (material.hlsl:31-35) from the rest.
struct TraceCounters { uint64_t closest = 0, shadow = 0; float minAlpha = 1e30f; };

inline bool shadowHit(const SceneData& sc, TraceCounters& tc, float3 origin, float3 dir, float tmax) {  // intersection.hlsl:34-47
    tc.shadow++;
    return sc.accel.trace(origin, dir, 0.0f, tmax, true, sc.brute).hit();
}

inline float areaMeasureToSolidAngleMeasure(float3 pos1, float3 pos2, float3 dir1, float3 dir2) {       // :105-110
    float r2 = dot(pos1 - pos2, pos1 - pos2);
    float lightCos = dot(-dir1, dir2);
    return lightCos > 0.0f ? r2 / lightCos : 0.0f;
}

struct EnvMap {                                                                    // :34-103
    const EnvMapData* d;
    // luminanceTexture.Load(uint3(x,y,level)); out-of-bounds reads return 0 (SURVEY.md §8 a17 hazard)
    float lumLoad(uint32_t x, uint32_t y, uint32_t level) const {
        uint32_t sz = d->size >> level;
        if (x >= sz || y >= sz) return 0.0f;
        return d->lum[level][(size_t)y * sz + x];
    }
    float3 rgbLoad(uint32_t x, uint32_t y) const {
        if (x >= d->size || y >= d->size) return f3(0.0f);
        float4 v = d->rgb[(size_t)y * d->size + x]; return {v.x, v.y, v.z};
    }
    static uint32_t ilog2(uint32_t v) { uint32_t l = 0; while (v >>= 1) l++; return l; }
    LightSample sample(const SceneData& sc, TraceCounters& tc, float3 positionWs, float3 normalWs, float2 rand) const {  // :47-80
        const uint32_t size = d->size;
        const uint32_t mipCount = ilog2(size) + 1;
        uint32_t ix = 0, iy = 0;
        for (uint32_t level = mipCount; level-- > 0;) {
            ix *= 2; iy *= 2;
            const float px0 = lumLoad(ix + 0, iy + 0, level) + lumLoad(ix + 0, iy + 1, level);
            const float px1 = lumLoad(ix + 1, iy + 0, level) + lumLoad(ix + 1, iy + 1, level);
            ix += coinFlipRemap(px1 / (px0 + px1), rand.x) ? 1u : 0u;
            const float py0 = lumLoad(ix, iy + 0, level);
            const float py1 = lumLoad(ix, iy + 1, level);
            iy += coinFlipRemap(py1 / (py0 + py1), rand.y) ? 1u : 0u;
        }
        const float integral = lumLoad(0, 0, mipCount - 1);
        const float discretePdf = lumLoad(ix, iy, 0) * float(size * size) / integral;
        const float2 uv = (f2((float)ix, (float)iy) + rand) / f2((float)size, (float)size);
        LightSample ls;
        ls.pdf = discretePdf / (4.0f * PI);
        ls.dirWs = squareToEqualAreaSphere(uv);
        ls.radiance = rgbLoad(ix, iy);
        if (ls.pdf > 0.0f && shadowHit(sc, tc, offsetAlongNormal(positionWs, faceForward(normalWs, ls.dirWs)), ls.dirWs, INFINITY_F)) ls.pdf = 0.0f;
        return ls;
    }
    LightEval eval(float3 dirWs) const {                                           // :83-97
        const uint32_t size = d->size;
        const uint32_t mipCount = ilog2(size) + 1;
        const float2 uv = squareToEqualAreaSphereInverse(dirWs);
        const float integral = lumLoad(0, 0, mipCount - 1);
        // clamp(uint2(uv*size), 0, size): upper clamp is `size`, not size-1 (hazard; OOB -> 0)
        float fx = uv.x * (float)size, fy = uv.y * (float)size;
        uint32_t ix = std::min((uint32_t)fx, size), iy = std::min((uint32_t)fy, size);
        const float discretePdf = lumLoad(ix, iy, 0) * float(size * size) / integral;
        LightEval l;
        l.pdf = discretePdf / (4.0f * PI);
        l.radiance = rgbLoad(ix, iy);
        return l;
    }
    float3 incomingRadiance(float3 dirWs) const {                                  // :99-102
        float2 uv = squareToEqualAreaSphereInverse(dirWs);
        Texture view;  // bilinear with the mirrored-repeat background sampler
        const uint32_t n = d->size;
        if (n == 1) { float4 v = d->rgb[0]; return {v.x, v.y, v.z}; }
        float u = uv.x * (float)n - 0.5f, v = uv.y * (float)n - 0.5f;
        float fu = std::floor(u), fv = std::floor(v);
        float a = u - fu, b = v - fv;
        int i0 = wrap_mirror((int)fu, n), i1 = wrap_mirror((int)fu + 1, n), j0 = wrap_mirror((int)fv, n), j1 = wrap_mirror((int)fv + 1, n);
        float4 r = lerp4(lerp4(d->rgb[(size_t)j0 * n + i0], d->rgb[(size_t)j0 * n + i1], a),
                         lerp4(d->rgb[(size_t)j1 * n + i0], d->rgb[(size_t)j1 * n + i1], a), b);
        return {r.x, r.y, r.z};
    }
};

struct MeshLights {                                                                // :119-158
    LightSample sample(const SceneData& sc, TraceCounters& tc, float3 positionWs, float3 triangleNormalDirWs, float2 rand) const {
        LightSample ls; ls.pdf = 0.0f; ls.dirWs = f3(0.0f); ls.radiance = f3(0.0f);
        if (sc.aliasTable.empty()) return ls;
        uint32_t entryCount = sc.aliasTable[0].alias;
        float sum = sc.aliasTable[0].select;
        if (entryCount == 0 || sum == 0) return ls;
        // sampleAlias (mappings.hlsl:114-126), offset 1
        float scaled = rand.x * (float)entryCount;
        uint32_t idx = (uint32_t)scaled;
        // hazard: rand.x*count can round up to count (reference would read one entry past the table);
        // restated as a clamp to the last entry.
        if (idx >= entryCount) idx = entryCount - 1;
        rand.x = scaled - std::floor(scaled);
        AliasEntry e = sc.aliasTable[1 + idx];
        if (!coinFlipRemap(e.select, rand.x)) { idx = e.alias; e = sc.aliasTable[1 + idx]; }
        uint32_t instanceID = sc.instances[e.instance].customIndex;
        float2 bary = squareToTriangle(rand);
        MeshAttributes attrs = MeshAttributes::lookupAndInterpolate(sc, e.instance, e.geometry, e.primitive, bary).inWorld(sc, e.instance);
        ls.radiance = getEmissive(sc, sc.geometries[instanceID + e.geometry].material, attrs.texcoord);
        ls.dirWs = normalize(attrs.position - positionWs);
        ls.pdf = areaMeasureToSolidAngleMeasure(attrs.position, positionWs, ls.dirWs, attrs.triangleFrame.n) / sum;
        float3 offsetLight = offsetAlongNormal(attrs.position, attrs.triangleFrame.n);
        float3 offsetShading = offsetAlongNormal(positionWs, faceForward(triangleNormalDirWs, ls.dirWs));
        float tmax = distance(offsetLight, offsetShading);
        if (ls.pdf > 0.0f && shadowHit(sc, tc, offsetShading, normalize(offsetLight - offsetShading), tmax)) ls.pdf = 0.0f;
        return ls;
    }
};

// ---- camera.hlsl:14-43 ------------------------------------------------------------------
struct RayDesc { float3 Origin, Direction; float TMin, TMax; };
inline RayDesc generateRay(const Lens& c, uint32_t W, uint32_t H, float2 uv, float2 rand) {
    float aspect = float(W) / float(H);
    float3 w = c.forward * -1.0f;
    float3 u = normalize(cross(c.up, w));
    float3 v = cross(w, u);
    float h = std::tan(c.vfov / 2.0f);
    float viewport_height = 2.0f * h * c.focus_distance;
    float viewport_width = aspect * viewport_height;
    float3 horizontal = u * viewport_width;
    float3 vertical = v * viewport_height;
    float3 llc = c.origin - (horizontal / 2.0f) - (vertical / 2.0f) - (w * c.focus_distance);
    float2 sampled_rand = squareToUniformDiskConcentric(rand);
    float2 rd = c.aperture * sampled_rand / 2.0f;
    float3 defocusOffset = u * rd.x + v * rd.y;
    RayDesc r;
    r.Origin = c.origin + defocusOffset;
    r.Direction = normalize(llc + uv.x * horizontal + uv.y * vertical - defocusOffset - c.origin);
    r.TMin = 0; r.TMax = INFINITY_F;
    return r;
}

// ---- integrator.hlsl --------------------------------------------------------------------
inline float powerHeuristic(uint32_t numf, float fPdf, uint32_t numg, float gPdf) {  // :10-16
    float f = numf * fPdf, g = numg * gPdf;
    float f2_ = f * f;
    return f2_ / (f2_ + g * g);
}

template <class Light>
inline float3 estimateDirectMISLight(const SceneData& sc, TraceCounters& tc, const Frame& frame, const Light& light, const MaterialVariant& material,
                                     float3 outgoingDirFs, float3 positionWs, float3 triangleNormalDirWs, float2 rand, uint32_t samplesTaken) {  // :20-35
    LightSample ls = light.sample(sc, tc, positionWs, triangleNormalDirWs, rand);
    if (ls.pdf > 0.0f) {
        float3 lightDirFs = frame.worldToFrame(ls.dirWs);
        float scatteringPdf = material.pdf(lightDirFs, outgoingDirFs);
        if (scatteringPdf > 0.0f) {
            float3 brdf = material.eval(lightDirFs, outgoingDirFs);
            float weight = powerHeuristic(samplesTaken, ls.pdf, 1, scatteringPdf);
            return ls.radiance * brdf * std::fabs(Frame::cosTheta(lightDirFs)) * weight / ls.pdf;
        }
    }
    return f3(0.0f);
}

static thread_local bool g_debugPath = false;   // diagnostics: OracleDebugPath prints every bounce of one path
inline float3 incomingRadiance(const SceneData& sc, TraceCounters& tc, RayDesc initialRay, Rng& rng) {   // :68-183
    const Constants& k = sc.consts;
    EnvMap envMap{&sc.env};
    MeshLights meshLights;
    float3 accumulatedColor = f3(0.0f);
    RayDesc ray = initialRay;
    float3 throughput = f3(1.0f);
    uint32_t bounceCount = 0;
    float lastMaterialPdf = 0.0f;  // read only when bounceCount>0
    bool isLastMaterialDelta = false;

    for (;;) {
        tc.closest++;
        Intersection its = sc.accel.trace(ray.Origin, ray.Direction, ray.TMin, ray.TMax, false, sc.brute);
        if (!its.hit()) break;

        uint32_t instanceID = sc.instances[its.instanceIndex].customIndex;
        const GeometryRec& geometry = sc.geometries[instanceID + its.geometryIndex];
        if (g_debugPath)
            fprintf(stderr, "bounce %u: o=(%.9g %.9g %.9g) d=(%.9g %.9g %.9g) -> inst %u geom %u prim %u bary (%.9g %.9g) material %u type %u | throughput (%g %g %g) acc (%g %g %g)\n",
                    bounceCount, ray.Origin.x, ray.Origin.y, ray.Origin.z, ray.Direction.x, ray.Direction.y, ray.Direction.z, its.instanceIndex, its.geometryIndex,
                    its.primitiveIndex, its.barycentrics.x, its.barycentrics.y, geometry.material, sc.materials[geometry.material].type, throughput.x, throughput.y, throughput.z,
                    accumulatedColor.x, accumulatedColor.y, accumulatedColor.z);
        MeshAttributes attrs = MeshAttributes::lookupAndInterpolate(sc, its.instanceIndex, its.geometryIndex, its.primitiveIndex, its.barycentrics).inWorld(sc, its.instanceIndex);
        Frame textureFrame = getTextureFrame(sc, geometry.material, attrs.texcoord, attrs.frame);
        float3 emissiveLight = getEmissive(sc, geometry.material, attrs.texcoord);
        MaterialVariant material = MaterialVariant::load(sc, sc.materials[geometry.material], attrs.texcoord);
        if (material.type == MAT_STANDARD_PBR && material.pbr.distr.a < tc.minAlpha) tc.minAlpha = material.pbr.distr.a;   // diagnostic

        float3 outgoingDirWs = -ray.Direction;

        bool frontfacing = dot(attrs.triangleFrame.n, outgoingDirWs) > 0;           // :93-104
        Frame shadingFrame;
        if ((frontfacing && dot(outgoingDirWs, textureFrame.n) > 0) || (!frontfacing && -dot(outgoingDirWs, textureFrame.n) > 0)) shadingFrame = textureFrame;
        else if ((frontfacing && dot(outgoingDirWs, attrs.frame.n) > 0) || (!frontfacing && -dot(outgoingDirWs, attrs.frame.n) > 0)) shadingFrame = attrs.frame;
        else shadingFrame = attrs.triangleFrame;

        float3 outgoingDirSs = shadingFrame.worldToFrame(outgoingDirWs);

        if (k.mesh_samples_per_bounce == 0 || bounceCount == 0 || !geometry.sampled || isLastMaterialDelta) {  // :109-124
            if (dot(outgoingDirWs, attrs.triangleFrame.n) > 0.0f) accumulatedColor += throughput * emissiveLight;
        } else if (geometry.sampled) {
            float sum = sc.aliasTable[0].select;
            float lightPdf = areaMeasureToSolidAngleMeasure(attrs.position, ray.Origin, ray.Direction, attrs.triangleFrame.n) / sum;
            if (lightPdf > 0.0f) {
                float weight = powerHeuristic(1, lastMaterialPdf, k.mesh_samples_per_bounce, lightPdf);
                accumulatedColor += throughput * emissiveLight * weight;
            }
        }

        if (bounceCount >= k.max_bounces + 1) {                                     // :128-135
            return accumulatedColor;
        } else if (bounceCount > 3) {
            float pSurvive = std::fmin(0.95f, luminance(throughput));
            if (rng.getFloat() > pSurvive) return accumulatedColor;
            throughput /= pSurvive;
        }

        bool isCurrentMaterialDelta = material.isDelta();
        if (!isCurrentMaterialDelta) {                                              // :139-151
            for (uint32_t i = 0; i < k.env_samples_per_bounce; i++) {
                float2 rand = rng.getFloat2();
                accumulatedColor += throughput * estimateDirectMISLight(sc, tc, shadingFrame, envMap, material, outgoingDirSs, attrs.position, attrs.triangleFrame.n, rand, k.env_samples_per_bounce) / (float)k.env_samples_per_bounce;
            }
            for (uint32_t i = 0; i < k.mesh_samples_per_bounce; i++) {
                float2 rand = rng.getFloat2();
                accumulatedColor += throughput * estimateDirectMISLight(sc, tc, shadingFrame, meshLights, material, outgoingDirSs, attrs.position, attrs.triangleFrame.n, rand, k.mesh_samples_per_bounce) / (float)k.mesh_samples_per_bounce;
            }
        }

        MaterialSample sample = material.sample(outgoingDirSs, rng.getFloat2());   // :154-163
        if (g_debugPath)
            fprintf(stderr, "          position (%.9g %.9g %.9g) n_tri (%g %g %g) n_shade (%g %g %g) w_o (%g %g %g) sample dir (%g %g %g) pdf %g acc (%g %g %g)\n",
                    attrs.position.x, attrs.position.y, attrs.position.z, attrs.triangleFrame.n.x, attrs.triangleFrame.n.y, attrs.triangleFrame.n.z,
                    shadingFrame.n.x, shadingFrame.n.y, shadingFrame.n.z, outgoingDirSs.x, outgoingDirSs.y, outgoingDirSs.z, sample.dirFs.x, sample.dirFs.y, sample.dirFs.z,
                    sample.pdf, accumulatedColor.x, accumulatedColor.y, accumulatedColor.z);
        if (sample.pdf == 0.0f) return accumulatedColor;
        lastMaterialPdf = sample.pdf;
        ray.Direction = shadingFrame.frameToWorld(sample.dirFs);
        ray.Origin = offsetAlongNormal(attrs.position, faceForward(attrs.triangleFrame.n, ray.Direction));
        throughput *= material.eval(sample.dirFs, outgoingDirSs) * std::fabs(Frame::cosTheta(sample.dirFs)) / sample.pdf;
        bounceCount += 1;
        isLastMaterialDelta = isCurrentMaterialDelta;
    }

    if (k.env_samples_per_bounce == 0 || bounceCount == 0 || isLastMaterialDelta) {  // :169-180
        accumulatedColor += throughput * envMap.incomingRadiance(ray.Direction);
    } else {
        LightEval l = envMap.eval(ray.Direction);
        if (l.pdf > 0.0f) {
            float weight = powerHeuristic(1, lastMaterialPdf, k.env_samples_per_bounce, l.pdf);
            accumulatedColor += throughput * l.radiance * weight;
        }
    }
    return accumulatedColor;
}

// ---- main.hlsl --------------------------------------------------------------------------
inline float2 dispatchUV(const Constants& k, uint32_t px, uint32_t py, uint32_t W, uint32_t H, float2 rand) {  // :54-59
    float2 randomCenter = f2(0.5f, 0.5f) + 0.5f * squareToGaussian(rand);
    float2 uv = (f2((float)px, (float)py) + randomCenter) / f2((float)W, (float)H);
    if (k.flip_image) uv.y = 1.0f - uv.y;
    return uv;
}

// raygen (main.hlsl:61-95) + storeColor (:43-51) for one pixel and one launch
inline void raygenPixel(const SceneData& sc, TraceCounters& tc, const Lens& lens, uint32_t W, uint32_t H, uint32_t px, uint32_t py,
                        uint32_t pushSampleCount, float* texel /*RGBA*/) {
    const Constants& k = sc.consts;
    float3 color = f3(0.0f);
    for (uint32_t s = 0; s < k.samples_per_run; s++) {
        Rng rng = Rng::fromSeed(pushSampleCount + s, px, py);
        float2 r0 = rng.getFloat2();
        float2 r1 = rng.getFloat2();
        RayDesc initialRay = generateRay(lens, W, H, dispatchUV(k, px, py, W, H, r0), r1);
        color += incomingRadiance(sc, tc, initialRay, rng);
    }
    if (pushSampleCount == 0) {
        float3 c = color / (float)k.samples_per_run;
        texel[0] = c.x; texel[1] = c.y; texel[2] = c.z; texel[3] = 1.0f;
    } else {
        float3 prior = {texel[0], texel[1], texel[2]};
        float3 d = (color - prior) / (float)(pushSampleCount + k.samples_per_run);
        texel[0] += d.x; texel[1] += d.y; texel[2] += d.z; texel[3] += 1.0f;
    }
}

}  // namespace orc
