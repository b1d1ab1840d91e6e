// ORACLE — TEST INFRASTRUCTURE ONLY (see vec.h).
// Scene tables of the reference (shaders/hrtsystem/world.hlsl:6-72) held
// as plain host arrays, plus a software model of the two Vulkan samplers the reference creates:
//   material textures: linear mag/min, REPEAT, LOD 0   (engine/hrtsystem/MaterialManager.zig:425-444)
//   background:        linear mag/min, MIRRORED_REPEAT (engine/hrtsystem/BackgroundManager.zig:78-95)
// Bilinear filtering follows the Vulkan texel-coordinate rules (unnormalised u = s*W,
// i0 = floor(u - 0.5), alpha = frac(u - 0.5)) in exact FP32 — the hardware's reduced-precision
// filter weights and the driver's arithmetic order are "parity unpinned" (SURVEY.md §8c), so the
// filter is *defined* as:  lerp(lerp(t00,t10,a), lerp(t01,t11,a), b),  lerp(x,y,t)=x+t*(y-x).
#pragma once
#include <vector>
#include "utils.h"

namespace orc {

enum TexFormat : uint32_t {
    TEX_R32F = 0, TEX_RG32F = 1, TEX_RGBA32F = 2,      // constant/solid textures (MaterialManager.zig:364-388)
    TEX_R8_UNORM = 3, TEX_RG8_UNORM = 4,               // World.zig:72,184
    TEX_RGBA8_SRGB = 5,                                // World.zig:101; hydra.zig:47-52
    TEX_RGBA16F = 6,                                   // hydra.zig:47-52
};

struct Texture {
    uint32_t w = 0, h = 0;
    std::vector<float4> texels;  // decoded to linear FP32 at upload (sRGB decode precedes filtering)
};

inline float srgb_to_linear(uint8_t c) {
    // IEC 61966-2-1 decode evaluated in double then rounded once to binary32
    double s = c / 255.0;
    double l = s <= 0.04045 ? s / 12.92 : std::pow((s + 0.055) / 1.055, 2.4);
    return (float)l;
}
inline float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 31, man = h & 1023;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { int e = -1; do { man <<= 1; e++; } while (!(man & 1024)); bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 1023) << 13); }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f; std::memcpy(&f, &bits, 4); return f;
}

inline Texture decode_texture(const void* data, uint32_t w, uint32_t h, uint32_t fmt) {
    Texture t; t.w = w; t.h = h; t.texels.resize((size_t)w * h);
    const float* f = (const float*)data; const uint8_t* b = (const uint8_t*)data; const uint16_t* hf = (const uint16_t*)data;
    for (size_t i = 0; i < (size_t)w * h; i++) {
        float4 v = {0, 0, 0, 1};
        switch (fmt) {
            case TEX_R32F: v.x = f[i]; break;
            case TEX_RG32F: v.x = f[2 * i]; v.y = f[2 * i + 1]; break;
            case TEX_RGBA32F: v = {f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]}; break;
            case TEX_R8_UNORM: v.x = b[i] / 255.0f; break;
            case TEX_RG8_UNORM: v.x = b[2 * i] / 255.0f; v.y = b[2 * i + 1] / 255.0f; break;
            case TEX_RGBA8_SRGB: v = {srgb_to_linear(b[4 * i]), srgb_to_linear(b[4 * i + 1]), srgb_to_linear(b[4 * i + 2]), b[4 * i + 3] / 255.0f}; break;
            case TEX_RGBA16F: v = {half_to_float(hf[4 * i]), half_to_float(hf[4 * i + 1]), half_to_float(hf[4 * i + 2]), half_to_float(hf[4 * i + 3])}; break;
        }
        t.texels[i] = v;
    }
    return t;
}

inline int wrap_repeat(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
inline int wrap_mirror(int i, int n) {
    int m = i % (2 * n); if (m < 0) m += 2 * n;
    int k = m - n;
    int mir = k >= 0 ? k : -(1 + k);
    return (n - 1) - mir;
}
inline float4 lerp4(float4 a, float4 b, float t) {
    return {a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z), a.w + t * (b.w - a.w)};
}
// SampleLevel(sampler, uv, 0)
inline float4 sample_bilinear(const Texture& t, float2 uv, bool mirror) {
    if (t.w == 1 && t.h == 1) return t.texels[0];
    float u = uv.x * (float)t.w - 0.5f, v = uv.y * (float)t.h - 0.5f;
    float fu = std::floor(u), fv = std::floor(v);
    float a = u - fu, b = v - fv;
    int i0 = (int)fu, j0 = (int)fv;
    int i1 = i0 + 1, j1 = j0 + 1;
    if (mirror) { i0 = wrap_mirror(i0, t.w); i1 = wrap_mirror(i1, t.w); j0 = wrap_mirror(j0, t.h); j1 = wrap_mirror(j1, t.h); }
    else        { i0 = wrap_repeat(i0, t.w); i1 = wrap_repeat(i1, t.w); j0 = wrap_repeat(j0, t.h); j1 = wrap_repeat(j1, t.h); }
    const float4 t00 = t.texels[(size_t)j0 * t.w + i0], t10 = t.texels[(size_t)j0 * t.w + i1];
    const float4 t01 = t.texels[(size_t)j1 * t.w + i0], t11 = t.texels[(size_t)j1 * t.w + i1];
    return lerp4(lerp4(t00, t10, a), lerp4(t01, t11, a), b);
}

// world.hlsl:31-47; MaterialManager.zig:45-77
enum MaterialType : uint32_t { MAT_GLASS = 0, MAT_LAMBERT = 1, MAT_PERFECT_MIRROR = 2, MAT_STANDARD_PBR = 3 };
struct MaterialRec {
    uint32_t normal, emissive;  // texture handles
    uint32_t type;
    uint32_t color, metalness, roughness;  // StandardPBR (material.hlsl:186-190); Lambert uses color
    float ior;                              // StandardPBR / Glass
};

struct MeshRec {  // world.hlsl:23-29 (addresses -> host arrays)
    std::vector<float3> positions;
    std::vector<float3> normals;    // empty = none
    std::vector<float2> texcoords;  // empty = none
    std::vector<uint3> indices;
};

struct GeometryRec { uint32_t mesh, material; bool sampled; };  // world.hlsl:17-21; Accel.zig:40-44

struct InstanceRec {  // world.hlsl:6-15 (the fields the shaders read) + visibility mask
    Mat3x4 transform;
    Mat3x4 worldToInstance;          // vector.zig:350-362 via Accel.zig:430-432
    uint32_t customIndex;            // offset into flat geometries[] (Accel.zig:395-411)
    uint32_t geometryCount;
    bool visible;
};

struct AliasEntry { uint32_t alias; float select; uint32_t instance, geometry, primitive; };  // light.hlsl:17-22,105-109

struct EnvMapData {  // BackgroundManager.zig:154-163
    uint32_t size = 0;
    std::vector<float4> rgb;                  // size*size RGBA32F equal-area image
    std::vector<std::vector<float>> lum;      // lum[level], level 0 = size*size ... last = 1x1 (sum mips)
};

struct Constants {  // main.hlsl:34-40; pipeline.zig:319-327
    uint32_t samples_per_run = 1, max_bounces = 4, env_samples_per_bounce = 1, mesh_samples_per_bounce = 1;
    bool flip_image = true, indexed_attributes = true, two_component_normal_texture = true;
};

struct Lens { float3 origin, forward, up; float vfov, aperture, focus_distance; };  // camera.hlsl:6-12

struct Intersection {  // intersection.hlsl:5-9
    uint32_t instanceIndex, geometryIndex, primitiveIndex;
    float2 barycentrics;
    bool hit() const { return instanceIndex != MAX_UINT; }
};

}  // namespace orc
