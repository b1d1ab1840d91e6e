// ORACLE — TEST INFRASTRUCTURE ONLY (see vec.h).
// CPU restatement of the host-side pieces the hot path's data depends on:
//   engine/hrtsystem/alias_table.zig:12-92      (Vose alias table, FP32)
//   engine/vector.zig:350-362,501-521           (Mat3x4.inverse_affine via Mat3.inverse)
//   engine/hrtsystem/Accel.zig:491-538          (emitter table: weights, entry order, header)
//   engine/hrtsystem/BackgroundManager.zig:142-394 + shaders/background/*.hlsl (env pre-process)
// All paths relative to /root/reference/.
#pragma once
#include "shading.h"

namespace orc {

// alias_table.zig:28-91. `data` carried through untouched.
inline float buildAliasTable(const std::vector<float>& w, std::vector<AliasEntry>& entries /* size n, data prefilled */) {
    const uint32_t NONE = 0xFFFFFFFFu;
    const uint32_t n = (uint32_t)w.size();
    float weight_sum = 0.0f;
    for (float x : w) weight_sum += x;
    uint32_t less_head = NONE, more_head = NONE;
    for (uint32_t i = 0; i < n; i++) {
        const float adjusted = (w[i] * (float)n) / weight_sum;
        entries[i].select = adjusted;
        if (adjusted < 1.0f) { entries[i].alias = less_head; less_head = i; }
        else { entries[i].alias = more_head; more_head = i; }
    }
    while (less_head != NONE && more_head != NONE) {
        const uint32_t less = less_head; less_head = entries[less].alias;
        const uint32_t more = more_head; more_head = entries[more].alias;
        entries[less].alias = more;
        entries[more].select = (entries[more].select + entries[less].select) - 1.0f;
        if (entries[more].select < 1.0f) { entries[more].alias = less_head; less_head = more; }
        else { entries[more].alias = more_head; more_head = more; }
    }
    while (less_head != NONE) {  // "should only happen due to floating point"
        const uint32_t less = less_head; less_head = entries[less].alias;
        entries[less].select = 1.0f;
    }
    return weight_sum;
}

// vector.zig:350-362 (inverse_affine) with Mat3.inverse (:514-521), determinant (:501-503),
// Vec3.cross (:127-132: x = a.y*b.z - b.y*a.z ...), Vec3.dot (:123-125).
inline Mat3x4 inverseAffine(const Mat3x4& M) {
    auto zcross = [](float3 a, float3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); };
    auto zdot = [](float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
    // p = Mat3(rows of A).transpose(): p.x/y/z are the COLUMNS of the linear part A
    float3 px = {M.m[0][0], M.m[1][0], M.m[2][0]}, py = {M.m[0][1], M.m[1][1], M.m[2][1]}, pz = {M.m[0][2], M.m[1][2], M.m[2][2]};
    float3 v = {M.m[0][3], M.m[1][3], M.m[2][3]};
    float det = zdot(px, zcross(py, pz));
    float3 v1 = zcross(py, pz) * (1 / det), v2 = zcross(pz, px) * (1 / det), v3 = zcross(px, py) * (1 / det);
    // Self.new(v1,v2,v3).transpose(): inv.x = (v1.x, v2.x, v3.x), ...
    float3 ix = {v1.x, v2.x, v3.x}, iy = {v1.y, v2.y, v3.y}, iz = {v1.z, v2.z, v3.z};
    // inv_p.mul_scalar(-1).mul_vec(v): res = x*v.x; res = y*v.y + res; res = z*v.z + res  (:487-492)
    float3 nx = ix * -1.0f, ny = iy * -1.0f, nz = iz * -1.0f;
    float3 res = nx * v.x; res = ny * v.y + res; res = nz * v.z + res;
    Mat3x4 R;
    R.m[0][0] = ix.x; R.m[0][1] = iy.x; R.m[0][2] = iz.x; R.m[0][3] = res.x;
    R.m[1][0] = ix.y; R.m[1][1] = iy.y; R.m[1][2] = iz.y; R.m[1][3] = res.y;
    R.m[2][0] = ix.z; R.m[2][1] = iy.z; R.m[2][2] = iz.z; R.m[2][3] = res.z;
    return R;
}

// Accel.zig:491-538: one entry per triangle of every `sampled` geometry, in (instance, geometry,
// primitive) order; weight = world-space area |(p1-p0) x (p2-p0)| / 2; header entry 0.
inline void buildEmitterTable(SceneData& sc) {
    std::vector<float> weights;
    std::vector<AliasEntry> entries;
    auto zcross = [](float3 a, float3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); };
    for (uint32_t i = 0; i < sc.instances.size(); i++) {
        const InstanceRec& in = sc.instances[i];
        for (uint32_t j = 0; j < in.geometryCount; j++) {
            const GeometryRec& g = sc.geometries[in.customIndex + j];
            if (!g.sampled) continue;
            const MeshRec& mesh = sc.meshes[g.mesh];
            for (uint32_t k = 0; k < mesh.indices.size(); k++) {
                uint3 idx = mesh.indices[k];
                // Mat3x4.mul_point (vector.zig:283-288): row.dot(v.extend(1)) = x*x' + y*y' + z*z' + w*1
                auto mp = [&](float3 p) {
                    return f3(in.transform.m[0][0] * p.x + in.transform.m[0][1] * p.y + in.transform.m[0][2] * p.z + in.transform.m[0][3] * 1.0f,
                              in.transform.m[1][0] * p.x + in.transform.m[1][1] * p.y + in.transform.m[1][2] * p.z + in.transform.m[1][3] * 1.0f,
                              in.transform.m[2][0] * p.x + in.transform.m[2][1] * p.y + in.transform.m[2][2] * p.z + in.transform.m[2][3] * 1.0f);
                };
                float3 p0 = mp(mesh.positions[idx.x]), p1 = mp(mesh.positions[idx.y]), p2 = mp(mesh.positions[idx.z]);
                float3 c = zcross(p1 - p0, p2 - p0);
                float area = std::sqrt(c.x * c.x + c.y * c.y + c.z * c.z) / 2.0f;
                weights.push_back(area);
                entries.push_back({0, 0.0f, i, j, k});
            }
        }
    }
    float sum = buildAliasTable(weights, entries);
    sc.aliasTable.clear();
    sc.aliasTable.push_back({(uint32_t)entries.size(), sum, 0, 0, 0});
    sc.aliasTable.insert(sc.aliasTable.end(), entries.begin(), entries.end());
}

// BackgroundManager.zig:154 + equirectangular_to_equal_area.hlsl:9-29 + luminance.hlsl:7-14 + fold.hlsl:6-16
inline void buildEnvMap(EnvMapData& env, const float* rgba, uint32_t W, uint32_t H) {
    Texture src; src.w = W; src.h = H; src.texels.resize((size_t)W * H);
    for (size_t i = 0; i < (size_t)W * H; i++) src.texels[i] = {rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]};
    uint32_t p2 = 1; while (p2 * 2 <= H) p2 *= 2;       // floorPowerOfTwo(height)
    const uint32_t S = std::min(p2, 1024u);              // maximum_equal_area_map_size
    env.size = S;
    env.rgb.assign((size_t)S * S, {0, 0, 0, 1});
    for (uint32_t y = 0; y < S; y++)
        for (uint32_t x = 0; x < S; x++) {
            float3 color = f3(0.0f);
            for (uint32_t i = 0; i < 3; i++)
                for (uint32_t j = 0; j < 3; j++) {
                    const float2 subpixel = f2((float)(1 + i), (float)(1 + j)) / f2(4.0f, 4.0f);
                    const float2 dst = (f2((float)x, (float)y) + subpixel) / f2((float)S, (float)S);
                    const float3 dir = squareToEqualAreaSphere(dst);
                    const float2 sc2 = cartesianToSpherical(dir) / f2(2 * PI, PI);
                    float4 t = sample_bilinear(src, sc2, true);  // background sampler: mirrored repeat
                    color += f3(t.x, t.y, t.z);
                }
            float3 c = color / 9.0f;
            env.rgb[(size_t)y * S + x] = {c.x, c.y, c.z, 1.0f};
        }
    uint32_t levels = EnvMap::ilog2(S) + 1;
    env.lum.assign(levels, {});
    env.lum[0].resize((size_t)S * S);
    for (size_t i = 0; i < (size_t)S * S; i++) env.lum[0][i] = luminance({env.rgb[i].x, env.rgb[i].y, env.rgb[i].z});
    for (uint32_t l = 1; l < levels; l++) {
        uint32_t sz = S >> l, ps = S >> (l - 1);
        env.lum[l].resize((size_t)sz * sz);
        const std::vector<float>& p = env.lum[l - 1];
        for (uint32_t y = 0; y < sz; y++)
            for (uint32_t x = 0; x < sz; x++)
                env.lum[l][(size_t)y * sz + x] = p[(size_t)(2 * y) * ps + 2 * x] + p[(size_t)(2 * y) * ps + 2 * x + 1] +
                                                 p[(size_t)(2 * y + 1) * ps + 2 * x] + p[(size_t)(2 * y + 1) * ps + 2 * x + 1];
    }
}

}  // namespace orc
