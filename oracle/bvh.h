// ORACLE — TEST INFRASTRUCTURE ONLY (see vec.h).
// Software stand-in for the Vulkan acceleration structure + TraceRay of the reference
// (shaders/hrtsystem/intersection.hlsl:18-22,34-47; engine/hrtsystem/Accel.zig:94-184,312-563).
// The reference's ray/triangle arithmetic lives in the Vulkan driver/RT hardware (no source:
// "parity unpinned"). It is therefore DEFINED here, and the CUDA path implements
// the same definition independently:
//
//   * two levels: per-instance object-space rays  o' = mul_point(worldToInstance, o),
//     d' = mul_vec(worldToInstance, d) (not renormalised, so t is shared between spaces);
//   * triangle test = WATERTIGHT signed-volume (Pluecker / scalar-triple-product) edge test in FP32,
//     no FMA. With A=v0-o, B=v1-o, C=v2-o (object space):
//        U = dot(d, cross(C,B));  V = dot(d, cross(A,C));  W = dot(d, cross(B,A));
//        miss if (U<0||V<0||W<0) && (U>0||V>0||W>0);   det = (U+V)+W;  det==0 -> miss;
//        rdet = 1/det;  u = V*rdet;  v = W*rdet;
//        T = (U*dot(A,d) + V*dot(B,d)) + W*dot(C,d);   t = (T*rdet) * (1/dot(d,d));
//        hit iff tmin<t && t<tmax (NaN -> miss).
//     Every edge function is computed from the two translated vertices of that edge only, and
//     cross(X,Y) = -cross(Y,X), dot(d,-x) = -dot(d,x) hold EXACTLY in IEEE arithmetic without
//     contraction, so two triangles sharing an edge always take consistent decisions: no ray can
//     slip between them (the reference's RT hardware is watertight too, and its furnace test
//     engine/tests.zig:340-344 needs it: one leaked path in 5e5 breaks the 1e-5 bound).
//     No face culling (RAY_FLAG_FORCE_OPAQUE only), barycentrics = (u,v) (weights 1-u-v,u,v);
//   * closest hit = smallest t; equal t -> lexicographically smallest
//     (instanceIndex, geometryIndex, primitiveIndex) so that the answer is independent of the
//     BVH shape and traversal order;
//   * invisible instances (mask 0x00, hydra.zig:499-505) are skipped.
//
// The BVH itself (median split, float boxes padded outward) only has to be conservative; a
// brute-force path (`brute=true`) bypasses it so tests can check that it is.
#pragma once
#include <atomic>
#include <cstdio>
#include <algorithm>
#include <limits>
#include "scene.h"

namespace orc {

struct Tri { float3 v0, v1, v2; uint32_t geom, prim; };

struct AABB {
    float3 lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
    void grow(float3 p) {
        lo = {std::fmin(lo.x, p.x), std::fmin(lo.y, p.y), std::fmin(lo.z, p.z)};
        hi = {std::fmax(hi.x, p.x), std::fmax(hi.y, p.y), std::fmax(hi.z, p.z)};
    }
    void grow(const AABB& b) { grow(b.lo); grow(b.hi); }
};

struct BvhNode { AABB box; int32_t left, right; uint32_t first, count; };  // count>0 => leaf

// Generic median-split BVH2 over items with a box and a centroid.
struct Bvh {
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> order;  // item indices in leaf order

    void build(const std::vector<AABB>& boxes, uint32_t leafSize) {
        nodes.clear(); order.resize(boxes.size());
        for (uint32_t i = 0; i < boxes.size(); i++) order[i] = i;
        if (boxes.empty()) return;
        nodes.reserve(2 * boxes.size());
        nodes.push_back({});
        struct Job { int node; uint32_t first, count; };
        std::vector<Job> stack{{0, 0, (uint32_t)boxes.size()}};
        while (!stack.empty()) {
            Job j = stack.back(); stack.pop_back();
            AABB b, cb;
            for (uint32_t i = j.first; i < j.first + j.count; i++) {
                const AABB& ib = boxes[order[i]];
                b.grow(ib);
                cb.grow((ib.lo + ib.hi) * 0.5f);
            }
            // conservative outward padding (relative + absolute)
            float3 ext = b.hi - b.lo;
            float pad = 1e-5f * std::fmax(std::fmax(ext.x, ext.y), std::fmax(ext.z, 1e-3f)) +
                        1e-6f * std::fmax(std::fmax(std::fabs(b.lo.x), std::fabs(b.hi.x)),
                                          std::fmax(std::fmax(std::fabs(b.lo.y), std::fabs(b.hi.y)),
                                                    std::fmax(std::fabs(b.lo.z), std::fabs(b.hi.z))));
            b.lo = b.lo - f3(pad); b.hi = b.hi + f3(pad);
            BvhNode n; n.box = b; n.left = n.right = -1; n.first = j.first; n.count = j.count;
            float3 ce = cb.hi - cb.lo;
            int axis = ce.x > ce.y ? (ce.x > ce.z ? 0 : 2) : (ce.y > ce.z ? 1 : 2);
            float extent = axis == 0 ? ce.x : axis == 1 ? ce.y : ce.z;
            if (j.count > leafSize && extent > 0.0f) {
                uint32_t mid = j.first + j.count / 2;
                auto key = [&](uint32_t it) {
                    const AABB& ib = boxes[it];
                    return axis == 0 ? ib.lo.x + ib.hi.x : axis == 1 ? ib.lo.y + ib.hi.y : ib.lo.z + ib.hi.z;
                };
                std::nth_element(order.begin() + j.first, order.begin() + mid, order.begin() + j.first + j.count,
                                 [&](uint32_t a, uint32_t c) { return key(a) < key(c); });
                n.count = 0;
                n.left = (int)nodes.size(); nodes.push_back({});
                n.right = (int)nodes.size(); nodes.push_back({});
                stack.push_back({n.left, j.first, mid - j.first});
                stack.push_back({n.right, mid, j.first + j.count - mid});
            }
            nodes[j.node] = n;
        }
    }
};

// Conservative slab test evaluated in double precision. The boxes bound the triangles geometrically, but the triangle test
// (tri_test, the definition of a hit) works on vertices translated by o in binary32, i.e. on geometry moved by up to an ulp of
// |v - o|, and reports t with a few ulp of error: every axis' interval is widened by 1e-6 of |o / d| on that axis (the scale of
// those errors in t) and the comparison relatively by 1e-6. The walk then returns exactly what testing every triangle returns
// (`brute`, cross-check mode), which is what "closest hit" means here.
struct RayBoxEps { double e[3]; };
inline RayBoxEps ray_box_eps(const double o[3], const double inv[3]) {
    RayBoxEps r;
    for (int a = 0; a < 3; a++) { const double m = std::fabs(o[a] * inv[a]); r.e[a] = (m == m && m < 1e300) ? m * 1e-6 : 0.0; }
    return r;
}
inline bool ray_box(const AABB& b, const double o[3], const double inv[3], double tmax, const RayBoxEps& eps) {
    double t0 = 0.0, t1 = tmax;
    const float lo[3] = {b.lo.x, b.lo.y, b.lo.z}, hi[3] = {b.hi.x, b.hi.y, b.hi.z};
    for (int a = 0; a < 3; a++) {
        double ta = (lo[a] - o[a]) * inv[a], tb = (hi[a] - o[a]) * inv[a];
        if (ta != ta || tb != tb) continue;  // 0*inf: origin on the slab plane, direction parallel -> no constraint
        if (ta > tb) std::swap(ta, tb);
        t0 = std::max(t0, ta - eps.e[a]); t1 = std::min(t1, tb + eps.e[a]);
    }
    return t0 <= t1 * 1.000001 + 1e-30;
}

struct Blas {
    std::vector<Tri> tris;  // in (geometry, primitive) order
    Bvh bvh;
};

// Ray/triangle test = the definition of a hit. The triangle is translated to the ray origin (A, B, C = v - o,
// rounded to binary32) and sheared onto the ray's dominant axis, which leaves three 2-D edge functions of the projected vertices
// (Woop, Benthin, Wald 2013). Their signs are EXACT: each is certified in binary32 (|value| above the rounding bound of its two
// products) or decided in binary64, where both products are exact and their difference, rounded once, has the exact sign. So two
// triangles sharing an edge decide consistently (no ray slips between them) and a triangle is accepted only if the ray passes
// through its projection: walking any conservative BVH returns what testing every triangle returns (`brute`, cross-check mode).
struct RayShear { uint32_t kx, ky, kz; float Sx, Sy, Sz; };
inline RayShear ray_shear(float3 d) {
    const float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    RayShear r;
    r.kz = ax > ay ? (ax > az ? 0u : 2u) : (ay > az ? 1u : 2u);
    r.kx = r.kz == 2u ? 0u : r.kz + 1u; r.ky = r.kx == 2u ? 0u : r.kx + 1u;
    const float dv[3] = {d.x, d.y, d.z};
    if (dv[r.kz] < 0.0f) std::swap(r.kx, r.ky);
    r.Sz = 1.0f / dv[r.kz]; r.Sx = dv[r.kx] * r.Sz; r.Sy = dv[r.ky] * r.Sz;
    return r;
}
inline bool tri_test(const Tri& tr, float3 o, const RayShear& r, float tmin, float tmax, float& t, float& u, float& v) {
    const float3 A3 = tr.v0 - o, B3 = tr.v1 - o, C3 = tr.v2 - o;
    const float A[3] = {A3.x, A3.y, A3.z}, B[3] = {B3.x, B3.y, B3.z}, C[3] = {C3.x, C3.y, C3.z};
    const float Ax = A[r.kx] - r.Sx * A[r.kz], Ay = A[r.ky] - r.Sy * A[r.kz];
    const float Bx = B[r.kx] - r.Sx * B[r.kz], By = B[r.ky] - r.Sy * B[r.kz];
    const float Cx = C[r.kx] - r.Sx * C[r.kz], Cy = C[r.ky] - r.Sy * C[r.kz];
    const float u0 = Cx * By, u1 = Cy * Bx, v0 = Ax * Cy, v1 = Ay * Cx, w0 = Bx * Ay, w1 = By * Ax;
    const float U = u0 - u1, V = v0 - v1, W = w0 - w1;
    const float Az = r.Sz * A[r.kz], Bz = r.Sz * B[r.kz], Cz = r.Sz * C[r.kz];
    const float kBound = 2.3841858e-07f;   // 2^-22
    if (std::fabs(U) > kBound * (std::fabs(u0) + std::fabs(u1)) && std::fabs(V) > kBound * (std::fabs(v0) + std::fabs(v1)) &&
        std::fabs(W) > kBound * (std::fabs(w0) + std::fabs(w1))) {
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
        const float det = (U + V) + W;
        const float rdet = 1.0f / det;
        u = V * rdet;
        v = W * rdet;
        const float T = (U * Az + V * Bz) + W * Cz;
        t = T * rdet;
        return tmin < t && t < tmax;
    }
    const double Ud = (double)Cx * (double)By - (double)Cy * (double)Bx;
    const double Vd = (double)Ax * (double)Cy - (double)Ay * (double)Cx;
    const double Wd = (double)Bx * (double)Ay - (double)By * (double)Ax;
    if ((Ud < 0.0 || Vd < 0.0 || Wd < 0.0) && (Ud > 0.0 || Vd > 0.0 || Wd > 0.0)) return false;
    const double det = (Ud + Vd) + Wd;
    if (det == 0.0) return false;
    const double T = (Ud * (double)Az + Vd * (double)Bz) + Wd * (double)Cz;
    t = (float)(T / det);
    u = (float)(Vd / det);
    v = (float)(Wd / det);
    return tmin < t && t < tmax;
}

struct Accel {
    std::vector<Blas> blases;
    std::vector<uint32_t> instanceBlas;  // per instance
    Bvh tlas;
    const std::vector<InstanceRec>* instances = nullptr;
    std::vector<uint32_t> tlasItems;  // visible instance ids (tlas items index into this)

    static AABB transformed_box(const AABB& b, const Mat3x4& M) {
        AABB r;
        for (int c = 0; c < 8; c++) {
            float3 p = {(c & 1) ? b.hi.x : b.lo.x, (c & 2) ? b.hi.y : b.lo.y, (c & 4) ? b.hi.z : b.lo.z};
            r.grow(mul_point(M, p));
        }
        float3 ext = r.hi - r.lo;
        float pad = 1e-5f * std::fmax(std::fmax(ext.x, ext.y), std::fmax(ext.z, 1e-3f));
        r.lo = r.lo - f3(pad); r.hi = r.hi + f3(pad);
        return r;
    }

    void buildTlas() {
        std::vector<AABB> boxes;
        tlasItems.clear();
        for (uint32_t i = 0; i < instances->size(); i++) {
            const InstanceRec& in = (*instances)[i];
            if (!in.visible) continue;
            const Blas& b = blases[instanceBlas[i]];
            if (b.bvh.nodes.empty()) continue;
            tlasItems.push_back(i);
            boxes.push_back(transformed_box(b.bvh.nodes[0].box, in.transform));
        }
        tlas.build(boxes, 1);
    }

    // diagnostic: every query is answered by the walk AND by testing every triangle; disagreements are counted and the first few printed
    mutable std::atomic<uint64_t> crossCheckMismatches{0};
    bool crossCheck = false;
    Intersection trace(float3 o, float3 d, float tmin, float tmax, bool anyHit, bool brute) const {
        if (!crossCheck) return traceImpl(o, d, tmin, tmax, anyHit, brute);
        const Intersection a = traceImpl(o, d, tmin, tmax, anyHit, false), b = traceImpl(o, d, tmin, tmax, anyHit, true);
        const bool same = anyHit ? (a.instanceIndex != MAX_UINT) == (b.instanceIndex != MAX_UINT)
                                 : (a.instanceIndex == b.instanceIndex && (a.instanceIndex == MAX_UINT || (a.geometryIndex == b.geometryIndex && a.primitiveIndex == b.primitiveIndex)));
        if (!same && crossCheckMismatches.fetch_add(1) < 8)
            fprintf(stderr, "oracle cross-check: %s ray o=(%.9g %.9g %.9g) d=(%.9g %.9g %.9g) tmax=%.9g  walk -> inst %u geom %u prim %u | all triangles -> inst %u geom %u prim %u\n",
                    anyHit ? "any-hit" : "closest", o.x, o.y, o.z, d.x, d.y, d.z, tmax, a.instanceIndex, a.geometryIndex, a.primitiveIndex, b.instanceIndex, b.geometryIndex, b.primitiveIndex);
        return b;
    }
    // closest hit when anyHit==false; returns true/false occlusion in its.instanceIndex!=MAX when anyHit
    Intersection traceImpl(float3 o, float3 d, float tmin, float tmax, bool anyHit, bool brute) const {
        Intersection best; best.instanceIndex = MAX_UINT; best.geometryIndex = best.primitiveIndex = 0; best.barycentrics = {0, 0};
        float bestT = tmax;
        auto consider = [&](uint32_t inst, const Tri& tr, float3 oo, const RayShear& shear) -> bool {
            float t, u, v;
            // closest: accept t <= bestT so ties can be resolved by id; any: plain interval test
            if (!tri_test(tr, oo, shear, tmin, tmax, t, u, v)) return false;
            if (anyHit) { best.instanceIndex = inst; return true; }
            bool better = t < bestT;
            if (!better && t == bestT && best.instanceIndex != MAX_UINT) {
                if (inst != best.instanceIndex) better = inst < best.instanceIndex;
                else if (tr.geom != best.geometryIndex) better = tr.geom < best.geometryIndex;
                else better = tr.prim < best.primitiveIndex;
            }
            if (better) { bestT = t; best = {inst, tr.geom, tr.prim, {u, v}}; }
            return false;
        };
        auto traceInstance = [&](uint32_t inst) -> bool {
            const InstanceRec& in = (*instances)[inst];
            const Blas& b = blases[instanceBlas[inst]];
            float3 oo = mul_point(in.worldToInstance, o), od = mul_vec(in.worldToInstance, d);
            const RayShear shear = ray_shear(od);      // once per object-space ray
            if (brute) {
                for (const Tri& tr : b.tris) if (consider(inst, tr, oo, shear)) return true;
                return false;
            }
            const double O[3] = {oo.x, oo.y, oo.z};
            const double I[3] = {1.0 / (double)od.x, 1.0 / (double)od.y, 1.0 / (double)od.z};
            const RayBoxEps eps = ray_box_eps(O, I);
            int stack[128]; int sp = 0; stack[sp++] = 0;
            while (sp) {
                const BvhNode& n = b.bvh.nodes[stack[--sp]];
                if (!ray_box(n.box, O, I, (double)bestT, eps)) continue;
                if (n.count) {
                    for (uint32_t i = n.first; i < n.first + n.count; i++)
                        if (consider(inst, b.tris[b.bvh.order[i]], oo, shear)) return true;
                } else { stack[sp++] = n.left; stack[sp++] = n.right; }
            }
            return false;
        };
        if (brute) {
            for (uint32_t i = 0; i < instances->size(); i++)
                if ((*instances)[i].visible && traceInstance(i)) return best;
            return best;
        }
        if (tlas.nodes.empty()) return best;
        const double O[3] = {o.x, o.y, o.z};
        const double I[3] = {1.0 / (double)d.x, 1.0 / (double)d.y, 1.0 / (double)d.z};
        const RayBoxEps eps = ray_box_eps(O, I);
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const BvhNode& n = tlas.nodes[stack[--sp]];
            if (!ray_box(n.box, O, I, (double)bestT, eps)) continue;
            if (n.count) {
                for (uint32_t i = n.first; i < n.first + n.count; i++)
                    if (traceInstance(tlasItems[tlas.order[i]])) return best;
            } else { stack[sp++] = n.left; stack[sp++] = n.right; }
        }
        return best;
    }
};

}  // namespace orc
