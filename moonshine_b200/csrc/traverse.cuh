// Software replacement of TraceRay (shaders/hrtsystem/intersection.hlsl:18-22 closest, :34-47 any-hit)
// for a GPU without RT cores: two-level BVH2 traversal + watertight ray/triangle test.
//
// The ray/triangle and instancing arithmetic is the project's definition of the (driver-opaque)
// reference behaviour and is stated in DESIGN.md §"Intersection"; the CPU oracle implements the same
// definition independently (oracle/bvh.h):
//   object-space ray   o' = mulPoint(w2o, o), d' = mulVec(w2o, d)   (t shared between spaces)
//   A=v0-o, B=v1-o, C=v2-o;  U=dot(d,cross(C,B)), V=dot(d,cross(A,C)), W=dot(d,cross(B,A))
//   miss if mixed signs; det=(U+V)+W, det==0 -> miss; rdet=1/det; u=V*rdet; v=W*rdet
//   T=(U*dot(A,d)+V*dot(B,d))+W*dot(C,d);  t=(T*rdet)*(1/dot(d,d));  hit iff tmin<t<tmax
//   closest = min t, ties -> smallest (instance, geometry, primitive); no culling; mask 0 skipped.
// The slab test is only required to be conservative (boxes are padded at build time, bvh_build.h).
#pragma once
#include "device_math.cuh"

namespace msne {

#ifndef MSNE_STACK
#define MSNE_STACK 64
#endif

struct HitRec { uint32_t inst, tri; float u, v; };

struct SpaceRay {
    float3 o, d, idir, ood;
    float rdd;
    DI void set(float3 o_, float3 d_) {
        o = o_; d = d_;
        idir.x = fabsf(d.x) > 1e-20f ? 1.0f / d.x : copysignf(1e20f, d.x);
        idir.y = fabsf(d.y) > 1e-20f ? 1.0f / d.y : copysignf(1e20f, d.y);
        idir.z = fabsf(d.z) > 1e-20f ? 1.0f / d.z : copysignf(1e20f, d.z);
        ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        rdd = 1.0f / dot(d, d);
    }
};

// returns true when the child's box overlaps [tmin, tmax] along the ray (conservative)
DI bool slab(const SpaceRay& r, float lox, float hix, float loy, float hiy, float loz, float hiz, float tmin, float tmax, float& tnear) {
    const float tx0 = fmaf(lox, r.idir.x, -r.ood.x), tx1 = fmaf(hix, r.idir.x, -r.ood.x);
    const float ty0 = fmaf(loy, r.idir.y, -r.ood.y), ty1 = fmaf(hiy, r.idir.y, -r.ood.y);
    const float tz0 = fmaf(loz, r.idir.z, -r.ood.z), tz1 = fmaf(hiz, r.idir.z, -r.ood.z);
    tnear = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
    const float tfar = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
    return tnear <= tfar * 1.0000004f;
}

DI bool triTest(const float4 t0, const float4 t1, const float4 t2, const SpaceRay& r, float tmin, float tmax, float& t, float& u, float& v) {
    const float3 A = mk3(t0.x, t0.y, t0.z) - r.o, B = mk3(t1.x, t1.y, t1.z) - r.o, C = mk3(t2.x, t2.y, t2.z) - r.o;
    const float U = dot(r.d, cross(C, B));
    const float V = dot(r.d, cross(A, C));
    const float W = dot(r.d, cross(B, A));
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = (U + V) + W;
    if (det == 0.0f) return false;
    const float rdet = 1.0f / det;
    u = V * rdet;
    v = W * rdet;
    const float T = (U * dot(A, r.d) + V * dot(B, r.d)) + W * dot(C, r.d);
    t = (T * rdet) * r.rdd;
    return tmin < t && t < tmax;
}

// stack entry encoding: >= 0 inner node (of the level currently traversed); -1 end; -2 leave instance;
// <= -3 TLAS item slot (-e-3)
template <bool ANY_HIT>
DI bool traceRay(const DeviceScene& sc, float3 o, float3 d, float tmin, float tmax, HitRec& hit) {
    hit.inst = MISS_ID; hit.tri = 0; hit.u = 0.0f; hit.v = 0.0f;
    if (sc.tlasNodes == nullptr) return false;
    float bestT = tmax;
    int stack[MSNE_STACK];
    int sp = 0;
    stack[sp++] = -1;
    SpaceRay world; world.set(o, d);
    SpaceRay ray = world;
    const BvhNode2* nodes = sc.tlasNodes;
    bool inBlas = false;
    uint32_t curInst = 0;
    int node = 0;
    for (;;) {
        while (node >= 0) {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            const int4 c = __ldg(reinterpret_cast<const int4*>(np + 3));
            float tn0, tn1;
            bool h0 = slab(ray, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin, bestT, tn0) && (c.x != -1 || c.z > 0);
            bool h1 = slab(ray, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin, bestT, tn1) && (c.y != -1 || c.w > 0);
            int next = -1, deferred = -1;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const bool h = s == 0 ? h0 : h1;
                const int child = s == 0 ? c.x : c.y;
                const int count = s == 0 ? c.z : c.w;
                if (!h) continue;
                if (count > 0) {
                    if (inBlas) {
                        for (int i = 0; i < count; i++) {
                            const float4* tp = reinterpret_cast<const float4*>(sc.tris + (child + i));
                            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                            float t, u, v;
                            if (!triTest(t0, t1, t2, ray, tmin, tmax, t, u, v)) continue;
                            if (ANY_HIT) { hit.inst = curInst; return true; }
                            bool better = t < bestT;
                            if (!better && t == bestT && hit.inst != MISS_ID) {
                                const uint32_t tri = (uint32_t)(child + i);
                                if (curInst != hit.inst) better = curInst < hit.inst;
                                else {
                                    const uint32_t g = __float_as_uint(t0.w), p = __float_as_uint(t1.w);
                                    const uint32_t bg = sc.tris[hit.tri].geom, bp = sc.tris[hit.tri].prim;
                                    better = g != bg ? g < bg : p < bp;
                                }
                                (void)tri;
                            }
                            if (better) { bestT = t; hit.inst = curInst; hit.tri = (uint32_t)(child + i); hit.u = u; hit.v = v; }
                        }
                    } else {
                        for (int i = 0; i < count; i++) if (sp < MSNE_STACK) stack[sp++] = -(child + i) - 3;
                    }
                } else {
                    if (next < 0) next = child; else deferred = child;
                }
            }
            if (deferred >= 0) {
                // both children are inner nodes: visit the nearer first
                if (tn1 < tn0) { const int tmp = next; next = deferred; deferred = tmp; }
                if (sp < MSNE_STACK) stack[sp++] = deferred;
            }
            node = next >= 0 ? next : -1;
            if (node < 0) break;
        }
        // pop
        for (;;) {
            const int e = stack[--sp];
            if (e == -1) return hit.inst != MISS_ID;
            if (e == -2) { inBlas = false; ray = world; nodes = sc.tlasNodes; continue; }
            if (e <= -3) {
                const uint32_t inst = __ldg(sc.tlasItems + (-e - 3));
                const DInstance* in = sc.instances + inst;
                const float4* mp = reinterpret_cast<const float4*>(in->w2o);
                const float4 r0 = __ldg(mp), r1 = __ldg(mp + 1), r2 = __ldg(mp + 2);
                const float3 oo = mk3(((r0.x * o.x + r0.y * o.y) + r0.z * o.z) + r0.w, ((r1.x * o.x + r1.y * o.y) + r1.z * o.z) + r1.w,
                                      ((r2.x * o.x + r2.y * o.y) + r2.z * o.z) + r2.w);
                const float3 od = mk3((r0.x * d.x + r0.y * d.y) + r0.z * d.z, (r1.x * d.x + r1.y * d.y) + r1.z * d.z, (r2.x * d.x + r2.y * d.y) + r2.z * d.z);
                ray.set(oo, od);
                curInst = inst;
                inBlas = true;
                nodes = sc.blasNodes;
                stack[sp++] = -2;
                node = (int)__ldg(&in->blasRoot);
                break;
            }
            node = e;
            break;
        }
    }
}

}  // namespace msne
