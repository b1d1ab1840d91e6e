// Software replacement of TraceRay (shaders/hrtsystem/intersection.hlsl:18-22 closest, :34-47 any-hit)
// for a GPU without RT cores: two-level BVH2 traversal + watertight ray/triangle test.
//
// The ray/triangle and instancing arithmetic is the project's definition of the (driver-opaque)
// reference behaviour and is stated in DESIGN.md §"Intersection"; the CPU oracle implements the same
// definition independently (oracle/bvh.h):
//   object-space ray   o' = mulPoint(w2o, o), d' = mulVec(w2o, d)   (t shared between spaces)
//   A=v0-o, B=v1-o, C=v2-o;  U=dot(d,cross(C,B)), V=dot(d,cross(A,C)), W=dot(d,cross(B,A)), each with a rounding-error bound;
//   any |value| <= its bound -> all three (and u, v, t) are recomputed in binary64 (exact products): signs are the EXACT signs
//   miss if mixed signs; det=(U+V)+W, det==0 -> miss; rdet=1/det; u=V*rdet; v=W*rdet
//   T=(U*dot(A,d)+V*dot(B,d))+W*dot(C,d);  t=(T*rdet)*(1/dot(d,d));  hit iff tmin<t<tmax
//   closest = min t, ties -> smallest (instance, geometry, primitive); no culling; mask 0 skipped.
// The slab test is only required to be conservative (boxes are padded at build time, bvh_build.h).
//
// Loop structure: classic while-while. A lane walks inner nodes until it reaches a leaf, tests the leaf's triangles at
// once (so the next slab tests already see the shorter ray), pops, and goes back to walking. Postponing leaves (park one
// or two and keep descending) was measured: parking one costs 2 %, parking two 12 % (profiles/r01_sweeps.md) - the extra
// speculative node visits outweigh the fuller triangle loop.
// All visible identity-transform instances live in ONE world-space BLAS (runtime.cpp buildGeometry): an identity
// transform leaves the ray bit-identical, so flattening changes no result, and their triangles carry the
// instance index. TWO_LEVEL = false is the specialisation for scenes whose TLAS holds nothing else
// (DeviceScene::staticOnly): the walk starts at that BLAS root and the instance entry / leave code, the registers that
// keep the world-space ray, and the stack-overflow test (bvh_build.cpp bounds the depth) go away.
#pragma once
#include "device_math.cuh"
#include "kernels.h"

namespace msne {

struct HitRec { uint32_t inst, tri; float u, v; };

struct SpaceRay {
    float3 o, d, idir, ood;
    float rdd, eps;
    DI void set(float3 o_, float3 d_) {
        o = o_; d = d_;
        idir.x = fabsf(d.x) > 1e-20f ? 1.0f / d.x : copysignf(1e20f, d.x);
        idir.y = fabsf(d.y) > 1e-20f ? 1.0f / d.y : copysignf(1e20f, d.y);
        idir.z = fabsf(d.z) > 1e-20f ? 1.0f / d.z : copysignf(1e20f, d.z);
        ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        rdd = 1.0f / dot(d, d);
        // The slab distances are lo * idir - ood: their ABSOLUTE error is a few ulp of |ood| = |o / d| (the two products cancel), which
        // for a ray that starts far from the box, or runs nearly parallel to a slab, exceeds any fixed padding of the box. The
        // triangle test has no such cancellation (vertices are translated by o first) and accepts rays within a few ulp of |o| of an
        // edge. eps (16 ulp of the largest |ood|) widens the slab comparison by both, so a box is never skipped that holds a
        // triangle the exact-as-written triangle test would accept: traversal finds what testing every triangle would find.
        eps = 4e-7f * fmaxf(fmaxf(fabsf(ood.x), fabsf(ood.y)), fabsf(ood.z));
    }
};

// returns true when the child's box overlaps [tmin, tmax] along the ray (conservative)
DI bool slab(const SpaceRay& r, float lox, float hix, float loy, float hiy, float loz, float hiz, float tmin, float tmax, float& tnear) {
    const float tx0 = fmaf(lox, r.idir.x, -r.ood.x), tx1 = fmaf(hix, r.idir.x, -r.ood.x);
    const float ty0 = fmaf(loy, r.idir.y, -r.ood.y), ty1 = fmaf(hiy, r.idir.y, -r.ood.y);
    const float tz0 = fmaf(loz, r.idir.z, -r.ood.z), tz1 = fmaf(hiz, r.idir.z, -r.ood.z);
    tnear = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
    const float tfar = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
    return tnear <= fmaf(tfar, 1.0000004f, r.eps);
}

// Edge function d . (P x Q) in binary32 with a bound of its rounding error (8 * 2^-24 times the sum of the magnitudes of its six
// products): when |value| > bound, the sign is that of the exact value.
DI float edgeFn(const float3 d, const float3 P, const float3 Q, float& bound) {
    const float px = P.y * Q.z, qx = P.z * Q.y, py = P.z * Q.x, qy = P.x * Q.z, pz = P.x * Q.y, qz = P.y * Q.x;
    const float cx = px - qx, cy = py - qy, cz = pz - qz;
    const float tx = d.x * cx, ty = d.y * cy, tz = d.z * cz;
    const float mag = (fabsf(d.x) * (fabsf(px) + fabsf(qx)) + fabsf(d.y) * (fabsf(py) + fabsf(qy))) + fabsf(d.z) * (fabsf(pz) + fabsf(qz));
    bound = mag * 4.7683716e-07f;
    return (tx + ty) + tz;
}
DI double edgeFnExact(const float3 d, const float3 P, const float3 Q) {
    const double cx = (double)P.y * (double)Q.z - (double)P.z * (double)Q.y;
    const double cy = (double)P.z * (double)Q.x - (double)P.x * (double)Q.z;
    const double cz = (double)P.x * (double)Q.y - (double)P.y * (double)Q.x;
    return ((double)d.x * cx + (double)d.y * cy) + (double)d.z * cz;
}
DI double dotExact(const float3 a, const float3 b) { return ((double)a.x * (double)b.x + (double)a.y * (double)b.y) + (double)a.z * (double)b.z; }
// The rare path of triTest (a few per cent of the tests): the ray grazes the triangle's plane or passes within rounding distance of
// an edge, so binary32 cannot tell the signs of the edge functions. In binary64 every product of two binary32 numbers is exact
// and the sums are good to 1e-15: the decision is the exact one. Out of line: it must not cost the traversal loop registers.
// Returns (t, u, v, hit ? 1 : 0) in registers.
__device__ __noinline__ float4 triTestExact(const float3 A, const float3 B, const float3 C, const float3 d) {
    const double Ud = edgeFnExact(d, C, B), Vd = edgeFnExact(d, A, C), Wd = edgeFnExact(d, B, A);
    if ((Ud < 0.0 || Vd < 0.0 || Wd < 0.0) && (Ud > 0.0 || Vd > 0.0 || Wd > 0.0)) return make_float4(0.f, 0.f, 0.f, 0.f);
    const double det = (Ud + Vd) + Wd;
    if (det == 0.0) return make_float4(0.f, 0.f, 0.f, 0.f);
    const double T = (Ud * dotExact(A, d) + Vd * dotExact(B, d)) + Wd * dotExact(C, d);
    return make_float4((float)((T / det) / dotExact(d, d)), (float)(Vd / det), (float)(Wd / det), 1.0f);
}
// A triangle is hit iff the ray passes through the translated triangle (A, B, C = v - o rounded to binary32) EXACTLY: the signs of
// the three edge functions are certified in binary32 or decided in binary64. No ray can slip between two triangles that share an
// edge (X x Y = -(Y x X) holds exactly in both precisions), and none is accepted by a triangle it does not pass through - so
// walking any conservative BVH finds exactly what testing every triangle would find (oracle/bvh.h states the same test).
DI bool triTest(const float4 t0, const float4 t1, const float4 t2, const SpaceRay& r, float tmin, float tmax, float& t, float& u, float& v) {
    const float3 A = mk3(t0.x, t0.y, t0.z) - r.o, B = mk3(t1.x, t1.y, t1.z) - r.o, C = mk3(t2.x, t2.y, t2.z) - r.o;
    float eU, eV, eW;
    const float U = edgeFn(r.d, C, B, eU);
    const float V = edgeFn(r.d, A, C, eV);
    const float W = edgeFn(r.d, B, A, eW);
    if (fabsf(U) > eU && fabsf(V) > eV && fabsf(W) > eW) {
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
        const float det = (U + V) + W;
        const float rdet = 1.0f / det;
        u = V * rdet;
        v = W * rdet;
        const float T = (U * dot(A, r.d) + V * dot(B, r.d)) + W * dot(C, r.d);
        t = (T * rdet) * r.rdd;
        return tmin < t && t < tmax;
    }
    const float4 e = triTestExact(A, B, C, r.d);
    t = e.x; u = e.y; v = e.z;
    return e.w != 0.0f && tmin < t && t < tmax;
}

template <bool ANY_HIT, bool TWO_LEVEL>
DI bool traceRay(const DeviceScene& sc, const float3 o, const float3 d, const float tmin, const float tmax, HitRec& hit) {
    hit.inst = MISS_ID; hit.tri = 0; hit.u = 0.0f; hit.v = 0.0f;
    if (sc.tlasNodes == nullptr) return false;
    float bestT = tmax;
    int stack[MSNE_STACK];
    int sp = 0;
    stack[sp++] = REF_END;
    SpaceRay ray; ray.set(o, d);
    const BvhNode2* nodes = TWO_LEVEL ? sc.tlasNodes : sc.blasNodes;
    uint32_t curInst = TWO_LEVEL ? MISS_ID : STATIC_INST;
    int node = TWO_LEVEL ? 0 : (int)sc.staticRoot;   // current reference (dtypes.h): inner node, leaf (<0), TLAS item, or sentinel
    while (node != REF_END) {
        // ---- inner nodes: descend until a leaf (or a TLAS item / sentinel) comes up ------------------------------
        while ((uint32_t)node < (uint32_t)REF_ITEM) {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            const int2 c = __ldg(reinterpret_cast<const int2*>(np + 3));
            float tn0, tn1;
            const bool h0 = slab(ray, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin, bestT, tn0);   // an empty child's box is unreachable (dtypes.h)
            const bool h1 = slab(ray, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin, bestT, tn1);
            if (h0 && h1) {
                const bool swap = tn1 < tn0;
                node = swap ? c.y : c.x;
                const int far = swap ? c.x : c.y;
                if (TWO_LEVEL ? sp < MSNE_STACK : true) stack[sp++] = far;
            } else if (h0) node = c.x;
            else if (h1) node = c.y;
            else node = stack[--sp];
        }
        // ---- leaves ---------------------------------------------------------------------------------------------
        while (node < 0) {
            const uint32_t x = ~(uint32_t)node;
            const uint32_t first = x >> 3, count = (x & 7u) + 1u;
            for (uint32_t i = 0; i < count; i++) {
                const float4* tp = reinterpret_cast<const float4*>(sc.tris + (first + i));
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                float t, u, v;
                if (!triTest(t0, t1, t2, ray, tmin, tmax, t, u, v)) continue;
                const uint32_t inst = curInst == STATIC_INST ? __float_as_uint(t2.w) : curInst;
                // a member of the flattened static BLAS that was hidden after the BLAS was built (runtime.cpp: no rebuild for a
                // visibility toggle); sc.staticHidden is a launch-uniform flag, 0 in every scene that never hides one
                if (sc.staticHidden && curInst == STATIC_INST && !__ldg(&sc.instances[inst].visible)) continue;
                // structured exit (no `return` from inside the loops): with an early return ptxas gives up the convergence
                // barriers of the whole any-hit walk and its lanes never re-join (-15 % on k_trace, profiles/r01_sweeps.md)
                if (ANY_HIT) { hit.inst = inst; node = REF_END; break; }
                bool better = t < bestT;
                if (!better && t == bestT && hit.inst != MISS_ID) {
                    if (inst != hit.inst) better = inst < hit.inst;
                    else {
                        const uint32_t g = __float_as_uint(t0.w), p = __float_as_uint(t1.w);
                        const uint32_t bg = sc.tris[hit.tri].geom, bp = sc.tris[hit.tri].prim;
                        better = g != bg ? g < bg : p < bp;
                    }
                }
                if (better) { bestT = t; hit.inst = inst; hit.tri = first + i; hit.u = u; hit.v = v; }
            }
            if (!(ANY_HIT && node == REF_END)) node = stack[--sp];
        }
        // ---- level changes (rare relative to node steps) -----------------------------------------------------------
        if (TWO_LEVEL && (uint32_t)node >= (uint32_t)REF_ITEM && node != REF_END) {
            if (node == REF_LEAVE) {
                ray.set(o, d); nodes = sc.tlasNodes; curInst = MISS_ID;
                node = stack[--sp];
            } else if (node == REF_EMPTY) {
                // only reachable from the probes (caller-supplied tmax / direction scale can make the slab test accept the
                // far-away point box of an empty child, dtypes.h): nothing there
                node = stack[--sp];
            } else {
                const uint32_t inst = __ldg(sc.tlasItems + (node & ~REF_ITEM));
                const DInstance* in = sc.instances + inst;
                if (!__ldg(&in->visible)) {
                    node = stack[--sp];          // hidden after the TLAS was built: mask 0 (Accel.zig:401), nothing to enter
                    continue;
                }
                if (__ldg(&in->isStatic)) {
                    curInst = STATIC_INST;   // world space == object space
                } else {
                    const float4* mp = reinterpret_cast<const float4*>(in->w2o);
                    const float4 r0 = __ldg(mp), r1 = __ldg(mp + 1), r2 = __ldg(mp + 2);
                    const float3 oo = mk3(((r0.x * o.x + r0.y * o.y) + r0.z * o.z) + r0.w, ((r1.x * o.x + r1.y * o.y) + r1.z * o.z) + r1.w,
                                          ((r2.x * o.x + r2.y * o.y) + r2.z * o.z) + r2.w);
                    const float3 od = mk3((r0.x * d.x + r0.y * d.y) + r0.z * d.z, (r1.x * d.x + r1.y * d.y) + r1.z * d.z, (r2.x * d.x + r2.y * d.y) + r2.z * d.z);
                    ray.set(oo, od);
                    curInst = inst;
                }
                nodes = sc.blasNodes;
                if (sp < MSNE_STACK) stack[sp++] = REF_LEAVE;
                node = (int)__ldg(&in->blasRoot);
            }
        }
    }
    return hit.inst != MISS_ID;
}

}  // namespace msne
