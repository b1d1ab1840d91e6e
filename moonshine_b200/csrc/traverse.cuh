// Software replacement of TraceRay (shaders/hrtsystem/intersection.hlsl:18-22 closest, :34-47 any-hit)
// for a GPU without RT cores: two-level BVH2 traversal + watertight ray/triangle test.
//
// The ray/triangle and instancing arithmetic is the project's definition of the (driver-opaque)
// reference behaviour and is stated in DESIGN.md §"Intersection"; the CPU oracle implements the same
// definition independently (oracle/bvh.h):
//   object-space ray   o' = mulPoint(w2o, o), d' = mulVec(w2o, d)   (t shared between spaces)
//   kz = dominant axis of d, (kx, ky) the other two in cyclic order (swapped if d[kz] < 0); Sz=1/d[kz], Sx=d[kx]*Sz, Sy=d[ky]*Sz
//   A=v0-o, B=v1-o, C=v2-o;  Ax=A[kx]-Sx*A[kz], Ay=A[ky]-Sy*A[kz] (same for B, C);  Az=Sz*A[kz] ...
//   U=Cx*By-Cy*Bx, V=Ax*Cy-Ay*Cx, W=Bx*Ay-By*Ax; any |value| <= 2^-22 (|p|+|q|) of its two products -> U, V, W, u, v, t are
//   recomputed in binary64 (exact products, exact sign of the difference): signs are the EXACT signs of the 2-D edge functions
//   miss if mixed signs; det=(U+V)+W, det==0 -> miss; rdet=1/det; u=V*rdet; v=W*rdet; t=((U*Az+V*Bz)+W*Cz)*rdet; hit iff tmin<t<tmax
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

#ifndef MSNE_SLAB_EPS
#define MSNE_SLAB_EPS 4e-7f   // measurement variants only; 0 makes the slab test unsafe for far / axis-parallel rays
#endif
#ifndef MSNE_SLAB_SUB
#define MSNE_SLAB_SUB 0   // 1: slab distances as (plane - o) * idir (two instructions per plane, relative error only, no margin constants)
#endif
#ifndef MSNE_STASH_WORLD_RAY
#define MSNE_STASH_WORLD_RAY 1   // keep the world-space ray constants (local memory) instead of recomputing them on leaving an instance: +1.4 % on cfg5
#endif
#ifndef MSNE_TRI_FALLBACK_INLINE
#define MSNE_TRI_FALLBACK_INLINE 0
#endif

struct SpaceRay {
    float3 o, idir;
#if !MSNE_SLAB_SUB
    float3 oodLo, oodHi;   // o * idir, pushed outwards per axis: subtracted from lo * idir / hi * idir respectively (slab)
#endif
    float Sx, Sy, Sz;      // shear of the ray onto its dominant axis (triTest)
    uint32_t k;            // kx | ky << 2 | kz << 4: the axis permutation of that shear
    DI void set(float3 o_, float3 d) {
        o = o_;
        idir.x = fabsf(d.x) > 1e-20f ? 1.0f / d.x : copysignf(1e20f, d.x);
        idir.y = fabsf(d.y) > 1e-20f ? 1.0f / d.y : copysignf(1e20f, d.y);
        idir.z = fabsf(d.z) > 1e-20f ? 1.0f / d.z : copysignf(1e20f, d.z);
        // The slab distances are plane * idir - o * idir: the two products cancel, so the ABSOLUTE error of a distance is a few ulp of
        // |o * idir| - which, for a ray that starts far from the box or runs nearly parallel to the slab, exceeds any fixed padding of
        // the box. Each axis therefore carries its own margin (MSNE_SLAB_EPS ~ 7 ulp of |o * idir|), applied outwards: the plane the
        // ray enters through moves towards the ray, the one it leaves through away from it. No box is skipped that holds a triangle
        // the ray passes through - traversal finds what testing every triangle would find - and the margin costs no instruction.
#if !MSNE_SLAB_SUB
        const float3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        const float3 e = mk3(MSNE_SLAB_EPS * fabsf(ood.x), MSNE_SLAB_EPS * fabsf(ood.y), MSNE_SLAB_EPS * fabsf(ood.z));
        // idir > 0: lo is the entry plane (distance must not come out too late: subtract MORE), hi the exit plane (subtract LESS)
        oodLo = mk3(idir.x > 0.0f ? ood.x + e.x : ood.x - e.x, idir.y > 0.0f ? ood.y + e.y : ood.y - e.y, idir.z > 0.0f ? ood.z + e.z : ood.z - e.z);
        oodHi = mk3(idir.x > 0.0f ? ood.x - e.x : ood.x + e.x, idir.y > 0.0f ? ood.y - e.y : ood.y + e.y, idir.z > 0.0f ? ood.z - e.z : ood.z + e.z);
#endif
        // dominant axis kz, the other two in cyclic order, swapped when d[kz] < 0 so that the winding is preserved
        const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
        const uint32_t kz = ax > ay ? (ax > az ? 0u : 2u) : (ay > az ? 1u : 2u);
        uint32_t kx = kz == 2u ? 0u : kz + 1u, ky = kx == 2u ? 0u : kx + 1u;
        const float dz = kz == 0u ? d.x : (kz == 1u ? d.y : d.z);
        if (dz < 0.0f) { const uint32_t tmp = kx; kx = ky; ky = tmp; }
        const float dx = kx == 0u ? d.x : (kx == 1u ? d.y : d.z), dy = ky == 0u ? d.x : (ky == 1u ? d.y : d.z);
        Sz = 1.0f / dz; Sx = dx * Sz; Sy = dy * Sz;   // one division: set() runs at every instance entry and exit of the two-level walk
        k = kx | (ky << 2) | (kz << 4);
    }
};

// returns true when the child's box overlaps [tmin, tmax] along the ray (conservative)
DI bool slab(const SpaceRay& r, float lox, float hix, float loy, float hiy, float loz, float hiz, float tmin, float tmax, float& tnear) {
#if MSNE_SLAB_SUB
    const float tx0 = (lox - r.o.x) * r.idir.x, tx1 = (hix - r.o.x) * r.idir.x;
    const float ty0 = (loy - r.o.y) * r.idir.y, ty1 = (hiy - r.o.y) * r.idir.y;
    const float tz0 = (loz - r.o.z) * r.idir.z, tz1 = (hiz - r.o.z) * r.idir.z;
#else
    const float tx0 = fmaf(lox, r.idir.x, -r.oodLo.x), tx1 = fmaf(hix, r.idir.x, -r.oodHi.x);
    const float ty0 = fmaf(loy, r.idir.y, -r.oodLo.y), ty1 = fmaf(hiy, r.idir.y, -r.oodHi.y);
    const float tz0 = fmaf(loz, r.idir.z, -r.oodLo.z), tz1 = fmaf(hiz, r.idir.z, -r.oodHi.z);
#endif
    tnear = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
    const float tfar = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
    return tnear <= tfar * 1.0000004f;
}

DI float pick(const float3 v, uint32_t k) { return k == 0u ? v.x : (k == 1u ? v.y : v.z); }
// The rare path of triTest: binary32 cannot certify the sign of an edge function (the ray passes within rounding distance of an
// edge line). In binary64 the two products of an edge function are exact and their difference, rounded once, has the exact sign.
// Returns (t, u, v, hit ? 1 : 0) in registers; out of line so that it costs the traversal loop no registers.
#if MSNE_TRI_FALLBACK_INLINE
DI
#else
__device__ __noinline__
#endif
float4 triTestExact(float Ax, float Ay, float Bx, float By, float Cx, float Cy, float Az, float Bz, float Cz) {
    const double U = (double)Cx * (double)By - (double)Cy * (double)Bx;
    const double V = (double)Ax * (double)Cy - (double)Ay * (double)Cx;
    const double W = (double)Bx * (double)Ay - (double)By * (double)Ax;
    if ((U < 0.0 || V < 0.0 || W < 0.0) && (U > 0.0 || V > 0.0 || W > 0.0)) return make_float4(0.f, 0.f, 0.f, 0.f);
    const double det = (U + V) + W;
    if (det == 0.0) return make_float4(0.f, 0.f, 0.f, 0.f);
    const double T = (U * (double)Az + V * (double)Bz) + W * (double)Cz;
    return make_float4((float)(T / det), (float)(V / det), (float)(W / det), 1.0f);
}
// Ray/triangle test: the project's definition of the driver-opaque reference behaviour; the CPU checker
// states the same test independently (oracle/bvh.h). The triangle is translated to the ray origin and sheared onto the ray's dominant axis, which
// turns the test into three 2-D edge functions of the projected vertices (Woop, Benthin, Wald 2013). Their SIGNS are exact: each is
// certified in binary32 (|value| above the rounding bound of its two products) or decided in binary64. Consequences:
//  * two triangles that share an edge see the same projected end points and opposite edge functions, exactly: no ray slips between;
//  * a triangle is accepted only if the ray passes through its projection, so a conservative BVH walk returns exactly what testing
//    every triangle returns - results do not depend on the tree, on traversal order, or on which boxes happened to be visited.
DI bool triTest(const float4 t0, const float4 t1, const float4 t2, const SpaceRay& r, float tmin, float tmax, float& t, float& u, float& v) {
    const float3 A = mk3(t0.x, t0.y, t0.z) - r.o, B = mk3(t1.x, t1.y, t1.z) - r.o, C = mk3(t2.x, t2.y, t2.z) - r.o;
    const uint32_t kx = r.k & 3u, ky = (r.k >> 2) & 3u, kz = r.k >> 4;
    const float Akz = pick(A, kz), Bkz = pick(B, kz), Ckz = pick(C, kz);
    const float Ax = pick(A, kx) - r.Sx * Akz, Ay = pick(A, ky) - r.Sy * Akz;
    const float Bx = pick(B, kx) - r.Sx * Bkz, By = pick(B, ky) - r.Sy * Bkz;
    const float Cx = pick(C, kx) - r.Sx * Ckz, Cy = pick(C, ky) - r.Sy * Ckz;
    const float u0 = Cx * By, u1 = Cy * Bx, v0 = Ax * Cy, v1 = Ay * Cx, w0 = Bx * Ay, w1 = By * Ax;
    const float U = u0 - u1, V = v0 - v1, W = w0 - w1;
    const float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    constexpr float kBound = 2.3841858e-07f;   // 2^-22 > 3 ulp: two rounded products and their rounded difference
    if (fabsf(U) > kBound * (fabsf(u0) + fabsf(u1)) && fabsf(V) > kBound * (fabsf(v0) + fabsf(v1)) && fabsf(W) > kBound * (fabsf(w0) + fabsf(w1))) {
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
        const float det = (U + V) + W;
        const float rdet = 1.0f / det;
        u = V * rdet;
        v = W * rdet;
        const float T = (U * Az + V * Bz) + W * Cz;
        t = T * rdet;
        return tmin < t && t < tmax;
    }
    const float4 e = triTestExact(Ax, Ay, Bx, By, Cx, Cy, Az, Bz, Cz);
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
#if MSNE_STASH_WORLD_RAY
    const SpaceRay worldRay = ray;   // two-level: restored when the walk leaves an instance (the compiler keeps it in local memory)
#endif
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
#if MSNE_STASH_WORLD_RAY
                ray = worldRay;
#else
                ray.set(o, d);
#endif
                nodes = sc.tlasNodes; curInst = MISS_ID;
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

// ------------------------------------------------------------------------------------------------
// Resumable single-level walk (kernels.cu k_trace_rs: suspend-and-regroup for sorted bounces of static-only scenes).
// The node step, the leaf test and the tie-break are those of traceRay<ANY_HIT, false>, statement for statement, so a ray that is
// walked in one go, or suspended and resumed any number of times, visits the same nodes in the same order and returns the same
// bits. What is added:
//  * the bottom of the stack is REF_FINISH, a leaf-coded reference no BLAS holds: a lane whose walk runs out pops it, lands in the
//    leaf section (not in the hot node loop) and signs off from its warp's count of walking lanes in shared memory;
//  * after every leaf visit a lane looks at that count: once at most `threshold` lanes of the warp are still walking - the others
//    only wait for them (while-while traversal keeps a warp until its longest ray ends: profiles/r02_sweeps.md section 8) - it
//    stops with its state (next reference, stack, closest hit so far) intact; the kernel files those lanes as continuation
//    records and a later pass walks 32 of them per warp. threshold = 0: never.
// ------------------------------------------------------------------------------------------------
static const int32_t REF_FINISH = (int32_t)0x80000000;
#ifndef MSNE_RS_MAX_SAVE
#define MSNE_RS_MAX_SAVE 24   // stack entries a continuation record holds (deeper walks are not suspended)
#endif
struct WalkState { int node, sp; float bestT; uint32_t inst, tri; float u, v; };

template <bool ANY_HIT>
DI bool walkResumable(const DeviceScene& sc, const SpaceRay& ray, const float tmin, const float tmax, WalkState& w, int* stack, uint32_t* live, const uint32_t threshold) {
    const BvhNode2* nodes = sc.blasNodes;
    int node = w.node, sp = w.sp;
    float bestT = w.bestT;
    bool suspended = false;
    while (node != REF_END) {
        while ((uint32_t)node < (uint32_t)REF_ITEM) {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            const int2 c = __ldg(reinterpret_cast<const int2*>(np + 3));
            float tn0, tn1;
            const bool h0 = slab(ray, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin, bestT, tn0);
            const bool h1 = slab(ray, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin, bestT, tn1);
            if (h0 && h1) {
                const bool swap = tn1 < tn0;
                node = swap ? c.y : c.x;
                stack[sp++] = swap ? c.x : c.y;
            } else if (h0) node = c.x;
            else if (h1) node = c.y;
            else node = stack[--sp];
        }
        while (node < 0) {
            if (node == REF_FINISH) { atomicSub(live, 1u); node = REF_END; break; }
            const uint32_t x = ~(uint32_t)node;
            const uint32_t first = x >> 3, count = (x & 7u) + 1u;
            for (uint32_t i = 0; i < count; i++) {
                const float4* tp = reinterpret_cast<const float4*>(sc.tris + (first + i));
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                float t, u, v;
                if (!triTest(t0, t1, t2, ray, tmin, tmax, t, u, v)) continue;
                const uint32_t inst = __float_as_uint(t2.w);
                if (sc.staticHidden && !__ldg(&sc.instances[inst].visible)) continue;
                if (ANY_HIT) { w.inst = inst; node = REF_END; break; }
                bool better = t < bestT;
                if (!better && t == bestT && w.inst != MISS_ID) {
                    if (inst != w.inst) better = inst < w.inst;
                    else {
                        const uint32_t g = __float_as_uint(t0.w), p = __float_as_uint(t1.w);
                        const uint32_t bg = sc.tris[w.tri].geom, bp = sc.tris[w.tri].prim;
                        better = g != bg ? g < bg : p < bp;
                    }
                }
                if (better) { bestT = t; w.inst = inst; w.tri = first + i; w.u = u; w.v = v; }
            }
            if (ANY_HIT && node == REF_END) atomicSub(live, 1u);
            else {
                node = stack[--sp];
                if (threshold != 0u && node != REF_FINISH && sp <= MSNE_RS_MAX_SAVE && *reinterpret_cast<volatile uint32_t*>(live) <= threshold) {
                    w.node = node; w.sp = sp;
                    atomicSub(live, 1u);
                    suspended = true;
                    node = REF_END;
                }
            }
        }
    }
    w.bestT = bestT;
    return suspended;
}

}  // namespace msne
