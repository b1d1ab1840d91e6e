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
//
// Loop structure ("while-while with postponed leaves"): a lane walks inner nodes until it holds a leaf,
// parks it and keeps descending until a second leaf shows up; only then does it run triangle tests. Lanes
// of a warp therefore meet in the triangle loop instead of executing it one lane at a time (the v1 kernel
// ran at 4.8/32 active lanes per instruction, profiles/r01a_ncu_summary_baseline.txt).
// All visible identity-transform instances live in ONE world-space BLAS (runtime.cpp buildAccel): an identity
// transform leaves the ray bit-identical, so flattening changes no result, and their triangles carry the
// instance index.
#pragma once
#include "device_math.cuh"

namespace msne {

#ifndef MSNE_STACK
#define MSNE_STACK 96
#endif
#ifndef DYN_FETCH_MIN
#define DYN_FETCH_MIN 0   // 0 = off. Measured on cfg3: 12/20/24 were 10-18 % slower than off (profiles/r01_sweeps.md)
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

template <class RayT>
DI bool triTest(const float4 t0, const float4 t1, const float4 t2, const RayT& r, float tmin, float tmax, float& t, float& u, float& v) {
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

template <bool ANY_HIT>
DI bool traceRay(const DeviceScene& sc, const float3 o, const float3 d, const float tmin, const float tmax, HitRec& hit) {
    hit.inst = MISS_ID; hit.tri = 0; hit.u = 0.0f; hit.v = 0.0f;
    if (sc.tlasNodes == nullptr) return false;
    float bestT = tmax;
    int stack[MSNE_STACK];
    int sp = 0;
    stack[sp++] = REF_END;
    SpaceRay ray; ray.set(o, d);
    const BvhNode2* nodes = sc.tlasNodes;
    uint32_t curInst = MISS_ID;
    int node = 0;   // current reference (see dtypes.h): inner node, leaf (<0), TLAS item, or sentinel
    int leaf = 0;   // parked leaf reference (0 = none; 0 is never a leaf code)
    while (node != REF_END) {
        // ---- inner nodes: descend until two leaves are pending (or the stack yields a non-inner entry) ----
        while ((uint32_t)node < (uint32_t)REF_ITEM) {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            const int2 c = __ldg(reinterpret_cast<const int2*>(np + 3));
            float tn0, tn1;
            const bool h0 = slab(ray, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin, bestT, tn0) && c.x != REF_EMPTY;
            const bool h1 = slab(ray, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin, bestT, tn1) && c.y != REF_EMPTY;
            if (h0 && h1) {
                const bool swap = tn1 < tn0;
                node = swap ? c.y : c.x;
                const int far = swap ? c.x : c.y;
                if (sp < MSNE_STACK) stack[sp++] = far;
            } else if (h0) node = c.x;
            else if (h1) node = c.y;
            else node = stack[--sp];
            if (node < 0 && leaf == 0) { leaf = node; node = stack[--sp]; }   // park the first leaf, keep walking
        }
        // ---- triangles of the parked leaf, then of `node` if that is a leaf too --------------------------
        while (leaf != 0) {
            const uint32_t x = ~(uint32_t)leaf;
            const uint32_t first = x >> 3, count = (x & 7u) + 1u;
            for (uint32_t i = 0; i < count; i++) {
                const float4* tp = reinterpret_cast<const float4*>(sc.tris + (first + i));
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                float t, u, v;
                if (!triTest(t0, t1, t2, ray, tmin, tmax, t, u, v)) continue;
                const uint32_t inst = curInst == STATIC_INST ? __float_as_uint(t2.w) : curInst;
                if (ANY_HIT) { hit.inst = inst; return true; }
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
            leaf = 0;
            if (node < 0) { leaf = node; node = stack[--sp]; }
        }
        // ---- level changes (rare relative to node steps) -------------------------------------------------
        if ((uint32_t)node >= (uint32_t)REF_ITEM && node != REF_END) {
            if (node == REF_LEAVE) {
                ray.set(o, d); nodes = sc.tlasNodes; curInst = MISS_ID;
                node = stack[--sp];
            } else {
                const uint32_t inst = __ldg(sc.tlasItems + (node & ~REF_ITEM));
                const DInstance* in = sc.instances + inst;
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


// Resumable form of traceRay for the persistent kernel: the same loop, with its state in a struct so a lane can
// leave the loop (dynamic fetch) and continue later. `run` returns true when the ray is finished.
struct Traversal {
    float3 o, d;
    float tmin, tmax, bestT;
    SpaceRay ray;
    HitRec hit;
    const BvhNode2* nodes;
    uint32_t curInst, target;
    int sp, node, leaf;

    DI void begin(const DeviceScene& sc, int* stack, float3 o_, float3 d_, float tmin_, float tmax_) {
        o = o_; d = d_; tmin = tmin_; tmax = tmax_; bestT = tmax_;
        hit.inst = MISS_ID; hit.tri = 0; hit.u = 0.0f; hit.v = 0.0f;
        ray.set(o, d);
        nodes = sc.tlasNodes; curInst = MISS_ID; target = 0;
        sp = 0; stack[sp++] = REF_END;
        node = sc.tlasNodes ? 0 : REF_END; leaf = 0;
    }

    template <bool ANY_HIT>
    DI bool run(const DeviceScene& sc, int* stack, bool mayRefill) {
        while (node != REF_END) {
            while ((uint32_t)node < (uint32_t)REF_ITEM) {
                const float4* np = reinterpret_cast<const float4*>(nodes + node);
                const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
                const int2 c = __ldg(reinterpret_cast<const int2*>(np + 3));
                float tn0, tn1;
                const bool h0 = slab(ray, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin, bestT, tn0) && c.x != REF_EMPTY;
                const bool h1 = slab(ray, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin, bestT, tn1) && c.y != REF_EMPTY;
                if (h0 && h1) {
                    const bool swap = tn1 < tn0;
                    node = swap ? c.y : c.x;
                    const int far = swap ? c.x : c.y;
                    if (sp < MSNE_STACK) stack[sp++] = far;
                } else if (h0) node = c.x;
                else if (h1) node = c.y;
                else node = stack[--sp];
                if (node < 0 && leaf == 0) { leaf = node; node = stack[--sp]; }
            }
            while (leaf != 0) {
                const uint32_t x = ~(uint32_t)leaf;
                const uint32_t first = x >> 3, count = (x & 7u) + 1u;
                for (uint32_t i = 0; i < count; i++) {
                    const float4* tp = reinterpret_cast<const float4*>(sc.tris + (first + i));
                    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                    float t, u, v;
                    if (!triTest(t0, t1, t2, ray, tmin, tmax, t, u, v)) continue;
                    const uint32_t inst = curInst == STATIC_INST ? __float_as_uint(t2.w) : curInst;
                    if (ANY_HIT) { hit.inst = inst; node = REF_END; leaf = 0; return true; }
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
                leaf = 0;
                if (node < 0) { leaf = node; node = stack[--sp]; }
            }
            if ((uint32_t)node >= (uint32_t)REF_ITEM && node != REF_END) {
                if (node == REF_LEAVE) {
                    ray.set(o, d); nodes = sc.tlasNodes; curInst = MISS_ID;
                    node = stack[--sp];
                } else {
                    const uint32_t inst = __ldg(sc.tlasItems + (node & ~REF_ITEM));
                    const DInstance* in = sc.instances + inst;
                    if (__ldg(&in->isStatic)) {
                        curInst = STATIC_INST;
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
            // dynamic fetch: too few lanes of this warp are still traversing -> go refill the idle ones
            if (mayRefill && __popc(__activemask()) < DYN_FETCH_MIN) break;
        }
        return node == REF_END;
    }
};


// ================================================================================================================
// Compressed wide BVH (BvhNode8, dtypes.h) traversal. Same two levels, same triangle test, same closest-hit rule
// (min t, ties -> smallest ids), so hit records are identical to the BVH2 path and to the oracle's own BVH.
//
// Traversal state = one "node group" (base index of the children of the last node visited + an 8-bit mask of those
// still to visit, already ordered front-to-back for this ray's octant: the highest set bit is the nearest child) and
// one "item group" (base index + 24-bit mask of the leaf items whose child box was hit). The stack holds groups, so
// one entry covers up to 8 pending children; a node visit costs five 16-byte loads for 8 boxes (10 B per box instead
// of 32 B), and the dependent-load chain is log8 instead of log2 deep.
// ================================================================================================================
#ifndef MSNE_STACK8
#define MSNE_STACK8 40
#endif

DI uint32_t signExtendS8x4(uint32_t x) { uint32_t r; asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x)); return r; }
DI uint32_t highBit(uint32_t x) { return 31u - (uint32_t)__clz((int)x); }
// byte j of x as a float without the (1/8-rate) I2F unit: 0x4B0000bb is the float 2^23 + bb, exactly
DI float byteToFloat(uint32_t x, int j) { return __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650u + (uint32_t)j)) - 8388608.0f; }

struct WideRay {
    float3 o, d, idir;
    float rdd;
    uint32_t octInv4;   // (7 - octant) replicated into 4 bytes; octant bit a = ray runs towards -axis a
    DI void set(float3 o_, float3 d_) {
        o = o_; d = d_;
        idir.x = fabsf(d.x) > 1e-20f ? 1.0f / d.x : copysignf(1e20f, d.x);
        idir.y = fabsf(d.y) > 1e-20f ? 1.0f / d.y : copysignf(1e20f, d.y);
        idir.z = fabsf(d.z) > 1e-20f ? 1.0f / d.z : copysignf(1e20f, d.z);
        rdd = 1.0f / dot(d, d);
        const uint32_t oct = (idir.x < 0.0f ? 1u : 0u) | (idir.y < 0.0f ? 2u : 0u) | (idir.z < 0.0f ? 4u : 0u);
        octInv4 = (7u - oct) * 0x01010101u;
    }
};

// child-box tests of one wide node -> hit mask (bits 24..31: inner children in traversal priority, bits 0..23: items)
DI uint32_t wideNodeHits(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4, const WideRay& r, float tmin, float tmax) {
    const uint32_t eim = __float_as_uint(n0.w);
    const float ax = __uint_as_float((eim & 0xFFu) << 23) * r.idir.x;
    const float ay = __uint_as_float(((eim >> 8) & 0xFFu) << 23) * r.idir.y;
    const float az = __uint_as_float(((eim >> 16) & 0xFFu) << 23) * r.idir.z;
    const float bx = (n0.x - r.o.x) * r.idir.x, by = (n0.y - r.o.y) * r.idir.y, bz = (n0.z - r.o.z) * r.idir.z;
    const bool nx = r.idir.x < 0.0f, ny = r.idir.y < 0.0f, nz = r.idir.z < 0.0f;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t innerMask4 = signExtendS8x4(isInner4 << 3);
        const uint32_t bitIndex4 = (meta4 ^ (r.octInv4 & innerMask4)) & 0x1F1F1F1Fu;
        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlx = __float_as_uint(half ? n2.y : n2.x), qly = __float_as_uint(half ? n2.w : n2.z), qlz = __float_as_uint(half ? n3.y : n3.x);
        const uint32_t qhx = __float_as_uint(half ? n3.w : n3.z), qhy = __float_as_uint(half ? n4.y : n4.x), qhz = __float_as_uint(half ? n4.w : n4.z);
        const uint32_t x0 = nx ? qhx : qlx, x1 = nx ? qlx : qhx;
        const uint32_t y0 = ny ? qhy : qly, y1 = ny ? qly : qhy;
        const uint32_t z0 = nz ? qhz : qlz, z1 = nz ? qlz : qhz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float tx0 = fmaf(byteToFloat(x0, j), ax, bx), tx1 = fmaf(byteToFloat(x1, j), ax, bx);
            const float ty0 = fmaf(byteToFloat(y0, j), ay, by), ty1 = fmaf(byteToFloat(y1, j), ay, by);
            const float tz0 = fmaf(byteToFloat(z0, j), az, bz), tz1 = fmaf(byteToFloat(z1, j), az, bz);
            const float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tmin));
            const float tf = fminf(fminf(tx1, ty1), fminf(tz1, tmax));
            if (tn <= tf * 1.0000004f) hitmask |= ((childBits4 >> (8 * j)) & 0xFFu) << ((bitIndex4 >> (8 * j)) & 0xFFu);
        }
    }
    return hitmask;
}

template <bool ANY_HIT>
DI bool traceRay8(const DeviceScene& sc, const float3 o, const float3 d, const float tmin, const float tmax, HitRec& hit) {
    hit.inst = MISS_ID; hit.tri = 0; hit.u = 0.0f; hit.v = 0.0f;
    if (sc.tlasNodes8 == nullptr) return false;
    float bestT = tmax;
    uint2 stack[MSNE_STACK8];
    int sp = 0;
    WideRay ray; ray.set(o, d);
    const BvhNode8* nodes;
    uint32_t curInst;
    bool inBlas;
    uint2 ng, tg = make_uint2(0u, 0u);
    if (sc.staticOnly) { nodes = sc.blasNodes8; curInst = STATIC_INST; inBlas = true; ng = make_uint2(sc.staticRoot, 0x80000000u); }
    else { nodes = sc.tlasNodes8; curInst = MISS_ID; inBlas = false; ng = make_uint2(0u, 0x80000000u); }
    for (;;) {
        if (ng.y > 0x00FFFFFFu) {
            // ---- visit the nearest pending child of the node group --------------------------------------------
            const uint32_t imask = ng.y & 0xFFu;
            const uint32_t bit = highBit(ng.y);
            ng.y &= ~(1u << bit);
            if (ng.y > 0x00FFFFFFu && sp < MSNE_STACK8) stack[sp++] = ng;
            const uint32_t slot = (bit - 24u) ^ (ray.octInv4 & 0xFFu);
            const uint32_t rel = (uint32_t)__popc(imask & ~(0xFFFFFFFFu << slot));
            const float4* np = reinterpret_cast<const float4*>(nodes + (ng.x + rel));
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            const uint32_t hits = wideNodeHits(n0, n1, n2, n3, n4, ray, tmin, bestT);
            ng.x = __float_as_uint(n1.x);
            ng.y = (hits & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
            tg.x = __float_as_uint(n1.y);
            tg.y = hits & 0x00FFFFFFu;
        } else {
            tg = ng;
            ng = make_uint2(0u, 0u);
        }
        if (inBlas) {
            // ---- triangles ------------------------------------------------------------------------------------
            while (tg.y != 0u) {
                const uint32_t bit = highBit(tg.y);
                tg.y &= ~(1u << bit);
                const uint32_t ti = tg.x + bit;
                const float4* tp = reinterpret_cast<const float4*>(sc.tris + ti);
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                float t, u, v;
                if (!triTest(t0, t1, t2, ray, tmin, tmax, t, u, v)) continue;
                const uint32_t inst = curInst == STATIC_INST ? __float_as_uint(t2.w) : curInst;
                if (ANY_HIT) { hit.inst = inst; return true; }
                bool better = t < bestT;
                if (!better && t == bestT && hit.inst != MISS_ID) {
                    if (inst != hit.inst) better = inst < hit.inst;
                    else {
                        const uint32_t g = __float_as_uint(t0.w), p = __float_as_uint(t1.w);
                        const uint32_t bg = sc.tris[hit.tri].geom, bp = sc.tris[hit.tri].prim;
                        better = g != bg ? g < bg : p < bp;
                    }
                }
                if (better) { bestT = t; hit.inst = inst; hit.tri = ti; hit.u = u; hit.v = v; }
            }
        } else if (tg.y != 0u) {
            // ---- TLAS leaf: enter one instance, park the rest of both groups under a LEAVE marker (y == 0) ------
            const uint32_t bit = highBit(tg.y);
            tg.y &= ~(1u << bit);
            const uint32_t inst = __ldg(sc.tlasItems + (tg.x + bit));
            if (tg.y != 0u && sp < MSNE_STACK8) stack[sp++] = tg;
            if (ng.y > 0x00FFFFFFu && sp < MSNE_STACK8) stack[sp++] = ng;
            if (sp < MSNE_STACK8) stack[sp++] = make_uint2(0u, 0u);
            const DInstance* in = sc.instances + inst;
            if (__ldg(&in->isStatic)) {
                curInst = STATIC_INST;
            } else {
                const float4* mp = reinterpret_cast<const float4*>(in->w2o);
                const float4 r0 = __ldg(mp), r1 = __ldg(mp + 1), r2 = __ldg(mp + 2);
                const float3 oo = mk3(((r0.x * o.x + r0.y * o.y) + r0.z * o.z) + r0.w, ((r1.x * o.x + r1.y * o.y) + r1.z * o.z) + r1.w,
                                      ((r2.x * o.x + r2.y * o.y) + r2.z * o.z) + r2.w);
                const float3 od = mk3((r0.x * d.x + r0.y * d.y) + r0.z * d.z, (r1.x * d.x + r1.y * d.y) + r1.z * d.z, (r2.x * d.x + r2.y * d.y) + r2.z * d.z);
                ray.set(oo, od);
                curInst = inst;
            }
            nodes = sc.blasNodes8; inBlas = true;
            ng = make_uint2(__ldg(&in->blasRoot), 0x80000000u);
            tg = make_uint2(0u, 0u);
            continue;
        }
        // ---- node group exhausted: pop ------------------------------------------------------------------------
        if (ng.y <= 0x00FFFFFFu) {
            for (;;) {
                if (sp == 0) return hit.inst != MISS_ID;
                ng = stack[--sp];
                if (ng.y != 0u) break;
                ray.set(o, d); nodes = sc.tlasNodes8; inBlas = false; curInst = MISS_ID;   // LEAVE marker
            }
        }
    }
}

}  // namespace msne
