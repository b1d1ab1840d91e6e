// sm_100a wavefront path-tracing kernels — the B200-native replacement of the reference's ray-tracing
// pipeline (shaders/hrtsystem/main.hlsl raygen/closesthit/miss/shadowmiss driven by vkCmdTraceRaysKHR,
// engine/hrtsystem/pipeline.zig:269-271).
//
// One camera sample of one pixel = one path. Paths live in float4 queues in HBM (dtypes.h) and move
// through two kernels per bounce:
//
//   k_trace  persistent threads; warps pull 32-ray tickets from ONE index space that holds this bounce's
//            extension rays (closest hit -> H[q]) followed by the previous bounce's shadow rays (any hit ->
//            zero the pending NEE contribution of path q when occluded). Long queues are traced in sorted
//            order (radix sort of the 15-bit direction-octant + origin-cell keys the shade pass wrote), and
//            k_bin then re-lists the hits per material in queue order.
//   k_shade  integrator.hlsl:68-183 for one bounce: fold pending NEE, emission (+MIS), termination /
//            Russian roulette, env + mesh NEE sample generation (-> shadow queue), BSDF sample -> next
//            extension ray; survivors are compacted into the other ping-pong buffer with warp-ballot
//            aggregated atomics, finished paths write their radiance to result[slot]. One launch per
//            material bin (specialised code), one generic launch for short queues, k_tail for stragglers.
//
// k_generate starts a wave (main.hlsl:85-88), k_accumulate applies storeColor (main.hlsl:43-51) in sample
// order. Compiled with -fmad=false (see device_math.cuh).
#include "kernels.h"
#include <cstdlib>
#include "shading.cuh"
#include "traverse.cuh"

namespace msne {

// Queue traffic (path state, hit records, shadow rays, bin lists) is written once and read once, a whole wave (GBs) apart:
// MSNE_STREAM_HINTS = 1 marks it evict-first (ld/st.global.cs) so that it does not displace the scene data in L2.
#ifndef MSNE_STREAM_HINTS
#define MSNE_STREAM_HINTS 1
#endif
#ifndef MSNE_SORT_KEY
#define MSNE_SORT_KEY 0   // 0: octant-major key, 1: origin-cell-major key (measurement variant)
#endif
// (2 = loads only, 3 = stores only: measurement variants)
#if MSNE_STREAM_HINTS == 1 || MSNE_STREAM_HINTS == 2
DI float4 qld(const float4* p) { return __ldcs(p); }
DI uint4 qld(const uint4* p) { return __ldcs(p); }
DI uint32_t qld(const uint32_t* p) { return __ldcs(p); }
#else
DI float4 qld(const float4* p) { return *p; }
DI uint4 qld(const uint4* p) { return *p; }
DI uint32_t qld(const uint32_t* p) { return __ldg(p); }
#endif
#if MSNE_STREAM_HINTS == 1 || MSNE_STREAM_HINTS == 3
DI void qst(float4* p, float4 v) { __stcs(p, v); }
DI void qst(uint4* p, uint4 v) { __stcs(p, v); }
DI void qst(uint32_t* p, uint32_t v) { __stcs(p, v); }
#else
DI void qst(float4* p, float4 v) { *p = v; }
DI void qst(uint4* p, uint4 v) { *p = v; }
DI void qst(uint32_t* p, uint32_t v) { *p = v; }
#endif

// ------------------------------------------------------------------------------------------------
// generate
// ------------------------------------------------------------------------------------------------
// Camera::generateRay (camera.hlsl:31-41); basis hoisted to the host (runtime.cpp makeCamera)
DI void cameraRayUV(const DCamera& cam, float2 uv, float2 r1, float3& origin, float3& dir);
DI void cameraRay(const DCamera& cam, const DConstants& k, uint32_t px, uint32_t py, float2 r0, float2 r1, float3& origin, float3& dir) {
    // dispatchUV (main.hlsl:54-59)
    const float2 randomCenter = mk2(0.5f, 0.5f) + 0.5f * squareToGaussian(r0);
    float2 uv = (mk2((float)px, (float)py) + randomCenter) / mk2((float)cam.width, (float)cam.height);
    if (k.flip_image) uv.y = 1.0f - uv.y;
    cameraRayUV(cam, uv, r1, origin, dir);
}
DI void cameraRayUV(const DCamera& cam, float2 uv, float2 r1, float3& origin, float3& dir) {
    const float3 u = mk3(cam.u[0], cam.u[1], cam.u[2]), v = mk3(cam.v[0], cam.v[1], cam.v[2]);
    const float3 horizontal = mk3(cam.horizontal[0], cam.horizontal[1], cam.horizontal[2]);
    const float3 vertical = mk3(cam.vertical[0], cam.vertical[1], cam.vertical[2]);
    const float3 llc = mk3(cam.llc[0], cam.llc[1], cam.llc[2]);
    const float3 camOrigin = mk3(cam.origin[0], cam.origin[1], cam.origin[2]);
    const float2 sampled = squareToUniformDiskConcentric(r1);
    const float2 rd = (cam.aperture * sampled) / 2.0f;
    const float3 defocusOffset = u * rd.x + v * rd.y;
    origin = camOrigin + defocusOffset;
    dir = normalize(llc + uv.x * horizontal + uv.y * vertical - defocusOffset - camOrigin);
}

__global__ void __launch_bounds__(256) k_generate(DCamera cam, DConstants k, PathBuffers pb, Counters* counters, const uint32_t* __restrict__ pixIds,
                                                  uint32_t nPix, uint32_t nSamples, uint32_t firstSample) {
    const uint32_t total = nPix * nSamples;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) {
        counters[0].nLive = total; counters[0].nDead = 0; counters[0].nShadow = 0;
        counters[1].nLive = 0; counters[1].nDead = 0; counters[1].nShadow = 0;
    }
    if (p >= total) return;
    const uint32_t s = p / nPix, pi = p - s * nPix;
    const uint32_t pix = __ldg(pixIds + pi);
    const uint32_t px = pix % cam.width, py = pix / cam.width;
    Rng rng = Rng::fromSeed(firstSample + s, px, py);
    const float2 r0 = rng.getFloat2();
    const float2 r1 = rng.getFloat2();
    float3 o, d;
    cameraRay(cam, k, px, py, r0, r1, o, d);
    pb.R0[0][RAY_STRIDE * (size_t)p] = make_float4(o.x, o.y, o.z, 0.0f);
    pb.R1[0][RAY_STRIDE * (size_t)p] = make_float4(d.x, d.y, d.z, __uint_as_float(0u));
    pb.S0[0][p] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(rng.state));
    pb.S1[0][p] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(p));
    pb.N0[0][p] = make_float4(0.f, 0.f, 0.f, 0.f);
    pb.N1[0][p] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// debug probe: primary rays + rng state (HdMoonshineExtPrimaryRays)
__global__ void k_primary_probe(DCamera cam, DConstants k, uint32_t sample, float* rays6, uint32_t* rngState) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cam.width * cam.height) return;
    const uint32_t px = p % cam.width, py = p / cam.width;
    Rng rng = Rng::fromSeed(sample, px, py);
    const float2 r0 = rng.getFloat2();
    const float2 r1 = rng.getFloat2();
    float3 o, d;
    cameraRay(cam, k, px, py, r0, r1, o, d);
    float* r = rays6 + 6 * (size_t)p;
    r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
    rngState[p] = rng.state;
}

// ------------------------------------------------------------------------------------------------
// trace (persistent threads)
// ------------------------------------------------------------------------------------------------
// material type of a hit (MaterialType 0..3) or BIN_MISS: world.hlsl:61-71 lookups, done once here so the shade
// pass can run one specialised kernel per type ("sorting rays by material id")
DI uint32_t hitBin(const DeviceScene& sc, const HitRec& h) {
    if (h.inst == MISS_ID) return BIN_MISS;
    const uint32_t instanceID = __ldg(&sc.instances[h.inst].customIndex);
    const uint32_t geom = __ldg(&sc.tris[h.tri].geom);
    const uint32_t mat = __ldg(&sc.geometries[instanceID + geom].material);
    return __ldg(&sc.materials[mat].type);
}

// Persistent-threads trace kernel: warps pull 32-ray tickets from the joint index space [extension rays | shadow rays]
// of this bounce, every lane walks its ray to the end (traverse.cuh), then the warp sorts its hits by material type.
// Refilling idle lanes before the whole ticket is done ("dynamic fetch") was measured in three variants and is not
// used (profiles/r01_sweeps.md).
// SORTED: the tickets walk the permutations of the extension and the shadow queue (counting sort of the keys the shade pass wrote:
// rays ordered by direction octant and origin cell) instead of the queue order, so the 32 rays of a warp visit the same part of
// the tree; the material binning is then done by k_bin in queue order (the shade kernels keep reading contiguous path state).
//
// Ticket staging (MSNE_TRACE_STAGING): the rays of a warp's NEXT ticket are copied into shared memory while the warp traverses the
// current one, so no ticket starts with every lane waiting a DRAM round trip for its ray (queues are streamed evict-first: always
// a miss), and without holding the next ray in registers - registers are what bounds this kernel's occupancy.
//   queue order: a ticket is 32 consecutive rays = ONE contiguous kilobyte -> one bulk tensor-memory-accelerator copy per ticket
//                (cp.async.bulk.shared.global, completion counted in bytes on a per-warp mbarrier; SASS: UBLKCP + SYNCS);
//   sorted:      the rays are gathered through the permutation -> two 16-byte cp.async (LDGSTS) per lane, the permutation entries
//                themselves fetched one ticket further ahead into a register.
// The index space is padded so that a ticket never straddles the two queues: extension tickets cover [0, roundup32(nLive)), shadow
// tickets follow.
#ifndef MSNE_TRACE_STAGING
#define MSNE_TRACE_STAGING 0
#endif
DI uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DI void mbarInit(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
DI void mbarExpectTx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
DI void mbarWait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
DI void bulkCopyG2S(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
DI void cpAsync16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
DI void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DI void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool TWO_LEVEL, bool SORTED>
__global__ void __launch_bounds__(TRACE_THREADS, TWO_LEVEL ? TRACE_BLOCKS_PER_SM : TRACE1_BLOCKS_PER_SM) k_trace(DeviceScene sc, PathBuffers pb, WaveState ws, uint32_t buf,
                                                                                                               const uint32_t* __restrict__ permExt,
                                                                                                               const uint32_t* __restrict__ permShadow) {
    const Counters cin = ws.counters[buf];
    const uint32_t nLive = cin.nLive, nShadow = cin.nShadow;
    const uint32_t live32 = (nLive + 31u) & ~31u;          // shadow tickets start at a ticket boundary
    const uint32_t total = live32 + nShadow;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ltMask = (1u << lane) - 1u;
    const float4* __restrict__ extRays = pb.R0[buf];       // ray q = the two float4 lanes at [2 q], [2 q + 1] (dtypes.h RAY_STRIDE)
    const float4* __restrict__ shadowRays = pb.Q0;
    // queue slot of index-space position i (0xFFFFFFFF: padding / past the end), and whether it is an extension ray
    auto slotOf = [&](uint32_t i, uint32_t permEntry) -> uint32_t {
        if (i < live32) return i < nLive ? (SORTED ? permEntry : i) : 0xFFFFFFFFu;
        return i < total ? (SORTED ? permEntry : i - live32) : 0xFFFFFFFFu;
    };
    auto permLoad = [&](uint32_t i) -> uint32_t {
        if (!SORTED) return 0u;
        if (i < nLive) return __ldg(permExt + i);
        if (i >= live32 && i < total) return __ldg(permShadow + (i - live32));
        return 0u;
    };
    auto grab = [&]() -> uint32_t {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(ws.traceTicket, 32u);
        return __shfl_sync(0xFFFFFFFFu, b, 0);
    };
#if MSNE_TRACE_STAGING
    __shared__ __align__(128) float4 stage[TRACE_THREADS / 32][64];   // one kilobyte per warp: the next ticket's 32 rays
    __shared__ __align__(8) unsigned long long stageBar[TRACE_THREADS / 32];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t myStage = smemAddr(&stage[warp][0]), myBar = smemAddr(&stageBar[warp]);
    uint32_t parity = 0;
    if (!SORTED) {
        if (lane == 0) { mbarInit(myBar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
    }
    // starts the copy of ticket `b` (its rays' queue slots: `slot` per lane) into this warp's stage
    auto stageTicket = [&](uint32_t b, uint32_t slot) {
        if (b >= total) return;
        const bool ext = b < live32;
        if (SORTED) {
            if (slot != 0xFFFFFFFFu) {
                const float4* src = (ext ? extRays : shadowRays) + RAY_STRIDE * (size_t)slot;
                cpAsync16(myStage + lane * 32u, src);
                cpAsync16(myStage + lane * 32u + 16u, src + 1);
            }
            cpAsyncCommit();
        } else if (lane == 0) {
            const uint32_t first = ext ? b : b - live32, n = ext ? nLive : nShadow;
            const uint32_t count = n - first < 32u ? n - first : 32u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was last read through the generic proxy
            mbarExpectTx(myBar, count * 32u);
            bulkCopyG2S(myStage, (ext ? extRays : shadowRays) + RAY_STRIDE * (size_t)first, count * 32u, myBar);
        }
    };
    // software pipeline: ticket k is traversed while ticket k+1's rays are in flight to shared memory and (SORTED) ticket k+2's
    // permutation entries are in flight to a register. Holding tickets ahead costs load balance at the end of a queue (a warp
    // then works through its two or three tickets while others idle), so a warp only grabs ahead while every warp of the grid
    // still has several tickets to come; short queues and queue ends run one ticket at a time.
    constexpr uint32_t NONE = 0xFFFFFFFEu;
    const uint32_t margin = gridDim.x * (TRACE_THREADS / 32u) * 32u * 4u;
    uint32_t base = grab(), perm = permLoad(base + lane);
    bool staged = false;
    uint32_t baseNext = NONE, permNext = 0, baseAfter = NONE, permAfter = 0;
    while (base < total) {
        const bool ahead = base + margin < total && base + margin > base;
        if (!staged) stageTicket(base, slotOf(base + lane, perm));
        if (ahead && baseNext == NONE) { baseNext = grab(); permNext = permLoad(baseNext + lane); }
        if (SORTED && ahead && baseAfter == NONE) { baseAfter = grab(); permAfter = permLoad(baseAfter + lane); }
        const uint32_t idx = slotOf(base + lane, perm);
        const bool ext = base < live32;
        // this ticket's rays have landed
        if (SORTED) cpAsyncWaitAll(); else { mbarWait(myBar, parity); parity ^= 1u; }
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        if (idx != 0xFFFFFFFFu) { r0 = stage[warp][2 * lane]; r1 = stage[warp][2 * lane + 1]; }
        __syncwarp();
        bool nextStaged = false;
        if (baseNext != NONE && baseNext < total) { stageTicket(baseNext, slotOf(baseNext + lane, permNext)); nextStaged = true; }
#else
    {
    uint32_t base = 0, perm = 0;
    for (;;) {
        base = grab();
        if (base >= total) break;
        perm = permLoad(base + lane);
        const uint32_t idx = slotOf(base + lane, perm);
        const bool ext = base < live32;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        if (idx != 0xFFFFFFFFu) {
            const float4* src = (ext ? extRays : shadowRays) + RAY_STRIDE * (size_t)idx;
            r0 = qld(src); r1 = qld(src + 1);
        }
#endif
        uint32_t bin = 0xFFu;
        if (idx != 0xFFFFFFFFu) {
            HitRec h;
            if (ext) {
                traceRay<false, TWO_LEVEL>(sc, mk3(r0.x, r0.y, r0.z), mk3(r1.x, r1.y, r1.z), 0.0f, MS_INFINITY, h);
                bin = hitBin(sc, h);
                // the material bin rides in the top bits of the triangle index (H_TRI_MASK, dtypes.h): k_bin reads it back without
                // repeating the instance -> geometry -> material chain
                qst(pb.H + idx, make_uint4(h.inst, h.tri | (bin << H_BIN_SHIFT), __float_as_uint(h.u), __float_as_uint(h.v)));
            } else if (traceRay<true, TWO_LEVEL>(sc, mk3(r0.x, r0.y, r0.z), mk3(r1.x, r1.y, r1.z), 0.0f, r0.w, h)) {
                const uint32_t target = __float_as_uint(r1.w);
                float4* n = (target & 1u) ? pb.N1[buf] : pb.N0[buf];
                n[target >> 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // ---- warp-level sort by material id: lanes with the same bin elect a leader that reserves their slots
        if (!SORTED) {
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, bin);
            if (bin != 0xFFu) {
                const int leader = __ffs(peers) - 1;
                uint32_t slot = 0;
                if ((int)lane == leader) slot = atomicAdd(ws.typeCount + bin, (uint32_t)__popc(peers));
                slot = __shfl_sync(peers, slot, leader) + __popc(peers & ltMask);
                qst(pb.typeList + ((size_t)bin * pb.capacity + slot), idx);
            }
        }
#if MSNE_TRACE_STAGING
        if (baseNext != NONE) { base = baseNext; perm = permNext; staged = nextStaged; baseNext = baseAfter; permNext = permAfter; baseAfter = NONE; }
        else { base = grab(); perm = permLoad(base + lane); staged = false; }
    }
#else
    }
    }
#endif
    // last block out: re-arm the ticket for the next launch and publish ray counts
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(ws.traceDone, 1u);
        if (done == gridDim.x - 1) {
            *ws.traceTicket = 0; *ws.traceDone = 0;
            ws.rayCounts[0] += nLive; ws.rayCounts[1] += nShadow;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_trace_rs: sorted bounce of a static-only scene with suspend-and-regroup (kernels.h ContBuffers, traverse.cuh walkResumable).
// A while-while warp is kept until its longest ray ends; in a sorted bounce 26 % of a warp's lane-steps are lanes whose ray is
// finished (tools/bvh_sim.cpp). Here the last few walking lanes of a warp write their traversal state to a continuation record
// and the warp takes its next ticket; the records - the long rays of the bounce - are walked together, 32 per warp, by the next
// pass. Results are those of k_trace bit for bit (same walk, same tie-break; hit records land in the same slots).
// PASS 0: tickets over the sorted queues (as k_trace<false, true>), PASS 1 / 2: tickets over the records of recA / recB.
// ------------------------------------------------------------------------------------------------
#define RS_REC_U4 8   // uint4 per record: 2 header + 6 x 4 stack entries
static_assert(2 + (MSNE_RS_MAX_SAVE + 3) / 4 <= RS_REC_U4, "a continuation record must hold every stack entry a suspended walk may save");
static_assert(MSNE_RS_MAX_SAVE + 4 <= MSNE_STACK, "restoring a record writes whole uint4 groups into the traversal stack");
template <int PASS>
__global__ void __launch_bounds__(TRACE_THREADS, TRACE1_BLOCKS_PER_SM) k_trace_rs(DeviceScene sc, PathBuffers pb, WaveState ws, uint32_t buf, const uint32_t* __restrict__ permExt,
                                                                                  const uint32_t* __restrict__ permShadow, ContBuffers cb, uint32_t threshold) {
    __shared__ uint32_t liveCount[TRACE_THREADS / 32];
    const Counters cin = ws.counters[buf];
    const uint32_t nLive = cin.nLive, nShadow = cin.nShadow;
    const uint32_t live32 = (nLive + 31u) & ~31u;
    const uint32_t total = PASS == 0 ? live32 + nShadow : cb.counts[PASS - 1];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    const float4* __restrict__ extRays = pb.R0[buf];
    const float4* __restrict__ shadowRays = pb.Q0;
    const uint4* __restrict__ inRec = PASS == 1 ? cb.recA : cb.recB;
    uint4* outRec = PASS == 0 ? cb.recA : cb.recB;
    uint32_t* outCount = cb.counts + (PASS == 0 ? 0 : 1);
    const uint32_t thr = PASS == 2 ? 0u : threshold;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(ws.traceTicket, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= total) break;
        const uint32_t i = base + lane;
        uint32_t idx = 0xFFFFFFFFu;
        bool ext = false;
        WalkState w;
        w.node = (int)sc.staticRoot; w.sp = 1; w.bestT = 0.0f; w.inst = MISS_ID; w.tri = 0; w.u = 0.0f; w.v = 0.0f;
        int stack[MSNE_STACK];
        stack[0] = REF_FINISH;
        if (PASS == 0) {
            ext = base < live32;
            if (ext) { if (i < nLive) idx = __ldg(permExt + i); }
            else if (i < total) idx = __ldg(permShadow + (i - live32));
        } else if (i < total) {
            const uint4* rec = inRec + (size_t)i * RS_REC_U4;
            const uint4 h0 = rec[0], h1 = rec[1];
            idx = h0.x & 0x7FFFFFFFu; ext = (h0.x >> 31) != 0u;
            w.bestT = __uint_as_float(h0.y); w.inst = h0.z; w.tri = h0.w;
            w.u = __uint_as_float(h1.x); w.v = __uint_as_float(h1.y); w.node = (int)h1.z; w.sp = (int)h1.w;
            for (int j = 0; j < w.sp; j += 4) {
                const uint4 e = rec[2 + (j >> 2)];
                stack[j] = (int)e.x; stack[j + 1] = (int)e.y; stack[j + 2] = (int)e.z; stack[j + 3] = (int)e.w;
            }
        }
        const bool valid = idx != 0xFFFFFFFFu;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        if (valid) {
            const float4* src = (ext ? extRays : shadowRays) + RAY_STRIDE * (size_t)idx;
            r0 = qld(src); r1 = qld(src + 1);
        }
        const uint32_t validMask = __ballot_sync(0xFFFFFFFFu, valid);
        if (lane == 0) liveCount[warp] = (uint32_t)__popc(validMask);
        __syncwarp();
        bool suspended = false;
        if (valid) {
            SpaceRay ray; ray.set(mk3(r0.x, r0.y, r0.z), mk3(r1.x, r1.y, r1.z));
            if (ext) {
                if (PASS == 0) w.bestT = MS_INFINITY;
                suspended = walkResumable<false>(sc, ray, 0.0f, MS_INFINITY, w, stack, &liveCount[warp], thr);
            } else {
                if (PASS == 0) w.bestT = r0.w;
                suspended = walkResumable<true>(sc, ray, 0.0f, r0.w, w, stack, &liveCount[warp], thr);
            }
        }
        __syncwarp();
        if (valid && !suspended) {
            if (ext) {
                HitRec h; h.inst = w.inst; h.tri = w.tri; h.u = w.u; h.v = w.v;
                const uint32_t bin = hitBin(sc, h);
                qst(pb.H + idx, make_uint4(h.inst, h.tri | (bin << H_BIN_SHIFT), __float_as_uint(h.u), __float_as_uint(h.v)));
            } else if (w.inst != MISS_ID) {
                const uint32_t target = __float_as_uint(r1.w);
                float4* n = (target & 1u) ? pb.N1[buf] : pb.N0[buf];
                n[target >> 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // ---- the lanes that stopped: one reservation per warp, one record per lane
        const uint32_t susMask = __ballot_sync(0xFFFFFFFFu, suspended);
        if (susMask) {
            uint32_t rb = 0;
            if (lane == 0) rb = atomicAdd(outCount, (uint32_t)__popc(susMask));
            rb = __shfl_sync(0xFFFFFFFFu, rb, 0);
            if (suspended) {
                uint4* rec = outRec + (size_t)(rb + __popc(susMask & ltMask)) * RS_REC_U4;
                rec[0] = make_uint4(idx | (ext ? 0x80000000u : 0u), __float_as_uint(w.bestT), w.inst, w.tri);
                rec[1] = make_uint4(__float_as_uint(w.u), __float_as_uint(w.v), (uint32_t)w.node, (uint32_t)w.sp);
                for (int j = 0; j < w.sp; j += 4) rec[2 + (j >> 2)] = make_uint4((uint32_t)stack[j], (uint32_t)stack[j + 1], (uint32_t)stack[j + 2], (uint32_t)stack[j + 3]);
            }
        }
    }
    // last block out: re-arm the ticket; pass 0 publishes the ray counts, passes 1 and 2 clear the record count they consumed
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(ws.traceDone, 1u);
        if (done == gridDim.x - 1) {
            *ws.traceTicket = 0; *ws.traceDone = 0;
            if (PASS == 0) { ws.rayCounts[0] += nLive; ws.rayCounts[1] += nShadow; }
            else cb.counts[PASS - 1] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ray reordering for incoherent bounces (runtime.cpp: only queues above the sort threshold)
// ------------------------------------------------------------------------------------------------
// 16-bit key of a queued ray: [14:12] direction octant, [11:0] origin cell (z, y, x cell coordinates concatenated; thousands of
// rays share a key, so the order BETWEEN cells is immaterial). Only the order in which rays are traced depends on it - never
// a result. Written by the shade kernels next to the ray itself (2 B per ray instead of a pass that re-reads 32 B per ray).
DI uint16_t rayKey(const SortFrame& f, float4 o, float4 d) {
    const uint32_t cx = (uint32_t)fminf(fmaxf((o.x - f.lo[0]) * f.scale[0], 0.0f), (float)((1u << f.bits[0]) - 1u));
    const uint32_t cy = (uint32_t)fminf(fmaxf((o.y - f.lo[1]) * f.scale[1], 0.0f), (float)((1u << f.bits[1]) - 1u));
    const uint32_t cz = (uint32_t)fminf(fmaxf((o.z - f.lo[2]) * f.scale[2], 0.0f), (float)((1u << f.bits[2]) - 1u));
    const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    const uint32_t cell = (((cz << f.bits[1]) | cy) << f.bits[0]) | cx;
#if MSNE_SORT_KEY == 1
    return (uint16_t)((cell << 3) | oct);
#else
    return (uint16_t)((oct << 12) | cell);
#endif
}
// ------------------------------------------------------------------------------------------------
// Counting sort of the 15-bit ray keys -> permutation (queue indices grouped by key). Hand-written instead of a library radix
// sort because (a) it takes its item counts from DEVICE memory (the counters the shade pass just wrote), so a bounce needs no
// host round trip, (b) the values are the identity, so nothing but 2-byte keys is read, and (c) the order inside a key - and in
// fact any order - is immaterial to the results: only traversal coherence depends on it, so the scatter can use atomics.
//   k_sort_count    every block histograms its contiguous slice of the keys in shared memory (2^15 bins, 128 KB; lanes with equal
//                   keys are aggregated with __match_any_sync first: coherent rays share keys) and adds its non-zero bins to the
//                   global histogram
//   k_sort_scan     exclusive prefix sum of the global histogram -> global bin cursors; the histogram is re-zeroed for the next sort
//   k_sort_scatter  every block reserves its run in every bin (its own histogram, kept by the count pass) with one atomic on the
//                   bin cursor, then writes the queue indices of its slice into its runs
// blockIdx.y = 0: extension rays of `buf`, 1: shadow rays. Keys are read twice (4 B per ray), the permutation written once.
// ------------------------------------------------------------------------------------------------
static constexpr uint32_t SORT_BINS = 1u << 15;
static constexpr uint32_t SORT_THREADS = 1024;
struct SortJob {
    const uint16_t* keys[2]; uint32_t* perm[2];
    uint32_t* hist;      // [2][SORT_BINS], zero between sorts
    uint32_t* cursor;    // [2][SORT_BINS]
    uint32_t* blockHist; // [2][gridDim.x][SORT_BINS]: every block's histogram of its slice, kept from the count to the scatter pass
    const Counters* counters;   // this bounce's input counts
};
DI void sortSlice(const SortJob& j, uint32_t& lo, uint32_t& hi) {
    const uint32_t n = blockIdx.y == 0 ? j.counters->nLive : j.counters->nShadow;
    const uint32_t per = ((n + gridDim.x - 1) / gridDim.x + 31u) & ~31u;
    const uint64_t l = (uint64_t)blockIdx.x * per;
    lo = (uint32_t)(l < n ? l : n);
    hi = (uint32_t)(l + per < n ? l + per : n);
}
// shared histogram of keys[lo, hi): every lane of a warp stays in the loop until the whole warp is done (full-mask match)
DI void sortCountSlice(const uint16_t* __restrict__ keys, uint32_t lo, uint32_t hi, uint32_t* hist) {
    for (uint32_t b = threadIdx.x; b < SORT_BINS; b += SORT_THREADS) hist[b] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t base = lo + (threadIdx.x & ~31u); base < hi; base += SORT_THREADS) {
        const uint32_t i = base + lane;
        const uint32_t key = i < hi ? (uint32_t)__ldg(keys + i) & (SORT_BINS - 1u) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        if (i < hi && (int)lane == __ffs(peers) - 1) atomicAdd(hist + key, (uint32_t)__popc(peers));
    }
    __syncthreads();
}
__global__ void __launch_bounds__(SORT_THREADS, 1) k_sort_count(SortJob j) {
    extern __shared__ uint32_t sortHist[];
    uint32_t lo, hi;
    sortSlice(j, lo, hi);
    if (lo >= hi) return;
    sortCountSlice(j.keys[blockIdx.y], lo, hi, sortHist);
    uint32_t* g = j.hist + blockIdx.y * SORT_BINS;
    uint32_t* mine = j.blockHist + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * SORT_BINS;
    for (uint32_t b = threadIdx.x; b < SORT_BINS; b += SORT_THREADS) { const uint32_t c = sortHist[b]; mine[b] = c; if (c) atomicAdd(g + b, c); }
}
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scan(SortJob j) {
    // one block per queue; warp w owns bins [1024 w, 1024 w + 1024) and walks them 32 at a time (coalesced)
    __shared__ uint32_t warpBase[32];
    uint32_t* g = j.hist + blockIdx.x * SORT_BINS;
    uint32_t* cur = j.cursor + blockIdx.x * SORT_BINS;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t v[32], total = 0;
#pragma unroll
    for (uint32_t p = 0; p < 32; p++) { v[p] = g[warp * 1024u + p * 32u + lane]; total += v[p]; }
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
    if (lane == 0) warpBase[warp] = total;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warpBase[lane];
        uint32_t wi = w;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o); if ((int)lane >= o) wi += t; }
        warpBase[lane] = wi - w;
    }
    __syncthreads();
    uint32_t carry = warpBase[warp];
#pragma unroll
    for (uint32_t p = 0; p < 32; p++) {
        uint32_t incl = v[p];
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += t; }
        const uint32_t idx = warp * 1024u + p * 32u + lane;
        cur[idx] = carry + incl - v[p];
        g[idx] = 0;
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
}
__global__ void __launch_bounds__(SORT_THREADS, 1) k_sort_scatter(SortJob j) {
    extern __shared__ uint32_t sortHist[];
    uint32_t lo, hi;
    sortSlice(j, lo, hi);
    if (lo >= hi) return;
    const uint16_t* __restrict__ keys = j.keys[blockIdx.y];
    uint32_t* cur = j.cursor + blockIdx.y * SORT_BINS;
    const uint32_t* mine = j.blockHist + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * SORT_BINS;
    for (uint32_t b = threadIdx.x; b < SORT_BINS; b += SORT_THREADS) { const uint32_t c = mine[b]; sortHist[b] = c ? atomicAdd(cur + b, c) : 0u; }
    __syncthreads();
    uint32_t* __restrict__ perm = j.perm[blockIdx.y];
    const uint32_t lane = threadIdx.x & 31u, ltMask = (1u << lane) - 1u;
    for (uint32_t base = lo + (threadIdx.x & ~31u); base < hi; base += SORT_THREADS) {
        const uint32_t i = base + lane;
        const uint32_t key = i < hi ? (uint32_t)__ldg(keys + i) & (SORT_BINS - 1u) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        const int leader = __ffs(peers) - 1;
        uint32_t pos = 0;
        if (i < hi && (int)lane == leader) pos = atomicAdd(sortHist + key, (uint32_t)__popc(peers));
        pos = __shfl_sync(0xFFFFFFFFu, pos, leader) + (uint32_t)__popc(peers & ltMask);
        if (i < hi) perm[pos] = i;
    }
}
// Material binning in queue order for sorted bounces (the epilogue of the unsorted k_trace as a streaming pass over the bins
// k_trace left in H.y). A warp owns 512 consecutive slots: it counts its bins first, reserves each bin's range with ONE atomic
// (per-warp-iteration atomics on the five counters made this pass contention-bound), then writes the slots in queue order.
__global__ void __launch_bounds__(256) k_bin(PathBuffers pb, WaveState ws, uint32_t buf) {
    const uint32_t nLive = ws.counters[buf].nLive;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5), warpId = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    constexpr uint32_t ITERS = 16, CHUNK = 32 * ITERS;
    for (uint64_t chunk = (uint64_t)warpId * CHUNK; chunk < nLive; chunk += (uint64_t)warpsTotal * CHUNK) {
        unsigned long long packed = 0;   // this lane's 16 bins, 4 bits each (0xF = past the end)
        uint32_t cnt[NUM_BINS];
#pragma unroll
        for (uint32_t b = 0; b < NUM_BINS; b++) cnt[b] = 0;
#pragma unroll
        for (uint32_t it = 0; it < ITERS; it++) {
            const uint64_t idx = chunk + it * 32u + lane;
            const uint32_t bin = idx < nLive ? (qld(&pb.H[idx].y) >> H_BIN_SHIFT) : 0xFu;
            packed |= (unsigned long long)bin << (4u * it);
#pragma unroll
            for (uint32_t b = 0; b < NUM_BINS; b++) cnt[b] += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, bin == b));
        }
        uint32_t mine = 0;
#pragma unroll
        for (uint32_t b = 0; b < NUM_BINS; b++) if (lane == b) mine = cnt[b];
        uint32_t reserved = 0;
        if (lane < NUM_BINS && mine) reserved = atomicAdd(ws.typeCount + lane, mine);
        uint32_t next[NUM_BINS];
#pragma unroll
        for (uint32_t b = 0; b < NUM_BINS; b++) next[b] = __shfl_sync(0xFFFFFFFFu, reserved, (int)b);
#pragma unroll
        for (uint32_t it = 0; it < ITERS; it++) {
            const uint32_t bin = (uint32_t)(packed >> (4u * it)) & 0xFu;
            uint32_t slot = 0;
#pragma unroll
            for (uint32_t b = 0; b < NUM_BINS; b++) {
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, bin == b);
                if (bin == b) slot = next[b] + (uint32_t)__popc(m & ltMask);
                next[b] += (uint32_t)__popc(m);
            }
            if (bin != 0xFu) qst(pb.typeList + ((size_t)bin * pb.capacity + slot), (uint32_t)(chunk + it * 32u + lane));
        }
    }
}

// debug probe: arbitrary rays (HdMoonshineExtTraceRays); rays8 = {o, tmin, d, tmax}
__global__ void k_trace_probe(DeviceScene sc, const float* __restrict__ rays8, uint32_t n, int anyHit, uint32_t* hits5) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = rays8 + 8 * (size_t)i;
    HitRec h;
    uint32_t geom = 0, prim = 0;
    // the renderer only ever traces [0, INFINITY = 1e12] (math.hlsl:5); a caller's +inf would make the slab test accept the
    // unreachable point box of an empty child (dtypes.h)
    const float tmin = fmaxf(r[3], 0.0f), tmax = fminf(r[7], MS_INFINITY);
    if (anyHit) traceRay<true, true>(sc, mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), tmin, tmax, h);
    else {
        traceRay<false, true>(sc, mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), tmin, tmax, h);
        if (h.inst != MISS_ID) { geom = sc.tris[h.tri].geom; prim = sc.tris[h.tri].prim; }
    }
    uint32_t* o = hits5 + 5 * (size_t)i;
    o[0] = h.inst; o[1] = geom; o[2] = prim; o[3] = __float_as_uint(h.u); o[4] = __float_as_uint(h.v);
}

// debug probes: the BSDF and environment-map device functions on caller-supplied inputs (HdMoonshineExtBsdfProbe / EnvProbe),
// so the tests can pin material.hlsl / light.hlsl function by function against the oracle instead of only through images.
// query = {type, color rgb, metalness, roughness, ior, op, a.xyz, w_o.xyz}: op 0 eval(a, w_o), 1 pdf(a, w_o), 2 sample(w_o, square = a.xy)
__global__ void k_bsdf_probe(const float* __restrict__ q14, uint32_t n, float* out4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = q14 + 14 * (size_t)i;
    MaterialParams m;
    m.type = (uint32_t)q[0]; m.color = mk3(q[1], q[2], q[3]); m.metalness = q[4];
    m.alpha = fmaxf(q[5] * q[5], 0.001f);   // StandardPBR::load (material.hlsl:196)
    m.ior = q[6];
    const int op = (int)q[7];
    const float3 a = mk3(q[8], q[9], q[10]), w_o = mk3(q[11], q[12], q[13]);
    prepareMaterial(m, w_o);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (op == 0) { const float3 e = matEval(m, a, w_o); r = make_float4(e.x, e.y, e.z, 0.f); }
    else if (op == 1) r.x = matPdf(m, a, w_o);
    else { const MatSample sm = matSample(m, w_o, mk2(a.x, a.y)); r = make_float4(sm.dirFs.x, sm.dirFs.y, sm.dirFs.z, sm.pdf); }
    reinterpret_cast<float4*>(out4)[i] = r;
}
// the defined transcendentals (device_math.cuh): fn 0 sin(a), 1 cos(a), 2 log(a), 3 atan2(a, b), 4 acos(a)
__global__ void k_math_probe(int fn, const float* __restrict__ a, const float* __restrict__ b, uint32_t n, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r;
    switch (fn) {
        case 0: r = sinDef(a[i]); break;
        case 1: r = cosDef(a[i]); break;
        case 2: r = logDef(a[i]); break;
        case 3: r = atan2Def(a[i], b[i]); break;
        default: r = acosDef(a[i]); break;
    }
    out[i] = r;
}
// op 0: EnvMap::sample(rand = in.xy) -> {dir, radiance, pdf}; 1: EnvMap::eval(dir) -> {radiance, pdf}; 2: incomingRadiance(dir) -> {rgb}
__global__ void k_env_probe(DeviceScene sc, int op, const float* __restrict__ in3, uint32_t n, float* out7) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = in3 + 3 * (size_t)i;
    float* r = out7 + 7 * (size_t)i;
    for (int k = 0; k < 7; k++) r[k] = 0.0f;
    if (op == 0) {
        const LightSample ls = envSample(sc, mk2(a[0], a[1]));
        r[0] = ls.dirWs.x; r[1] = ls.dirWs.y; r[2] = ls.dirWs.z; r[3] = ls.radiance.x; r[4] = ls.radiance.y; r[5] = ls.radiance.z; r[6] = ls.pdf;
    } else if (op == 1) {
        float3 rad; float pdf;
        envEval(sc, mk3(a[0], a[1], a[2]), rad, pdf);
        r[0] = rad.x; r[1] = rad.y; r[2] = rad.z; r[3] = pdf;
    } else {
        const float3 c = envIncomingRadiance(sc, mk3(a[0], a[1], a[2]));
        r[0] = c.x; r[1] = c.y; r[2] = c.z;
    }
}

// object picker (shaders/hrtsystem/input.hlsl:24-57): one closest-hit query through normalized sensor coordinates
// with lens sample (0,0). The HLSL edits a local copy of the camera (perfect focus) but then generates the ray
// from the unmodified push constant (input.hlsl:28-32) - restated as written, i.e. the lens is used as is.
__global__ void k_pick(DeviceScene sc, DCamera cam, float2 coords, MsneClick* out) {
    float2 uv = coords;
    uv.y -= 1.0f;
    uv.y *= -1.0f;
    float3 o, d;
    cameraRayUV(cam, uv, mk2(0.0f, 0.0f), o, d);
    HitRec h;
    traceRay<false, true>(sc, o, d, 0.0f, MS_INFINITY, h);
    MsneClick c;
    c.instance_index = -1; c.geometry_index = 0; c.primitive_index = 0; c.bary[0] = 0.0f; c.bary[1] = 0.0f;
    if (h.inst != MISS_ID) {
        c.instance_index = (int32_t)h.inst; c.geometry_index = sc.tris[h.tri].geom; c.primitive_index = sc.tris[h.tri].prim;
        c.bary[0] = h.u; c.bary[1] = h.v;
    }
    *out = c;
}

// ------------------------------------------------------------------------------------------------
// shade (one bounce of PathTracingIntegrator::incomingRadiance, integrator.hlsl:68-183)
// ------------------------------------------------------------------------------------------------
struct ShadeOut {
    int kind;               // 0 = finished (result written), 1 = live, 2 = dead-with-pending-NEE
    float4 R0, R1, S0, S1, N0, N1;
    bool envRay, meshRay;
    float4 envQ0, envQ1, meshQ0, meshQ1;   // Q1.w (target) filled in after the slot is known
    uint32_t inlineShadow;                  // INLINE_NEE: shadow rays traced inside shadeOne
};

// One light sample of estimateDirectMISLight (integrator.hlsl:20-35) up to, but not including, the occlusion test: the
// contribution throughput * L * f * |cos| * w / p_light / n as the integrator adds it when the shadow ray finds nothing, and that
// shadow ray. Returns false when the sample contributes nothing whatever the ray finds (light pdf or BSDF response zero: no ray).
DI bool neeEnvSample(const DeviceScene& sc, const DConstants& k, const MaterialParams& material, const Frame& shadingFrame, float3 outgoingDirSs,
                     float3 position, float3 triangleNormal, float3 throughput, float2 rand, float3& contribution, float4& q0, float4& q1) {
    // EnvMap::sample (light.hlsl:47-80)
    const LightSample ls = envSample(sc, rand);
    if (!(ls.pdf > 0.0f)) return false;
    const float3 lightDirFs = shadingFrame.worldToFrame(ls.dirWs);
    const float scatteringPdf = matPdf(material, lightDirFs, outgoingDirSs);
    if (!(scatteringPdf > 0.0f)) return false;
    const float3 brdf = matEval(material, lightDirFs, outgoingDirSs);
    const float weight = powerHeuristic(k.env_samples, ls.pdf, 1, scatteringPdf);
    const float3 est = ls.radiance * brdf * fabsf(cosTheta(lightDirFs)) * weight / ls.pdf;
    contribution = (throughput * est) / (float)k.env_samples;
    const float3 so = offsetAlongNormal(position, faceForward(triangleNormal, ls.dirWs));
    q0 = make_float4(so.x, so.y, so.z, MS_INFINITY);
    q1 = make_float4(ls.dirWs.x, ls.dirWs.y, ls.dirWs.z, 0.0f);
    return true;
}
DI bool neeMeshSample(const DeviceScene& sc, const DConstants& k, const MaterialParams& material, const Frame& shadingFrame, float3 outgoingDirSs,
                      float3 position, float3 triangleNormal, float3 throughput, float2 rand, float3& contribution, float4& q0, float4& q1) {
    // MeshLights::sample (light.hlsl:130-158)
    if (!(sc.aliasEntries > 1)) return false;
    const uint32_t entryCount = __ldg(&sc.alias[0].alias);
    const float sum = __ldg(&sc.alias[0].select);
    if (entryCount == 0 || sum == 0) return false;
    // sampleAlias (mappings.hlsl:114-126), offset 1
    const float scaled = rand.x * (float)entryCount;
    uint32_t idx = (uint32_t)scaled;
    if (idx >= entryCount) idx = entryCount - 1;   // hazard: float rounding past the table end
    rand.x = scaled - floorf(scaled);
    DAlias e = sc.alias[1 + idx];
    if (!coinFlipRemap(e.select, rand.x)) { idx = e.alias; e = sc.alias[1 + idx]; }
    const uint32_t lInstanceID = __ldg(&sc.instances[e.inst].customIndex);
    const DGeometry lg = sc.geometries[lInstanceID + e.geom];
    const float2 lbary = squareToTriangle(rand);
#if MSNE_EMITTER_RECORDS
    const LightPoint lp = emitterPoint(sc, idx, e.inst, lbary);
    struct { float3 position; float2 texcoord; struct { float3 n; } triangleFrame; } la = {lp.position, lp.texcoord, {lp.n}};
#else
    const MeshAttributes la = lookupAttributes(sc, k, e.inst, lg.mesh, e.prim, lbary);
#endif
    const DMaterial* lm = sc.materials + lg.material;
    float3 radiance;
    if (MSNE_MATERIAL_CONSTANTS && (__ldg(&lm->constMask) >> MAT_SLOT_EMISSIVE) & 1u)
        radiance = mk3(__ldg(&lm->constant[MAT_SLOT_EMISSIVE][0]), __ldg(&lm->constant[MAT_SLOT_EMISSIVE][1]), __ldg(&lm->constant[MAT_SLOT_EMISSIVE][2]));
    else radiance = xyz(sampleTexture(sc, __ldg(&lm->emissive), la.texcoord));
    const float3 dirWs = normalize(la.position - position);
    const float pdf = areaMeasureToSolidAngleMeasure(la.position, position, dirWs, la.triangleFrame.n) / sum;
    if (!(pdf > 0.0f)) return false;
    const float3 lightDirFs = shadingFrame.worldToFrame(dirWs);
    const float scatteringPdf = matPdf(material, lightDirFs, outgoingDirSs);
    if (!(scatteringPdf > 0.0f)) return false;
    const float3 brdf = matEval(material, lightDirFs, outgoingDirSs);
    const float weight = powerHeuristic(k.mesh_samples, pdf, 1, scatteringPdf);
    const float3 est = radiance * brdf * fabsf(cosTheta(lightDirFs)) * weight / pdf;
    contribution = (throughput * est) / (float)k.mesh_samples;
    const float3 offL = offsetAlongNormal(la.position, la.triangleFrame.n);
    const float3 offS = offsetAlongNormal(position, faceForward(triangleNormal, dirWs));
    const float tmax = length(offL - offS);
    const float3 sd = normalize(offL - offS);
    q0 = make_float4(offS.x, offS.y, offS.z, tmax);
    q1 = make_float4(sd.x, sd.y, sd.z, 0.0f);
    return true;
}

// MT: -1 = any (runtime dispatch), 0..3 = MaterialType known at compile time, 4 = miss / dead paths only
// INLINE_NEE = false: at most one env and one mesh light sample per bounce (the offline and Hydra constants); their shadow rays are
// queued (out.envRay / out.meshRay) and the contributions wait in N0 / N1 for the next trace pass.
// INLINE_NEE = true: any env_samples_per_bounce / mesh_samples_per_bounce (integrator.hlsl:141-150 loops over them): every light
// sample's shadow ray is traced right here and its contribution added in the reference's order, so nothing is pending afterwards.
template <int MT, bool INLINE_NEE = false>
DI void shadeOne(const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, uint32_t buf, uint32_t q, ShadeOut& out, uint32_t& bounceOut) {
    const float4 R0 = qld(pb.R0[buf] + RAY_STRIDE * (size_t)q), R1 = qld(pb.R1[buf] + RAY_STRIDE * (size_t)q), S0 = qld(pb.S0[buf] + q), S1 = qld(pb.S1[buf] + q);
    const float4 N0 = qld(pb.N0[buf] + q), N1 = qld(pb.N1[buf] + q);
    const uint32_t flags = __float_as_uint(R1.w);
    const uint32_t bounceCount = flags & BOUNCE_MASK;
    const uint32_t slot = __float_as_uint(S1.w);
    bounceOut = bounceCount;
    out.kind = 0; out.envRay = false; out.meshRay = false; out.inlineShadow = 0;

    float3 acc = mk3(S1.x, S1.y, S1.z);
    // NEE of the previous bounce, in the reference's order: env first, then mesh (integrator.hlsl:141-150)
    acc = acc + mk3(N0.x, N0.y, N0.z);
    acc = acc + mk3(N1.x, N1.y, N1.z);

    if (MT < 0 || MT == (int)BIN_MISS) {
        if (flags & FLAG_DEAD) { pb.result[slot] = make_float4(acc.x, acc.y, acc.z, 1.0f); return; }
    }

    const float3 rayOrigin = mk3(R0.x, R0.y, R0.z), rayDir = mk3(R1.x, R1.y, R1.z);
    const float lastMaterialPdf = R0.w;
    const bool isLastMaterialDelta = (flags & FLAG_DELTA) != 0;
    float3 throughput = mk3(S0.x, S0.y, S0.z);
    Rng rng; rng.state = __float_as_uint(S0.w);

    const uint4 H = qld(pb.H + q);
    if ((MT < 0 || MT == (int)BIN_MISS) && H.x == MISS_ID) {
        // integrator.hlsl:169-180
        if (k.env_samples == 0 || bounceCount == 0 || isLastMaterialDelta) {
            acc = acc + throughput * envIncomingRadiance(sc, rayDir);
        } else {
            float3 rad; float pdf;
            envEval(sc, rayDir, rad, pdf);
            if (pdf > 0.0f) {
                const float weight = powerHeuristic(1, lastMaterialPdf, k.env_samples, pdf);
                acc = acc + throughput * rad * weight;
            }
        }
        pb.result[slot] = make_float4(acc.x, acc.y, acc.z, 1.0f);
        return;
    }
    if (MT == (int)BIN_MISS) return;   // the miss bin holds only misses and dead paths

    // decode hit (integrator.hlsl:82-88)
    const uint32_t instanceIndex = H.x;
    const TriRec* tr = sc.tris + (H.y & H_TRI_MASK);
    const uint32_t geometryIndex = __ldg(&tr->geom), primitiveIndex = __ldg(&tr->prim);
    const float2 bary = mk2(__uint_as_float(H.z), __uint_as_float(H.w));
    const uint32_t instanceID = __ldg(&sc.instances[instanceIndex].customIndex);
    const DGeometry geometry = sc.geometries[instanceID + geometryIndex];
    const bool geomSampled = (geometry.sampled & 0xFFu) != 0;   // only the low byte is defined (Accel.zig:40-44)
    MeshAttributes attrs = lookupAttributes<true>(sc, k, instanceIndex, geometry.mesh, primitiveIndex, bary, tr, H.y & H_TRI_MASK);
    const DMaterial& dm = sc.materials[geometry.material];   // read in place (112 B: header + resolved constants), field by field
    const Frame textureFrame = getTextureFrame(sc, k, dm, attrs.texcoord, attrs.frame);
    const float3 emissiveLight = xyz(sampleMaterialSlot(sc, dm, MAT_SLOT_EMISSIVE, dm.emissive, attrs.texcoord));
    MaterialParams material = loadMaterial(sc, dm, attrs.texcoord);
    if (MT >= 0) material.type = (uint32_t)MT;   // compile-time constant: the other variants' code is pruned

    const float3 outgoingDirWs = -rayDir;

    // shading-frame selection (integrator.hlsl:93-104)
    const bool frontfacing = dot(attrs.triangleFrame.n, outgoingDirWs) > 0;
    Frame shadingFrame;
    if ((frontfacing && dot(outgoingDirWs, textureFrame.n) > 0) || (!frontfacing && -dot(outgoingDirWs, textureFrame.n) > 0)) shadingFrame = textureFrame;
    else if ((frontfacing && dot(outgoingDirWs, attrs.frame.n) > 0) || (!frontfacing && -dot(outgoingDirWs, attrs.frame.n) > 0)) shadingFrame = attrs.frame;
    else { completeTriangleFrame(sc, attrs, instanceIndex, H.y & H_TRI_MASK); shadingFrame = attrs.triangleFrame; }
    const float3 outgoingDirSs = shadingFrame.worldToFrame(outgoingDirWs);

    // emission (integrator.hlsl:109-124)
    if (k.mesh_samples == 0 || bounceCount == 0 || !geomSampled || isLastMaterialDelta) {
        if (dot(outgoingDirWs, attrs.triangleFrame.n) > 0.0f) acc = acc + throughput * emissiveLight;
    } else {
        const float sum = __ldg(&sc.alias[0].select);
        const float lightPdf = areaMeasureToSolidAngleMeasure(attrs.position, rayOrigin, rayDir, attrs.triangleFrame.n) / sum;
        if (lightPdf > 0.0f) {
            const float weight = powerHeuristic(1, lastMaterialPdf, k.mesh_samples, lightPdf);
            acc = acc + throughput * emissiveLight * weight;
        }
    }

    // termination (integrator.hlsl:128-135)
    if (bounceCount >= k.max_bounces + 1) { pb.result[slot] = make_float4(acc.x, acc.y, acc.z, 1.0f); return; }
    if (bounceCount > 3) {
        const float pSurvive = fminf(0.95f, luminance(throughput));
        if (rng.getFloat() > pSurvive) { pb.result[slot] = make_float4(acc.x, acc.y, acc.z, 1.0f); return; }
        throughput = throughput / pSurvive;
    }

    prepareMaterial(material, outgoingDirSs);
    const bool isCurrentMaterialDelta = matIsDelta(material.type);
    float3 neeEnv = mk3(0.0f), neeMesh = mk3(0.0f);
    if (!isCurrentMaterialDelta) {
        if (INLINE_NEE) {
            // integrator.hlsl:141-150 with any sample counts: draw, sample, occlusion test, add - in that order
            float3 c; float4 q0, q1; HitRec sh;
            for (uint32_t i = 0; i < k.env_samples; i++) {
                const float2 rand = rng.getFloat2();
                if (!neeEnvSample(sc, k, material, shadingFrame, outgoingDirSs, attrs.position, attrs.triangleFrame.n, throughput, rand, c, q0, q1)) continue;
                out.inlineShadow++;
                if (!traceRay<true, true>(sc, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), 0.0f, q0.w, sh)) acc = acc + c;
            }
            for (uint32_t i = 0; i < k.mesh_samples; i++) {
                const float2 rand = rng.getFloat2();
                if (!neeMeshSample(sc, k, material, shadingFrame, outgoingDirSs, attrs.position, attrs.triangleFrame.n, throughput, rand, c, q0, q1)) continue;
                out.inlineShadow++;
                if (!traceRay<true, true>(sc, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), 0.0f, q0.w, sh)) acc = acc + c;
            }
        } else {
            // env NEE (integrator.hlsl:141-144 -> estimateDirectMISLight :20-35 -> EnvMap::sample light.hlsl:47-80)
            if (k.env_samples) {
                const float2 rand = rng.getFloat2();
                out.envRay = neeEnvSample(sc, k, material, shadingFrame, outgoingDirSs, attrs.position, attrs.triangleFrame.n, throughput, rand, neeEnv, out.envQ0, out.envQ1);
                if (!out.envRay) neeEnv = mk3(0.0f);
            }
            // mesh NEE (integrator.hlsl:147-150 -> MeshLights::sample light.hlsl:130-158)
            if (k.mesh_samples) {
                const float2 rand = rng.getFloat2();
                out.meshRay = neeMeshSample(sc, k, material, shadingFrame, outgoingDirSs, attrs.position, attrs.triangleFrame.n, throughput, rand, neeMesh, out.meshQ0, out.meshQ1);
                if (!out.meshRay) neeMesh = mk3(0.0f);
            }
        }
    }

    // BSDF sample (integrator.hlsl:154-163)
    const MatSample sample = matSample(material, outgoingDirSs, rng.getFloat2());
    out.N0 = make_float4(neeEnv.x, neeEnv.y, neeEnv.z, 0.0f);
    out.N1 = make_float4(neeMesh.x, neeMesh.y, neeMesh.z, 0.0f);
    out.S1 = make_float4(acc.x, acc.y, acc.z, __uint_as_float(slot));
    if (sample.pdf == 0.0f) {
        if (!out.envRay && !out.meshRay) { pb.result[slot] = make_float4(acc.x, acc.y, acc.z, 1.0f); return; }
        out.kind = 2;
        out.R0 = make_float4(0.f, 0.f, 0.f, 0.f);
        out.R1 = make_float4(0.f, 0.f, 0.f, __uint_as_float(FLAG_DEAD | bounceCount));
        out.S0 = make_float4(throughput.x, throughput.y, throughput.z, __uint_as_float(rng.state));
        return;
    }
    const float3 newDir = shadingFrame.frameToWorld(sample.dirFs);
    const float3 newOrigin = offsetAlongNormal(attrs.position, faceForward(attrs.triangleFrame.n, newDir));
    throughput = throughput * (matEval(material, sample.dirFs, outgoingDirSs) * fabsf(cosTheta(sample.dirFs)) / sample.pdf);
    out.kind = 1;
    out.R0 = make_float4(newOrigin.x, newOrigin.y, newOrigin.z, sample.pdf);
    out.R1 = make_float4(newDir.x, newDir.y, newDir.z, __uint_as_float((bounceCount + 1) | (isCurrentMaterialDelta ? FLAG_DELTA : 0u)));
    out.S0 = make_float4(throughput.x, throughput.y, throughput.z, __uint_as_float(rng.state));
}

// Shared body of the shade kernels. BIN = -1: every queue slot [0, nLive) + the dead range, any material (used for
// small queues, one launch per bounce). BIN = 0..4: only the slots the trace kernel binned under that material type
// (4 = misses + the dead range); five such launches cover a bounce with material-specialised, convergent code.
template <int BIN, int MIN_BLOCKS, bool INLINE_NEE = false>
__global__ void __launch_bounds__(SHADE_THREADS, MIN_BLOCKS) k_shade(DeviceScene sc, DConstants k, PathBuffers pb, WaveState ws, uint32_t buf, int32_t historySlot, SortKeys keys) {
    const Counters cin = ws.counters[buf];
    Counters* cout = ws.counters + (buf ^ 1u);
    const uint32_t P = pb.capacity;
    const uint32_t nList = BIN < 0 ? cin.nLive : ws.typeCount[BIN];
    const uint32_t nDeadHere = (BIN < 0 || BIN == (int)BIN_MISS) ? cin.nDead : 0u;
    const uint32_t total = nList + nDeadHere;
    const uint32_t* list = BIN < 0 ? nullptr : pb.typeList + (size_t)BIN * P;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t ob = buf ^ 1u;
    uint32_t localMaxBounce = 0;
    uint32_t nInline = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    // every lane of a warp runs the same number of iterations so the ballots below are convergent
    // queue slot of grid-stride item i: list entry, or the dead range filled from the top of the buffer
    auto slotOf = [&](uint32_t i) -> uint32_t { return i < nList ? (BIN < 0 ? i : qld(list + i)) : (P - 1u - (i - nList)); };
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < total; base += stride) {
        const uint32_t i = base + lane;
        ShadeOut out;
        out.kind = 0; out.envRay = false; out.meshRay = false; out.inlineShadow = 0;
        if (i < total) {
            const uint32_t q = slotOf(i);
            uint32_t b;
            shadeOne<BIN, INLINE_NEE>(sc, k, pb, buf, q, out, b);
            localMaxBounce = max(localMaxBounce, b);
            if (INLINE_NEE) nInline += out.inlineShadow;
        }
        // ---- warp-ballot compaction into the other ping-pong buffer -----------------------------
        const uint32_t liveMask = __ballot_sync(0xFFFFFFFFu, out.kind == 1);
        const uint32_t deadMask = __ballot_sync(0xFFFFFFFFu, out.kind == 2);
        const uint32_t envMask = __ballot_sync(0xFFFFFFFFu, out.envRay);
        const uint32_t meshMask = __ballot_sync(0xFFFFFFFFu, out.meshRay);
        uint32_t liveBase = 0, deadBase = 0, shadowBase = 0;
        if (lane == 0) {
            if (liveMask) liveBase = atomicAdd(&cout->nLive, (uint32_t)__popc(liveMask));
            if (deadMask) deadBase = atomicAdd(&cout->nDead, (uint32_t)__popc(deadMask));
            if (envMask | meshMask) shadowBase = atomicAdd(&cout->nShadow, (uint32_t)(__popc(envMask) + __popc(meshMask)));
        }
        liveBase = __shfl_sync(0xFFFFFFFFu, liveBase, 0);
        deadBase = __shfl_sync(0xFFFFFFFFu, deadBase, 0);
        shadowBase = __shfl_sync(0xFFFFFFFFu, shadowBase, 0);
        if (out.kind != 0) {
            const uint32_t oq = out.kind == 1 ? liveBase + __popc(liveMask & ltMask) : P - 1u - (deadBase + __popc(deadMask & ltMask));
            qst(pb.R0[ob] + RAY_STRIDE * (size_t)oq, out.R0); qst(pb.R1[ob] + RAY_STRIDE * (size_t)oq, out.R1); qst(pb.S0[ob] + oq, out.S0); qst(pb.S1[ob] + oq, out.S1);
            qst(pb.N0[ob] + oq, out.N0); qst(pb.N1[ob] + oq, out.N1);
            if (keys.ext != nullptr && out.kind == 1) keys.ext[oq] = rayKey(keys.frame, out.R0, out.R1);
            if (out.envRay) {
                const uint32_t si = shadowBase + __popc(envMask & ltMask);
                out.envQ1.w = __uint_as_float(oq * 2u + 0u);
                qst(pb.Q0 + RAY_STRIDE * (size_t)si, out.envQ0); qst(pb.Q1 + RAY_STRIDE * (size_t)si, out.envQ1);
                if (keys.ext != nullptr) keys.shadow[si] = rayKey(keys.frame, out.envQ0, out.envQ1);
            }
            if (out.meshRay) {
                const uint32_t si = shadowBase + __popc(envMask) + __popc(meshMask & ltMask);
                out.meshQ1.w = __uint_as_float(oq * 2u + 1u);
                qst(pb.Q0 + RAY_STRIDE * (size_t)si, out.meshQ0); qst(pb.Q1 + RAY_STRIDE * (size_t)si, out.meshQ1);
                if (keys.ext != nullptr) keys.shadow[si] = rayKey(keys.frame, out.meshQ0, out.meshQ1);
            }
        }
    }
    // last block out: clear what this launch consumed, publish the deepest bounce seen
    if (localMaxBounce) atomicMax(ws.maxBounce, localMaxBounce);
    if (INLINE_NEE) {
        for (int o = 16; o > 0; o >>= 1) nInline += __shfl_down_sync(0xFFFFFFFFu, nInline, o);
        if (lane == 0 && nInline) atomicAdd(ws.rayCounts + 1, (unsigned long long)nInline);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        uint32_t* doneCtr = BIN < 0 ? ws.shadeDone : ws.binDone + BIN;
        const uint32_t done = atomicAdd(doneCtr, 1u);
        if (done == gridDim.x - 1) {
            if (BIN < 0) { for (uint32_t b = 0; b < NUM_BINS; b++) ws.typeCount[b] = 0; }
            else ws.typeCount[BIN] = 0;
            // historySlot >= 0: this is the bounce's last shade launch - log what the bounce produced (the host reads the log after
            // the wave, never during it) and clear what it consumed
            if (historySlot >= 0) {
                ws.history[historySlot] = *cout;
                ws.counters[buf].nLive = 0; ws.counters[buf].nDead = 0; ws.counters[buf].nShadow = 0;
            }
            *doneCtr = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tail: once only a few thousand paths survive, every further bounce costs two launches at the latency floor of
// a single warp (~0.25 ms, ~130 times per wave with max_bounces = 1024). k_tail finishes those stragglers in ONE
// launch: each thread owns a queue slot and alternates shadeOne / shadow rays / closest hit in place until its
// path ends. Same device functions, same per-path order of operations -> bit-identical results.
// Precondition: k_trace has just run on `buf` (H holds this bounce's hits, pending shadow rays are resolved).
// ------------------------------------------------------------------------------------------------
template <bool INLINE_NEE>
__global__ void __launch_bounds__(128) k_tail(DeviceScene sc, DConstants k, PathBuffers pb, WaveState ws, uint32_t buf) {
    const Counters cin = ws.counters[buf];
    const uint32_t nLive = cin.nLive, total = cin.nLive + cin.nDead, P = pb.capacity;
    unsigned long long nClosest = 0, nShadow = 0;
    uint32_t localMaxBounce = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t q = i < nLive ? i : (P - 1u - (i - nLive));
        for (;;) {
            ShadeOut out;
            out.kind = 0; out.envRay = false; out.meshRay = false; out.inlineShadow = 0;
            uint32_t b;
            shadeOne<-1, INLINE_NEE>(sc, k, pb, buf, q, out, b);
            localMaxBounce = max(localMaxBounce, b);
            if (INLINE_NEE) nShadow += out.inlineShadow;
            if (out.kind == 0) break;
            HitRec h;
            if (out.envRay) {
                nShadow++;
                if (traceRay<true, true>(sc, mk3(out.envQ0.x, out.envQ0.y, out.envQ0.z), mk3(out.envQ1.x, out.envQ1.y, out.envQ1.z), 0.0f, out.envQ0.w, h))
                    out.N0 = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (out.meshRay) {
                nShadow++;
                if (traceRay<true, true>(sc, mk3(out.meshQ0.x, out.meshQ0.y, out.meshQ0.z), mk3(out.meshQ1.x, out.meshQ1.y, out.meshQ1.z), 0.0f, out.meshQ0.w, h))
                    out.N1 = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            pb.R0[buf][RAY_STRIDE * (size_t)q] = out.R0; pb.R1[buf][RAY_STRIDE * (size_t)q] = out.R1; pb.S0[buf][q] = out.S0; pb.S1[buf][q] = out.S1;
            pb.N0[buf][q] = out.N0; pb.N1[buf][q] = out.N1;
            if (out.kind == 1) {
                nClosest++;
                traceRay<false, true>(sc, mk3(out.R0.x, out.R0.y, out.R0.z), mk3(out.R1.x, out.R1.y, out.R1.z), 0.0f, MS_INFINITY, h);
                pb.H[q] = make_uint4(h.inst, h.tri, __float_as_uint(h.u), __float_as_uint(h.v));
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nClosest += __shfl_down_sync(0xFFFFFFFFu, nClosest, o);
        nShadow += __shfl_down_sync(0xFFFFFFFFu, nShadow, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nClosest) atomicAdd(ws.rayCounts + 0, nClosest);
        if (nShadow) atomicAdd(ws.rayCounts + 1, nShadow);
    }
    if (localMaxBounce) atomicMax(ws.maxBounce, localMaxBounce);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(ws.shadeDone, 1u);
        if (done == gridDim.x - 1) {
            for (uint32_t b = 0; b < NUM_BINS; b++) ws.typeCount[b] = 0;
            ws.counters[0].nLive = 0; ws.counters[0].nDead = 0; ws.counters[0].nShadow = 0;
            ws.counters[1].nLive = 0; ws.counters[1].nDead = 0; ws.counters[1].nShadow = 0;
            *ws.shadeDone = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// accumulate: raygen's sample loop sum + storeColor (main.hlsl:79-94, :43-51), in sample-index order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_accumulate(const float4* __restrict__ result, float4* image, float4* runState, const uint32_t* __restrict__ pixIds,
                                                    uint32_t nPix, uint32_t nSamples, uint32_t firstSample, uint32_t spr) {
    const uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= nPix) return;
    const uint32_t pix = __ldg(pixIds + pi);
    float4 rs = runState[pi];   // (sum.rgb, samples summed so far in the current launch)
    float3 color = mk3(rs.x, rs.y, rs.z);
    uint32_t cnt = __float_as_uint(rs.w);
    float4 texel = image[pix];
    for (uint32_t s = 0; s < nSamples; s++) {
        const float4 c = result[(size_t)s * nPix + pi];
        color = color + mk3(c.x, c.y, c.z);
        cnt++;
        if (cnt == spr) {
            const uint32_t pushSampleCount = (firstSample + s + 1u) - spr;   // sampleCount of the launch that just completed
            if (pushSampleCount == 0) {
                const float3 o = color / (float)spr;
                texel = make_float4(o.x, o.y, o.z, 1.0f);
            } else {
                const float3 prior = mk3(texel.x, texel.y, texel.z);
                const float3 dlt = (color - prior) / (float)(pushSampleCount + spr);
                texel = make_float4(texel.x + dlt.x, texel.y + dlt.y, texel.z + dlt.z, texel.w + 1.0f);
            }
            color = mk3(0.0f); cnt = 0;
        }
    }
    image[pix] = texel;
    runState[pi] = make_float4(color.x, color.y, color.z, __uint_as_float(cnt));
}

// ------------------------------------------------------------------------------------------------
// environment pre-process (shaders/background/{equirectangular_to_equal_area,luminance,fold}.hlsl)
// ------------------------------------------------------------------------------------------------
__global__ void k_env_resample(const float4* __restrict__ src, uint32_t W, uint32_t H, float4* dstRgb, float* dstLum, uint32_t S) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= S || y >= S) return;
    float3 color = mk3(0.0f);
    for (uint32_t i = 0; i < 3; i++)
        for (uint32_t j = 0; j < 3; j++) {
            const float2 subpixel = mk2((float)(1 + i), (float)(1 + j)) / mk2(4.0f, 4.0f);
            const float2 dstCoords = (mk2((float)x, (float)y) + subpixel) / mk2((float)S, (float)S);
            const float3 dir = squareToEqualAreaSphere(dstCoords);
            const float2 uv = cartesianToSpherical(dir) / mk2(2 * MS_PI, MS_PI);
            float4 t;
            if (W == 1 && H == 1) t = src[0];
            else {
                const float u = uv.x * (float)W - 0.5f, v = uv.y * (float)H - 0.5f;
                const float fu = floorf(u), fv = floorf(v);
                const float a = u - fu, b = v - fv;
                const int i0 = wrapMirror((int)fu, W), i1 = wrapMirror((int)fu + 1, W), j0 = wrapMirror((int)fv, H), j1 = wrapMirror((int)fv + 1, H);
                t = lerp4(lerp4(src[(size_t)j0 * W + i0], src[(size_t)j0 * W + i1], a), lerp4(src[(size_t)j1 * W + i0], src[(size_t)j1 * W + i1], a), b);
            }
            color = color + mk3(t.x, t.y, t.z);
        }
    const float3 c = color / 9.0f;
    dstRgb[(size_t)y * S + x] = make_float4(c.x, c.y, c.z, 1.0f);
    dstLum[(size_t)y * S + x] = luminance(c);
}
// fold.hlsl: one mip of the SUM pyramid; also emits the four source texels as one float4 (quad-major copy of the source mip
// that EnvMap::sample walks with a single 16-byte load per level, shading.cuh envSample)
__global__ void k_env_fold(const float* __restrict__ srcMip, float* dstMip, float4* srcQuads, uint32_t dstSize) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dstSize || y >= dstSize) return;
    const uint32_t ps = dstSize * 2;
    const float l00 = srcMip[(size_t)(2 * y) * ps + 2 * x], l10 = srcMip[(size_t)(2 * y) * ps + 2 * x + 1];
    const float l01 = srcMip[(size_t)(2 * y + 1) * ps + 2 * x], l11 = srcMip[(size_t)(2 * y + 1) * ps + 2 * x + 1];
    dstMip[(size_t)y * dstSize + x] = l00 + l10 + l01 + l11;
    srcQuads[(size_t)y * dstSize + x] = make_float4(l00, l10, l01, l11);
}

// ------------------------------------------------------------------------------------------------
// TLAS refit on the device (Accel.zig:629-679: recordRebuild with mode = update_khr): after transform edits the instance
// level keeps its topology and only boxes change. k_tlas_boxes recomputes the world-space box of every edited instance from
// its (already uploaded) new transform and the object-space bounds of its BLAS and writes it into the instance's TLAS leaf;
// k_tlas_propagate then carries boxes upwards: one walker per TLAS leaf, the SECOND walker to arrive at an inner node (the
// first one stops) re-unites the node's child boxes and writes the union into the parent. Counters reset themselves.
// ------------------------------------------------------------------------------------------------
DI void writeChildBox(BvhNode2* nodes, uint32_t packed, const float lo[3], const float hi[3]) {
    BvhNode2* n = nodes + (packed >> 1);
    const uint32_t s = packed & 1u;
    float* xy = s == 0 ? n->n0 : n->n1;
    xy[0] = lo[0]; xy[1] = hi[0]; xy[2] = lo[1]; xy[3] = hi[1];
    n->n2[2 * s] = lo[2]; n->n2[2 * s + 1] = hi[2];
}
// ------------------------------------------------------------------------------------------------
// frame records: the triangle-only part of MeshAttributes::lookupAndInterpolate(...).inWorld(...) (world.hlsl:137-147, :160-176),
// evaluated once per pooled triangle with the functions the hit path would call (shading.cuh lookupAttributes) - same operands,
// same IEEE operations, so a hit that reads the record sees the bits it would have computed.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tri_frames(DeviceScene sc, DConstants k, const uint32_t* __restrict__ triMesh, uint32_t n, float4* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t tm = triMesh[i];
    const DMesh mesh = sc.meshes[tm & 0x7FFFFFFFu];
    const bool inWorld = (tm >> 31) != 0;
    const float4* tp = reinterpret_cast<const float4*>(sc.tris + i);
    const float4 r0 = __ldg(tp), r1 = __ldg(tp + 1), r2 = __ldg(tp + 2);
    const float3 p0 = mk3(r0.x, r0.y, r0.z), p1 = mk3(r1.x, r1.y, r1.z), p2 = mk3(r2.x, r2.y, r2.z);
    const uint32_t primitiveIndex = __float_as_uint(r1.w);
    const uint32_t ix = __ldg(mesh.idx + 3 * (size_t)primitiveIndex), iy = __ldg(mesh.idx + 3 * (size_t)primitiveIndex + 1), iz = __ldg(mesh.idx + 3 * (size_t)primitiveIndex + 2);
    uint32_t ax, ay, az;
    attributeIndices(k, primitiveIndex, ix, iy, iz, ax, ay, az);
    float2 t0, t1, t2;
    cornerTexcoords(mesh, ax, ay, az, t0, t1, t2);
    const Frame obj = objectTriangleFrame(p0, p1, p2, t0, t1, t2);
    Frame rec = obj;
    if (inWorld) {
        float m[12];
        loadRows(sc.instances[__float_as_uint(r2.w)].w2o, m);
        rec = obj.inSpaceTranspose(m);
    }
    float4* o = out + 4 * (size_t)i;
    o[0] = make_float4(obj.s.x, obj.s.y, obj.s.z, inWorld ? 1.0f : 0.0f);
    o[1] = make_float4(rec.n.x, rec.n.y, rec.n.z, 0.0f);
    o[2] = make_float4(rec.s.x, rec.s.y, rec.s.z, 0.0f);
    o[3] = make_float4(rec.t.x, rec.t.y, rec.t.z, 0.0f);
}

__global__ void k_tlas_boxes(TlasRefit r, const uint32_t* __restrict__ dirty, uint32_t nDirty) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nDirty) return;
    const DInstance* in = r.instances + dirty[k];
    const uint32_t slot = in->pad[1];
    if (slot == 0xFFFFFFFFu) return;   // not a TLAS item (identity or hidden at build time)
    const float* bb = r.blasBounds + 6 * (size_t)in->pad[0];
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int c = 0; c < 8; c++) {
        const float p[3] = {(c & 1) ? bb[3] : bb[0], (c & 2) ? bb[4] : bb[1], (c & 4) ? bb[5] : bb[2]};
        for (int a = 0; a < 3; a++) {
            const float q = ((in->o2w[4 * a] * p[0] + in->o2w[4 * a + 1] * p[1]) + in->o2w[4 * a + 2] * p[2]) + in->o2w[4 * a + 3];
            lo[a] = fminf(lo[a], q); hi[a] = fmaxf(hi[a], q);
        }
    }
    // bvh_build.h padBox: keeps the FP32 slab test conservative
    float ext = 1e-3f, mag = 0.0f;
    for (int a = 0; a < 3; a++) { ext = fmaxf(ext, hi[a] - lo[a]); mag = fmaxf(mag, fmaxf(fabsf(lo[a]), fabsf(hi[a]))); }
    const float pad = 1e-5f * ext + 1e-6f * mag;
    for (int a = 0; a < 3; a++) { lo[a] -= pad; hi[a] += pad; }
    writeChildBox(r.nodes, r.slotNode[slot], lo, hi);
}
__global__ void k_tlas_propagate(TlasRefit r, uint32_t nSlots) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nSlots) return;
    uint32_t node = r.slotNode[slot] >> 1;
    for (;;) {
        __threadfence();
        const uint32_t expected = r.expected[node];
        if (expected > 1u) {
            const uint32_t arrived = atomicAdd(r.arrived + node, 1u);
            if (arrived + 1u < expected) return;     // the other child's walker carries on from here
            r.arrived[node] = 0;
            __threadfence();
        }
        const uint32_t parent = r.parent[node];
        if (parent == 0xFFFFFFFFu) return;           // root
        const volatile BvhNode2* n = r.nodes + node;
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (int s = 0; s < 2; s++) {
            if (n->c[s] == REF_EMPTY) continue;
            const volatile float* xy = s == 0 ? n->n0 : n->n1;
            lo[0] = fminf(lo[0], xy[0]); hi[0] = fmaxf(hi[0], xy[1]); lo[1] = fminf(lo[1], xy[2]); hi[1] = fmaxf(hi[1], xy[3]);
            lo[2] = fminf(lo[2], n->n2[2 * s]); hi[2] = fmaxf(hi[2], n->n2[2 * s + 1]);
        }
        writeChildBox(r.nodes, parent, lo, hi);
        node = parent >> 1;
    }
}

// ------------------------------------------------------------------------------------------------
// host-callable launch wrappers
// ------------------------------------------------------------------------------------------------
static inline dim3 grid1(uint64_t n, uint32_t block) { return dim3((uint32_t)((n + block - 1) / block)); }

void launchGenerate(cudaStream_t st, const DCamera& cam, const DConstants& k, const PathBuffers& pb, Counters* counters, const uint32_t* pixIds,
                    uint32_t nPix, uint32_t nSamples, uint32_t firstSample) {
    k_generate<<<grid1((uint64_t)nPix * nSamples, 256), 256, 0, st>>>(cam, k, pb, counters, pixIds, nPix, nSamples, firstSample);
}
void launchTrace(cudaStream_t st, int blocks, const DeviceScene& sc, const PathBuffers& pb, const WaveState& ws, uint32_t buf) {
    if (sc.staticOnly) k_trace<false, false><<<blocks, TRACE_THREADS, 0, st>>>(sc, pb, ws, buf, nullptr, nullptr);
    else k_trace<true, false><<<blocks, TRACE_THREADS, 0, st>>>(sc, pb, ws, buf, nullptr, nullptr);
}
// sorted bounce: counting sort of the keys the previous shade pass wrote (extension and shadow queue in one set of launches,
// item counts read on the device) -> trace through the permutations -> binning pass
void launchSortRays(cudaStream_t st, int smCount, const SortBuffers& sb, const Counters* counters) {
    static const size_t smemBytes = SORT_BINS * sizeof(uint32_t);
    static const cudaError_t a0 = cudaFuncSetAttribute(k_sort_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
    static const cudaError_t a1 = cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
    (void)a0; (void)a1;
    SortJob j;
    j.keys[0] = sb.keysExt; j.keys[1] = sb.keysShadow; j.perm[0] = sb.permExt; j.perm[1] = sb.permShadow;
    j.hist = sb.hist; j.cursor = sb.cursor; j.blockHist = sb.blockHist; j.counters = counters;
    k_sort_count<<<dim3(smCount, 2), SORT_THREADS, smemBytes, st>>>(j);
    k_sort_scan<<<2, SORT_THREADS, 0, st>>>(j);
    k_sort_scatter<<<dim3(smCount, 2), SORT_THREADS, smemBytes, st>>>(j);
}
void launchTraceSorted(cudaStream_t st, int blocks, const DeviceScene& sc, const PathBuffers& pb, const WaveState& ws, uint32_t buf, const uint32_t* permExt, const uint32_t* permShadow) {
    if (sc.staticOnly) k_trace<false, true><<<blocks, TRACE_THREADS, 0, st>>>(sc, pb, ws, buf, permExt, permShadow);
    else k_trace<true, true><<<blocks, TRACE_THREADS, 0, st>>>(sc, pb, ws, buf, permExt, permShadow);
}
void launchTraceSuspend(cudaStream_t st, int blocks, const DeviceScene& sc, const PathBuffers& pb, const WaveState& ws, uint32_t buf, const uint32_t* permExt,
                        const uint32_t* permShadow, const ContBuffers& cb, uint32_t threshold, int passes) {
    // passes: 3 = suspend twice (records A, then B), 2 = suspend once (pass 1 walks the records to the end), 1 = never (measurement: the kernel alone)
    k_trace_rs<0><<<blocks, TRACE_THREADS, 0, st>>>(sc, pb, ws, buf, permExt, permShadow, cb, passes >= 2 ? threshold : 0u);
    if (passes >= 2) k_trace_rs<1><<<blocks, TRACE_THREADS, 0, st>>>(sc, pb, ws, buf, permExt, permShadow, cb, passes >= 3 ? threshold : 0u);
    if (passes >= 3) k_trace_rs<2><<<blocks, TRACE_THREADS, 0, st>>>(sc, pb, ws, buf, permExt, permShadow, cb, 0u);
}
void launchBin(cudaStream_t st, int smCount, const PathBuffers& pb, const WaveState& ws, uint32_t buf) { k_bin<<<smCount * 8, 256, 0, st>>>(pb, ws, buf); }
// Register caps (measured, profiles/r01_sweeps.md): the all-material kernel runs best at 80 registers
// (6 blocks/SM); the specialised kernels are smaller and take the same cap.
#ifndef SHADE_MINB
#define SHADE_MINB 4   // 128 registers, 16 warps per SM, next to no spill: 2575 vs 2543 (6: 80 registers, 460 B of spill stores) vs 2389 (3) Mrays/s on cfg3 (profiles/r02_sweeps.md section 7)
#endif
void launchShade(cudaStream_t st, int blocks, const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, const WaveState& ws, uint32_t buf, int32_t slot, const SortKeys& keys) {
    k_shade<-1, SHADE_MINB><<<blocks, SHADE_THREADS, 0, st>>>(sc, k, pb, ws, buf, slot, keys);
}
// env/mesh_samples_per_bounce > 1: one generic launch per bounce whose light samples trace their shadow rays in place (4 blocks/SM:
// the traversal stack and the shading state are live at the same time)
void launchShadeInlineNee(cudaStream_t st, int blocks, const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, const WaveState& ws, uint32_t buf, int32_t slot, const SortKeys& keys) {
    k_shade<-1, 4, true><<<blocks, SHADE_THREADS, 0, st>>>(sc, k, pb, ws, buf, slot, keys);
}
// one launch per material bin; the miss/dead bin goes last and clears the bounce's counters
void launchShadeBinned(cudaStream_t st, int blocks, const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, const WaveState& ws, uint32_t buf, int32_t slot, const SortKeys& keys) {
    k_shade<3, SHADE_MINB><<<blocks, SHADE_THREADS, 0, st>>>(sc, k, pb, ws, buf, -1, keys);   // StandardPBR
    k_shade<1, SHADE_MINB><<<blocks, SHADE_THREADS, 0, st>>>(sc, k, pb, ws, buf, -1, keys);   // Lambert
    k_shade<0, SHADE_MINB><<<blocks, SHADE_THREADS, 0, st>>>(sc, k, pb, ws, buf, -1, keys);   // Glass
    k_shade<2, SHADE_MINB><<<blocks, SHADE_THREADS, 0, st>>>(sc, k, pb, ws, buf, -1, keys);   // PerfectMirror
    k_shade<4, SHADE_MINB><<<blocks, SHADE_THREADS, 0, st>>>(sc, k, pb, ws, buf, slot, keys);   // miss + dead
}
void launchTail(cudaStream_t st, int blocks, int threads, const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, const WaveState& ws, uint32_t buf, bool inlineNee) {
    if (inlineNee) k_tail<true><<<blocks, threads, 0, st>>>(sc, k, pb, ws, buf);
    else k_tail<false><<<blocks, threads, 0, st>>>(sc, k, pb, ws, buf);
}
void launchAccumulate(cudaStream_t st, const float4* result, float4* image, float4* runState, const uint32_t* pixIds, uint32_t nPix, uint32_t nSamples,
                      uint32_t firstSample, uint32_t spr) {
    k_accumulate<<<grid1(nPix, 256), 256, 0, st>>>(result, image, runState, pixIds, nPix, nSamples, firstSample, spr);
}
void launchEnvResample(cudaStream_t st, const float4* src, uint32_t W, uint32_t H, float4* dstRgb, float* dstLum, uint32_t S) {
    dim3 b(8, 8), g((S + 7) / 8, (S + 7) / 8);
    k_env_resample<<<g, b, 0, st>>>(src, W, H, dstRgb, dstLum, S);
}
void launchEnvFold(cudaStream_t st, const float* srcMip, float* dstMip, float4* srcQuads, uint32_t dstSize) {
    dim3 b(8, 8), g((dstSize + 7) / 8, (dstSize + 7) / 8);
    k_env_fold<<<g, b, 0, st>>>(srcMip, dstMip, srcQuads, dstSize);
}
void launchTraceProbe(cudaStream_t st, const DeviceScene& sc, const float* rays8, uint32_t n, int anyHit, uint32_t* hits5) {
    k_trace_probe<<<grid1(n, 128), 128, 0, st>>>(sc, rays8, n, anyHit, hits5);
}
void launchMathProbe(cudaStream_t st, int fn, const float* a, const float* b, uint32_t n, float* out) { k_math_probe<<<grid1(n, 256), 256, 0, st>>>(fn, a, b, n, out); }
void launchBsdfProbe(cudaStream_t st, const float* q14, uint32_t n, float* out4) { k_bsdf_probe<<<grid1(n, 128), 128, 0, st>>>(q14, n, out4); }
void launchEnvProbe(cudaStream_t st, const DeviceScene& sc, int op, const float* in3, uint32_t n, float* out7) {
    k_env_probe<<<grid1(n, 128), 128, 0, st>>>(sc, op, in3, n, out7);
}
__global__ void k_commit_counts(unsigned long long* c) { c[0] += c[2]; c[1] += c[3]; c[2] = 0; c[3] = 0; }
void launchCommitCounts(cudaStream_t st, unsigned long long* counts) { k_commit_counts<<<1, 1, 0, st>>>(counts); }
void launchTriFrames(cudaStream_t st, const DeviceScene& sc, const DConstants& k, const uint32_t* triMesh, uint32_t n, float4* out) {
    if (n == 0) return;
    k_tri_frames<<<grid1(n, 128), 128, 0, st>>>(sc, k, triMesh, n, out);
}
void launchTlasRefit(cudaStream_t st, const TlasRefit& r, const uint32_t* dirty, uint32_t nDirty, uint32_t nSlots) {
    if (nDirty) k_tlas_boxes<<<grid1(nDirty, 128), 128, 0, st>>>(r, dirty, nDirty);
    if (nSlots) k_tlas_propagate<<<grid1(nSlots, 128), 128, 0, st>>>(r, nSlots);
}
void launchPick(cudaStream_t st, const DeviceScene& sc, const DCamera& cam, float x, float y, MsneClick* out) {
    k_pick<<<1, 1, 0, st>>>(sc, cam, make_float2(x, y), out);
}
void launchPrimaryProbe(cudaStream_t st, const DCamera& cam, const DConstants& k, uint32_t sample, float* rays6, uint32_t* rngState) {
    k_primary_probe<<<grid1((uint64_t)cam.width * cam.height, 256), 256, 0, st>>>(cam, k, sample, rays6, rngState);
}
int traceOccupancyBlocksPerSm() {
    int n = 0, n1 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<true, false>, TRACE_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k_trace<false, false>, TRACE_THREADS, 0);
    return n > n1 ? n : n1;
}
int shadeOccupancyBlocksPerSm() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_shade<-1, SHADE_MINB>, SHADE_THREADS, 0);
    return n;
}
int shadeInlineOccupancyBlocksPerSm() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_shade<-1, 4, true>, SHADE_THREADS, 0);
    return n;
}

}  // namespace msne
