// Launch wrappers of the sm_100a kernels (kernels.cu), callable from the host runtime (runtime.cpp).
#pragma once
#include <cuda_runtime.h>
#include "dtypes.h"

namespace msne {

#ifndef TRACE_THREADS
#define TRACE_THREADS 128
#endif
#ifndef TRACE_BLOCKS_PER_SM
#define TRACE_BLOCKS_PER_SM 8   // 64 registers: measured best (profiles/r01_sweeps.md)
#endif
#ifndef TRACE1_BLOCKS_PER_SM
#define TRACE1_BLOCKS_PER_SM 8  // single-level specialisation (no transformed instances): 64 registers; 8 vs 9 blocks: +2 % since the exact triangle test (profiles/r02_sweeps.md)
#endif
// Entries of the per-ray traversal stack (traverse.cuh): one per tree level (the far child) + the REF_END / REF_LEAVE sentinels.
// runtime.cpp holds depth(TLAS) + depth(BLAS) + 2 below it at build time.
#ifndef MSNE_STACK
#define MSNE_STACK 96
#endif
#ifndef SHADE_THREADS
#define SHADE_THREADS 128
#endif

void launchGenerate(cudaStream_t st, const DCamera& cam, const DConstants& k, const PathBuffers& pb, Counters* counters, const uint32_t* pixIds,
                    uint32_t nPix, uint32_t nSamples, uint32_t firstSample);
void launchTrace(cudaStream_t st, int blocks, const DeviceScene& sc, const PathBuffers& pb, const WaveState& ws, uint32_t buf);
// Ray reordering for incoherent bounces. The shade kernels emit a 16-bit key per ray they queue (kernels.cu rayKey: origin cell
// = (o - lo) * scale in a grid of 2^bits[a] cells per axis over the scene bounds, 3 bits per axis + 3 dealt out to the axes with
// the longest cells, runtime.cpp buildTlas); launchSortRays radix-sorts extension and shadow rays, k_trace walks the two
// permutations and k_bin re-lists the hits per material in queue order.
struct SortFrame { float lo[3], scale[3]; uint32_t bits[3]; };
struct SortKeys { uint16_t* ext; uint16_t* shadow; SortFrame frame; };   // ext == nullptr: the shade pass writes no keys
struct SortBuffers {          // sized for P extension + 2 P shadow rays (runtime.cpp ensureSortBuffers)
    uint16_t* keysExt; uint16_t* keysShadow;
    uint32_t* permExt; uint32_t* permShadow;
    uint32_t* hist; uint32_t* cursor;   // [2][2^15] each; hist is zero between sorts
    uint32_t* blockHist;                // [2][SM count][2^15]
};
// three launches (count, scan, scatter); the item counts are read from `counters` (device memory) when the kernels run
void launchSortRays(cudaStream_t st, int smCount, const SortBuffers& sb, const Counters* counters);
void launchTraceSorted(cudaStream_t st, int blocks, const DeviceScene& sc, const PathBuffers& pb, const WaveState& ws, uint32_t buf, const uint32_t* permExt, const uint32_t* permShadow);
// Suspend-and-regroup tracing of a sorted bounce (static-only scenes; kernels.cu k_trace_rs, traverse.cuh walkResumable): pass 0 walks
// the sorted tickets and files the last `threshold` walking lanes of every warp as 128-byte continuation records in recA, pass 1
// walks recA 32 records per warp (same rule, into recB), pass 2 walks recB to the end. capA >= ceil(rays / 32) * threshold,
// capB >= ceil(capA / 32) * threshold (a ticket never files more than `threshold` records); counts are zero between bounces.
struct ContBuffers { uint4* recA; uint4* recB; uint32_t* counts; };   // counts[0] = records in recA, counts[1] = in recB
void launchTraceSuspend(cudaStream_t st, int blocks, const DeviceScene& sc, const PathBuffers& pb, const WaveState& ws, uint32_t buf, const uint32_t* permExt,
                        const uint32_t* permShadow, const ContBuffers& cb, uint32_t threshold, int passes);
void launchBin(cudaStream_t st, int smCount, const PathBuffers& pb, const WaveState& ws, uint32_t buf);
// `slot` >= 0 on the launch that ends a bounce: the bounce's output counts are logged in ws.history[slot]
void launchShade(cudaStream_t st, int blocks, const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, const WaveState& ws, uint32_t buf, int32_t slot, const SortKeys& keys);
void launchShadeBinned(cudaStream_t st, int blocks, const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, const WaveState& ws, uint32_t buf, int32_t slot, const SortKeys& keys);
void launchShadeInlineNee(cudaStream_t st, int blocks, const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, const WaveState& ws, uint32_t buf, int32_t slot, const SortKeys& keys);
void launchTail(cudaStream_t st, int blocks, int threads, const DeviceScene& sc, const DConstants& k, const PathBuffers& pb, const WaveState& ws, uint32_t buf, bool inlineNee);
// counts[0..1] += counts[2..3]; counts[2..3] = 0 (ray counters of a wave that was rendered ahead and is now claimed)
void launchCommitCounts(cudaStream_t st, unsigned long long* counts);
void launchAccumulate(cudaStream_t st, const float4* result, float4* image, float4* runState, const uint32_t* pixIds, uint32_t nPix, uint32_t nSamples,
                      uint32_t firstSample, uint32_t spr);
void launchEnvResample(cudaStream_t st, const float4* src, uint32_t W, uint32_t H, float4* dstRgb, float* dstLum, uint32_t S);
void launchEnvFold(cudaStream_t st, const float* srcMip, float* dstMip, float4* srcQuads, uint32_t dstSize);
void launchTraceProbe(cudaStream_t st, const DeviceScene& sc, const float* rays8, uint32_t n, int anyHit, uint32_t* hits5);
void launchMathProbe(cudaStream_t st, int fn, const float* a, const float* b, uint32_t n, float* out);
void launchBsdfProbe(cudaStream_t st, const float* q14, uint32_t n, float* out4);
void launchEnvProbe(cudaStream_t st, const DeviceScene& sc, int op, const float* in3, uint32_t n, float* out7);
// TLAS refit on the device (kernels.cu k_tlas_boxes / k_tlas_propagate)
struct TlasRefit {
    BvhNode2* nodes;              // TLAS nodes (boxes rewritten in place)
    const DInstance* instances;   // pad[0] = BLAS id, pad[1] = TLAS leaf slot (0xFFFFFFFF: none)
    const float* blasBounds;      // per BLAS: lo.xyz, hi.xyz (object space, padded)
    const uint32_t* slotNode;     // per TLAS leaf slot: node * 2 + child
    const uint32_t* parent;       // per node: parent * 2 + child, 0xFFFFFFFF for the root
    const uint32_t* expected;     // per node: non-empty children (1 or 2)
    uint32_t* arrived;            // per node, zero between refits
};
void launchTlasRefit(cudaStream_t st, const TlasRefit& r, const uint32_t* dirty, uint32_t nDirty, uint32_t nSlots);
// per-pooled-triangle frame records (shading.cuh MSNE_TRI_FRAMES): triMesh[i] = mesh index of pooled triangle i, bit 31 set for members of
// the merged static BLAS (their instance - TriRec.inst - is known, so the record holds the world-space frame)
void launchTriFrames(cudaStream_t st, const DeviceScene& sc, const DConstants& k, const uint32_t* triMesh, uint32_t n, float4* out);
void launchPick(cudaStream_t st, const DeviceScene& sc, const DCamera& cam, float x, float y, MsneClick* out);
void launchPrimaryProbe(cudaStream_t st, const DCamera& cam, const DConstants& k, uint32_t sample, float* rays6, uint32_t* rngState);
int traceOccupancyBlocksPerSm();
int shadeOccupancyBlocksPerSm();
int shadeInlineOccupancyBlocksPerSm();

}  // namespace msne
