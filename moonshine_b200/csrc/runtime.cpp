// Host runtime of libmoonshine.so: the `HdMoonshine` object behind the reference's C ABI
// (include/moonshine.h <- hydra/moonshine.h:71-95, implemented in the reference by hydra/hydra.zig:107-558)
// plus the extensions of include/moonshine_ext.h.
//
// Replaces, for the hot path only: engine/hrtsystem/{Accel,MeshManager,MaterialManager,BackgroundManager,
// Camera,pipeline}.zig and engine/core (Vulkan context/allocator/commands) with CUDA-runtime equivalents:
// device buffers, one stream, events, and the wavefront scheduler that drives kernels.cu.
// There is NO CPU fallback: every entry point that renders requires a CUDA device and fails otherwise.
//
// Host FP32 arithmetic that feeds the device (camera basis, inverse transforms, alias table, emitter
// areas) is compiled with -ffp-contract=off and follows the reference's operation order
// (engine/vector.zig:350-362,501-521; engine/hrtsystem/alias_table.zig:28-91; Accel.zig:503-521).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library itself is dlopen'ed on first use (HdMoonshineExtComm*)

#include "../../include/moonshine_ext.h"
#include "bvh_build.h"
#include "dtypes.h"
#include "kernels.h"

using namespace msne;

namespace {

thread_local std::string g_lastError;
// Selected per calling thread (HdMoonshineExtSetDevice -> HdMoonshineCreate on the same thread): two threads that bring up
// renderers on different GPUs do not race on it.
thread_local int g_device = 0;

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char buf_[512];                                                                              \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw CudaError(buf_);                                                                       \
        }                                                                                                \
    } while (0)

[[noreturn]] void fatal(const char* what) {
    // hydra.zig: `catch unreachable` — everything but Create/Render/RebuildPipeline aborts on failure
    fprintf(stderr, "libmoonshine: fatal: %s\n", what);
    abort();
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count) { release(); if (count) { CK(cudaMalloc((void**)&p, count * sizeof(T))); n = count; } }
    void ensure(size_t count) { if (count > n) alloc(count); }
    void upload(const T* src, size_t count, cudaStream_t st) {
        ensure(count);
        if (count) CK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
};

struct HostMesh {
    std::vector<float> pos, nrm, uv;     // host copies (alias table + BLAS builds; MeshManager.zig:46-48)
    std::vector<uint32_t> idx;
    DevBuf<float> dPos, dNrm, dUv;
    DevBuf<uint32_t> dIdx;
};

struct HostTexture { DevBuf<uint8_t> data; uint32_t w, h, fmt; uint8_t first[16]; };   // first = host copy of texel (0,0)

struct HostInstance {
    float transform[12];
    float inverse[12];
    bool visible;
    uint32_t customIndex, geometryCount;
    std::vector<uint32_t> meshKey;   // BLAS dedupe key: the mesh ids of the geometry list (Accel.zig:315-343)
    uint32_t blas;                   // index into blases
    bool inStatic = false;           // its triangles are part of the current merged static BLAS
    uint32_t tlasSlot = 0xFFFFFFFFu; // its leaf in the current TLAS (transformed instances that were visible when it was built)
    uint32_t triCount = 0;
};

struct Blas { uint32_t rootNode; BuildBox bounds; uint32_t depth; };

struct Sensor {
    Extent2D extent{};
    uint32_t sample_count = 0;
    DevBuf<float4> image;
    float* host = nullptr;           // pinned, W*H*4 floats, returned by HdMoonshineGetSensorData
    DevBuf<uint32_t> pixIds;         // owned pixels in tile order
    DevBuf<float4> runState;
    uint32_t nPix = 0;
    uint32_t shardRank = 0, shardWorld = 0;   // shard the pixIds were built for
    DevBuf<float4> reduced;          // root rank of a communicator: the frame assembled from every rank's tiles
};

// engine/vector.zig:350-362 (inverse_affine) on row-major 3x4
void inverseAffine(const float* M, float* R) {
    auto cross = [](const float* a, const float* b, float* o) { o[0] = a[1] * b[2] - b[1] * a[2]; o[1] = a[2] * b[0] - b[2] * a[0]; o[2] = a[0] * b[1] - b[0] * a[1]; };
    const float px[3] = {M[0], M[4], M[8]}, py[3] = {M[1], M[5], M[9]}, pz[3] = {M[2], M[6], M[10]};
    const float v[3] = {M[3], M[7], M[11]};
    float c[3];
    cross(py, pz, c);
    const float det = px[0] * c[0] + px[1] * c[1] + px[2] * c[2];
    const float rd = 1 / det;
    float v1[3], v2[3], v3[3];
    cross(py, pz, v1); cross(pz, px, v2); cross(px, py, v3);
    for (int i = 0; i < 3; i++) { v1[i] *= rd; v2[i] *= rd; v3[i] *= rd; }
    const float ix[3] = {v1[0], v2[0], v3[0]}, iy[3] = {v1[1], v2[1], v3[1]}, iz[3] = {v1[2], v2[2], v3[2]};
    float res[3];
    for (int i = 0; i < 3; i++) {
        float r = (ix[i] * -1.0f) * v[0];
        r = (iy[i] * -1.0f) * v[1] + r;
        r = (iz[i] * -1.0f) * v[2] + r;
        res[i] = r;
    }
    for (int r = 0; r < 3; r++) { R[4 * r + 0] = ix[r]; R[4 * r + 1] = iy[r]; R[4 * r + 2] = iz[r]; R[4 * r + 3] = res[r]; }
}

// engine/hrtsystem/alias_table.zig:28-91 (Vose); entries carry their payload already
float buildAliasTable(const std::vector<float>& w, std::vector<DAlias>& e) {
    const uint32_t NONE = 0xFFFFFFFFu, n = (uint32_t)w.size();
    float sum = 0.0f;
    for (float x : w) sum += x;
    uint32_t less = NONE, more = NONE;
    for (uint32_t i = 0; i < n; i++) {
        const float adj = (w[i] * (float)n) / sum;
        e[i].select = adj;
        if (adj < 1.0f) { e[i].alias = less; less = i; } else { e[i].alias = more; more = i; }
    }
    while (less != NONE && more != NONE) {
        const uint32_t l = less; less = e[l].alias;
        const uint32_t m = more; more = e[m].alias;
        e[l].alias = m;
        e[m].select = (e[m].select + e[l].select) - 1.0f;
        if (e[m].select < 1.0f) { e[m].alias = less; less = m; } else { e[m].alias = more; more = m; }
    }
    while (less != NONE) { const uint32_t l = less; less = e[l].alias; e[l].select = 1.0f; }
    return sum;
}

float srgbToLinear(int c) {
    const double s = c / 255.0;
    return (float)(s <= 0.04045 ? s / 12.92 : std::pow((s + 0.055) / 1.055, 2.4));
}

// NCCL entry points, resolved at run time: a single-GPU deployment needs no NCCL installed, and inside a process that already
// loaded a libnccl.so.2 (e.g. under torch.distributed) the loader hands back that same copy.
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) getUniqueId = nullptr;
    decltype(&ncclCommInitRank) commInitRank = nullptr;
    decltype(&ncclCommDestroy) commDestroy = nullptr;
    decltype(&ncclReduce) reduce = nullptr;
    decltype(&ncclAllReduce) allReduce = nullptr;
    decltype(&ncclGetErrorString) getErrorString = nullptr;
};
NcclApi& ncclApi() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
        if (!a.lib) return a;
        a.getUniqueId = (decltype(a.getUniqueId))dlsym(a.lib, "ncclGetUniqueId");
        a.commInitRank = (decltype(a.commInitRank))dlsym(a.lib, "ncclCommInitRank");
        a.commDestroy = (decltype(a.commDestroy))dlsym(a.lib, "ncclCommDestroy");
        a.reduce = (decltype(a.reduce))dlsym(a.lib, "ncclReduce");
        a.allReduce = (decltype(a.allReduce))dlsym(a.lib, "ncclAllReduce");
        a.getErrorString = (decltype(a.getErrorString))dlsym(a.lib, "ncclGetErrorString");
        return a;
    }();
    if (!api.lib || !api.getUniqueId || !api.commInitRank || !api.commDestroy || !api.reduce || !api.allReduce || !api.getErrorString)
        throw std::runtime_error("NCCL (libnccl.so.2) is not available in this process");
    return api;
}
#define NK(expr)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (expr);                                                                        \
        if (r_ != ncclSuccess) {                                                                         \
            char buf_[512];                                                                              \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #expr, ncclApi().getErrorString(r_), __FILE__, __LINE__); \
            throw CudaError(buf_);                                                                       \
        }                                                                                                \
    } while (0)

size_t texelBytes(uint32_t fmt) {
    switch (fmt) {
        case TEX_R32F: return 4; case TEX_RG32F: return 8; case TEX_RGBA32F: return 16; case TEX_R8_UNORM: return 1;
        case TEX_RG8_UNORM: return 2; case TEX_RGBA8_SRGB: return 4; case TEX_RGBA16F: return 8; default: return 0;
    }
}

}  // namespace

struct HdMoonshine {
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    std::mutex mutex;   // hydra.zig:78 — every entry point is serialised

    // scene (host mirrors + device tables)
    std::vector<std::unique_ptr<HostMesh>> meshes;
    std::vector<std::unique_ptr<HostTexture>> textures;
    std::vector<DMaterial> materials;
    std::vector<HostInstance> instances;
    std::vector<DGeometry> geometries;
    std::vector<DAlias> alias;   // [0] header; always at least the header
    std::map<std::vector<uint32_t>, uint32_t> blasByKey;
    std::vector<Blas> blases;
    std::vector<BvhNode2> blasNodePool;
    std::vector<TriRec> triPool;
    std::vector<uint32_t> triMeshPool;   // per pooled triangle: mesh index, bit 31 = member of the merged static BLAS (frame records, kernels.cu k_tri_frames)
    DevBuf<float4> dTriFrames;
    bool triFramesDirty = true, triFramesIndexed = false;

    // geomDirty: the set of flattened (identity-transform, visible) instances or the BLAS pool changed -> rebuild BLASes + TLAS.
    // tlasDirty: only transforms / visibility of transformed instances changed -> rebuild the TLAS and instance records.
    bool tablesDirty = true, geomDirty = true, tlasDirty = true, aliasDirty = true;
    // tlasRefitOk: since the last TLAS build only transforms of transformed, visible instances changed -> the TLAS keeps its
    // topology and only its boxes are recomputed bottom-up (Accel.zig:629-679 recordRebuild with mode = update_khr).
    bool tlasRefitOk = false;
    uint32_t refitsSinceBuild = 0, maxRefits = 32;   // MSNE_TLAS_MAX_REFITS: rebuild after this many consecutive refits
    // Edits that only touch instance RECORDS (transform of a TLAS member, visibility of a member of either level) are served
    // without rebuilding anything: the records of `dirtyInstances` are re-uploaded (128 B each), transforms additionally refit
    // the TLAS boxes on the device. Hidden members stay in their tree (skipped at traversal time) until they make up a quarter
    // of it, then the next sync rebuilds without them.
    std::vector<uint32_t> dirtyInstances;
    uint64_t staticTris = 0, hiddenStaticTris = 0;
    uint32_t tlasItemCount = 0, hiddenTlasItems = 0;
    DevBuf<uint32_t> dSlotNode, dNodeParent, dNodeExpected, dNodeArrived, dDirty;
    DevBuf<float> dBlasBounds;
    std::vector<BvhNode2> tlasHostNodes;             // pre-encoded TLAS as uploaded
    std::vector<uint32_t> tlasHostSlots;             // TLAS leaf slot -> instance record
    // emitter alias table: rebuilt by the next render after an edit that changes it, once following is switched on
    bool emitterFollowsEdits = false, emitterStale = false;

    DevBuf<BvhNode2> dTlasNodes, dBlasNodes;
    DevBuf<uint32_t> dTlasItems;
    DevBuf<TriRec> dTris;
    DevBuf<DInstance> dInstances;
    DevBuf<DGeometry> dGeometries;
    DevBuf<DMesh> dMeshes;
    DevBuf<DMaterial> dMaterials;
    DevBuf<DTexture> dTextures;
    DevBuf<DAlias> dAlias;
    DevBuf<float> dSrgbLut;
    DevBuf<float4> dEnvRgb;
    DevBuf<float> dEnvLum;
    DevBuf<float4> dEnvQuads;
    uint32_t envSize = 0, envLevels = 0, envLevelOffset[12] = {0}, envQuadOffset[12] = {0};
    bool hasTlas = false, tlasStaticOnly = false;

    MsneSpecConstants consts{};   // hydra defaults (hydra.zig:97-105)
    std::vector<Lens> lenses;
    std::vector<std::unique_ptr<Sensor>> sensors;
    uint32_t shardRank = 0, shardWorld = 1;
    uint32_t tileEdge = 16;          // pixel tiles of tileEdge x tileEdge are dealt round-robin to the ranks (MSNE_TILE_EDGE)
    ncclComm_t comm = nullptr;       // HdMoonshineExtCommInit: one rank per GPU, tiles sharded by rank, frames assembled with ncclReduce
    uint32_t commRank = 0, commWorld = 1;

    // ---- wavefront execution ---------------------------------------------------------------------------------------------
    // A WAVE = up to P paths (consecutive samples of a chunk of owned pixels) carried from k_generate to k_accumulate. The host
    // enqueues a whole wave - every bounce's sort / trace / bin / shade launches, then k_tail - WITHOUT waiting for the device:
    // queue sizes live in device counters that the kernels read, and which launches a bounce gets (sorted or not, binned or
    // generic shading, how many wavefront bounces before k_tail) is a performance HINT taken from the survival profile of the
    // last finished wave. Whatever the hint, the image is the same bits (tests assert it). Two wave contexts (own stream, pool,
    // counters) alternate, so the latency-bound Russian-roulette tail of one wave (k_tail: a few thousand paths, ~150 bounces
    // deep) runs under the bandwidth-bound head of the next; accumulates are chained by events to stay in sample order.
    static constexpr int kWaveContexts = 4;
    static constexpr int kDefaultWaveContexts = 2;
    static constexpr uint64_t kMinSplitWave = 131072;   // paths: smaller renders stay one wave
    static constexpr uint32_t kMaxHistory = 192;   // wavefront bounces logged / scheduled per wave at most (k_tail finishes the rest)
    static const int kProfEvents = 96;
    struct WaveCtx {
        cudaStream_t stream = nullptr;
        PathBuffers pb{};
        std::vector<DevBuf<float4>> poolBufs;
        DevBuf<uint4> poolH;
        DevBuf<uint32_t> poolTypeList;
        DevBuf<Counters> dCounters;        // [2] ping-pong
        DevBuf<uint32_t> dScalars;         // [0] traceTicket [1] traceDone [2] shadeDone [3] maxBounce [4..8] typeCount [9..13] binDone
        DevBuf<unsigned long long> dRayCounts;
        DevBuf<Counters> dHistory;         // [kMaxHistory] queue sizes every bounce produced
        Counters* hostHistory = nullptr;   // pinned copy, valid once evDone has fired
        // ray reordering (kernels.h SortBuffers): keys of the extension (P) and the shadow queue (2 P), the two permutations,
        // global histogram + bin cursors of the counting sort
        DevBuf<uint16_t> sortKeysExt, sortKeysShadow;
        DevBuf<uint32_t> sortPermExt, sortPermShadow, sortHist, sortCursor, sortBlockHist;
        DevBuf<uint4> contA, contB;        // suspend-and-regroup tracing: continuation records (128 B each)
        DevBuf<uint32_t> contCounts;
        cudaEvent_t evDone = nullptr;      // recorded after the wave's k_accumulate
        bool pending = false;              // a wave was enqueued and its log not read yet
        uint32_t bounces = 0;              // wavefront bounces that wave scheduled
        uint64_t paths = 0;
        int profUsed = 0;
        cudaEvent_t profEv[kProfEvents][4]{};   // [0],[1] around k_trace; sorted bounces: [2] before the sort, [3] after the binning pass
        bool profSorted[kProfEvents]{};
    };
    WaveCtx ctx[kWaveContexts];
    uint64_t pathsInFlight = 1ull << 27;   // over both contexts: 2 x 64 Mi paths x (308 + 18) B = 41 GiB of the 180 GB HBM
    uint32_t suspendThreshold = 0;   // > 0: sorted bounces of static-only scenes are traced by k_trace_rs, whose warps hand their last N walking lanes to a later pass (MSNE_TRACE_SUSPEND)
    int suspendPasses = 3;           // MSNE_TRACE_SUSPEND_PASSES: 3 = suspend twice, 2 = once, 1 = never (the kernel alone; measurements)
    uint64_t sortThreshold = 262144; // extension-ray queues at least this long are traced in sorted order (0 = never)
    bool sortAnyScene = false;       // false: only scenes with >= kSortMinNodes BVH nodes (a 36-triangle Cornell box loses 4 % to the sort)
    static constexpr size_t kSortMinNodes = 2048;
    bool sortActive = false;         // decided per render (renderLaunches)
    SortFrame sortFrame{};           // scene bounds for the origin cells (buildTlas)
    cudaEvent_t evStart = nullptr, evStop = nullptr, evAccum[kWaveContexts]{};
    // `timelineOpen` from the first enqueue of a render call until finishWaves(). The machinery allows a render call to return with
    // its waves still in flight (renderLaunches(wait = false); every other entry point calls finishWaves() first, consecutive calls
    // keep alternating the contexts and chaining the accumulates); both public render entry points wait, see HdMoonshineExtRenderLaunches.
    bool timelineOpen = false;
    uint64_t waveSerial = 0;
    cudaEvent_t lastAccumEvent = nullptr;
    int traceBlocks = 0, shadeBlocks = 0, shadeInlineBlocks = 0;
    uint32_t maxLeaf = 3;   // triangles per BVH leaf (<= 8: 3-bit count in the leaf reference); 2-3 measured best
    int tailBlockThreads = 128;       // k_tail block size (32, 64 or 128; MSNE_TAIL_BLOCK): a block holds its registers until its longest path ends
    uint64_t tailThreshold = 32768;   // once fewer paths than this are expected to survive, k_tail finishes the wave
    bool tailExplicit = false;        // the caller chose tailThreshold (HdMoonshineExtSetSchedule / MSNE_TAIL_THRESHOLD)
    // render-ahead keeps the wavefront going four times longer: k_tail holds a quarter of an SM's registers per block until its longest
    // path ends, which is what the younger waves' kernels wait for (8 Ki: 9.5 ms per call on cfg3, 32 Ki: 9.9, 128 Ki: 11.6; plain calls
    // want 16-32 Ki: profiles/r02_sweeps.md section 9)
    static constexpr uint64_t kAheadTailThreshold = 8192;
    bool forceTwoLevel = false;       // MSNE_TWO_LEVEL=1: never use the single-level trace kernel (A/B measurements)
    bool logBounces = false;          // MSNE_LOG_BOUNCES=1: queue sizes after every bounce on stderr (printed when the wave's log is read)
    uint64_t binnedThreshold = 262144;   // bounces expected to shade fewer paths than this use the single generic shade launch
    int waveContextsInUse = kDefaultWaveContexts;   // MSNE_WAVE_CONTEXTS=1: one wave at a time (A/B measurements)
    // survival profile of the last finished wave: survivalIn[b] = live paths entering bounce b / paths of the wave (b = 0: 1)
    std::vector<double> survivalIn;

    // profiling
    bool profiling = false;
    MsneStats stats{};

    // ---- lifecycle ---------------------------------------------------------------------------
    void init() {
        device = g_device;
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        smCount = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&evStart)); CK(cudaEventCreate(&evStop));
        for (auto& e : evAccum) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (WaveCtx& c : ctx) {
            CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c.evDone, cudaEventDisableTiming));
            for (auto& p : c.profEv) for (auto& e : p) CK(cudaEventCreate(&e));
            CK(cudaMallocHost((void**)&c.hostHistory, kMaxHistory * sizeof(Counters)));
            c.dCounters.alloc(2); c.dScalars.alloc(16); c.dRayCounts.alloc(4); c.dHistory.alloc(kMaxHistory);
            c.sortHist.alloc(2 * 32768); c.sortCursor.alloc(2 * 32768);
            CK(cudaMemsetAsync(c.dCounters.p, 0, 2 * sizeof(Counters), stream));
            CK(cudaMemsetAsync(c.dScalars.p, 0, 16 * sizeof(uint32_t), stream));
            CK(cudaMemsetAsync(c.dRayCounts.p, 0, 4 * sizeof(unsigned long long), stream));
            CK(cudaMemsetAsync(c.dHistory.p, 0, kMaxHistory * sizeof(Counters), stream));
            CK(cudaMemsetAsync(c.sortHist.p, 0, 2 * 32768 * sizeof(uint32_t), stream));
        }
        float lut[512];
        for (int i = 0; i < 256; i++) { lut[i] = srgbToLinear(i); lut[256 + i] = (float)i / 255.0f; }   // UNORM: the same IEEE division a texel fetch would do
        dSrgbLut.upload(lut, 512, stream);
        alias.assign(1, DAlias{0, 0.0f, 0, 0, 0});
        // hydra.zig:97-105
        consts = MsneSpecConstants{1, 1024, 0, 0, 0, 0, 0};
        const float white[4] = {1, 1, 1, 1};
        setBackground(white, 1, 1);   // addDefaultBackground (BackgroundManager.zig:116-126)
        if (const char* e = getenv("MSNE_MAX_LEAF")) maxLeaf = (uint32_t)std::min(8, std::max(1, atoi(e)));
        if (const char* e = getenv("MSNE_TLAS_MAX_REFITS")) maxRefits = (uint32_t)std::max(0, atoi(e));
        if (const char* e = getenv("MSNE_TWO_LEVEL")) forceTwoLevel = atoi(e) != 0;
        if (const char* e = getenv("MSNE_LOG_BOUNCES")) logBounces = atoi(e) != 0;
        if (const char* e = getenv("MSNE_TRACE_SUSPEND")) suspendThreshold = (uint32_t)std::min(16, std::max(0, atoi(e)));
        if (const char* e = getenv("MSNE_TRACE_SUSPEND_PASSES")) suspendPasses = std::min(3, std::max(1, atoi(e)));
        if (const char* e = getenv("MSNE_TAIL_BLOCK")) { const int t = atoi(e); if (t == 32 || t == 64 || t == 128) tailBlockThreads = t; }
        if (const char* e = getenv("MSNE_TAIL_THRESHOLD")) { tailThreshold = strtoull(e, nullptr, 10); tailExplicit = true; }
        if (const char* e = getenv("MSNE_SORT_THRESHOLD")) { sortThreshold = strtoull(e, nullptr, 10); sortAnyScene = true; }
        if (const char* e = getenv("MSNE_TILE_EDGE")) tileEdge = (uint32_t)std::min(256, std::max(8, atoi(e) / 8 * 8));
        if (const char* e = getenv("MSNE_WAVE_CONTEXTS")) waveContextsInUse = std::min(kWaveContexts, std::max(1, atoi(e)));
        if (const char* e = getenv("MSNE_BINNED_THRESHOLD")) binnedThreshold = strtoull(e, nullptr, 10);
        if (const char* e = getenv("MSNE_RENDER_AHEAD")) { renderAhead = atoi(e) != 0; if (atoi(e) > 0) aheadDepth = (uint32_t)atoi(e); }
        traceBlocks = smCount * std::max(1, traceOccupancyBlocksPerSm());
        shadeBlocks = smCount * std::max(1, shadeOccupancyBlocksPerSm());
        // measurement knobs: launch fewer blocks per SM than fit, so that kernels of the OTHER wave context can share the SMs
        if (const char* e = getenv("MSNE_TRACE_GRID_PER_SM")) traceBlocks = smCount * std::max(1, atoi(e));
        if (const char* e = getenv("MSNE_SHADE_GRID_PER_SM")) shadeBlocks = smCount * std::max(1, atoi(e));
        shadeInlineBlocks = smCount * std::max(1, shadeInlineOccupancyBlocksPerSm());
        CK(cudaStreamSynchronize(stream));
    }

    ~HdMoonshine() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        if (comm) { try { ncclApi().commDestroy(comm); } catch (...) {} comm = nullptr; }
        for (WaveCtx& c : ctx) {
            if (c.stream) { cudaStreamSynchronize(c.stream); cudaStreamDestroy(c.stream); }
            if (c.evDone) cudaEventDestroy(c.evDone);
            for (auto& p : c.profEv) for (auto& e : p) if (e) cudaEventDestroy(e);
            if (c.hostHistory) cudaFreeHost(c.hostHistory);
        }
        for (auto& s : sensors) if (s && s->host) cudaFreeHost(s->host);
        for (auto& e : evAccum) if (e) cudaEventDestroy(e);
        if (evStart) cudaEventDestroy(evStart);
        if (evStop) cudaEventDestroy(evStop);
        if (stream) cudaStreamDestroy(stream);
    }

    void use() { CK(cudaSetDevice(device)); }

    // ---- environment (BackgroundManager.zig:142-394) -----------------------------------------
    void setBackground(const float* rgba, uint32_t W, uint32_t H) {
        DevBuf<float4> src;
        src.upload(reinterpret_cast<const float4*>(rgba), (size_t)W * H, stream);
        uint32_t p2 = 1;
        while (p2 * 2 <= H) p2 *= 2;
        const uint32_t S = std::min(p2, 1024u);
        envSize = S;
        envLevels = 0;
        size_t total = 0;
        for (uint32_t s = S; s >= 1; s >>= 1) { envLevelOffset[envLevels++] = (uint32_t)total; total += (size_t)s * s; if (s == 1) break; }
        dEnvRgb.ensure((size_t)S * S);
        dEnvLum.ensure(total);
        size_t quads = 0;
        for (uint32_t l = 1; l < envLevels; l++) { envQuadOffset[l - 1] = (uint32_t)quads; quads += (size_t)(S >> l) * (S >> l); }
        dEnvQuads.ensure(std::max<size_t>(quads, 1));
        launchEnvResample(stream, src.p, W, H, dEnvRgb.p, dEnvLum.p, S);
        stats.kernel_launches++;
        for (uint32_t l = 1; l < envLevels; l++) {
            launchEnvFold(stream, dEnvLum.p + envLevelOffset[l - 1], dEnvLum.p + envLevelOffset[l], dEnvQuads.p + envQuadOffset[l - 1], S >> l);
            stats.kernel_launches++;
        }
        CK(cudaStreamSynchronize(stream));   // src is freed on return
    }

    // ---- acceleration structure (replaces Accel.zig:94-184 makeBlases and :440-489 TLAS build) ------
    // pools a host BVH: children become the pre-encoded traversal references of dtypes.h
    uint32_t appendBlas(const BuiltBvh& bvh, const std::vector<TriRec>& tris, const std::vector<uint32_t>& triMesh) {
        // leaf references hold the pooled triangle index in 28 bits (dtypes.h), hit records in 29 (H_TRI_MASK); inner references must
        // stay below REF_ITEM. The merged static BLAS duplicates triangles per identity instance, so the pools can grow quickly.
        if (triPool.size() + tris.size() >= (1ull << 28)) throw std::runtime_error("acceleration structure: more than 2^28 pooled triangles");
        if (blasNodePool.size() + bvh.nodes.size() >= (size_t)REF_ITEM) throw std::runtime_error("acceleration structure: more than 2^30 pooled BVH nodes");
        const uint32_t nodeBase = (uint32_t)blasNodePool.size(), triBase = (uint32_t)triPool.size();
        for (BvhNode2 n : bvh.nodes) {
            for (int s = 0; s < 2; s++) {
                const int32_t child = n.c[s], count = n.c[2 + s];
                if (count > 0) n.c[s] = (int32_t)~((((uint32_t)child + triBase) << 3) | (uint32_t)(count - 1));
                else if (child < 0) { n.c[s] = REF_EMPTY; setEmptyChild(n, s); }
                else n.c[s] = child + (int32_t)nodeBase;
            }
            n.c[2] = n.c[3] = 0;
            blasNodePool.push_back(n);
        }
        for (uint32_t i = 0; i < bvh.order.size(); i++) { triPool.push_back(tris[bvh.order[i]]); triMeshPool.push_back(triMesh[bvh.order[i]]); }
        return nodeBase;
    }
    static void meshTriangles(const HostMesh& m, uint32_t geom, uint32_t inst, std::vector<TriRec>& tris, std::vector<BuildBox>& boxes) {
        const uint32_t nTri = (uint32_t)(m.idx.size() / 3);
        for (uint32_t p = 0; p < nTri; p++) {
            TriRec t{};
            const float* v[3] = {&m.pos[3 * (size_t)m.idx[3 * p]], &m.pos[3 * (size_t)m.idx[3 * p + 1]], &m.pos[3 * (size_t)m.idx[3 * p + 2]]};
            memcpy(t.v0, v[0], 12); memcpy(t.v1, v[1], 12); memcpy(t.v2, v[2], 12);
            t.geom = geom; t.prim = p; t.inst = inst;
            BuildBox b;
            for (int a = 0; a < 3; a++) {
                b.lo[a] = std::min(v[0][a], std::min(v[1][a], v[2][a]));
                b.hi[a] = std::max(v[0][a], std::max(v[1][a], v[2][a]));
            }
            tris.push_back(t); boxes.push_back(b);
        }
    }
    // The traversal stack holds one entry per level of the walk (far child), plus the REF_END and REF_LEAVE sentinels:
    // depth(TLAS) + depth(BLAS) + 2 <= MSNE_STACK. The builder's own bound (SAH cap + log2 n per tree) does not guarantee that for
    // two degenerate trees stacked on each other, so BLASes are held to kMaxBlasDepth levels and the TLAS to what is left
    // (buildTlas); a tree over the limit is rebuilt with a tighter SAH cap (median splits below it).
    static constexpr uint32_t kMaxBlasDepth = 64;
    uint32_t maxBlasDepth = 1;
    static uint32_t ceilLog2(uint32_t n) { uint32_t l = 0; while ((1ull << l) < n) l++; return l; }
    static void buildDepthLimited(const std::vector<BuildBox>& boxes, uint32_t maxLeaf, uint32_t depthLimit, BuiltBvh& out) {
        buildBvh2(boxes.data(), (uint32_t)boxes.size(), maxLeaf, out);
        if (out.depth <= depthLimit) return;
        const uint32_t lg = ceilLog2((uint32_t)boxes.size());
        if (depthLimit <= lg + 1) throw std::runtime_error("acceleration structure: hierarchy too deep for the traversal stack");
        buildBvh2(boxes.data(), (uint32_t)boxes.size(), maxLeaf, out, depthLimit - lg - 1);
        if (out.depth > depthLimit) throw std::runtime_error("acceleration structure: depth limit not met");
    }
    uint32_t getBlas(const std::vector<uint32_t>& key) {
        auto it = blasByKey.find(key);
        if (it != blasByKey.end()) return it->second;
        std::vector<BuildBox> boxes;
        std::vector<TriRec> tris;
        std::vector<uint32_t> triMesh;
        for (uint32_t g = 0; g < key.size(); g++) { meshTriangles(*meshes[key[g]], g, 0, tris, boxes); triMesh.resize(tris.size(), key[g]); }
        BuiltBvh bvh;
        buildDepthLimited(boxes, maxLeaf, kMaxBlasDepth, bvh);
        maxBlasDepth = std::max(maxBlasDepth, bvh.depth);
        Blas b; b.rootNode = appendBlas(bvh, tris, triMesh); b.bounds = bvh.bounds; b.depth = bvh.depth;
        blases.push_back(b);
        const uint32_t id = (uint32_t)blases.size() - 1;
        blasByKey[key] = id;
        return id;
    }
    static bool isIdentity(const float* m) {
        static const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        return memcmp(m, I, 48) == 0;
    }

    // (1) BLAS level: merged static BLAS over every visible identity-transform instance + one shared BLAS per mesh-key
    //     of the transformed instances. Only needed when that set changes (Accel.zig:94-184 equivalent).
    DInstance staticRec{};
    bool haveStatic = false;
    BuildBox staticBounds{};
    uint64_t stats_static_builds = 0, stats_tlas_builds = 0, stats_tlas_refits = 0;
    uint32_t tlasDepth = 0;
    void buildGeometry() {
        blasNodePool.clear(); triPool.clear(); triMeshPool.clear(); blasByKey.clear(); blases.clear();
        maxBlasDepth = 1;
        const uint32_t nInst = (uint32_t)instances.size();
        std::vector<TriRec> stTris;
        std::vector<BuildBox> stBoxes;
        std::vector<uint32_t> stMesh;
        for (uint32_t i = 0; i < nInst; i++) {
            HostInstance& in = instances[i];
            in.inStatic = in.visible && isIdentity(in.transform);
            if (!in.inStatic) continue;
            for (uint32_t g = 0; g < in.meshKey.size(); g++) { meshTriangles(*meshes[in.meshKey[g]], g, i, stTris, stBoxes); stMesh.resize(stTris.size(), in.meshKey[g] | 0x80000000u); }
        }
        staticTris = stTris.size(); hiddenStaticTris = 0;
        staticRec = DInstance{};
        const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        memcpy(staticRec.o2w, I, 48); memcpy(staticRec.w2o, I, 48);
        staticRec.isStatic = 1; staticRec.visible = 1;
        haveStatic = !stTris.empty();
        if (haveStatic) {
            BuiltBvh bvh;
            buildDepthLimited(stBoxes, maxLeaf, kMaxBlasDepth, bvh);
            maxBlasDepth = std::max(maxBlasDepth, bvh.depth);
            staticRec.blasRoot = appendBlas(bvh, stTris, stMesh);
            staticBounds = bvh.bounds;
        }
        for (uint32_t i = 0; i < nInst; i++) {
            HostInstance& in = instances[i];
            in.blas = isIdentity(in.transform) ? 0xFFFFFFFFu : getBlas(in.meshKey);   // invisible ones too: toggling visibility stays TLAS-only
        }
        dBlasNodes.upload(blasNodePool.data(), blasNodePool.size(), stream);
        dTris.upload(triPool.data(), triPool.size(), stream);
        CK(cudaStreamSynchronize(stream));
        triFramesDirty = true;
        stats_static_builds++;
    }

    // World-space bounds of a transformed instance (TLAS item). Small meshes (<= kHullVertices vertices): bounds of the
    // transformed vertices - for a rotated object up to 2x tighter than the transformed object-space box, which is what
    // decides how many instances a ray has to enter in instance-heavy scenes (cfg5: 10^6 icosahedra). Larger ones: the
    // eight corners of the (padded) BLAS bounds. Either way padded again in world space.
    static constexpr size_t kHullVertices = 4096;
    BuildBox instanceWorldBox(const HostInstance& in) const {
        const BuildBox& bb = blases[in.blas].bounds;
        BuildBox w; for (int a = 0; a < 3; a++) { w.lo[a] = 1e30f; w.hi[a] = -1e30f; }
        size_t nVerts = 0;
        for (uint32_t mk : in.meshKey) nVerts += meshes[mk]->pos.size() / 3;
        if (nVerts > 0 && nVerts <= kHullVertices) {
            for (uint32_t mk : in.meshKey) {
                const std::vector<float>& pos = meshes[mk]->pos;
                for (size_t v = 0; v + 2 < pos.size(); v += 3)
                    for (int r = 0; r < 3; r++) {
                        const float q = ((in.transform[4 * r] * pos[v] + in.transform[4 * r + 1] * pos[v + 1]) + in.transform[4 * r + 2] * pos[v + 2]) + in.transform[4 * r + 3];
                        w.lo[r] = std::min(w.lo[r], q); w.hi[r] = std::max(w.hi[r], q);
                    }
            }
            padBox(w); padBox(w);   // the object-space padding of the BLAS bounds is not inherited on this path
            return w;
        }
        for (int c = 0; c < 8; c++) {
            const float p[3] = {(c & 1) ? bb.hi[0] : bb.lo[0], (c & 2) ? bb.hi[1] : bb.lo[1], (c & 4) ? bb.hi[2] : bb.lo[2]};
            for (int r = 0; r < 3; r++) {
                const float q = ((in.transform[4 * r] * p[0] + in.transform[4 * r + 1] * p[1]) + in.transform[4 * r + 2] * p[2]) + in.transform[4 * r + 3];
                w.lo[r] = std::min(w.lo[r], q); w.hi[r] = std::max(w.hi[r], q);
            }
        }
        padBox(w);
        return w;
    }
    DInstance instanceRecord(const HostInstance& in) const {
        DInstance d{};
        memcpy(d.o2w, in.transform, 48); memcpy(d.w2o, in.inverse, 48);
        d.customIndex = in.customIndex; d.visible = in.visible ? 1 : 0;
        if (!isIdentity(in.transform) && in.blas != 0xFFFFFFFFu) { d.blasRoot = blases[in.blas].rootNode; d.pad[0] = in.blas; }
        d.pad[1] = in.tlasSlot;
        return d;
    }

    // re-uploads the device records of the edited instances (one 128-byte copy each; everything when there are many)
    void uploadDirtyRecords() {
        std::sort(dirtyInstances.begin(), dirtyInstances.end());
        dirtyInstances.erase(std::unique(dirtyInstances.begin(), dirtyInstances.end()), dirtyInstances.end());
        const uint32_t nInst = (uint32_t)instances.size();
        if (dirtyInstances.size() > 4096) {
            std::vector<DInstance> di(nInst + 1);
            for (uint32_t i = 0; i < nInst; i++) di[i] = instanceRecord(instances[i]);
            di[nInst] = staticRec;
            dInstances.upload(di.data(), di.size(), stream);
            CK(cudaStreamSynchronize(stream));
        } else {
            std::vector<DInstance> recs(dirtyInstances.size());
            for (size_t k = 0; k < dirtyInstances.size(); k++) {
                recs[k] = instanceRecord(instances[dirtyInstances[k]]);
                CK(cudaMemcpyAsync(dInstances.p + dirtyInstances[k], &recs[k], sizeof(DInstance), cudaMemcpyHostToDevice, stream));
            }
            CK(cudaStreamSynchronize(stream));   // recs die here
        }
    }
    // (2b) transform-only edits: same TLAS topology; the edited instances' boxes are recomputed and carried up the tree ON THE
    //      DEVICE (kernels.cu k_tlas_boxes / k_tlas_propagate) - what crosses the bus is 132 B per edited instance, whatever the
    //      size of the scene. Hit results do not depend on the tree (closest = min t with id tie-break), only traversal cost
    //      does; after `maxRefits` refits the TLAS is rebuilt.
    void refitTlas() {
        uploadDirtyRecords();
        dDirty.upload(dirtyInstances.data(), dirtyInstances.size(), stream);
        TlasRefit r{dTlasNodes.p, dInstances.p, dBlasBounds.p, dSlotNode.p, dNodeParent.p, dNodeExpected.p, dNodeArrived.p};
        launchTlasRefit(stream, r, dDirty.p, (uint32_t)dirtyInstances.size(), (uint32_t)tlasHostSlots.size());
        stats.kernel_launches += 2;
        CK(cudaStreamSynchronize(stream));
        refitsSinceBuild++;
        stats_tlas_refits++;
    }

    // (2) instance level: device instance records, inverse transforms and the TLAS (Accel.zig:394-489 equivalent)
    void buildTlas() {
        const uint32_t nInst = (uint32_t)instances.size();
        std::vector<DInstance> di(nInst + 1);
        std::vector<BuildBox> boxes;      // TLAS item bounds (world space)
        std::vector<uint32_t> items;      // TLAS item -> instance record
        di[nInst] = staticRec;
        if (haveStatic) { boxes.push_back(staticBounds); items.push_back(nInst); }
        // transformed instances created since the last BLAS-level build: append their BLAS (or reuse a shared one)
        const size_t poolNodes = blasNodePool.size();
        for (HostInstance& in : instances)
            if (in.blas == 0xFFFFFFFFu && !isIdentity(in.transform)) in.blas = getBlas(in.meshKey);
        if (blasNodePool.size() != poolNodes) {
            dBlasNodes.upload(blasNodePool.data(), blasNodePool.size(), stream);
            dTris.upload(triPool.data(), triPool.size(), stream);
        }
        for (uint32_t i = 0; i < nInst; i++) {
            HostInstance& in = instances[i];
            in.tlasSlot = 0xFFFFFFFFu;
            if (in.visible && !isIdentity(in.transform)) { boxes.push_back(instanceWorldBox(in)); items.push_back(i); }
        }
        tlasItemCount = (uint32_t)items.size(); hiddenTlasItems = 0;
        {   // bounds of everything visible: the 16^3 grid of the ray-sorting keys (ordering only, never a result)
            BuildBox all; for (int a = 0; a < 3; a++) { all.lo[a] = 1e30f; all.hi[a] = -1e30f; }
            for (const BuildBox& b : boxes) for (int a = 0; a < 3; a++) { all.lo[a] = std::min(all.lo[a], b.lo[a]); all.hi[a] = std::max(all.hi[a], b.hi[a]); }
            float ext[3];
            for (int a = 0; a < 3; a++) { ext[a] = boxes.empty() ? 1.0f : std::max(all.hi[a] - all.lo[a], 1e-20f); sortFrame.lo[a] = boxes.empty() ? 0.0f : all.lo[a]; sortFrame.bits[a] = 3; }
            // 3 bits per axis (even the thin axis of a flat scene separates ray populations: cfg5 loses 3 % without them), then
            // the remaining 3 bits go to the axes whose cells are still the longest
            for (int n = 0; n < 3; n++) {
                int best = 0;
                for (int a = 1; a < 3; a++) if (ext[a] / float(1u << sortFrame.bits[a]) > ext[best] / float(1u << sortFrame.bits[best])) best = a;
                sortFrame.bits[best]++;
            }
            for (int a = 0; a < 3; a++) sortFrame.scale[a] = float(1u << sortFrame.bits[a]) / ext[a];
        }
        hasTlas = !items.empty();
        tlasStaticOnly = haveStatic && items.size() == 1;
        if (hasTlas) {
            BuiltBvh tl;
            buildDepthLimited(boxes, 1, (uint32_t)MSNE_STACK - 2u - maxBlasDepth, tl);
            tlasDepth = tl.depth;
            std::vector<uint32_t> slots(tl.order.size());
            for (size_t s = 0; s < tl.order.size(); s++) { slots[s] = items[tl.order[s]]; if (slots[s] < nInst) instances[slots[s]].tlasSlot = (uint32_t)s; }
            // what the device-side refit needs: where every leaf slot lives, every node's parent, how many children it has
            std::vector<uint32_t> slotNode(slots.size(), 0), parent(tl.nodes.size(), 0xFFFFFFFFu), expected(tl.nodes.size(), 0);
            for (size_t k = 0; k < tl.nodes.size(); k++) {
                BvhNode2& n = tl.nodes[k];
                for (int s = 0; s < 2; s++) {
                    const int32_t child = n.c[s], count = n.c[2 + s];
                    if (count > 0) { n.c[s] = REF_ITEM | child; slotNode[(size_t)child] = (uint32_t)(2 * k + s); expected[k]++; }   // leaf size is 1: one instance per TLAS leaf
                    else if (child < 0) { n.c[s] = REF_EMPTY; setEmptyChild(n, s); }
                    else { parent[(size_t)child] = (uint32_t)(2 * k + s); expected[k]++; }
                }
            }
            for (BvhNode2& n : tl.nodes) n.c[2] = n.c[3] = 0;
            dTlasNodes.upload(tl.nodes.data(), tl.nodes.size(), stream);
            dTlasItems.upload(slots.data(), slots.size(), stream);
            dSlotNode.upload(slotNode.data(), slotNode.size(), stream);
            dNodeParent.upload(parent.data(), parent.size(), stream);
            dNodeExpected.upload(expected.data(), expected.size(), stream);
            dNodeArrived.ensure(tl.nodes.size());
            CK(cudaMemsetAsync(dNodeArrived.p, 0, tl.nodes.size() * sizeof(uint32_t), stream));
            std::vector<float> bb(6 * std::max<size_t>(1, blases.size()), 0.0f);
            for (size_t b = 0; b < blases.size(); b++) for (int a = 0; a < 3; a++) { bb[6 * b + a] = blases[b].bounds.lo[a]; bb[6 * b + 3 + a] = blases[b].bounds.hi[a]; }
            dBlasBounds.upload(bb.data(), bb.size(), stream);
            CK(cudaStreamSynchronize(stream));   // the staging vectors die here
            tlasHostNodes.swap(tl.nodes); tlasHostSlots.swap(slots);
        } else { tlasHostNodes.clear(); tlasHostSlots.clear(); }
        for (uint32_t i = 0; i < nInst; i++) di[i] = instanceRecord(instances[i]);   // after the slots are known
        dInstances.upload(di.data(), di.size(), stream);
        CK(cudaStreamSynchronize(stream));
        dirtyInstances.clear();
        refitsSinceBuild = 0;
        stats_tlas_builds++;
    }

    // fetchTexel (shading.cuh) of texel (0,0) on the host: same table values, same conversions -> the bits the sampler returns
    static void decodeFirstTexel(const HostTexture& t, float out[4]) {
        auto unorm = [](uint8_t c) { return (float)c / 255.0f; };
        auto half = [](uint16_t h) {
            const uint32_t sign = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
            uint32_t bits;
            if (e == 0) {
                if (m == 0) bits = sign;
                else { int sh = 0; uint32_t mm = m; while (!(mm & 1024u)) { mm <<= 1; sh++; } bits = sign | ((uint32_t)(113 - sh) << 23) | ((mm & 1023u) << 13); }
            } else if (e == 31) bits = sign | 0x7F800000u | (m << 13);
            else bits = sign | ((e + 112u) << 23) | (m << 13);
            float f; memcpy(&f, &bits, 4); return f;
        };
        float f[4] = {0, 0, 0, 0}; memcpy(f, t.first, 16);
        uint16_t hh[4]; memcpy(hh, t.first, 8);
        out[0] = out[1] = out[2] = 0.0f; out[3] = 1.0f;
        switch (t.fmt) {
            case TEX_R32F: out[0] = f[0]; break;
            case TEX_RG32F: out[0] = f[0]; out[1] = f[1]; break;
            case TEX_RGBA32F: out[0] = f[0]; out[1] = f[1]; out[2] = f[2]; out[3] = f[3]; break;
            case TEX_R8_UNORM: out[0] = unorm(t.first[0]); break;
            case TEX_RG8_UNORM: out[0] = unorm(t.first[0]); out[1] = unorm(t.first[1]); break;
            case TEX_RGBA8_SRGB: out[0] = srgbToLinear(t.first[0]); out[1] = srgbToLinear(t.first[1]); out[2] = srgbToLinear(t.first[2]); out[3] = unorm(t.first[3]); break;
            default: for (int i = 0; i < 4; i++) out[i] = half(hh[i]); break;   // TEX_RGBA16F
        }
    }
    void resolveMaterials() {
        for (DMaterial& m : materials) {
            const uint32_t slots[5] = {m.normal, m.emissive, m.color, m.metalness, m.roughness};
            m.constMask = 0;
            for (int s = 0; s < 5; s++) {
                for (int c = 0; c < 4; c++) m.constant[s][c] = 0.0f;
                if (slots[s] >= textures.size()) continue;
                const HostTexture& t = *textures[slots[s]];
                if (t.w != 1 || t.h != 1) continue;
                decodeFirstTexel(t, m.constant[s]);
                m.constMask |= 1u << s;
            }
        }
    }

    bool sceneDirty() const {
        return tablesDirty || geomDirty || tlasDirty || aliasDirty || !dirtyInstances.empty() || (emitterFollowsEdits && emitterStale) ||
               emitterRecIndexed != (consts.indexed_attributes != 0);
    }
    void syncScene() {
        if (tablesDirty) {
            std::vector<DMesh> dm(meshes.size());
            for (size_t i = 0; i < meshes.size(); i++) dm[i] = DMesh{meshes[i]->dPos.p, meshes[i]->dUv.p, meshes[i]->dNrm.p, meshes[i]->dIdx.p};
            dMeshes.upload(dm.data(), dm.size(), stream);
            std::vector<DTexture> dt(textures.size());
            for (size_t i = 0; i < textures.size(); i++) dt[i] = DTexture{textures[i]->data.p, textures[i]->w, textures[i]->h, textures[i]->fmt, 0};
            dTextures.upload(dt.data(), dt.size(), stream);
            resolveMaterials();
            dMaterials.upload(materials.data(), materials.size(), stream);
            dGeometries.upload(geometries.data(), geometries.size(), stream);
            CK(cudaStreamSynchronize(stream));   // staging vectors die here
            tablesDirty = false;
        }
        const auto t0 = std::chrono::steady_clock::now();
        const bool accelWork = geomDirty || tlasDirty || !dirtyInstances.empty();
        if (geomDirty) { buildGeometry(); geomDirty = false; tlasDirty = true; tlasRefitOk = false; }
        if (tlasDirty) {
            if (tlasRefitOk && hasTlas && refitsSinceBuild < maxRefits) refitTlas(); else buildTlas();
            tlasDirty = false; tlasRefitOk = false;
        } else if (!dirtyInstances.empty()) uploadDirtyRecords();   // visibility toggles: records only
        dirtyInstances.clear();
        if (accelWork) stats.accel_sync_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (triFramesDirty || triFramesIndexed != (consts.indexed_attributes != 0)) buildTriFrames();
        if (emitterFollowsEdits && emitterStale) { buildEmitterTable(); emitterStale = false; }
        if (aliasDirty || emitterRecIndexed != (consts.indexed_attributes != 0)) {
            dAlias.upload(alias.data(), alias.size(), stream);
            uploadEmitterRecords();
            CK(cudaStreamSynchronize(stream));
            aliasDirty = false;
        }
    }

    // frame records of the pooled triangles (shading.cuh MSNE_TRI_FRAMES): needs the mesh table, the pooled triangles and the instance
    // records on the device; depends on indexed_attributes (which texcoords a corner reads)
    void buildTriFrames() {
        triFramesIndexed = consts.indexed_attributes != 0;
        triFramesDirty = false;
        if (triPool.empty()) return;
        DevBuf<uint32_t> dTriMesh;
        dTriMesh.upload(triMeshPool.data(), triMeshPool.size(), stream);
        dTriFrames.ensure(4 * triPool.size());
        launchTriFrames(stream, deviceScene(), deviceConstants(), dTriMesh.p, (uint32_t)triPool.size(), dTriFrames.p);
        stats.kernel_launches += 1;
        CK(cudaStreamSynchronize(stream));   // dTriMesh dies here
    }

    DeviceScene deviceScene() const {
        DeviceScene sc{};
        sc.tlasNodes = hasTlas ? dTlasNodes.p : nullptr;
        sc.tlasItems = dTlasItems.p;
        sc.blasNodes = dBlasNodes.p; sc.tris = dTris.p; sc.triFrames = dTriFrames.p; sc.instances = dInstances.p; sc.geometries = dGeometries.p;
        sc.meshes = dMeshes.p; sc.materials = dMaterials.p; sc.textures = dTextures.p;
        sc.alias = dAlias.p; sc.aliasEntries = (uint32_t)alias.size(); sc.emitterTris = dEmitterTris.p;
        sc.envRgb = reinterpret_cast<const float*>(dEnvRgb.p); sc.envLum = dEnvLum.p; sc.envSize = envSize; sc.envLevels = envLevels;
        for (int i = 0; i < 12; i++) { sc.envLevelOffset[i] = envLevelOffset[i]; sc.envQuadOffset[i] = envQuadOffset[i]; }
        sc.envQuads = dEnvQuads.p;
        sc.texLuts = dSrgbLut.p;
        sc.staticOnly = (tlasStaticOnly && !forceTwoLevel) ? 1u : 0u; sc.staticRoot = staticRec.blasRoot;
        sc.staticHidden = hiddenStaticTris ? 1u : 0u;
        return sc;
    }
    DConstants deviceConstants() const {
        return DConstants{consts.samples_per_run, consts.max_bounces, consts.env_samples_per_bounce, consts.mesh_samples_per_bounce,
                          consts.flip_image, consts.indexed_attributes, consts.two_component_normal_texture};
    }

    // Accel.zig:491-538
    void buildEmitterTable() {
        std::vector<float> weights;
        std::vector<DAlias> entries;
        for (uint32_t i = 0; i < instances.size(); i++) {
            const HostInstance& in = instances[i];
            for (uint32_t j = 0; j < in.geometryCount; j++) {
                const DGeometry& g = geometries[in.customIndex + j];
                if (!(g.sampled & 0xFF)) continue;
                const HostMesh& m = *meshes[g.mesh];
                const uint32_t nTri = (uint32_t)(m.idx.size() / 3);
                for (uint32_t k = 0; k < nTri; k++) {
                    float p[3][3];
                    for (int v = 0; v < 3; v++) {
                        const float* q = &m.pos[3 * (size_t)m.idx[3 * k + v]];
                        // Mat3x4.mul_point (vector.zig:283-288): row.dot((q,1)) = x*x' + y*y' + z*z' + w*1
                        for (int r = 0; r < 3; r++) p[v][r] = in.transform[4 * r] * q[0] + in.transform[4 * r + 1] * q[1] + in.transform[4 * r + 2] * q[2] + in.transform[4 * r + 3] * 1.0f;
                    }
                    const float a[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]}, b[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
                    const float c[3] = {a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]};
                    const float area = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) / 2.0f;
                    weights.push_back(area);
                    entries.push_back(DAlias{0, 0.0f, i, j, k});
                }
            }
        }
        const float sum = buildAliasTable(weights, entries);
        alias.clear();
        alias.push_back(DAlias{(uint32_t)entries.size(), sum, 0, 0, 0});
        alias.insert(alias.end(), entries.begin(), entries.end());
        aliasDirty = true;
    }

    // Per alias entry: the triangle's object-space vertices and texcoords exactly as MeshAttributes::lookupAndInterpolate
    // (world.hlsl:114-136) would fetch them for (instance, geometry, primitive) - including the float round trip of the
    // attribute indices (:130) and the default texcoords of a mesh without uvs - packed into 64 B (dtypes.h emitterTris).
    DevBuf<float4> dEmitterTris;
    bool emitterRecIndexed = false;
    void uploadEmitterRecords() {
        emitterRecIndexed = consts.indexed_attributes != 0;
        const size_t n = alias.empty() ? 0 : alias.size() - 1;
        std::vector<float4> rec(4 * std::max<size_t>(n, 1), float4{0, 0, 0, 0});
        for (size_t e = 0; e < n; e++) {
            const DAlias& a = alias[1 + e];
            const HostInstance& in = instances[a.inst];
            const HostMesh& m = *meshes[geometries[in.customIndex + a.geom].mesh];
            float p[3][3] = {{0}}, t[3][2] = {{0, 0}, {1, 0}, {1, 1}};
            for (int v = 0; v < 3; v++) {
                const uint32_t vi = m.idx[3 * (size_t)a.prim + v];
                memcpy(p[v], &m.pos[3 * (size_t)vi], 12);
                const uint32_t ai = emitterRecIndexed ? (uint32_t)(float)vi : (uint32_t)(float)(a.prim * 3u + (uint32_t)v);
                if (!m.uv.empty()) { t[v][0] = t[v][1] = 0.0f; if (2 * (size_t)ai + 1 < m.uv.size()) { t[v][0] = m.uv[2 * (size_t)ai]; t[v][1] = m.uv[2 * (size_t)ai + 1]; } }
            }
            rec[4 * e + 0] = float4{p[0][0], p[0][1], p[0][2], t[0][0]};
            rec[4 * e + 1] = float4{p[1][0], p[1][1], p[1][2], t[0][1]};
            rec[4 * e + 2] = float4{p[2][0], p[2][1], p[2][2], t[1][0]};
            rec[4 * e + 3] = float4{t[1][1], t[2][0], t[2][1], 0.0f};
        }
        dEmitterTris.upload(rec.data(), rec.size(), stream);
        CK(cudaStreamSynchronize(stream));   // rec dies here
    }

    // ---- camera (camera.hlsl:14-30 hoisted) ----------------------------------------------------
    static DCamera makeCamera(const Lens& c, uint32_t W, uint32_t H) {
        auto cross = [](const float* a, const float* b, float* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
        DCamera cam{};
        const float aspect = float(W) / float(H);
        const float fw[3] = {c.forward.x, c.forward.y, c.forward.z}, up[3] = {c.up.x, c.up.y, c.up.z}, org[3] = {c.origin.x, c.origin.y, c.origin.z};
        float w[3] = {fw[0] * -1.0f, fw[1] * -1.0f, fw[2] * -1.0f};
        float u[3], v[3];
        cross(up, w, u);
        const float inv = 1.0f / std::sqrt((u[0] * u[0] + u[1] * u[1]) + u[2] * u[2]);
        for (int i = 0; i < 3; i++) u[i] = u[i] * inv;
        cross(w, u, v);
        const float h = std::tan(c.vfov / 2.0f);
        const float viewport_height = 2.0f * h * c.focus_distance;
        const float viewport_width = aspect * viewport_height;
        for (int i = 0; i < 3; i++) {
            cam.origin[i] = org[i]; cam.u[i] = u[i]; cam.v[i] = v[i];
            cam.horizontal[i] = u[i] * viewport_width;
            cam.vertical[i] = v[i] * viewport_height;
            cam.llc[i] = org[i] - (cam.horizontal[i] / 2.0f) - (cam.vertical[i] / 2.0f) - (w[i] * c.focus_distance);
        }
        cam.aperture = c.aperture; cam.width = W; cam.height = H;
        return cam;
    }

    // ---- sensors -------------------------------------------------------------------------------
    // owner of pixel tile (cx, cy): consecutive tiles of a row go to consecutive ranks, and every row starts `stagger` ranks
    // further, stagger = the smallest number >= 2 coprime to the world size (3 for 2, 4 and 8 ranks): neither a tile column
    // nor a tile row is served by one rank
    static uint32_t tileStagger(uint32_t world) {
        auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
        uint32_t s = 2;
        while (world > 1 && gcd(s, world) != 1) s++;
        return world == 2 ? 3u : s;
    }
    static uint32_t tileOwner(uint32_t cx, uint32_t cy, uint32_t world) { return (cx + tileStagger(world) * cy) % world; }
    void buildPixIds(Sensor& s) {
        // tileEdge x tileEdge pixel tiles dealt round-robin over ranks; inside a tile 8x4 blocks so that a warp of 32
        // consecutive queue slots is a compact 8x4 pixel block (coherent primary rays). Small tiles even out the per-rank work
        // (a 32-pixel tile can be all sky or all chrome); the edge stays a multiple of 8 so the 8x4 blocks are never split.
        const uint32_t W = s.extent.width, H = s.extent.height, E = tileEdge;
        const uint32_t tx = (W + E - 1) / E, ty = (H + E - 1) / E;
        std::vector<uint32_t> ids;
        ids.reserve((size_t)W * H / std::max(1u, shardWorld) + 1024);
        for (uint32_t t = 0; t < tx * ty; t++) {
            if (tileOwner(t % tx, t / tx, shardWorld) != shardRank) continue;
            const uint32_t x0 = (t % tx) * E, y0 = (t / tx) * E;
            for (uint32_t by = 0; by < E; by += 4)
                for (uint32_t bx = 0; bx < E; bx += 8)
                    for (uint32_t y = 0; y < 4; y++)
                        for (uint32_t x = 0; x < 8; x++) {
                            const uint32_t px = x0 + bx + x, py = y0 + by + y;
                            if (px < W && py < H) ids.push_back(py * W + px);
                        }
        }
        s.nPix = (uint32_t)ids.size();
        s.pixIds.upload(ids.data(), ids.size(), stream);
        s.runState.alloc(std::max<size_t>(1, ids.size()));
        CK(cudaMemsetAsync(s.runState.p, 0, std::max<size_t>(1, ids.size()) * sizeof(float4), stream));
        CK(cudaStreamSynchronize(stream));
        s.shardRank = shardRank; s.shardWorld = shardWorld;
    }
    void clearSensor(Sensor& s) {
        s.sample_count = 0;   // Sensor.clear (core/Sensor.zig:81-83); partial launch sums are dropped too
        if (s.runState.p) CK(cudaMemsetAsync(s.runState.p, 0, s.runState.n * sizeof(float4), stream));
    }
    void clearAllSensors() { for (auto& s : sensors) clearSensor(*s); }

    // ---- wavefront pool --------------------------------------------------------------------------
    void ensurePool(WaveCtx& c, uint64_t wanted) {
        const uint32_t P = (uint32_t)std::min<uint64_t>(wanted, 0x7FFFFFFFull / 4);
        if (c.pb.capacity >= P && c.pb.capacity != 0) return;
        finishWaves();
        // poolBufs: [0],[1] interleaved ray lanes R0/R1 of the two ping-pong buffers (dtypes.h RAY_STRIDE), [2..9] S0,S1,N0,N1 x 2,
        // [10] interleaved shadow-ray lanes Q0/Q1 (2 P rays), [11] result
        CK(cudaStreamSynchronize(c.stream));
        c.poolBufs.clear();
        c.poolBufs.resize(12);
        for (int i = 0; i < 2; i++) c.poolBufs[i].alloc((size_t)RAY_STRIDE * P);
        for (int i = 2; i < 10; i++) c.poolBufs[i].alloc(P);
        c.poolBufs[10].alloc((size_t)RAY_STRIDE * 2 * P);
        c.poolBufs[11].alloc(P);
        c.poolH.alloc(P);
        c.poolTypeList.alloc((size_t)NUM_BINS * P);
        c.sortKeysExt.release();   // sized with the pool (ensureSortBuffers)
        PathBuffers& pb = c.pb;
        for (int b = 0; b < 2; b++) {
            pb.R0[b] = c.poolBufs[b].p; pb.R1[b] = c.poolBufs[b].p + 1; pb.S0[b] = c.poolBufs[2 + b].p; pb.S1[b] = c.poolBufs[4 + b].p;
            pb.N0[b] = c.poolBufs[6 + b].p; pb.N1[b] = c.poolBufs[8 + b].p;
        }
        pb.Q0 = c.poolBufs[10].p; pb.Q1 = c.poolBufs[10].p + 1; pb.result = c.poolBufs[11].p; pb.H = c.poolH.p; pb.typeList = c.poolTypeList.p; pb.capacity = P;
    }
    // bytes of device memory held by the wave pools (path state, queues, sort and continuation buffers) of all contexts
    uint64_t poolBytes() const {
        uint64_t b = 0;
        for (const WaveCtx& c : ctx) {
            for (const auto& buf : c.poolBufs) b += buf.n * sizeof(float4);
            b += c.poolH.n * sizeof(uint4) + c.poolTypeList.n * 4 + c.sortKeysExt.n * 2 + c.sortKeysShadow.n * 2 + c.sortPermExt.n * 4 + c.sortPermShadow.n * 4 +
                 c.sortBlockHist.n * 4 + c.contA.n * sizeof(uint4) + c.contB.n * sizeof(uint4);
        }
        return b;
    }
    // gives the wave pools back to the driver (they are sized by the largest render so far and never shrink on their own); the next
    // render allocates what it needs
    void releasePools() {
        finishWaves();
        for (WaveCtx& c : ctx) {
            if (c.stream) CK(cudaStreamSynchronize(c.stream));
            c.poolBufs.clear(); c.poolH.release(); c.poolTypeList.release();
            c.sortKeysExt.release(); c.sortKeysShadow.release(); c.sortPermExt.release(); c.sortPermShadow.release(); c.sortBlockHist.release();
            c.contA.release(); c.contB.release();
            c.pb = PathBuffers{};
        }
    }
    void ensureSortBuffers(WaveCtx& c) {
        const size_t P = c.pb.capacity;
        if (c.sortKeysExt.n >= P) return;
        finishWaves();
        c.sortKeysExt.alloc(P); c.sortKeysShadow.alloc(2 * P);
        c.sortPermExt.alloc(P); c.sortPermShadow.alloc(2 * P);
        c.sortBlockHist.ensure((size_t)2 * smCount * 32768);
    }
    // continuation records of suspend-and-regroup tracing (kernels.h ContBuffers): a ticket files at most `suspendThreshold` records
    void ensureContBuffers(WaveCtx& c) {
        const size_t P = c.pb.capacity, T = suspendThreshold;
        const size_t capA = ((3 * P + 31) / 32 + 1) * T + 32, capB = ((capA + 31) / 32 + 1) * T + 32;
        if (c.contA.n >= capA * 8 && c.contB.n >= capB * 8 && c.contCounts.n) return;
        finishWaves();
        c.contA.alloc(capA * 8); c.contB.alloc(capB * 8);
        if (!c.contCounts.n) { c.contCounts.alloc(2); CK(cudaMemsetAsync(c.contCounts.p, 0, 2 * sizeof(uint32_t), stream)); CK(cudaStreamSynchronize(stream)); }
    }
    static SortBuffers sortBuffers(const WaveCtx& c) {
        return SortBuffers{c.sortKeysExt.p, c.sortKeysShadow.p, c.sortPermExt.p, c.sortPermShadow.p, c.sortHist.p, c.sortCursor.p, c.sortBlockHist.p};
    }
    static WaveState waveState(const WaveCtx& c) {
        WaveState ws{};
        uint32_t* sc = c.dScalars.p;
        ws.counters = c.dCounters.p; ws.traceTicket = sc + 0; ws.traceDone = sc + 1; ws.shadeDone = sc + 2;
        ws.maxBounce = sc + 3; ws.rayCounts = c.dRayCounts.p; ws.typeCount = sc + 4; ws.binDone = sc + 9; ws.history = c.dHistory.p;
        return ws;
    }

    // live paths expected to ENTER bounce b of a wave of n paths, from the last finished wave's log; with no log yet: everything
    // survives the first four bounces (Russian roulette starts after bounce 3, integrator.hlsl:130), then 72 % per bounce
    double expectedIn(uint32_t b, uint64_t n) const {
        if (b == 0) return (double)n;
        if (b < survivalIn.size()) return survivalIn[b] * (double)n;
        if (survivalIn.size() > 1) return survivalIn.back() * std::pow(0.8, (double)(b - (survivalIn.size() - 1))) * (double)n;
        return (double)n * (b < 4 ? 1.0 : std::pow(0.72, (double)(b - 3)));
    }

    // reads the log of the wave last enqueued on `c` (blocks until that wave is done): survival profile, per-kernel timings
    void harvest(WaveCtx& c) {
        if (!c.pending) return;
        CK(cudaEventSynchronize(c.evDone));
        c.pending = false;
        stats.waves++;
        if (c.bounces && c.paths) {
            survivalIn.assign(1, 1.0);
            for (uint32_t b = 0; b < c.bounces; b++) survivalIn.push_back((double)c.hostHistory[b].nLive / (double)c.paths);
            const Counters& last = c.hostHistory[c.bounces - 1];
            stats.tail_paths += (uint64_t)last.nLive + last.nDead;
        }
        if (logBounces)
            for (uint32_t b = 0; b < c.bounces; b++)
                fprintf(stderr, "libmoonshine: bounce %u -> live %u dead %u shadow %u\n", b, c.hostHistory[b].nLive, c.hostHistory[b].nDead, c.hostHistory[b].nShadow);
        for (int i = 0; i < c.profUsed; i++) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, c.profEv[i][0], c.profEv[i][1]));
            stats.trace_ms += ms; stats.trace_launches++;
            if (c.profSorted[i]) {
                CK(cudaEventElapsedTime(&ms, c.profEv[i][2], c.profEv[i][0])); stats.sort_ms += ms;
                CK(cudaEventElapsedTime(&ms, c.profEv[i][1], c.profEv[i][3])); stats.sort_ms += ms;
            }
        }
        c.profUsed = 0;
    }

    // one wave: `nSamples` consecutive samples of `nPix` owned pixels starting at pixIds + pixOffset, enqueued on c.stream without
    // any host synchronisation; `afterAccum` (may be null): the previous wave's accumulate, which this wave's must follow
    // `withAccumulate` = false: the wave stops short of adding its samples to the sensor (render-ahead: enqueueAccumulate commits it later)
    void enqueueWave(WaveCtx& c, Sensor& s, const DCamera& cam, const DeviceScene& sc, const DConstants& k, uint32_t pixOffset, uint32_t nPix,
                     uint32_t firstSample, uint32_t nSamples, cudaEvent_t afterAccum, bool withAccumulate = true) {
        WaveState ws = waveState(c);
        if (!withAccumulate) ws.rayCounts = c.dRayCounts.p + 2;   // a wave that may be dropped counts its rays aside until it is claimed
        const PathBuffers& pb = c.pb;
        cudaStream_t st = c.stream;
        const uint64_t nPaths = (uint64_t)nPix * nSamples;
        const bool inlineNee = k.env_samples > 1 || k.mesh_samples > 1;
        launchGenerate(st, cam, k, pb, c.dCounters.p, s.pixIds.p + pixOffset, nPix, nSamples, firstSample);
        stats.kernel_launches++;
        // how many bounces run as wavefront launches before k_tail finishes the stragglers
        const double tailAt = (double)std::max<uint64_t>(tailThreshold, 1);
        uint32_t K = 1;
        const uint32_t kCap = (uint32_t)std::min<uint64_t>((uint64_t)k.max_bounces + 2, kMaxHistory);
        while (K < kCap && expectedIn(K, nPaths) >= tailAt) K++;
        auto sortedAt = [&](uint32_t b) { return b >= 1 && sortActive && expectedIn(b, nPaths) >= (double)sortThreshold; };
        c.profUsed = 0;
        for (uint32_t b = 0; b < K; b++) {
            const uint32_t buf = b & 1u;
            const bool prof = profiling && c.profUsed < kProfEvents;
            const bool sorted = sortedAt(b);
            if (prof) c.profSorted[c.profUsed] = sorted;
            if (sorted) {
                if (prof) CK(cudaEventRecord(c.profEv[c.profUsed][2], st));
                launchSortRays(st, smCount, sortBuffers(c), c.dCounters.p + buf);
                if (prof) CK(cudaEventRecord(c.profEv[c.profUsed][0], st));
                if (suspendThreshold && sc.staticOnly && c.contCounts.n) {
                    launchTraceSuspend(st, traceBlocks, sc, pb, ws, buf, c.sortPermExt.p, c.sortPermShadow.p, ContBuffers{c.contA.p, c.contB.p, c.contCounts.p}, suspendThreshold, suspendPasses);
                    stats.kernel_launches += (uint64_t)suspendPasses - 1;
                } else launchTraceSorted(st, traceBlocks, sc, pb, ws, buf, c.sortPermExt.p, c.sortPermShadow.p);
                if (prof) CK(cudaEventRecord(c.profEv[c.profUsed][1], st));
                launchBin(st, smCount, pb, ws, buf);
                if (prof) { CK(cudaEventRecord(c.profEv[c.profUsed][3], st)); c.profUsed++; }
                stats.kernel_launches += 5;   // count, scan, scatter, trace, bin
            } else {
                if (prof) CK(cudaEventRecord(c.profEv[c.profUsed][0], st));
                launchTrace(st, traceBlocks, sc, pb, ws, buf);
                if (prof) { CK(cudaEventRecord(c.profEv[c.profUsed][1], st)); c.profUsed++; }
                stats.kernel_launches += 1;
            }
            // the rays this pass queues get their sort keys from it when the next bounce is going to be traced in sorted order
            SortKeys keys{};
            if (b + 1 < K && sortedAt(b + 1)) { keys.ext = c.sortKeysExt.p; keys.shadow = c.sortKeysShadow.p; keys.frame = sortFrame; }
            // big queues: one material-specialised launch per bin; small queues: one generic launch
            if (inlineNee) { launchShadeInlineNee(st, shadeInlineBlocks, sc, k, pb, ws, buf, (int32_t)b, keys); stats.kernel_launches += 1; }
            else if (expectedIn(b, nPaths) >= (double)binnedThreshold) { launchShadeBinned(st, shadeBlocks, sc, k, pb, ws, buf, (int32_t)b, keys); stats.kernel_launches += 5; }
            else { launchShade(st, shadeBlocks, sc, k, pb, ws, buf, (int32_t)b, keys); stats.kernel_launches += 1; }
        }
        // whatever is still alive (normally < tailThreshold paths; more only when the profile was off): trace once more and let
        // k_tail finish every remaining path in a single launch instead of ~130 launch pairs at the latency floor
        launchTrace(st, traceBlocks, sc, pb, ws, K & 1u);
        launchTail(st, smCount * 4 * (128 / tailBlockThreads), tailBlockThreads, sc, k, pb, ws, K & 1u, inlineNee);
        stats.kernel_launches += 2;
        CK(cudaMemcpyAsync(c.hostHistory, c.dHistory.p, K * sizeof(Counters), cudaMemcpyDeviceToHost, st));
        c.bounces = K; c.paths = nPaths;
        if (withAccumulate) enqueueAccumulate(c, s, k, pixOffset, nPix, firstSample, nSamples, afterAccum);
    }
    // the last launch of a wave: its samples join the sensor's running mean (main.hlsl:43-51), in sample order
    void enqueueAccumulate(WaveCtx& c, Sensor& s, const DConstants& k, uint32_t pixOffset, uint32_t nPix, uint32_t firstSample, uint32_t nSamples, cudaEvent_t afterAccum, bool deferred = false) {
        if (afterAccum) CK(cudaStreamWaitEvent(c.stream, afterAccum, 0));
        if (deferred) { launchCommitCounts(c.stream, c.dRayCounts.p); stats.kernel_launches++; }   // the rays the wave counted aside
        launchAccumulate(c.stream, c.pb.result, s.image.p, s.runState.p + pixOffset, s.pixIds.p + pixOffset, nPix, nSamples, firstSample, k.samples_per_run);
        stats.kernel_launches++;
        CK(cudaEventRecord(c.evDone, c.stream));
        c.pending = true;
    }

    // ---- render-ahead (HdMoonshineExtSetRenderAhead) ------------------------------------------------------------------------------
    // The Hydra delegate renders ONE sample per HdMoonshineRender call and reads the sensor back after each (hydra.zig:316-358). A wave
    // of one sample ends in a latency-bound tail - a few thousand long paths, ~150 bounces deep - that leaves most of the GPU idle
    // for a third of the call. With render-ahead on, a call also enqueues the NEXT sample's wave (same sensor, lens and scene) on the
    // other wave context before it waits for its own, so that wave's bandwidth-bound head runs under this wave's tail; its samples
    // stay in the context's pool (no accumulate) until the next call claims them. A sample is a pure function of (scene, lens, sample
    // index), so the claimed wave is bit for bit what the call would have rendered itself; any other entry point, another sensor or
    // lens, or an edited scene drops the speculative wave unseen (one wave of GPU time lost).
    bool renderAhead = false;
    uint32_t aheadDepth = 2;   // waves kept in flight beyond the one a call waits for (1..kWaveContexts-1); MSNE_RENDER_AHEAD=<depth>
    struct AheadWave { int ctx; uint32_t firstSample; };
    std::deque<AheadWave> aheadQ;   // speculative waves, oldest first: rendered, not accumulated
    const Sensor* aheadSensor = nullptr;
    Lens aheadLens{};
    uint32_t aheadSpr = 0;
    void discardAhead() {
        if (aheadQ.empty()) return;
        use();
        for (const AheadWave& w : aheadQ) {
            WaveCtx& c = ctx[w.ctx];
            CK(cudaMemsetAsync(c.dRayCounts.p + 2, 0, 2 * sizeof(unsigned long long), c.stream));
            CK(cudaStreamSynchronize(c.stream));   // never accumulated: nothing to undo, the pool is simply reused
            stats.ahead_discarded++;
        }
        aheadQ.clear();
    }
    // true when this call pattern can be served by renderAheadCall: one wave of samples_per_run samples over the sensor's pixels
    bool aheadEligible(const Sensor& s) const {
        if (!renderAhead) return false;
        const uint64_t paths = (uint64_t)s.extent.width * s.extent.height * consts.samples_per_run;
        return paths > 0 && paths <= std::max<uint64_t>(1, pathsInFlight / 2) && paths <= 0x7FFFFFFFull / 4;
    }
    void renderAheadCall(Sensor& s, const Lens& lens) {
        use();
        const bool reshard = s.shardRank != shardRank || s.shardWorld != shardWorld || s.pixIds.p == nullptr;
        if (!aheadQ.empty() && (aheadSensor != &s || memcmp(&aheadLens, &lens, sizeof(Lens)) != 0 || aheadQ.front().firstSample != s.sample_count ||
                                aheadSpr != consts.samples_per_run || sceneDirty() || reshard))
            discardAhead();
        if (sceneDirty() || reshard) finishWaves();   // uploads must not race waves in flight
        if (reshard) { buildPixIds(s); clearSensor(s); }
        syncScene();
        const DeviceScene sc = deviceScene();
        const DConstants k = deviceConstants();
        const DCamera cam = makeCamera(lens, s.extent.width, s.extent.height);
        const uint32_t spr = k.samples_per_run;
        if (s.nPix == 0) { discardAhead(); s.sample_count += spr; return; }
        const uint64_t paths = (uint64_t)s.nPix * spr;
        const uint64_t wantP = (paths + 31) & ~31ull;
        const int depth = (int)std::min<uint32_t>(std::max<uint32_t>(aheadDepth, 1), kWaveContexts - 1), nCtx = depth + 1;
        for (int i = 0; i < nCtx; i++) if (ctx[i].pb.capacity < wantP) ensurePool(ctx[i], wantP);   // reallocation waits for (and drops) everything in flight
        size_t bvhNodes = blasNodePool.size() + tlasHostNodes.size();
        sortActive = sortThreshold != 0 && wantP >= sortThreshold && (sortAnyScene || bvhNodes >= kSortMinNodes);
        if (sortActive) for (int i = 0; i < nCtx; i++) { ensureSortBuffers(ctx[i]); if (suspendThreshold) ensureContBuffers(ctx[i]); }
        CK(cudaEventRecord(evStart, stream));
        for (int i = 0; i < nCtx; i++) CK(cudaStreamWaitEvent(ctx[i].stream, evStart, 0));   // uploads / clears queued on the main stream come first
        aheadSensor = &s; aheadLens = lens; aheadSpr = spr;
        struct TailGuard { uint64_t& ref; uint64_t saved; ~TailGuard() { ref = saved; } } tailGuard{tailThreshold, tailThreshold};
        if (!tailExplicit) tailThreshold = kAheadTailThreshold;
        auto launchNext = [&](uint32_t first, int busy) {   // sample `first` on a context that holds no unclaimed wave
            int c = -1;
            for (int i = 0; i < nCtx && c < 0; i++) {
                bool used = i == busy;
                for (const AheadWave& w : aheadQ) used = used || w.ctx == i;
                if (!used) c = i;
            }
            if (c < 0) throw std::runtime_error("render-ahead: no free wave context");
            harvest(ctx[c]);
            enqueueWave(ctx[c], s, cam, sc, k, 0, s.nPix, first, spr, nullptr, false);
            aheadQ.push_back(AheadWave{c, first});
        };
        if (aheadQ.empty()) launchNext(s.sample_count, -1);   // nothing rendered ahead: this call's own sample
        const AheadWave cur = aheadQ.front();
        aheadQ.pop_front();
        // keep `depth` samples in flight beyond this one before waiting: their bandwidth-bound heads run under this wave's tail
        uint32_t nextFirst = (aheadQ.empty() ? cur.firstSample : aheadQ.back().firstSample) + spr;
        while ((int)aheadQ.size() < depth) { launchNext(nextFirst, cur.ctx); nextFirst += spr; }
        // claim the current wave: its rays and samples count from here on
        enqueueAccumulate(ctx[cur.ctx], s, k, 0, s.nPix, s.sample_count, spr, nullptr, true);
        stats.samples += paths;
        s.sample_count += spr;
        CK(cudaStreamWaitEvent(stream, ctx[cur.ctx].evDone, 0));
        CK(cudaEventRecord(evStop, stream));
        CK(cudaStreamSynchronize(stream));
        stats.host_syncs++;
        harvest(ctx[cur.ctx]);
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, evStart, evStop));
        stats.device_ms += ms;
    }

    // offline/main.zig:131-165: n launches of samples_per_run samples each; sensor.sample_count advances
    void renderLaunches(Sensor& s, const Lens& lens, uint32_t nLaunches, bool wait) {
        use();
        discardAhead();   // this path reuses the wave contexts from the start
        if (sceneDirty() || s.shardRank != shardRank || s.shardWorld != shardWorld || s.pixIds.p == nullptr) finishWaves();   // uploads must not race the waves
        if (s.shardRank != shardRank || s.shardWorld != shardWorld || s.pixIds.p == nullptr) { buildPixIds(s); clearSensor(s); }
        syncScene();
        const DeviceScene sc = deviceScene();
        const DConstants k = deviceConstants();
        const DCamera cam = makeCamera(lens, s.extent.width, s.extent.height);
        const uint64_t totalSamples = (uint64_t)nLaunches * k.samples_per_run;
        if (s.nPix == 0 || totalSamples == 0) { s.sample_count += (uint32_t)totalSamples; if (wait) finishWaves(); return; }
        const uint64_t totalPaths = (uint64_t)s.nPix * totalSamples;
        // pool per context: half of the paths in flight; a render that fits one context uses only that one
        const int nCtx = waveContextsInUse;
        const uint64_t perCtx = std::max<uint64_t>(1, pathsInFlight / (uint64_t)nCtx);
        // A render that would fit one wave is still cut into one wave per context (when there is enough of it): the contexts'
        // streams run concurrently, so the latency-bound end of every wave - a handful of long paths - overlaps the others'
        // instead of being waited for in full. This is what a one-sample HdMoonshineRender call (hydra.zig:316-329) gets.
        const uint64_t split = totalPaths >= (uint64_t)nCtx * kMinSplitWave ? (totalPaths + (uint64_t)nCtx - 1) / (uint64_t)nCtx : totalPaths;
        const uint64_t wantP = std::min<uint64_t>(perCtx, (split + 31) & ~31ull);
        ensurePool(ctx[0], wantP);
        const uint32_t P = (uint32_t)std::min<uint64_t>(ctx[0].pb.capacity, wantP);
        const uint32_t pixChunk = std::min(s.nPix, P);
        const uint32_t samplesPerWave = std::max(1u, P / pixChunk);
        const uint64_t nWaves = (uint64_t)((s.nPix + pixChunk - 1) / pixChunk) * ((totalSamples + samplesPerWave - 1) / samplesPerWave);
        const int used = nWaves > 1 ? nCtx : 1;
        for (int i = 1; i < used; i++) ensurePool(ctx[i], P);
        size_t bvhNodes = blasNodePool.size() + tlasHostNodes.size();
        sortActive = sortThreshold != 0 && P >= sortThreshold && (sortAnyScene || bvhNodes >= kSortMinNodes);
        if (sortActive) for (int i = 0; i < used; i++) { ensureSortBuffers(ctx[i]); if (suspendThreshold) ensureContBuffers(ctx[i]); }
        // everything queued on the main stream so far (uploads, sensor clears) precedes the waves; the waves' streams start at evStart
        if (!timelineOpen) {
            CK(cudaEventRecord(evStart, stream));
            for (WaveCtx& c : ctx) CK(cudaStreamWaitEvent(c.stream, evStart, 0));
            timelineOpen = true;
            lastAccumEvent = nullptr;
        }
        for (uint32_t p0 = 0; p0 < s.nPix; p0 += pixChunk) {
            const uint32_t np = std::min(pixChunk, s.nPix - p0);
            for (uint64_t s0 = 0; s0 < totalSamples; s0 += samplesPerWave) {
                const uint32_t ns = (uint32_t)std::min<uint64_t>(samplesPerWave, totalSamples - s0);
                WaveCtx& c = ctx[waveSerial % (uint64_t)used];
                harvest(c);   // the wave before last: its log feeds this wave's schedule while the last one keeps the GPU busy
                enqueueWave(c, s, cam, sc, k, p0, np, s.sample_count + (uint32_t)s0, ns, lastAccumEvent);
                cudaEvent_t ev = evAccum[waveSerial % (uint64_t)kWaveContexts];
                CK(cudaEventRecord(ev, c.stream));
                lastAccumEvent = ev;
                waveSerial++;
            }
        }
        stats.samples += totalPaths;
        s.sample_count += (uint32_t)totalSamples;
        if (wait) finishWaves();
    }

    // waits for every wave in flight, reads their logs, closes the device-time interval opened by the first enqueue
    void finishWaves() {
        discardAhead();
        if (!timelineOpen) return;
        use();
        for (WaveCtx& c : ctx) if (c.pending) CK(cudaStreamWaitEvent(stream, c.evDone, 0));
        CK(cudaEventRecord(evStop, stream));
        CK(cudaStreamSynchronize(stream));
        stats.host_syncs++;
        for (WaveCtx& c : ctx) harvest(c);
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, evStart, evStop));
        stats.device_ms += ms;
        timelineOpen = false;
        lastAccumEvent = nullptr;
    }
    void finishWavesOrDie() noexcept {
        try { finishWaves(); } catch (const std::exception& e) { fatal(e.what()); }
    }

    void readSensor(Sensor& s) {
        use();
        CK(cudaMemcpyAsync(s.host, s.image.p, (size_t)s.extent.width * s.extent.height * sizeof(float4), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }

    // every rank holds a full-frame buffer that is zero outside its own tiles, so one ncclReduce(sum) assembles the
    // frame exactly (x + 0 + ... + 0). The root receives into a separate buffer - its own sensor keeps accumulating its tiles -
    // and copies the assembled frame to the host buffer of GetSensorData.
    void reduceSensor(Sensor& s, uint32_t root) {
        use();
        if (!comm) throw std::runtime_error("no communicator: call HdMoonshineExtCommInit first");
        const size_t n = (size_t)s.extent.width * s.extent.height;
        const bool isRoot = commRank == root;
        if (isRoot) s.reduced.ensure(std::max<size_t>(1, n));
        NK(ncclApi().reduce(s.image.p, isRoot ? (void*)s.reduced.p : nullptr, n * 4, ncclFloat, ncclSum, (int)root, comm, stream));
        if (isRoot) CK(cudaMemcpyAsync(s.host, s.reduced.p, n * sizeof(float4), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }

    void fetchStats(MsneStats& out) {
        use();
        out = stats;
        out.pool_bytes = poolBytes();
        out.rays_closest = 0; out.rays_shadow = 0; out.max_bounce_reached = 0;
        for (WaveCtx& c : ctx) {
            unsigned long long rc[2] = {0, 0};
            uint32_t scal[4] = {0, 0, 0, 0};
            CK(cudaMemcpyAsync(rc, c.dRayCounts.p, sizeof(rc), cudaMemcpyDeviceToHost, c.stream));
            CK(cudaMemcpyAsync(scal, c.dScalars.p, sizeof(scal), cudaMemcpyDeviceToHost, c.stream));
            CK(cudaStreamSynchronize(c.stream));
            out.rays_closest += rc[0]; out.rays_shadow += rc[1];
            out.max_bounce_reached = std::max<uint64_t>(out.max_bounce_reached, scal[3]);
        }
        out.static_blas_builds = stats_static_builds; out.tlas_builds = stats_tlas_builds; out.tlas_refits = stats_tlas_refits;
    }
};

// ================================================================================================
// C ABI
// ================================================================================================
// every entry point serialises on the object's mutex (hydra.zig:78) and, except the two that enqueue waves, first waits for the
// waves an earlier HdMoonshineExtRenderLaunches left in flight
#define LOCK_ASYNC(h) std::lock_guard<std::mutex> lock_((h)->mutex)
#define LOCK(h) std::lock_guard<std::mutex> lock_((h)->mutex); (h)->finishWavesOrDie()
#define ABORT_ON_THROW(...)                          \
    try { __VA_ARGS__ } catch (const std::exception& e) { fatal(e.what()); }

static uint32_t uploadTexture(HdMoonshine* h, const void* data, uint32_t w, uint32_t hgt, uint32_t fmt) {
    h->use();
    auto t = std::make_unique<HostTexture>();
    t->w = w; t->h = hgt; t->fmt = fmt;
    const size_t bytes = (size_t)w * hgt * texelBytes(fmt);
    memset(t->first, 0, sizeof(t->first));
    if (bytes) memcpy(t->first, data, std::min<size_t>(texelBytes(fmt), sizeof(t->first)));
    t->data.alloc(bytes);
    CK(cudaMemcpyAsync(t->data.p, data, bytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // `data` is borrowed for the duration of the call
    h->textures.push_back(std::move(t));
    h->tablesDirty = true;
    return (uint32_t)h->textures.size() - 1;
}

static uint32_t createMesh(HdMoonshine* h, const F32x3* pos, const F32x3* nrm, const F32x2* uv, size_t nPos, size_t nAttr, const U32x3* idx, size_t nTri) {
    h->use();
    auto m = std::make_unique<HostMesh>();
    m->pos.assign((const float*)pos, (const float*)pos + 3 * nPos);
    m->idx.assign((const uint32_t*)idx, (const uint32_t*)idx + 3 * nTri);
    for (uint32_t v : m->idx) if (v >= nPos) throw std::runtime_error("CreateMesh: vertex index out of range");
    m->dPos.upload(m->pos.data(), m->pos.size(), h->stream);
    m->dIdx.upload(m->idx.data(), m->idx.size(), h->stream);
    if (nrm) { m->nrm.assign((const float*)nrm, (const float*)nrm + 3 * nAttr); m->dNrm.upload(m->nrm.data(), m->nrm.size(), h->stream); }
    if (uv) { m->uv.assign((const float*)uv, (const float*)uv + 2 * nAttr); m->dUv.upload(m->uv.data(), m->uv.size(), h->stream); }
    CK(cudaStreamSynchronize(h->stream));
    h->meshes.push_back(std::move(m));
    h->tablesDirty = true;
    return (uint32_t)h->meshes.size() - 1;
}

extern "C" {

const char* HdMoonshineExtLastError(void) { return g_lastError.c_str(); }

bool HdMoonshineExtSetDevice(int ordinal) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || ordinal < 0 || ordinal >= n) { g_lastError = "no such CUDA device"; return false; }
    g_device = ordinal;
    return true;
}

HdMoonshine* HdMoonshineCreate(void) {
    HdMoonshine* h = nullptr;
    try {
        h = new HdMoonshine();
        h->init();
        return h;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        fprintf(stderr, "libmoonshine: HdMoonshineCreate failed: %s\n", e.what());
        delete h;
        return nullptr;   // hydra.zig:111-135: any failure -> null
    }
}

void HdMoonshineDestroy(HdMoonshine* h) {
    if (!h) return;
    { LOCK(h); }   // also waits for waves in flight   // hydra.zig:542-558 takes no lock; here an entry point still running on another thread finishes first
    delete h;
}

bool HdMoonshineRender(HdMoonshine* h, SensorHandle sensor, LensHandle lens) {
    LOCK_ASYNC(h);
    try {
        Sensor& s = *h->sensors.at(sensor);
        if (h->aheadEligible(s)) h->renderAheadCall(s, h->lenses.at(lens));   // same sample, plus the next one started early
        else h->renderLaunches(s, h->lenses.at(lens), 1, true);   // hydra.zig:316-329: one traceRays per call, then wait idle (:358)
        if (h->comm) h->reduceSensor(s, 0);            // one rank per GPU: assemble the tiles on rank 0, whose host buffer gets the frame
        else h->readSensor(s);                         // hydra.zig:332-358: copy to the persistent host buffer, wait idle
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}

bool HdMoonshineRebuildPipeline(HdMoonshine* h) {
    LOCK(h);
    try { h->use(); h->clearAllSensors(); return true; } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}

MeshHandle HdMoonshineCreateMesh(HdMoonshine* h, const F32x3* pos, const F32x3* nrm, const F32x2* uv, size_t nPos, const U32x3* idx, size_t nTri) {
    LOCK(h);
    // hydra.zig:379-380: normals/texcoords are per-corner arrays of length 3*triangle_count
    ABORT_ON_THROW(return createMesh(h, pos, nrm, uv, nPos, 3 * nTri, idx, nTri);)
}
MeshHandle HdMoonshineExtCreateMeshIndexed(HdMoonshine* h, const F32x3* pos, const F32x3* nrm, const F32x2* uv, size_t nPos, const U32x3* idx, size_t nTri) {
    LOCK(h);
    ABORT_ON_THROW(return createMesh(h, pos, nrm, uv, nPos, nPos, idx, nTri);)
}

ImageHandle HdMoonshineCreateSolidTexture1(HdMoonshine* h, float v, const char*) {
    LOCK(h);
    ABORT_ON_THROW(return uploadTexture(h, &v, 1, 1, TEX_R32F);)
}
ImageHandle HdMoonshineCreateSolidTexture2(HdMoonshine* h, F32x2 v, const char*) {
    LOCK(h);
    ABORT_ON_THROW(return uploadTexture(h, &v, 1, 1, TEX_RG32F);)
}
ImageHandle HdMoonshineCreateSolidTexture3(HdMoonshine* h, F32x3 v, const char*) {
    LOCK(h);
    const float rgba[4] = {v.x, v.y, v.z, 1.0f};   // stored as f32x4 (MaterialManager.zig:364-373)
    ABORT_ON_THROW(return uploadTexture(h, rgba, 1, 1, TEX_RGBA32F);)
}
ImageHandle HdMoonshineCreateRawTexture(HdMoonshine* h, uint8_t* data, Extent2D e, TextureFormat f, const char*) {
    LOCK(h);
    ABORT_ON_THROW(return uploadTexture(h, data, e.width, e.height, f == f16x4 ? TEX_RGBA16F : TEX_RGBA8_SRGB);)
}
ImageHandle HdMoonshineExtCreateTexture(HdMoonshine* h, const void* data, Extent2D e, MsneTexFormat f, const char*) {
    LOCK(h);
    if (texelBytes((uint32_t)f) == 0) fatal("HdMoonshineExtCreateTexture: unknown format");
    ABORT_ON_THROW(return uploadTexture(h, data, e.width, e.height, (uint32_t)f);)
}

MaterialHandle HdMoonshineCreateMaterial(HdMoonshine* h, Material m) {
    LOCK(h);
    ABORT_ON_THROW(
        h->materials.push_back(DMaterial{m.normal, m.emissive, MAT_STANDARD_PBR, m.color, m.metalness, m.roughness, m.ior, 0});
        h->tablesDirty = true;
        return (uint32_t)h->materials.size() - 1;
    )
}
MaterialHandle HdMoonshineExtCreateMaterial(HdMoonshine* h, const MsneMaterialInfo* m) {
    LOCK(h);
    if (m->type > MAT_STANDARD_PBR) fatal("HdMoonshineExtCreateMaterial: bad material type");
    ABORT_ON_THROW(
        h->materials.push_back(DMaterial{m->normal, m->emissive, m->type, m->color, m->metalness, m->roughness, m->ior, 0});
        h->tablesDirty = true;
        return (uint32_t)h->materials.size() - 1;
    )
}
// hydra.zig:435-481: queued in the reference and flushed by the next Render; applied to the host table
// here and uploaded by the next render — same observable order. (No sensor clear, as in the reference.)
// A handle that was never returned by Create* is a caller bug; like the reference's `catch unreachable` it aborts, but with a message
// (no exception may cross the C boundary).
static void setMaterialSlot(HdMoonshine* h, MaterialHandle m, ImageHandle i, uint32_t DMaterial::*slot, const char* what) {
    LOCK(h);
    if (m >= h->materials.size()) fatal(what);
    h->materials[m].*slot = i;
    h->tablesDirty = true;
}
void HdMoonshineSetMaterialNormal(HdMoonshine* h, MaterialHandle m, ImageHandle i) { setMaterialSlot(h, m, i, &DMaterial::normal, "HdMoonshineSetMaterialNormal: no such material"); }
void HdMoonshineSetMaterialEmissive(HdMoonshine* h, MaterialHandle m, ImageHandle i) { setMaterialSlot(h, m, i, &DMaterial::emissive, "HdMoonshineSetMaterialEmissive: no such material"); }
void HdMoonshineSetMaterialColor(HdMoonshine* h, MaterialHandle m, ImageHandle i) { setMaterialSlot(h, m, i, &DMaterial::color, "HdMoonshineSetMaterialColor: no such material"); }
void HdMoonshineSetMaterialMetalness(HdMoonshine* h, MaterialHandle m, ImageHandle i) { setMaterialSlot(h, m, i, &DMaterial::metalness, "HdMoonshineSetMaterialMetalness: no such material"); }
void HdMoonshineSetMaterialRoughness(HdMoonshine* h, MaterialHandle m, ImageHandle i) { setMaterialSlot(h, m, i, &DMaterial::roughness, "HdMoonshineSetMaterialRoughness: no such material"); }
void HdMoonshineSetMaterialIOR(HdMoonshine* h, MaterialHandle m, float ior) {
    LOCK(h);
    if (m >= h->materials.size()) fatal("HdMoonshineSetMaterialIOR: no such material");
    h->materials[m].ior = ior; h->tablesDirty = true;
}

InstanceHandle HdMoonshineCreateInstance(HdMoonshine* h, Mat3x4 t, const Geometry* geoms, size_t n, bool visible) {
    LOCK(h);
    ABORT_ON_THROW(
        h->use();
        HostInstance in{};
        memcpy(in.transform, &t, 48);
        inverseAffine(in.transform, in.inverse);
        in.visible = visible;
        in.customIndex = (uint32_t)h->geometries.size();
        in.geometryCount = (uint32_t)n;
        for (size_t i = 0; i < n; i++) {
            h->geometries.push_back(DGeometry{geoms[i].mesh, geoms[i].material, geoms[i].sampled ? 1u : 0u});
            in.meshKey.push_back(geoms[i].mesh);
        }
        in.blas = 0xFFFFFFFFu;
        for (uint32_t mk : in.meshKey) in.triCount += (uint32_t)(h->meshes.at(mk)->idx.size() / 3);
        h->instances.push_back(in);
        h->tablesDirty = true;
        if (HdMoonshine::isIdentity(in.transform)) h->geomDirty = true;   // joins the flattened static BLAS
        else { h->tlasDirty = true; h->tlasRefitOk = false; }             // own/shared BLAS appended lazily, TLAS rebuilt
        for (size_t i = 0; i < n; i++) if (geoms[i].sampled) h->emitterStale = true;
        h->clearAllSensors();   // hydra.zig:491
        return (uint32_t)h->instances.size() - 1;
    )
}
void HdMoonshineSetInstanceVisibility(HdMoonshine* h, InstanceHandle i, bool visible) {
    LOCK(h);
    ABORT_ON_THROW(
        h->use();
        HostInstance& in = h->instances.at(i);
        if (in.visible != visible) {
            in.visible = visible;
            // A member of the merged static BLAS or of the TLAS keeps its place in the tree: only its record changes (traversal
            // skips hidden members). Rebuild only to INSERT something the tree does not hold, or once a quarter of it is hidden.
            if (HdMoonshine::isIdentity(in.transform)) {
                if (in.inStatic) {
                    h->dirtyInstances.push_back(i);
                    if (visible) h->hiddenStaticTris -= in.triCount; else h->hiddenStaticTris += in.triCount;
                    if (h->hiddenStaticTris * 4 > h->staticTris) h->geomDirty = true;
                } else if (visible) h->geomDirty = true;
            } else {
                if (in.tlasSlot != 0xFFFFFFFFu) {
                    h->dirtyInstances.push_back(i);
                    if (visible) h->hiddenTlasItems--; else h->hiddenTlasItems++;
                    if (h->hiddenTlasItems * 4 > h->tlasItemCount) { h->tlasDirty = true; h->tlasRefitOk = false; }
                } else if (visible) { h->tlasDirty = true; h->tlasRefitOk = false; }
            }
        }
        h->clearAllSensors();
    )
}
void HdMoonshineDestroyInstance(HdMoonshine* h, InstanceHandle i) { HdMoonshineSetInstanceVisibility(h, i, false); }   // hydra.zig:495-497
void HdMoonshineSetInstanceTransform(HdMoonshine* h, InstanceHandle i, Mat3x4 t) {
    LOCK(h);
    ABORT_ON_THROW(
        h->use();
        HostInstance& in = h->instances.at(i);
        const bool wasStatic = HdMoonshine::isIdentity(in.transform);
        memcpy(in.transform, &t, 48);
        inverseAffine(in.transform, in.inverse);
        const bool isStatic = HdMoonshine::isIdentity(in.transform);
        if (wasStatic || isStatic) {
            // enters / leaves the flattened static BLAS (an invisible identity instance that was never part of it: nothing to do)
            if (in.inStatic || (isStatic && in.visible) || (!isStatic && in.visible)) h->geomDirty = true;
        } else if (in.tlasSlot != 0xFFFFFFFFu) {
            // TLAS + instance records only (Accel.zig:567-602 recordUpdateSingleTransform + :629-679 update-mode rebuild):
            // the set of TLAS items is unchanged, so the next render refits - on the device - instead of rebuilding
            h->dirtyInstances.push_back(i);
            if (!h->tlasDirty) h->tlasRefitOk = true;
            h->tlasDirty = true;
        } else if (in.visible) { h->tlasDirty = true; h->tlasRefitOk = false; }   // visible but not in the tree yet
        for (uint32_t g = 0; g < in.geometryCount; g++) if (h->geometries[in.customIndex + g].sampled & 0xFF) h->emitterStale = true;
        h->clearAllSensors();
    )
}

SensorHandle HdMoonshineCreateSensor(HdMoonshine* h, Extent2D e) {
    LOCK(h);
    ABORT_ON_THROW(
        h->use();
        auto s = std::make_unique<Sensor>();
        s->extent = e;
        const size_t n = (size_t)e.width * e.height;
        s->image.alloc(std::max<size_t>(1, n));
        CK(cudaMemsetAsync(s->image.p, 0, std::max<size_t>(1, n) * sizeof(float4), h->stream));
        CK(cudaMallocHost((void**)&s->host, std::max<size_t>(1, n) * sizeof(float4)));
        memset(s->host, 0, n * sizeof(float4));
        h->buildPixIds(*s);
        h->sensors.push_back(std::move(s));
        return (uint32_t)h->sensors.size() - 1;
    )
}
float* HdMoonshineGetSensorData(const HdMoonshine* h, SensorHandle s) {
    // hydra.zig:522-524 takes no lock (its sensor list is a fixed-capacity array); ours is a growing vector, so the lookup is
    // serialised with CreateSensor. The returned buffer itself is stable from CreateSensor until Destroy.
    HdMoonshine* m = const_cast<HdMoonshine*>(h);
    std::lock_guard<std::mutex> lock_(m->mutex);
    if (s >= m->sensors.size()) fatal("HdMoonshineGetSensorData: no such sensor");
    return m->sensors[s]->host;
}

LensHandle HdMoonshineCreateLens(HdMoonshine* h, Lens l) { LOCK(h); ABORT_ON_THROW(h->lenses.push_back(l); return (uint32_t)h->lenses.size() - 1;) }
void HdMoonshineSetLens(HdMoonshine* h, LensHandle i, Lens l) {
    LOCK(h);
    ABORT_ON_THROW(h->use(); h->lenses.at(i) = l; h->clearAllSensors();)
}

// ---- extensions ------------------------------------------------------------------------------------
bool HdMoonshineExtSetSpecConstants(HdMoonshine* h, const MsneSpecConstants* c) {
    LOCK(h);
    if (c->samples_per_run == 0) { g_lastError = "samples_per_run must be >= 1"; return false; }
    try { h->use(); h->consts = *c; h->clearAllSensors(); return true; } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
void HdMoonshineExtGetSpecConstants(const HdMoonshine* h, MsneSpecConstants* c) {
    HdMoonshine* m = const_cast<HdMoonshine*>(h);
    std::lock_guard<std::mutex> lock_(m->mutex);
    *c = m->consts;
}

bool HdMoonshineExtSetBackground(HdMoonshine* h, const float* rgba, Extent2D e) {
    LOCK(h);
    try { h->use(); h->setBackground(rgba, e.width, e.height); h->clearAllSensors(); return true; } catch (const std::exception& ex) { g_lastError = ex.what(); return false; }
}
bool HdMoonshineExtBuildEmitterTable(HdMoonshine* h) {
    LOCK(h);
    try { h->use(); h->buildEmitterTable(); h->clearAllSensors(); return true; } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtRenderLaunches(HdMoonshine* h, SensorHandle s, LensHandle l, uint32_t n) {
    LOCK_ASYNC(h);
    // Waits for its waves like HdMoonshineRender does. Returning right after the enqueue (wait = false: consecutive calls then overlap
    // tail to head, every other entry point waits through finishWaves) was measured: 2509 vs 2527 Mrays/s on the metric config -
    // two waves that are BOTH in their throughput-bound phase only take SMs and cache from each other (profiles/r02_sweeps.md).
    try { h->renderLaunches(*h->sensors.at(s), h->lenses.at(l), n, true); return true; } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtReadSensor(HdMoonshine* h, SensorHandle s) {
    LOCK(h);
    try { h->readSensor(*h->sensors.at(s)); return true; } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
void HdMoonshineExtClearSensor(HdMoonshine* h, SensorHandle s) {
    LOCK(h);
    ABORT_ON_THROW(h->use(); h->clearSensor(*h->sensors.at(s));)
}
uint32_t HdMoonshineExtGetSensorSampleCount(const HdMoonshine* h, SensorHandle s) {
    HdMoonshine* m = const_cast<HdMoonshine*>(h);
    std::lock_guard<std::mutex> lock_(m->mutex);
    if (s >= m->sensors.size()) fatal("HdMoonshineExtGetSensorSampleCount: no such sensor");
    return m->sensors[s]->sample_count;
}
void* HdMoonshineExtGetSensorDevicePtr(const HdMoonshine* h, SensorHandle s) {
    HdMoonshine* m = const_cast<HdMoonshine*>(h);
    std::lock_guard<std::mutex> lock_(m->mutex);
    m->finishWavesOrDie();
    if (s >= m->sensors.size()) fatal("HdMoonshineExtGetSensorDevicePtr: no such sensor");
    return m->sensors[s]->image.p;
}

bool HdMoonshineExtSetTileShard(HdMoonshine* h, uint32_t rank, uint32_t world) {
    LOCK(h);
    if (world == 0 || rank >= world) { g_lastError = "bad shard"; return false; }
    try {
        h->use();
        h->shardRank = rank; h->shardWorld = world;
        for (auto& s : h->sensors) {
            h->buildPixIds(*s);
            CK(cudaMemsetAsync(s->image.p, 0, (size_t)s->extent.width * s->extent.height * sizeof(float4), h->stream));
            h->clearSensor(*s);
        }
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
// ---- multi-GPU: one process (or thread) per GPU, NCCL over NVLink ----------------------------------------------------
bool HdMoonshineExtCommGetUniqueId(void* id128) {
    try {
        ncclUniqueId id;
        NK(ncclApi().getUniqueId(&id));
        static_assert(sizeof(id) == MSNE_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
        memcpy(id128, &id, sizeof(id));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtCommInit(HdMoonshine* h, const void* id128, uint32_t rank, uint32_t world) {
    LOCK(h);
    if (world == 0 || rank >= world) { g_lastError = "bad rank / world size"; return false; }
    try {
        h->use();
        if (h->comm) { ncclApi().commDestroy(h->comm); h->comm = nullptr; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        NK(ncclApi().commInitRank(&h->comm, (int)world, id, (int)rank));
        h->commRank = rank; h->commWorld = world;
        h->shardRank = rank; h->shardWorld = world;
        for (auto& s : h->sensors) {
            h->buildPixIds(*s);
            CK(cudaMemsetAsync(s->image.p, 0, (size_t)s->extent.width * s->extent.height * sizeof(float4), h->stream));
            h->clearSensor(*s);
        }
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtCommReduceSensor(HdMoonshine* h, SensorHandle s, uint32_t root) {
    LOCK(h);
    try { h->reduceSensor(*h->sensors.at(s), root); return true; } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtCommAllReduceU64(HdMoonshine* h, uint64_t* values, size_t n) {
    LOCK(h);
    try {
        h->use();
        if (!h->comm) throw std::runtime_error("no communicator: call HdMoonshineExtCommInit first");
        DevBuf<unsigned long long> d;
        d.upload(reinterpret_cast<const unsigned long long*>(values), n, h->stream);
        NK(ncclApi().allReduce(d.p, d.p, n, ncclUint64, ncclSum, h->comm, h->stream));
        CK(cudaMemcpyAsync(values, d.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
void HdMoonshineExtCommDestroy(HdMoonshine* h) {
    LOCK(h);
    if (h->comm) { try { h->use(); ncclApi().commDestroy(h->comm); } catch (...) {} h->comm = nullptr; }
}

void HdMoonshineExtSetRaySorting(HdMoonshine* h, uint64_t minRays) {
    LOCK(h);
    if (minRays == UINT64_MAX) { h->sortThreshold = 262144; h->sortAnyScene = false; }
    else { h->sortThreshold = minRays; h->sortAnyScene = true; }
}
void HdMoonshineExtSetPathsInFlight(HdMoonshine* h, uint64_t n) {
    LOCK(h);
    h->pathsInFlight = n ? n : (1ull << 27);
    for (HdMoonshine::WaveCtx& c : h->ctx) c.pb.capacity = 0;   // pools are re-sized by the next render
}
void HdMoonshineExtSetProfiling(HdMoonshine* h, bool on) { LOCK(h); h->profiling = on; }
void HdMoonshineExtSetSchedule(HdMoonshine* h, uint64_t tailPaths, uint64_t binnedPaths, uint32_t waveContexts) {
    LOCK(h);
    h->tailThreshold = tailPaths == UINT64_MAX ? 32768 : tailPaths;
    h->tailExplicit = tailPaths != UINT64_MAX;
    h->binnedThreshold = binnedPaths == UINT64_MAX ? 262144 : binnedPaths;
    h->waveContextsInUse = waveContexts == 0 ? HdMoonshine::kDefaultWaveContexts : (int)std::min<uint32_t>(waveContexts, HdMoonshine::kWaveContexts);
    h->survivalIn.clear();
}
bool HdMoonshineExtReleasePools(HdMoonshine* h) {
    LOCK(h);
    try { h->use(); h->releasePools(); return true; } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
void HdMoonshineExtSetTraceSuspend(HdMoonshine* h, uint32_t lanes) {
    LOCK(h);
    h->suspendThreshold = std::min<uint32_t>(lanes, 16);
}
void HdMoonshineExtSetRenderAhead(HdMoonshine* h, uint32_t wavesAhead) {
    LOCK(h);
    h->renderAhead = wavesAhead != 0;
    if (wavesAhead) h->aheadDepth = std::min<uint32_t>(wavesAhead, HdMoonshine::kWaveContexts - 1);
}
void HdMoonshineExtGetStats(const HdMoonshine* h, MsneStats* out) {
    HdMoonshine* m = const_cast<HdMoonshine*>(h);
    std::lock_guard<std::mutex> lock_(m->mutex);
    m->finishWavesOrDie();
    ABORT_ON_THROW(m->fetchStats(*out);)
}
void HdMoonshineExtResetStats(HdMoonshine* h) {
    LOCK(h);
    ABORT_ON_THROW(
        h->use();
        h->stats = MsneStats{};
        for (HdMoonshine::WaveCtx& c : h->ctx) {
            CK(cudaMemsetAsync(c.dRayCounts.p, 0, 4 * sizeof(unsigned long long), c.stream));
            CK(cudaMemsetAsync(c.dScalars.p + 3, 0, sizeof(uint32_t), c.stream));
            CK(cudaStreamSynchronize(c.stream));
        }
    )
}

// Accel.recordUpdateSingleMaterial (Accel.zig:609-627; online/main.zig:231): geometry `g` of instance `i` gets another material
void HdMoonshineExtSetGeometryMaterial(HdMoonshine* h, InstanceHandle i, uint32_t g, MaterialHandle m) {
    LOCK(h);
    if (i >= h->instances.size()) fatal("HdMoonshineExtSetGeometryMaterial: no such instance");
    const HostInstance& in = h->instances[i];
    if (g >= in.geometryCount || m >= h->materials.size()) fatal("HdMoonshineExtSetGeometryMaterial: index out of range");
    h->geometries[in.customIndex + g].material = m;
    h->tablesDirty = true;
    if (h->geometries[in.customIndex + g].sampled & 0xFF) h->emitterStale = true;   // table entries are per triangle, radiance is read at sample time; kept for symmetry
}
// The reference builds the emitter alias table only in Accel.create (Accel.zig:491-538); instances added or moved
// afterwards are never re-sampled (Accel.zig:187). With following on, the table is rebuilt by the next render after any
// edit that changes it (instance added / moved with a sampled geometry).
void HdMoonshineExtSetEmitterTableFollowsEdits(HdMoonshine* h, bool on) {
    LOCK(h);
    h->emitterFollowsEdits = on;
    if (on) h->emitterStale = true;
}

// ---- object picker (ObjectPicker.zig:89-128) ---------------------------------------------------------
bool HdMoonshineExtPick(HdMoonshine* h, SensorHandle s, LensHandle l, F32x2 coords, MsneClickData* out) {
    LOCK(h);
    try {
        h->use(); h->syncScene();
        const Sensor& se = *h->sensors.at(s);
        DevBuf<MsneClick> d;
        d.alloc(1);
        launchPick(h->stream, h->deviceScene(), HdMoonshine::makeCamera(h->lenses.at(l), se.extent.width, se.extent.height), coords.x, coords.y, d.p);
        h->stats.kernel_launches++;
        MsneClick c{};
        CK(cudaMemcpyAsync(&c, d.p, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        out->instance_index = c.instance_index; out->geometry_index = c.geometry_index; out->primitive_index = c.primitive_index;
        out->barycentrics.x = c.bary[0]; out->barycentrics.y = c.bary[1];
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}

// ---- debug probes ----------------------------------------------------------------------------------
bool HdMoonshineExtTraceRays(HdMoonshine* h, const float* rays8, size_t n, bool anyHit, uint32_t* hits5) {
    LOCK(h);
    try {
        h->use(); h->syncScene();
        DevBuf<float> dr; DevBuf<uint32_t> dh;
        dr.upload(rays8, 8 * n, h->stream); dh.alloc(5 * n);
        launchTraceProbe(h->stream, h->deviceScene(), dr.p, (uint32_t)n, anyHit ? 1 : 0, dh.p);
        h->stats.kernel_launches++;
        CK(cudaMemcpyAsync(hits5, dh.p, 5 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtPrimaryRays(HdMoonshine* h, SensorHandle s, LensHandle l, uint32_t sample, float* rays6, uint32_t* rngState) {
    LOCK(h);
    try {
        h->use();
        const Sensor& se = *h->sensors.at(s);
        const size_t n = (size_t)se.extent.width * se.extent.height;
        DevBuf<float> dr; DevBuf<uint32_t> ds;
        dr.alloc(6 * n); ds.alloc(n);
        launchPrimaryProbe(h->stream, HdMoonshine::makeCamera(h->lenses.at(l), se.extent.width, se.extent.height), h->deviceConstants(), sample, dr.p, ds.p);
        h->stats.kernel_launches++;
        CK(cudaMemcpyAsync(rays6, dr.p, 6 * n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(rngState, ds.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtMathProbe(HdMoonshine* h, int fn, const float* a, const float* b, size_t n, float* out) {
    LOCK(h);
    try {
        h->use();
        DevBuf<float> da, db, dout;
        da.upload(a, n, h->stream); db.upload(b, n, h->stream); dout.alloc(n);
        launchMathProbe(h->stream, fn, da.p, db.p, (uint32_t)n, dout.p);
        h->stats.kernel_launches++;
        CK(cudaMemcpyAsync(out, dout.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtBsdfProbe(HdMoonshine* h, const float* queries14, size_t n, float* out4) {
    LOCK(h);
    try {
        h->use();
        DevBuf<float> dq, dout;
        dq.upload(queries14, 14 * n, h->stream); dout.alloc(4 * n);
        launchBsdfProbe(h->stream, dq.p, (uint32_t)n, dout.p);
        h->stats.kernel_launches++;
        CK(cudaMemcpyAsync(out4, dout.p, 4 * n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
bool HdMoonshineExtEnvProbe(HdMoonshine* h, int op, const float* in3, size_t n, float* out7) {
    LOCK(h);
    try {
        h->use(); h->syncScene();
        DevBuf<float> din, dout;
        din.upload(in3, 3 * n, h->stream); dout.alloc(7 * n);
        launchEnvProbe(h->stream, h->deviceScene(), op, din.p, (uint32_t)n, dout.p);
        h->stats.kernel_launches++;
        CK(cudaMemcpyAsync(out7, dout.p, 7 * n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
uint32_t HdMoonshineExtGetEnvSize(const HdMoonshine* h) { return h->envSize; }
bool HdMoonshineExtGetEnvData(HdMoonshine* h, float* rgba, float* lum) {
    LOCK(h);
    try {
        h->use();
        size_t total = 0;
        for (uint32_t s = h->envSize; s >= 1; s >>= 1) { total += (size_t)s * s; if (s == 1) break; }
        CK(cudaMemcpyAsync(rgba, h->dEnvRgb.p, (size_t)h->envSize * h->envSize * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(lum, h->dEnvLum.p, total * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}
uint32_t HdMoonshineExtGetEmitterTableSize(const HdMoonshine* h) { return (uint32_t)h->alias.size(); }
bool HdMoonshineExtGetEmitterTable(HdMoonshine* h, void* entries20) {
    LOCK(h);
    try {
        h->use(); h->syncScene();
        CK(cudaMemcpyAsync(entries20, h->dAlias.p, h->alias.size() * sizeof(DAlias), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return true;
    } catch (const std::exception& e) { g_lastError = e.what(); return false; }
}

}  // extern "C"
