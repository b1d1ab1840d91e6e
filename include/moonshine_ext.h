/* moonshine_ext.h — additions to the reference's C ABI (SURVEY.md §8b "Needed ABI additions").
 *
 * The Hydra ABI (moonshine.h) only covers what hydra/hydra.zig exports. The offline entry point and
 * the reference's tests reach the rest of the integrator through Zig; those Zig interfaces are
 * mirrored here as plain C so a cgo/ctypes/Zig `extern fn` caller can drive the same path:
 *
 *   specialization constants  engine/hrtsystem/pipeline.zig:319-327  (shaders/hrtsystem/main.hlsl:34-40)
 *   BackgroundManager.addBackground   engine/hrtsystem/BackgroundManager.zig:142-394
 *   MaterialManager.upload (all four variants)  engine/hrtsystem/MaterialManager.zig:140-170
 *   TextureManager.upload raw formats  engine/hrtsystem/MaterialManager.zig:351-413; World.zig:72,101,184
 *   MeshManager.upload with indexed attributes  engine/hrtsystem/MeshManager.zig:70-156
 *   Accel.create's emitter alias table  engine/hrtsystem/Accel.zig:491-538
 *   offline/main.zig:131-194 (spp launches, then one read-back)
 *
 * plus what a multi-GPU deployment needs (device selection, pixel-tile sharding, ray counters,
 * device-resident sensor access for an NCCL reduce) and debug probes used by the parity tests.
 */
#pragma once
#include "moonshine.h"

typedef struct MsneSpecConstants {      /* pipeline.zig:319-327 — identical field order */
    uint32_t samples_per_run;
    uint32_t max_bounces;
    uint32_t env_samples_per_bounce;    /* any count (integrator.hlsl:141-144); 0 and 1 take the queued-shadow-ray fast path, */
    uint32_t mesh_samples_per_bounce;   /* larger counts trace their shadow rays inside the shade pass (integrator.hlsl:147-150) */
    uint32_t flip_image;                /* Bool32 */
    uint32_t indexed_attributes;        /* Bool32 */
    uint32_t two_component_normal_texture; /* Bool32 */
} MsneSpecConstants;

typedef enum MsneTexFormat {
    MSNE_TEX_R32F = 0, MSNE_TEX_RG32F = 1, MSNE_TEX_RGBA32F = 2,   /* MaterialManager.zig:364-388 */
    MSNE_TEX_R8_UNORM = 3, MSNE_TEX_RG8_UNORM = 4,                 /* World.zig:72,184 */
    MSNE_TEX_RGBA8_SRGB = 5, MSNE_TEX_RGBA16F = 6,                 /* World.zig:101; hydra.zig:47-52 */
} MsneTexFormat;

typedef enum MsneMaterialType {         /* shaders/hrtsystem/world.hlsl:31-36 */
    MSNE_MAT_GLASS = 0, MSNE_MAT_LAMBERT = 1, MSNE_MAT_PERFECT_MIRROR = 2, MSNE_MAT_STANDARD_PBR = 3,
} MsneMaterialType;

typedef struct MsneMaterialInfo {       /* MaterialManager.zig:22-77 flattened */
    uint32_t type;                      /* MsneMaterialType */
    ImageHandle normal, emissive;
    ImageHandle color, metalness, roughness;  /* StandardPBR; Lambert uses color */
    float ior;                          /* StandardPBR / Glass */
} MsneMaterialInfo;

typedef struct MsneStats {
    uint64_t rays_closest;              /* closest-hit queries traced (intersection.hlsl:18-22) */
    uint64_t rays_shadow;               /* any-hit queries traced   (intersection.hlsl:34-47) */
    uint64_t samples;                   /* camera samples started */
    uint64_t kernel_launches;           /* this library's kernels launched */
    double   device_ms;                 /* CUDA-event time of the render loops (first generate .. last accumulate) */
    double   trace_ms;                  /* CUDA-event time spent in the trace kernel (valid when profiling is on) */
    uint64_t trace_launches;
    uint64_t max_bounce_reached;
    uint64_t static_blas_builds;        /* rebuilds of the BLAS level (merged static BLAS + shared BLASes) */
    uint64_t tlas_builds;               /* rebuilds of the instance level (TLAS + instance records) */
    uint64_t tlas_refits;               /* transform-only edits served by a refit of the existing TLAS on the device */
    double   sort_ms;                   /* CUDA-event time of the ray-reordering passes (counting sort + binning; profiling on) */
    uint64_t waves;                     /* waves rendered (generate .. accumulate); two are in flight at a time */
    uint64_t host_syncs;                /* host<->device synchronisations inside render loops (one per render call, none per bounce) */
    uint64_t tail_paths;                /* paths left to k_tail after the wavefront bounces, summed over waves */
    double   accel_sync_ms;             /* host wall time spent bringing the acceleration structure up to date after edits (builds, refits, record uploads) */
    uint64_t ahead_discarded;           /* render-ahead: speculative waves dropped unseen because the next call did not ask for that sample */
    uint64_t pool_bytes;                /* device memory held by the wavefront pools right now (path state, queues, sort buffers); see HdMoonshineExtReleasePools */
} MsneStats;

/* must be called before HdMoonshineCreate; selects the CUDA device of objects created afterwards */
MSNE_API bool HdMoonshineExtSetDevice(int ordinal);

MSNE_API bool HdMoonshineExtSetSpecConstants(HdMoonshine*, const MsneSpecConstants*);    /* = Pipeline.recreate, pipeline.zig:180-243 */
MSNE_API void HdMoonshineExtGetSpecConstants(const HdMoonshine*, MsneSpecConstants*);
MSNE_API bool HdMoonshineExtSetBackground(HdMoonshine*, const float* rgba_equirect, Extent2D); /* addBackground (replaces the current one) */
MSNE_API ImageHandle HdMoonshineExtCreateTexture(HdMoonshine*, const void* data, Extent2D, MsneTexFormat, const char* name);
MSNE_API MaterialHandle HdMoonshineExtCreateMaterial(HdMoonshine*, const MsneMaterialInfo*);
/* attributes of length position_count, looked up through the index buffer (indexed_attributes = true) */
MSNE_API MeshHandle HdMoonshineExtCreateMeshIndexed(HdMoonshine*, const F32x3* positions, const F32x3* normals_or_null,
                                                    const F32x2* texcoords_or_null, size_t position_count,
                                                    const U32x3* indices, size_t triangle_count);
/* Accel.create semantics: rebuild the emitter alias table over every `sampled` geometry of every instance */
MSNE_API bool HdMoonshineExtBuildEmitterTable(HdMoonshine*);

/* offline loop: n_launches x (push constants, trace rays), sensor.sample_count advances; no read-back */
MSNE_API bool HdMoonshineExtRenderLaunches(HdMoonshine*, SensorHandle, LensHandle, uint32_t n_launches);
MSNE_API bool HdMoonshineExtReadSensor(HdMoonshine*, SensorHandle);      /* device image -> host buffer of GetSensorData */
MSNE_API void HdMoonshineExtClearSensor(HdMoonshine*, SensorHandle);     /* Sensor.clear (core/Sensor.zig:81-83) */
MSNE_API uint32_t HdMoonshineExtGetSensorSampleCount(const HdMoonshine*, SensorHandle);
MSNE_API void* HdMoonshineExtGetSensorDevicePtr(const HdMoonshine*, SensorHandle);  /* RGBA32F W*H, for NCCL */

/* multi-GPU: this object renders only the pixels of the 16x16 tiles (cx, cy) with (cx + s cy) % world == rank, s = the smallest
 * number >= 2 coprime to world (3 for 2, 4, 8 ranks); the rest of its frame stays 0 */
MSNE_API bool HdMoonshineExtSetTileShard(HdMoonshine*, uint32_t rank, uint32_t world);
/* One rank per GPU over NCCL (the reference is single-GPU, so this has no reference twin). Rank 0 calls
 * CommGetUniqueId and hands the 128 bytes to the other ranks by any channel (file, socket, MPI, torch.distributed); every rank
 * then calls CommInit, which also shards the pixel tiles by rank. From then on HdMoonshineRender renders the rank's tiles and ends
 * with one ncclReduce(sum) of the RGBA32F frame to rank 0, whose GetSensorData buffer receives the assembled frame (exact: every
 * texel has one non-zero contributor). CommReduceSensor does the same after HdMoonshineExtRenderLaunches. Collective calls: every
 * rank must make them in the same order. */
#define MSNE_COMM_ID_BYTES 128
MSNE_API bool HdMoonshineExtCommGetUniqueId(void* id128);
MSNE_API bool HdMoonshineExtCommInit(HdMoonshine*, const void* id128, uint32_t rank, uint32_t world);
MSNE_API bool HdMoonshineExtCommReduceSensor(HdMoonshine*, SensorHandle, uint32_t root);
MSNE_API bool HdMoonshineExtCommAllReduceU64(HdMoonshine*, uint64_t* values, size_t n);   /* sum over ranks, in place (ray counters) */
MSNE_API void HdMoonshineExtCommDestroy(HdMoonshine*);
MSNE_API void HdMoonshineExtSetPathsInFlight(HdMoonshine*, uint64_t max_paths);   /* wavefront pool size (0 = default) */
/* The pools grow to what the largest render so far needed (326 B per path in flight: 0.7 GB for a one-sample 1080p call, 41 GiB for
 * the default 2 x 64 Mi paths of a long offline render) and stay allocated between renders. This returns them to the driver; the
 * next render allocates what it needs again. MsneStats.pool_bytes reports the current holding. */
MSNE_API bool HdMoonshineExtReleasePools(HdMoonshine*);
/* extension-ray queues of at least `min_rays` are traced in sorted order (direction octant + origin cell); 0 = never,
 * UINT64_MAX = library default (262144, and only for scenes whose BVHs hold >= 2048 nodes). Ordering only: images are
 * bitwise identical either way. */
MSNE_API void HdMoonshineExtSetRaySorting(HdMoonshine*, uint64_t min_rays);
MSNE_API void HdMoonshineExtSetProfiling(HdMoonshine*, bool per_kernel_events);
/* Scheduling hints (performance only - the image is the same bits for any values): k_tail finishes a wave once fewer than
 * `tail_paths` paths are expected to survive; bounces expected to shade fewer than `binned_paths` paths use one generic shade
 * launch instead of one per material; `wave_contexts` = waves in flight (1 or 2). UINT64_MAX / 0 = library defaults
 * (32768, 262144, 2). Also forgets the survival profile learnt from earlier waves. */
MSNE_API void HdMoonshineExtSetSchedule(HdMoonshine*, uint64_t tail_paths, uint64_t binned_paths, uint32_t wave_contexts);

/* Suspend-and-regroup tracing (performance hint, same bits): in sorted bounces of scenes without transformed instances a warp of the
 * trace kernel hands its last `lanes` walking rays (0 = off, at most 16) to a later pass that walks such long rays together, instead
 * of idling its other lanes until they end. MSNE_TRACE_SUSPEND=<lanes> sets it at creation. */
MSNE_API void HdMoonshineExtSetTraceSuspend(HdMoonshine*, uint32_t lanes);
/* Render-ahead for one-sample-per-call rendering (the Hydra delegate's pattern, hydra/hydra.zig:316-358). waves_ahead > 0:
 * HdMoonshineRender keeps that many FURTHER samples' waves (same sensor, lens, scene; at most 3) in flight when it waits for its own,
 * holds their samples back from the sensor, and the following HdMoonshineRender calls claim them one by one if they ask for exactly
 * those samples - otherwise (any other entry point in between, another sensor / lens, an edited scene) the waves are dropped unseen.
 * Frames and sample counts are bit for bit those of render-ahead off; ray counters cover claimed waves only
 * (MsneStats.ahead_discarded counts the dropped ones). 0 = off (default); MSNE_RENDER_AHEAD=<n> sets it at creation. */
MSNE_API void HdMoonshineExtSetRenderAhead(HdMoonshine*, uint32_t waves_ahead);
MSNE_API void HdMoonshineExtGetStats(const HdMoonshine*, MsneStats*);
MSNE_API void HdMoonshineExtResetStats(HdMoonshine*);

/* ---- incremental edits beyond the hydra ABI (online/main.zig:225-270) ---- */
/* Accel.recordUpdateSingleMaterial (Accel.zig:609-627): geometry `geometry` of `instance` switches to `material` */
MSNE_API void HdMoonshineExtSetGeometryMaterial(HdMoonshine*, InstanceHandle instance, uint32_t geometry, MaterialHandle material);
/* rebuild the emitter alias table automatically (at the next render) after edits that change it; the reference never does (Accel.zig:187) */
MSNE_API void HdMoonshineExtSetEmitterTableFollowsEdits(HdMoonshine*, bool on);

/* ---- object picker: ObjectPicker.getClickedObject (engine/hrtsystem/ObjectPicker.zig:89-128) + its ray-generation
 * program (shaders/hrtsystem/input.hlsl:1-57). `normalized_coords` in [0,1]^2 with the origin at the top left as the
 * online viewer passes them (online/main.zig:287-305); instance_index is -1 when nothing was hit. ---- */
typedef struct MsneClickData {
    int32_t instance_index;
    uint32_t geometry_index;
    uint32_t primitive_index;
    F32x2 barycentrics;
} MsneClickData;
MSNE_API bool HdMoonshineExtPick(HdMoonshine*, SensorHandle, LensHandle, F32x2 normalized_coords, MsneClickData* out);

/* ---- debug probes used by tests/ (stage-level parity against the oracle) ---- */
/* rays: n x {origin.xyz, tmin, dir.xyz, tmax}; hits: n x {u32 instance, u32 geometry, u32 primitive, f32 u, f32 v} */
MSNE_API bool HdMoonshineExtTraceRays(HdMoonshine*, const float* rays8, size_t n, bool any_hit, uint32_t* hits5);
/* primary rays of launch sample `sample`: rays6 = W*H x {origin, dir}; rng_state = W*H (after the 4 camera draws) */
MSNE_API bool HdMoonshineExtPrimaryRays(HdMoonshine*, SensorHandle, LensHandle, uint32_t sample, float* rays6, uint32_t* rng_state);
/* the renderer's DEFINED binary32 transcendentals: fn 0 sin(a), 1 cos(a), 2 log(a), 3 atan2(a, b), 4 acos(a); n floats each */
MSNE_API bool HdMoonshineExtMathProbe(HdMoonshine*, int fn, const float* a, const float* b, size_t n, float* out);
/* material.hlsl MaterialVariant::{eval,pdf,sample} on caller-supplied inputs: queries = n x 14 floats
 * {type, color r,g,b, metalness, roughness, ior, op, a.x,a.y,a.z, w_o.x,w_o.y,w_o.z}; op 0 = eval(w_i = a, w_o) -> rgb, 1 = pdf(a, w_o) -> x,
 * 2 = sample(w_o, square = a.xy) -> {dir.xyz, pdf}; directions in the shading frame. out = n x 4 floats. */
MSNE_API bool HdMoonshineExtBsdfProbe(HdMoonshine*, const float* queries14, size_t n, float* out4);
/* light.hlsl EnvMap against the current background: op 0 = sample(rand = in.xy) -> {dir.xyz, radiance.rgb, pdf} (no shadow ray),
 * 1 = eval(dir = in.xyz) -> {radiance.rgb, pdf}, 2 = incomingRadiance(dir) -> {rgb}. in = n x 3 floats, out = n x 7 floats. */
MSNE_API bool HdMoonshineExtEnvProbe(HdMoonshine*, int op, const float* in3, size_t n, float* out7);
MSNE_API uint32_t HdMoonshineExtGetEnvSize(const HdMoonshine*);
MSNE_API bool HdMoonshineExtGetEnvData(HdMoonshine*, float* rgba, float* lum_all_levels);
MSNE_API uint32_t HdMoonshineExtGetEmitterTableSize(const HdMoonshine*);            /* entries incl. header */
MSNE_API bool HdMoonshineExtGetEmitterTable(HdMoonshine*, void* entries20);
MSNE_API const char* HdMoonshineExtLastError(void);
