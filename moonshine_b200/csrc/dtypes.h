// Plain-old-data layouts shared by the host runtime (runtime.cpp) and the sm_100a kernels (kernels.cu).
//
// The reference binds 7 storage buffers + a bindless texture array + 2 env images + the TLAS
// (shaders/hrtsystem/main.hlsl:10-24). Here the same tables live in HBM as flat arrays reachable
// through one `DeviceScene` parameter block; "device addresses" become raw pointers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define MSNE_HD __host__ __device__
#else
#define MSNE_HD
#endif

namespace msne {

// ---- acceleration structure -----------------------------------------------------------------
// BVH2 node, 64 B (four 16-byte lanes, one 128-bit load each):
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   c  = (ref0, ref1, -, -)
// The host builder (bvh_build.cpp) emits c = (child0, child1, count0, count1) with local indices; when a BVH is
// pooled for the device (runtime.cpp) the children become pre-encoded traversal references:
//   0 <= ref < REF_ITEM      inner node index in the pooled node array
//   ref < 0                  BLAS leaf: x = ~ref, first triangle = x >> 3, count = (x & 7) + 1
//   REF_ITEM | slot          TLAS leaf: one instance, tlasItems[slot]
//   REF_EMPTY                no child: its box is the point (3e38, 3e38, 3e38), which no slab test can reach, so the
//                            kernels never look at the reference;  REF_LEAVE / REF_END are stack sentinels
// Every BVH has at least one inner node (the root).
struct BvhNode2 {
    float n0[4], n1[4], n2[4];
    int32_t c[4];
};
static_assert(sizeof(BvhNode2) == 64, "BvhNode2 must be 64 B");
inline void setEmptyChild(BvhNode2& n, int s) {
    float* xy = s == 0 ? n.n0 : n.n1;
    xy[0] = xy[1] = xy[2] = xy[3] = 3e38f; n.n2[2 * s] = n.n2[2 * s + 1] = 3e38f;
}

static const int32_t REF_ITEM = 0x40000000, REF_EMPTY = 0x7FFFFFFD, REF_LEAVE = 0x7FFFFFFE, REF_END = 0x7FFFFFFF;

// Triangle record, 48 B: three vertices in object space + ids (watertight edge test needs all three)
//   t0 = (v0.xyz, geometryIndex)  t1 = (v1.xyz, primitiveIndex)  t2 = (v2.xyz, instanceIndex)
// instanceIndex is only meaningful inside the merged static BLAS (all visible identity-transform instances
// flattened into one world-space BVH: an identity transform leaves the ray bit-identical, so this is exact).
struct TriRec { float v0[3]; uint32_t geom; float v1[3]; uint32_t prim; float v2[3]; uint32_t inst; };
static_assert(sizeof(TriRec) == 48, "TriRec must be 48 B");

// world.hlsl:6-15 + worldToInstance[] (main.hlsl:13) + our BLAS references, 128 B
struct DInstance {
    float o2w[12];         // instance transform, row-major 3x4
    float w2o[12];         // inverse_affine(transform) (engine/vector.zig:350-362)
    uint32_t customIndex;  // offset into geometries[] (low 24 bits of instanceCustomIndexAndMask)
    uint32_t blasRoot;     // index of the BLAS root in the pooled node array
    uint32_t visible;      // mask != 0
    uint32_t isStatic;     // 1: pseudo-instance holding the merged static BLAS (no transform; hit.inst comes from TriRec.inst)
    uint32_t pad[4];       // [0] BLAS id, [1] TLAS leaf slot or 0xFFFFFFFF (device-side refit, kernels.cu k_tlas_boxes)
};
static_assert(sizeof(DInstance) == 128, "DInstance must be 128 B");

struct DGeometry { uint32_t mesh, material, sampled; };                      // world.hlsl:17-21
struct DMesh { const float* pos; const float* uv; const float* nrm; const uint32_t* idx; };  // world.hlsl:23-29
struct DMaterial {                                                            // world.hlsl:38-47 + variant payload
    uint32_t normal, emissive, type;
    uint32_t color, metalness, roughness;                                     // material.hlsl:186-190 / :146-147
    float ior;                                                                // material.hlsl:189 / :348-351
    // Resolved on the host (runtime.cpp resolveMaterials): bit s of constMask = texture slot s (MAT_SLOT_*) is a 1x1 image, whose
    // SampleLevel is its only texel whatever the uv - that texel, decoded exactly as the sampler would, sits in constant[s].
    // Most materials are built from solid textures (hydra's defaults, glTF factors), so a hit costs one 112-byte record instead
    // of five descriptor -> texel -> table chains.
    uint32_t constMask;
    float constant[5][4];
};
static_assert(sizeof(DMaterial) == 112, "DMaterial is 112 B");
enum MatSlot : uint32_t { MAT_SLOT_NORMAL = 0, MAT_SLOT_EMISSIVE, MAT_SLOT_COLOR, MAT_SLOT_METALNESS, MAT_SLOT_ROUGHNESS };
struct DTexture { const void* data; uint32_t w, h, fmt, pad; };               // bindless dTextures[] (material.hlsl:3)
struct DAlias { uint32_t alias; float select; uint32_t inst, geom, prim; };   // light.hlsl:17-22,105-109
static_assert(sizeof(DAlias) == 20, "alias entry is 20 B (Accel.zig:51-55)");

struct DeviceScene {
    const BvhNode2* tlasNodes;     // may be null when there are no visible instances
    const uint32_t* tlasItems;     // TLAS leaf slot -> instance index
    const BvhNode2* blasNodes;     // pooled BLAS nodes
    const TriRec* tris;            // pooled triangles (leaf order)
    const float4* triFrames;       // per pooled triangle 64 B: its tangent frame, evaluated once per build (shading.cuh MSNE_TRI_FRAMES, kernels.cu k_tri_frames)
    const DInstance* instances;
    const DGeometry* geometries;
    const DMesh* meshes;
    const DMaterial* materials;
    const DTexture* textures;
    const DAlias* alias;           // [0] = header {alias = count, select = sum of areas}
    uint32_t aliasEntries;         // 0 when no table was built
    const float4* emitterTris;     // per alias entry i (alias[1 + i]) 64 B: (p0, uv0.x) (p1, uv0.y) (p2, uv1.x) (uv1.y, uv2.x, uv2.y, -): the object-space
                                   // vertices + texcoords MeshLights::sample would fetch through instance -> geometry -> mesh -> index
    // environment (BackgroundManager.zig:154-163)
    const float* envRgb;           // S*S float4
    const float* envLum;           // mip 0 .. log2(S), tightly packed, SUM pyramid
    uint32_t envSize;
    uint32_t envLevels;
    uint32_t envLevelOffset[12];   // float offset of each mip inside envLum
    const float4* envQuads;        // the same pyramid, quad-major: for mip L the 2x2 texels under texel (x, y) of mip L+1 as ONE float4
    uint32_t envQuadOffset[12];    // (l00, l10, l01, l11) at envQuads[envQuadOffset[L] + y * (size >> (L+1)) + x]: one 16-byte load per level of EnvMap::sample
    const float* texLuts;          // 512 floats: [0,256) sRGB -> linear, [256,512) c / 255 (shading.cuh)
    uint32_t staticOnly;           // 1: the TLAS holds nothing but the merged static BLAS -> single-level traversal from staticRoot
    uint32_t staticRoot;
    uint32_t staticHidden;         // 1: some member of the merged static BLAS is hidden (checked per candidate hit, traverse.cuh)
};

// specialization constants (main.hlsl:34-40) as runtime values
struct DConstants {
    uint32_t samples_per_run, max_bounces, env_samples, mesh_samples;
    uint32_t flip_image, indexed_attributes, two_component_normal;
};

// camera basis hoisted out of Camera::generateRay (camera.hlsl:14-43); computed once per launch on
// the host with the arithmetic of the HLSL (runtime.cpp: makeCamera)
struct DCamera {
    float origin[3], u[3], v[3], horizontal[3], vertical[3], llc[3];
    float aperture;
    uint32_t width, height;
};

// ClickData of the object picker (input.hlsl:1-6): instance_index = -1 when nothing was hit
struct MsneClick { int32_t instance_index; uint32_t geometry_index, primitive_index; float bary[2]; };

// ---- wavefront pool ---------------------------------------------------------------------------
// One path = six 16-byte lanes per ping-pong buffer (SoA arrays of float4, index q = queue slot):
//   R0 = (origin.xyz, lastMaterialPdf)      R1 = (dir.xyz, bits: bounce | DELTA<<30 | DEAD<<31)   (interleaved, RAY_STRIDE)
//   S0 = (throughput.xyz, rng state bits)   S1 = (accumulated.xyz, result-slot bits)
//   N0 = pending env-NEE contribution       N1 = pending mesh-NEE contribution (zeroed by the trace
//        kernel when the shadow ray is occluded; folded into `accumulated` by the next shade pass)
//   H  = hit record written by the trace kernel: (instanceIndex, triangle index | bin << 29, u bits, v bits)
// Shadow queue: Q0 = (origin.xyz, tmax)  Q1 = (dir.xyz, bits: target q*2 + light)
// The two lanes of a ray (R0/R1, and Q0/Q1 of a shadow ray) are interleaved in ONE array - R1 = R0 + 1, element q at index
// RAY_STRIDE * q - so that a ray is one aligned 32-byte sector: sorted bounces gather rays by queue index, and a gathered
// sector costs the same DRAM access whether 16 or 32 of its bytes are used. Streaming passes touch the same sectors as with
// separate arrays.
static const uint32_t RAY_STRIDE = 2;
struct Counters { uint32_t nLive, nDead, nShadow, pad; };

struct PathBuffers {
    float4* R0[2]; float4* R1[2]; float4* S0[2]; float4* S1[2]; float4* N0[2]; float4* N1[2];
    uint4* H;
    float4* Q0; float4* Q1;
    float4* result;        // per sample-slot final radiance
    uint32_t* typeList;    // [NUM_BINS][P] queue slots binned by material type (4 = miss) by the trace kernel
    uint32_t capacity;     // P
};

struct WaveState {
    Counters* counters;        // [2] ping-pong
    uint32_t* traceTicket;     // work-stealing ticket of the trace kernel (self-resetting)
    uint32_t* traceDone;       // finished-block counter of the trace kernel
    uint32_t* shadeDone;       // finished-block counter of the shade kernel
    unsigned long long* rayCounts;  // [0] closest, [1] shadow
    uint32_t* maxBounce;
    uint32_t* typeCount;       // [NUM_BINS] entries in typeList bins (reset by the shade kernel that consumes the bin)
    uint32_t* binDone;         // [NUM_BINS] finished-block counters of the specialised shade kernels
    Counters* history;         // [bounce] queue sizes each bounce produced, logged by its last shade launch (read by the host after the wave)
};

static const uint32_t NUM_BINS = 5;   // MaterialType 0..3 + 4 = miss (and dead paths)
static const uint32_t BIN_MISS = 4;
static const uint32_t FLAG_DEAD = 0x80000000u;
static const uint32_t FLAG_DELTA = 0x40000000u;
static const uint32_t BOUNCE_MASK = 0x3FFFFFFFu;
static const uint32_t MISS_ID = 0xFFFFFFFFu;
// H.y = pooled triangle index (< 2^29) | material bin << 29 (written by k_trace, read back by k_bin; shading masks it off)
static const uint32_t H_BIN_SHIFT = 29, H_TRI_MASK = (1u << 29) - 1u;
static const uint32_t STATIC_INST = 0xFFFFFFFEu;

enum TexFormat : uint32_t { TEX_R32F = 0, TEX_RG32F, TEX_RGBA32F, TEX_R8_UNORM, TEX_RG8_UNORM, TEX_RGBA8_SRGB, TEX_RGBA16F };
enum MatType : uint32_t { MAT_GLASS = 0, MAT_LAMBERT, MAT_PERFECT_MIRROR, MAT_STANDARD_PBR };

}  // namespace msne
