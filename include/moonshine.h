/* moonshine.h — the reference's C ABI, re-exported by the B200-native libmoonshine.so.
 *
 * This header restates /root/reference/hydra/moonshine.h:10-95 (types :10-69, entry points :71-95)
 * so that the reference's USD Hydra delegate (hydra/ *.cpp) links against the replacement library
 * unchanged. Semantics of every entry point follow the Zig implementation it replaces,
 * hydra/hydra.zig:107-558 (cited per function below). Differences are limited to what is behind the
 * boundary: CUDA (sm_100a) instead of Vulkan ray tracing.
 *
 * Ownership / errors / threading (hydra/hydra.zig): input pointers are borrowed for the duration of
 * the call; the library owns all device memory and the host sensor buffers; Create returns NULL on
 * failure, Render/RebuildPipeline return false, everything else aborts on failure; every entry point
 * except GetSensorData takes one per-object mutex.
 */
#pragma once

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define MSNE_API extern "C"
#else
#define MSNE_API extern
#endif

typedef uint32_t MeshHandle;      /* moonshine.h:10 */
typedef uint32_t ImageHandle;     /* :11 */
typedef uint32_t MaterialHandle;  /* :12 */
typedef uint32_t SensorHandle;    /* :13 */
typedef uint32_t LensHandle;      /* :14 */
typedef uint32_t InstanceHandle;  /* :15 */

typedef struct F32x2 { float x, y; } F32x2;             /* :17-19 */
typedef struct F32x3 { float x, y, z; } F32x3;          /* :21-23 */
typedef struct F32x4 { float x, y, z, w; } F32x4;       /* :25-27 */
typedef struct U32x3 { uint32_t x, y, z; } U32x3;       /* :29-31 */
typedef struct Mat3x4 { F32x4 x, y, z; } Mat3x4;        /* :33-35, three rows */

typedef struct Geometry {                               /* :37-41; engine/hrtsystem/Accel.zig:40-44 */
    MeshHandle mesh;
    MaterialHandle material;
    bool sampled;
} Geometry;

typedef struct Extent2D { uint32_t width, height; } Extent2D;   /* :43-46 */

typedef struct Lens {                                   /* :48-55; shaders/hrtsystem/camera.hlsl:6-12 */
    F32x3 origin;
    F32x3 forward;
    F32x3 up;
    float vfov;
    float aperture;
    float focus_distance;
} Lens;

typedef struct Material {                               /* :57-64; hydra.zig:36-40 (always StandardPBR) */
    ImageHandle normal;
    ImageHandle emissive;
    ImageHandle color;
    ImageHandle metalness;
    ImageHandle roughness;
    float ior;
} Material;

typedef enum TextureFormat {                            /* :66-69; hydra.zig:42-59 */
    f16x4,
    u8x4_srgb,
} TextureFormat;

typedef struct HdMoonshine HdMoonshine;

MSNE_API HdMoonshine* HdMoonshineCreate(void);                                                  /* :72; hydra.zig:107-141 */
MSNE_API void HdMoonshineDestroy(HdMoonshine*);                                                 /* :73; hydra.zig:542-558 */
MSNE_API bool HdMoonshineRender(HdMoonshine*, SensorHandle, LensHandle);                        /* :74; hydra.zig:145-363 */
MSNE_API bool HdMoonshineRebuildPipeline(HdMoonshine*);                                         /* :75; hydra.zig:365-372 */
MSNE_API MeshHandle HdMoonshineCreateMesh(HdMoonshine*, const F32x3* positions, const F32x3* normals_or_null,
                                          const F32x2* texcoords_or_null, size_t position_count,
                                          const U32x3* indices, size_t triangle_count);        /* :76; hydra.zig:374-384 */
MSNE_API ImageHandle HdMoonshineCreateSolidTexture1(HdMoonshine*, float, const char* name);     /* :77; hydra.zig:386-392 */
MSNE_API ImageHandle HdMoonshineCreateSolidTexture2(HdMoonshine*, F32x2, const char* name);     /* :78; hydra.zig:394-400 */
MSNE_API ImageHandle HdMoonshineCreateSolidTexture3(HdMoonshine*, F32x3, const char* name);     /* :79; hydra.zig:402-408 */
MSNE_API ImageHandle HdMoonshineCreateRawTexture(HdMoonshine*, uint8_t* data, Extent2D, TextureFormat, const char* name); /* :80; hydra.zig:410-421 */
MSNE_API MaterialHandle HdMoonshineCreateMaterial(HdMoonshine*, Material);                      /* :81; hydra.zig:423-433 */
MSNE_API void HdMoonshineSetMaterialNormal(HdMoonshine*, MaterialHandle, ImageHandle);          /* :82; hydra.zig:435-441 */
MSNE_API void HdMoonshineSetMaterialEmissive(HdMoonshine*, MaterialHandle, ImageHandle);        /* :83; hydra.zig:443-449 */
MSNE_API void HdMoonshineSetMaterialColor(HdMoonshine*, MaterialHandle, ImageHandle);           /* :84; hydra.zig:451-457 */
MSNE_API void HdMoonshineSetMaterialMetalness(HdMoonshine*, MaterialHandle, ImageHandle);       /* :85; hydra.zig:459-465 */
MSNE_API void HdMoonshineSetMaterialRoughness(HdMoonshine*, MaterialHandle, ImageHandle);       /* :86; hydra.zig:467-473 */
MSNE_API void HdMoonshineSetMaterialIOR(HdMoonshine*, MaterialHandle, float);                   /* :87; hydra.zig:475-481 */
MSNE_API InstanceHandle HdMoonshineCreateInstance(HdMoonshine*, Mat3x4, const Geometry*, size_t geometry_count, bool visible); /* :88; hydra.zig:483-493 */
MSNE_API void HdMoonshineDestroyInstance(HdMoonshine*, InstanceHandle);                         /* :89; hydra.zig:495-497 */
MSNE_API void HdMoonshineSetInstanceTransform(HdMoonshine*, InstanceHandle, Mat3x4);            /* :90; hydra.zig:507-513 */
MSNE_API void HdMoonshineSetInstanceVisibility(HdMoonshine*, InstanceHandle, bool);             /* :91; hydra.zig:499-505 */
MSNE_API SensorHandle HdMoonshineCreateSensor(HdMoonshine*, Extent2D);                          /* :92; hydra.zig:515-520 */
MSNE_API float* HdMoonshineGetSensorData(const HdMoonshine*, SensorHandle);                     /* :93; hydra.zig:522-524 */
MSNE_API LensHandle HdMoonshineCreateLens(HdMoonshine*, Lens);                                  /* :94; hydra.zig:526-530 */
MSNE_API void HdMoonshineSetLens(HdMoonshine*, LensHandle, Lens);                               /* :95; hydra.zig:532-540 */
