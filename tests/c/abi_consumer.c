/* A compiled C11 consumer of the drop-in boundary (include/moonshine.h + moonshine_ext.h), the way the reference's Hydra
 * delegate (hydra/ *.cpp, renderDelegate.cpp:56-65, mesh.cpp:169-264, renderPass.cpp:25) consumes hydra/moonshine.h: it links
 * libmoonshine.so, passes `bool` and 48-byte structs BY VALUE across the ABI, and never sees a C++ or Python type.
 *
 *   abi_consumer render <out.f32>   cfg1 (BASELINE.json configs[0]: one Lambert triangle, 1x1 white env, 64x64, 1 spp, offline
 *                                   constants) through HdMoonshineRender; writes the 64*64*4 float sensor buffer
 *   abi_consumer stress [n]         (n: render-ahead depth, default off) 8 threads create textures / materials / meshes / instances and edit them while the main
 *                                   thread renders: the per-object mutex (hydra.zig:78) must serialise everything
 * Test infrastructure (tests/test_c_consumer.py builds and runs it); not part of the product.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "moonshine_ext.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "abi_consumer: %s failed at line %d: %s\n", #c, __LINE__, HdMoonshineExtLastError()); exit(2); } } while (0)

static const Mat3x4 kIdentity = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};

static MaterialHandle lambert(HdMoonshine* m, float r, float g, float b) {
    const float black[4] = {0, 0, 0, 1}, color[4] = {r, g, b, 1}, flat[2] = {0.5f, 0.5f};
    const Extent2D one = {1, 1};
    MsneMaterialInfo info;
    memset(&info, 0, sizeof info);
    info.type = MSNE_MAT_LAMBERT;
    info.normal = HdMoonshineExtCreateTexture(m, flat, one, MSNE_TEX_RG32F, "flat normal");
    info.emissive = HdMoonshineExtCreateTexture(m, black, one, MSNE_TEX_RGBA32F, "black");
    info.color = HdMoonshineExtCreateTexture(m, color, one, MSNE_TEX_RGBA32F, "albedo");
    info.ior = 1.5f;
    return HdMoonshineExtCreateMaterial(m, &info);
}

static int render_cfg1(const char* out_path) {
    HdMoonshine* m = HdMoonshineCreate();
    CHECK(m != NULL);
    const MsneSpecConstants offline = {1, 1024, 1, 1, 1, 1, 1};   /* offline/main.zig:106-111 */
    CHECK(HdMoonshineExtSetSpecConstants(m, &offline));
    const MaterialHandle mat = lambert(m, 0.8f, 0.8f, 0.8f);
    const F32x3 pos[3] = {{0, -1, -1}, {0, 1, -1}, {0, 0, 1}};
    const U32x3 idx[1] = {{0, 1, 2}};
    const MeshHandle mesh = HdMoonshineExtCreateMeshIndexed(m, pos, NULL, NULL, 3, idx, 1);
    const Geometry geom = {mesh, mat, false};
    const InstanceHandle inst = HdMoonshineCreateInstance(m, kIdentity, &geom, 1, true);
    CHECK(inst == 0);
    const float white[4] = {1, 1, 1, 1};
    const Extent2D one = {1, 1}, extent = {64, 64};
    CHECK(HdMoonshineExtSetBackground(m, white, one));
    CHECK(HdMoonshineExtBuildEmitterTable(m));
    const SensorHandle sensor = HdMoonshineCreateSensor(m, extent);
    const Lens lens = {{-3, 0, 0}, {1, 0, 0}, {0, 0, 1}, 3.14159265358979323846f / 4.0f, 0.0f, 1.0f};   /* engine/tests.zig:302-309 */
    const LensHandle lh = HdMoonshineCreateLens(m, lens);
    const float* pixels = HdMoonshineGetSensorData(m, sensor);
    CHECK(pixels != NULL);
    CHECK(HdMoonshineRender(m, sensor, lh));
    CHECK(HdMoonshineExtGetSensorSampleCount(m, sensor) == 1);
    MsneStats st;
    HdMoonshineExtGetStats(m, &st);
    printf("rays_closest=%llu rays_shadow=%llu launches=%llu center=%.9g\n", (unsigned long long)st.rays_closest, (unsigned long long)st.rays_shadow,
           (unsigned long long)st.kernel_launches, pixels[4 * (32 * 64 + 32)]);
    FILE* f = fopen(out_path, "wb");
    CHECK(f != NULL);
    CHECK(fwrite(pixels, sizeof(float), 64 * 64 * 4, f) == 64 * 64 * 4);
    fclose(f);
    HdMoonshineDestroy(m);
    return 0;
}

/* ---- stress ---------------------------------------------------------------------------------------------- */
enum { kThreads = 8, kPerThread = 24 };
typedef struct Worker { HdMoonshine* m; int id; InstanceHandle made[kPerThread]; } Worker;

static void* worker(void* arg) {
    Worker* w = (Worker*)arg;
    for (int k = 0; k < kPerThread; k++) {
        const float s = 0.05f + 0.01f * (float)w->id, x = 0.3f * (float)k - 3.0f, y = 0.4f * (float)w->id - 1.4f;
        const MaterialHandle mat = lambert(w->m, 0.1f * (float)(w->id + 1), 0.5f, 0.05f * (float)(k % 10));
        const F32x3 pos[4] = {{x, y, 0}, {x + s, y, 0}, {x + s, y + s, 0.1f}, {x, y + s, 0.1f}};
        const F32x3 nrm[6] = {{0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}};   /* per corner (hydra.zig:379-380) */
        const U32x3 idx[2] = {{0, 1, 2}, {0, 2, 3}};
        const MeshHandle mesh = HdMoonshineCreateMesh(w->m, pos, nrm, NULL, 4, idx, 2);
        const Geometry geom = {mesh, mat, (k % 5) == 0};
        Mat3x4 t = kIdentity;
        if (k % 2) { t.x.w = 0.01f * (float)k; t.z.w = 0.2f; }   /* every other one is a transformed instance (own BLAS under the TLAS) */
        w->made[k] = HdMoonshineCreateInstance(w->m, t, &geom, 1, true);
        if (k % 3 == 0) HdMoonshineSetInstanceVisibility(w->m, w->made[k], false);
        if (k % 4 == 1) { t.y.w = 0.05f; HdMoonshineSetInstanceTransform(w->m, w->made[k], t); }
        if (k % 6 == 2) HdMoonshineSetMaterialIOR(w->m, mat, 1.33f);
    }
    return NULL;
}

static int stress(int render_ahead) {
    HdMoonshine* m = HdMoonshineCreate();
    CHECK(m != NULL);
    if (render_ahead) HdMoonshineExtSetRenderAhead(m, (uint32_t)render_ahead);   /* edits from the workers keep dropping the waves rendered ahead */
    const Extent2D extent = {96, 64};
    const SensorHandle sensor = HdMoonshineCreateSensor(m, extent);
    const Lens lens = {{0, -6, 3}, {0, 0.8944272f, -0.4472136f}, {0, 0, 1}, 0.9f, 0.0f, 1.0f};
    const LensHandle lh = HdMoonshineCreateLens(m, lens);
    pthread_t th[kThreads];
    Worker w[kThreads];
    for (int i = 0; i < kThreads; i++) { w[i].m = m; w[i].id = i; CHECK(pthread_create(&th[i], NULL, worker, &w[i]) == 0); }
    int renders = 0;
    for (int r = 0; r < 40; r++) { CHECK(HdMoonshineRender(m, sensor, lh)); renders++; (void)HdMoonshineGetSensorData(m, sensor); }
    for (int i = 0; i < kThreads; i++) pthread_join(th[i], NULL);
    CHECK(HdMoonshineRender(m, sensor, lh));
    /* handles are dense indices in creation order (hydra.zig): all kThreads*kPerThread instances got distinct handles */
    char seen[kThreads * kPerThread];
    memset(seen, 0, sizeof seen);
    for (int i = 0; i < kThreads; i++)
        for (int k = 0; k < kPerThread; k++) {
            CHECK(w[i].made[k] < (InstanceHandle)(kThreads * kPerThread));
            CHECK(!seen[w[i].made[k]]);
            seen[w[i].made[k]] = 1;
        }
    const float* px = HdMoonshineGetSensorData(m, sensor);
    double sum = 0;
    for (int i = 0; i < 96 * 64; i++) { CHECK(px[4 * i] == px[4 * i]); sum += px[4 * i]; }   /* no NaN */
    MsneStats st;
    HdMoonshineExtGetStats(m, &st);
    printf("stress ok: %d renders during edits, %d instances, mean red %.6f, rays %llu, waves dropped %llu\n", renders, kThreads * kPerThread, sum / (96 * 64),
           (unsigned long long)(st.rays_closest + st.rays_shadow), (unsigned long long)st.ahead_discarded);
    HdMoonshineDestroy(m);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 3 && strcmp(argv[1], "render") == 0) return render_cfg1(argv[2]);
    if (argc >= 2 && strcmp(argv[1], "stress") == 0) return stress(argc >= 3 ? atoi(argv[2]) : 0);
    fprintf(stderr, "usage: abi_consumer render <out.f32> | stress [render_ahead_waves]\n");
    return 64;
}
