// Host-side model of k_trace's traversal (moonshine_b200/csrc/traverse.cuh, single-level specialisation) used to
// rank BVH builder variants WITHOUT GPU time: builds the BVH with the product's builder (bvh_build.cpp), generates a
// realistic ray population by diffuse path tracing through the scene (camera rays, bounce rays, shadow rays towards a
// sun direction), walks every ray with the same while-while loop as the kernel and reports
//   * per-ray inner-node visits, leaf visits and triangle tests (closest-hit and any-hit rays separately),
//   * a lock-step SIMT estimate for warps of 32 consecutive rays: per outer iteration the warp pays the MAXIMUM over
//     its lanes of the inner-node steps, then the maximum of the triangle tests (what divergence costs on the GPU).
// Hit results never depend on the tree (closest = min t, id tie-break), so only these costs change between builders.
//
//   g++ -O2 -std=c++17 -I/usr/local/cuda/include -Imoonshine_b200/csrc tools/bvh_sim.cpp moonshine_b200/csrc/bvh_build.cpp -o /tmp/bvh_sim
//   /tmp/bvh_sim tris.bin [maxLeaf] [paths]        (tris.bin = float32 [n][9] world-space triangle soup)
// Builder options come from the same MSNE_BVH_* environment variables the runtime reads (bvh_build.cpp).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "bvh_build.h"

using namespace msne;

struct V3 { float x, y, z; };
static V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static V3 norm(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }

struct Ray { V3 o, d; float tmax; bool any; };
struct SpaceRay { V3 o, d, idir, ood; float rdd; };
static SpaceRay mk(const Ray& r) {
    SpaceRay s; s.o = r.o; s.d = r.d;
    auto inv = [](float d) { return std::fabs(d) > 1e-20f ? 1.0f / d : std::copysign(1e20f, d); };
    s.idir = {inv(r.d.x), inv(r.d.y), inv(r.d.z)};
    s.ood = {r.o.x * s.idir.x, r.o.y * s.idir.y, r.o.z * s.idir.z};
    s.rdd = 1.0f / dot(r.d, r.d);
    return s;
}
static bool slab(const SpaceRay& r, const BvhNode2& n, int s, float tmin, float tmax, float& tnear) {
    const float* xy = s == 0 ? n.n0 : n.n1;
    const float lox = xy[0], hix = xy[1], loy = xy[2], hiy = xy[3], loz = n.n2[2 * s], hiz = n.n2[2 * s + 1];
    const float tx0 = std::fma(lox, r.idir.x, -r.ood.x), tx1 = std::fma(hix, r.idir.x, -r.ood.x);
    const float ty0 = std::fma(loy, r.idir.y, -r.ood.y), ty1 = std::fma(hiy, r.idir.y, -r.ood.y);
    const float tz0 = std::fma(loz, r.idir.z, -r.ood.z), tz1 = std::fma(hiz, r.idir.z, -r.ood.z);
    tnear = std::fmax(std::fmax(std::fmin(tx0, tx1), std::fmin(ty0, ty1)), std::fmax(std::fmin(tz0, tz1), tmin));
    const float tfar = std::fmin(std::fmin(std::fmax(tx0, tx1), std::fmax(ty0, ty1)), std::fmin(std::fmax(tz0, tz1), tmax));
    return tnear <= tfar * 1.0000004f;
}
static bool triTest(const float* v, const SpaceRay& r, float tmin, float tmax, float& t) {
    const V3 A = V3{v[0], v[1], v[2]} - r.o, B = V3{v[3], v[4], v[5]} - r.o, C = V3{v[6], v[7], v[8]} - r.o;
    const float U = dot(r.d, cross(C, B)), V = dot(r.d, cross(A, C)), W = dot(r.d, cross(B, A));
    if ((U < 0 || V < 0 || W < 0) && (U > 0 || V > 0 || W > 0)) return false;
    const float det = (U + V) + W;
    if (det == 0) return false;
    const float rdet = 1.0f / det;
    const float T = (U * dot(A, r.d) + V * dot(B, r.d)) + W * dot(C, r.d);
    t = (T * rdet) * r.rdd;
    return tmin < t && t < tmax;
}

// One lane of the while-while walk, resumable so that 32 of them can be stepped in lock step.
struct Lane {
    SpaceRay r; bool any; float bestT; int hitTri = -1;
    int stack[128]; int sp = 0; int node; bool done = false;
    uint64_t nodes = 0, leaves = 0, tris = 0;
};
static const int END = 0x7FFFFFFF;
struct Scene { const BuiltBvh* bvh; const float* tri; };   // tri in ORIGINAL order; leaves index bvh->order

static void start(Lane& L, const Ray& ray) { L.r = mk(ray); L.any = ray.any; L.bestT = ray.tmax; L.sp = 0; L.stack[L.sp++] = END; L.node = 0; L.done = false; L.hitTri = -1; L.nodes = L.leaves = L.tris = 0; }
// encode: inner >= 0; leaf: -(1 + (first << 3 | (count-1)))
static int innerPhase(Lane& L, const Scene& sc) {   // returns steps taken
    int steps = 0;
    while (L.node >= 0 && L.node != END) {
        const BvhNode2& n = sc.bvh->nodes[L.node];
        steps++;
        float t0, t1;
        const bool h0 = n.c[0] != -1 || n.c[2] != 0 ? slab(L.r, n, 0, 0.0f, L.bestT, t0) : false;
        const bool h1 = n.c[1] != -1 || n.c[3] != 0 ? slab(L.r, n, 1, 0.0f, L.bestT, t1) : false;
        auto ref = [&](int s) { return n.c[2 + s] > 0 ? -(1 + ((n.c[s] << 3) | (n.c[2 + s] - 1))) : n.c[s]; };
        if (h0 && h1) { const bool sw = t1 < t0; L.node = sw ? ref(1) : ref(0); L.stack[L.sp++] = sw ? ref(0) : ref(1); }
        else if (h0) L.node = ref(0);
        else if (h1) L.node = ref(1);
        else L.node = L.stack[--L.sp];
    }
    L.nodes += steps;
    return steps;
}
static int leafPhase(Lane& L, const Scene& sc) {   // returns triangle tests
    int tests = 0;
    while (L.node < 0) {
        const uint32_t x = (uint32_t)(-(L.node + 1)), first = x >> 3, count = (x & 7u) + 1;
        L.leaves++;
        bool stop = false;
        for (uint32_t i = 0; i < count; i++) {
            tests++;
            float t;
            const uint32_t id = sc.bvh->order[first + i];
            if (!triTest(sc.tri + 9 * (size_t)id, L.r, 0.0f, L.any ? L.bestT : 3e38f, t)) continue;
            if (L.any) { L.hitTri = (int)id; stop = true; break; }
            if (t < L.bestT || (t == L.bestT && (int)id < L.hitTri)) { L.bestT = t; L.hitTri = (int)id; }
        }
        if (stop) { L.node = END; break; }
        L.node = L.stack[--L.sp];
    }
    L.tris += tests;
    return tests;
}
static void walk(Lane& L, const Scene& sc) { while (L.node != END) { innerPhase(L, sc); leafPhase(L, sc); } }

struct Rng { uint32_t s; float f() { s = s * 747796405u + 2891336453u; uint32_t w = ((s >> ((s >> 28) + 4)) ^ s) * 277803737u; return float(((w >> 22) ^ w) >> 8) * 0x1p-24f; } };

// --selftest: builder invariants on a random triangle soup (run by tests/test_bvh_host.py, no GPU involved)
//   every item sits in exactly one leaf, leaves hold 1..maxLeaf items, every child box contains its items, depth <= 72 (the
//   single-level kernel's stack needs no overflow test), the while-while walk finds the brute-force closest hit for random
//   rays, and reorderTopBfs is a pure relabelling (same hits, first k nodes in breadth-first order).
static int selftest() {
    Rng rng{4242};
    int failures = 0;
    for (uint32_t maxLeaf : {1u, 3u, 8u}) {
        for (uint32_t n : {1u, 2u, 7u, 500u, 20000u}) {
            std::vector<float> tri((size_t)n * 9);
            for (uint32_t i = 0; i < n; i++) {
                const float c[3] = {rng.f() * 10 - 5, rng.f() * 10 - 5, rng.f() * 2 - 1};
                const float sz = i % 97 == 0 ? 3.0f : 0.15f;   // a few large triangles among many small ones
                for (int v = 0; v < 3; v++) for (int a = 0; a < 3; a++) tri[(size_t)i * 9 + v * 3 + a] = c[a] + (rng.f() - 0.5f) * sz;
            }
            if (n >= 500) for (int a = 0; a < 9; a++) tri[9 * 5 + a] = tri[9 * 4 + a];   // an exact duplicate (coinciding centroids)
            std::vector<BuildBox> boxes(n);
            for (uint32_t i = 0; i < n; i++) for (int a = 0; a < 3; a++) {
                const float* v = &tri[(size_t)i * 9];
                boxes[i].lo[a] = std::min(v[a], std::min(v[3 + a], v[6 + a]));
                boxes[i].hi[a] = std::max(v[a], std::max(v[3 + a], v[6 + a]));
            }
            BuiltBvh bvh;
            buildBvh2(boxes.data(), n, maxLeaf, bvh);
            // structure
            std::vector<int> seen(n, 0);
            uint32_t maxDepth = 0;
            std::vector<std::pair<int, uint32_t>> st{{0, 1u}};
            while (!st.empty()) {
                auto [ni, depth] = st.back(); st.pop_back();
                maxDepth = std::max(maxDepth, depth);
                const BvhNode2& nd = bvh.nodes[ni];
                for (int s2 = 0; s2 < 2; s2++) {
                    const float* xy = s2 == 0 ? nd.n0 : nd.n1;
                    const float lo[3] = {xy[0], xy[2], nd.n2[2 * s2]}, hi[3] = {xy[1], xy[3], nd.n2[2 * s2 + 1]};
                    if (nd.c[2 + s2] > 0) {
                        if ((uint32_t)nd.c[2 + s2] > maxLeaf) { printf("leaf of %d items > maxLeaf %u\n", nd.c[2 + s2], maxLeaf); failures++; }
                        for (int k = 0; k < nd.c[2 + s2]; k++) {
                            const uint32_t id = bvh.order[nd.c[s2] + k];
                            seen[id]++;
                            for (int a = 0; a < 3; a++) if (boxes[id].lo[a] < lo[a] || boxes[id].hi[a] > hi[a]) { printf("item %u outside its leaf box\n", id); failures++; }
                        }
                    } else if (nd.c[s2] >= 0) st.push_back({nd.c[s2], depth + 1});
                }
            }
            for (uint32_t i = 0; i < n; i++) if (seen[i] != 1) { printf("n=%u leaf=%u: item %u referenced %d times\n", n, maxLeaf, i, seen[i]); failures++; break; }
            if (maxDepth > 72) { printf("depth %u > 72\n", maxDepth); failures++; }
            // traversal == brute force, before and after the breadth-first relabelling
            std::vector<Ray> rays;
            for (int r = 0; r < 300; r++) {
                const V3 o{rng.f() * 12 - 6, rng.f() * 12 - 6, rng.f() * 6 - 3};
                rays.push_back(Ray{o, norm(V3{rng.f() - 0.5f, rng.f() - 0.5f, rng.f() - 0.5f}), 3e38f, false});
            }
            std::vector<int> want(rays.size());
            for (size_t r = 0; r < rays.size(); r++) {
                const SpaceRay sr = mk(rays[r]);
                float best = 3e38f; int hit = -1;
                for (uint32_t i = 0; i < n; i++) { float t; if (triTest(&tri[(size_t)i * 9], sr, 0.0f, 3e38f, t) && (t < best || (t == best && (int)i < hit))) { best = t; hit = (int)i; } }
                want[r] = hit;
            }
            for (int pass = 0; pass < 2; pass++) {
                if (pass == 1) {
                    const std::vector<BvhNode2> before = bvh.nodes;
                    reorderTopBfs(bvh, 63);
                    if (bvh.nodes.size() != before.size()) { printf("reorderTopBfs changed the node count\n"); failures++; }
                    // breadth-first prefix: children of node i (i < k) that are inner nodes carry increasing indices > i
                    int last = 0;
                    for (size_t i = 0; i < std::min<size_t>(63, bvh.nodes.size()); i++)
                        for (int s2 = 0; s2 < 2; s2++) if (bvh.nodes[i].c[2 + s2] == 0 && bvh.nodes[i].c[s2] >= 0 && (size_t)bvh.nodes[i].c[s2] < std::min<size_t>(63, bvh.nodes.size())) {
                            if (bvh.nodes[i].c[s2] <= last) { printf("top of the tree is not in breadth-first order\n"); failures++; }
                            last = bvh.nodes[i].c[s2];
                        }
                }
                Scene sc{&bvh, tri.data()};
                for (size_t r = 0; r < rays.size(); r++) {
                    Lane L; start(L, rays[r]); walk(L, sc);
                    if (L.hitTri != want[r]) { printf("n=%u leaf=%u pass=%d ray %zu: bvh hit %d, brute force %d\n", n, maxLeaf, pass, r, L.hitTri, want[r]); failures++; break; }
                }
            }
        }
    }
    printf(failures ? "SELFTEST FAILED (%d)\n" : "selftest ok\n", failures);
    return failures ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "--selftest") == 0) return selftest();
    if (argc < 2) { fprintf(stderr, "usage: bvh_sim tris.bin [maxLeaf] [paths] | bvh_sim --selftest\n"); return 2; }
    const uint32_t maxLeaf = argc > 2 ? atoi(argv[2]) : 3;
    const uint32_t nPaths = argc > 3 ? atoi(argv[3]) : 65536;
    FILE* f = fopen(argv[1], "rb"); if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END); const size_t bytes = ftell(f); fseek(f, 0, SEEK_SET);
    const uint32_t n = (uint32_t)(bytes / 36);
    std::vector<float> tri((size_t)n * 9);
    if (fread(tri.data(), 36, n, f) != n) return 1;
    fclose(f);
    std::vector<BuildBox> boxes(n);
    for (uint32_t i = 0; i < n; i++) for (int a = 0; a < 3; a++) {
        const float* v = &tri[(size_t)i * 9];
        boxes[i].lo[a] = std::min(v[a], std::min(v[3 + a], v[6 + a]));
        boxes[i].hi[a] = std::max(v[a], std::max(v[3 + a], v[6 + a]));
    }
    BuiltBvh bvh;
    auto t0 = std::chrono::steady_clock::now();
    buildBvh2(boxes.data(), n, maxLeaf, bvh);
    const double buildS = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // SAH cost of the tree (root-normalised): sum over inner nodes of area + sum over leaves of area * count
    auto area = [](const BvhNode2& nd, int s) {
        const float* xy = s == 0 ? nd.n0 : nd.n1;
        const double dx = xy[1] - xy[0], dy = xy[3] - xy[2], dz = nd.n2[2 * s + 1] - nd.n2[2 * s];
        return dx * dy + dy * dz + dz * dx;
    };
    double sahInner = 0, sahLeaf = 0; size_t nLeaves = 0, refs = 0;
    for (const BvhNode2& nd : bvh.nodes) for (int s = 0; s < 2; s++) {
        if (nd.c[2 + s] > 0) { sahLeaf += area(nd, s) * nd.c[2 + s]; nLeaves++; refs += nd.c[2 + s]; }
        else if (nd.c[s] >= 0) sahInner += area(nd, s);
    }
    const double rootA = [&] { double dx = bvh.bounds.hi[0] - bvh.bounds.lo[0], dy = bvh.bounds.hi[1] - bvh.bounds.lo[1], dz = bvh.bounds.hi[2] - bvh.bounds.lo[2]; return dx * dy + dy * dz + dz * dx; }();
    printf("tris %u  nodes %zu  leaves %zu  refs %zu  build %.1f s  SAH inner %.2f leaf %.2f\n", n, bvh.nodes.size(), nLeaves, refs, buildS, sahInner / rootA, sahLeaf / rootA);

    Scene sc{&bvh, tri.data()};
    // ---- ray population: camera rays of the cfg3 lens over 8x4 pixel blocks, diffuse bounces, sun shadow rays ----------
    const V3 eye{0.35f, 0.55f, 1.55f}, fw = norm(V3{0.93f, 0.28f, -0.22f}), up{0, 0, 1};
    const V3 right = norm(cross(fw, up)), upv = cross(right, fw);
    const float th = std::tan(0.5f * 1.0122909f), aspect = 1920.0f / 1080.0f;
    const V3 sun = norm(V3{0.15f, 0.75f, 0.64f});
    const int maxB = 6;
    std::vector<std::vector<Ray>> byBounce(maxB), shadowBy(maxB);
    Rng rng{12345};
    const uint32_t side = (uint32_t)std::sqrt((double)nPaths);
    for (uint32_t p = 0; p < side * side; p++) {
        // 8x4 blocks in raster order like k_generate's pixel blocks
        const uint32_t blk = p / 32, in = p % 32, bw = side / 8;
        const uint32_t px = (blk % bw) * 8 + in % 8, py = (blk / bw) * 4 + in / 8;
        if (py >= side) continue;
        const float u = ((px + rng.f()) / side * 2 - 1) * th * aspect, v = (1 - (py + rng.f()) / side * 2) * th;
        Ray r{eye, norm(fw + right * u + upv * v), 3e38f, false};
        for (int b = 0; b < maxB; b++) {
            byBounce[b].push_back(r);
            Lane L; start(L, r); walk(L, sc);
            if (L.hitTri < 0) break;
            const float* tv = &tri[(size_t)L.hitTri * 9];
            V3 ng = norm(cross(V3{tv[3], tv[4], tv[5]} - V3{tv[0], tv[1], tv[2]}, V3{tv[6], tv[7], tv[8]} - V3{tv[0], tv[1], tv[2]}));
            if (dot(ng, r.d) > 0) ng = ng * -1.0f;
            const V3 pos = r.o + r.d * L.bestT + ng * 1e-4f;
            shadowBy[b].push_back(Ray{pos, sun, 3e38f, true});
            // cosine hemisphere
            const float r1 = rng.f(), r2 = rng.f(), rad = std::sqrt(r1), phi = 6.2831853f * r2;
            const V3 tng = norm(std::fabs(ng.x) > 0.5f ? cross(ng, V3{0, 1, 0}) : cross(ng, V3{1, 0, 0})), btg = cross(ng, tng);
            r = Ray{pos, norm(tng * (rad * std::cos(phi)) + btg * (rad * std::sin(phi)) + ng * std::sqrt(std::fmax(0.0f, 1 - r1))), 3e38f, false};
            if (b >= 3 && rng.f() > 0.7f) break;
        }
    }
    // ---- measurement --------------------------------------------------------------------------------------------------
    const double CN = 40, CT = 75;   // warp instructions per inner step / per triangle test (SASS of k_trace<false>)
    double totalCost = 0; uint64_t totalRays = 0;
    auto measure = [&](const char* name, const std::vector<Ray>& rays) {
        uint64_t nodes = 0, leaves = 0, tris = 0, warpNode = 0, warpTri = 0, warps = 0;
        uint64_t laneHist[33] = {0}; uint64_t idleDone = 0, idleWait = 0;
        std::vector<Lane> L(32);
        for (size_t base = 0; base < rays.size(); base += 32) {
            const size_t cnt = std::min<size_t>(32, rays.size() - base);
            for (size_t i = 0; i < cnt; i++) start(L[i], rays[base + i]);
            warps++;
            for (;;) {
                int mN = 0, mT = 0; bool any = false;
                int steps[32]; int nAct = 0;
                for (size_t i = 0; i < cnt; i++) { steps[i] = 0; if (L[i].node != END) { any = true; steps[i] = innerPhase(L[i], sc); mN = std::max(mN, steps[i]); nAct++; } }
                if (!any) break;
                // active-lane histogram of the node steps of this phase: step k of the phase is executed by the lanes with steps > k
                for (int k = 0; k < mN; k++) { int a = 0; for (size_t i = 0; i < cnt; i++) a += steps[i] > k; laneHist[a]++; }
                // idle lane-steps of this phase, split by cause: the lane's ray is finished / the lane waits at a leaf for the longest walk
                idleDone += (uint64_t)mN * (32 - nAct);
                for (size_t i = 0; i < cnt; i++) if (steps[i] || L[i].node != END) idleWait += (uint64_t)(mN - steps[i]);
                for (size_t i = 0; i < cnt; i++) if (L[i].node != END) mT = std::max(mT, leafPhase(L[i], sc));
                warpNode += mN; warpTri += mT;
            }
            for (size_t i = 0; i < cnt; i++) { nodes += L[i].nodes; leaves += L[i].leaves; tris += L[i].tris; }
        }
        if (getenv("MSNE_SIM_HIST")) {
            uint64_t tot = 0; for (int a = 0; a <= 32; a++) tot += laneHist[a];
            uint64_t acc = 0; printf("   node steps by active lanes (cumulative %%):");
            for (int a = 1; a <= 32; a++) { acc += laneHist[a]; if (a == 1 || a == 2 || a == 4 || a == 8 || a == 16 || a == 24 || a == 32) printf(" <=%d: %.0f", a, 100.0 * acc / tot); }
            printf("\n   lane-steps: useful %.0f %%, idle because the ray is finished %.0f %%, idle at a leaf waiting for the longest walk %.0f %%\n",
                   100.0 * nodes / (32.0 * warpNode), 100.0 * idleDone / (32.0 * warpNode), 100.0 * idleWait / (32.0 * warpNode));
        }
        const double nr = (double)rays.size();
        const double cost = (warpNode * CN + warpTri * CT);
        printf("%-10s rays %8zu | per ray: nodes %6.1f leaves %5.2f tris %5.2f | per warp: node-steps %6.1f tri-steps %5.1f | cost/ray %7.1f\n", name, rays.size(),
               nodes / nr, leaves / nr, tris / nr, (double)warpNode / warps, (double)warpTri / warps, cost / nr);
        totalCost += cost; totalRays += rays.size();
    };
    char nm[32];
    // optional ray reordering before the warps are formed (MSNE_SIM_SORT): 1 = stable by direction octant, 2 = octant then
    // origin cell (4 bits per axis, Morton), 3 = origin cell then octant, 4 = origin cell (5 bits) + octahedral direction (4x4)
    const int sortMode = getenv("MSNE_SIM_SORT") ? atoi(getenv("MSNE_SIM_SORT")) : 0;
    if (sortMode) {
        auto part = [](uint32_t v) { v &= 1023; v = (v | (v << 16)) & 0x30000FF; v = (v | (v << 8)) & 0x300F00F; v = (v | (v << 4)) & 0x30C30C3; v = (v | (v << 2)) & 0x9249249; return v; };
        auto key = [&](const Ray& r) -> uint64_t {
            const uint32_t oct = (r.d.x < 0) | ((r.d.y < 0) << 1) | ((r.d.z < 0) << 2);
            auto cell = [&](int bits) {
                const float sc[3] = {(r.o.x - bvh.bounds.lo[0]) / (bvh.bounds.hi[0] - bvh.bounds.lo[0]), (r.o.y - bvh.bounds.lo[1]) / (bvh.bounds.hi[1] - bvh.bounds.lo[1]),
                                     (r.o.z - bvh.bounds.lo[2]) / (bvh.bounds.hi[2] - bvh.bounds.lo[2])};
                uint32_t c[3];
                for (int a = 0; a < 3; a++) c[a] = (uint32_t)std::min<float>((1 << bits) - 1, std::max(0.0f, sc[a] * (1 << bits)));
                return (uint64_t)(part(c[0]) | (part(c[1]) << 1) | (part(c[2]) << 2));
            };
            if (sortMode == 1) return oct;
            if (sortMode == 2) return ((uint64_t)oct << 32) | cell(4);
            if (sortMode == 3) return (cell(4) << 3) | oct;
            const float an = std::fabs(r.d.x) + std::fabs(r.d.y) + std::fabs(r.d.z);
            const float u = r.d.x / an, v = r.d.y / an;
            const uint32_t du = (uint32_t)std::min(3.0f, (u * 0.5f + 0.5f) * 4), dv = (uint32_t)std::min(3.0f, (v * 0.5f + 0.5f) * 4);
            return (cell(5) << 8) | (oct << 4) | (du << 2) | dv;
        };
        for (int b = 1; b < maxB; b++)
            for (auto* v : {&byBounce[b], &shadowBy[b]})
                std::stable_sort(v->begin(), v->end(), [&](const Ray& a, const Ray& c) { return key(a) < key(c); });
    }
    for (int b = 0; b < maxB; b++) { snprintf(nm, sizeof nm, "ext b%d", b); measure(nm, byBounce[b]); }
    for (int b = 0; b < maxB; b++) { snprintf(nm, sizeof nm, "shadow b%d", b); measure(nm, shadowBy[b]); }
    printf("TOTAL warp cost per ray %.1f over %llu rays\n", totalCost / totalRays, (unsigned long long)totalRays);
    // MSNE_SIM_VOTE: model of a vote-scheduled loop - every iteration the warp runs EITHER one node step (for the lanes that
    // hold an inner node) OR one leaf visit (for the lanes that hold a leaf), chosen by lane counts: leaf visit when
    // nLeaf >= T or no lane holds an inner node. T = 33: today's while-while (leaves only when nobody can step).
    if (getenv("MSNE_SIM_VOTE")) {
        const double CB = 6;   // ballots + branch of the vote
        for (int b = 2; b <= 3; b++) {
            const std::vector<Ray>& rays = byBounce[b];
            for (int T : {33, 24, 16, 12, 8, 4, 1}) {
                uint64_t nodeIt = 0, triIt = 0, its = 0;
                std::vector<Lane> L(32);
                for (size_t base = 0; base < rays.size(); base += 32) {
                    const size_t cnt = std::min<size_t>(32, rays.size() - base);
                    for (size_t i = 0; i < cnt; i++) start(L[i], rays[base + i]);
                    for (;;) {
                        int nI = 0, nL = 0;
                        for (size_t i = 0; i < cnt; i++) if (L[i].node != END) { if (L[i].node >= 0) nI++; else nL++; }
                        if (nI + nL == 0) break;
                        its++;
                        if (nL >= T || nI == 0) {
                            int mT = 0;
                            for (size_t i = 0; i < cnt; i++) if (L[i].node != END && L[i].node < 0) {
                                // one leaf visit: the body of leafPhase for a single leaf
                                Lane& l = L[i];
                                const uint32_t x = (uint32_t)(-(l.node + 1)), first = x >> 3, count = (x & 7u) + 1;
                                int tests = 0; bool stop = false;
                                for (uint32_t k = 0; k < count; k++) {
                                    tests++; float t; const uint32_t id = sc.bvh->order[first + k];
                                    if (!triTest(sc.tri + 9 * (size_t)id, l.r, 0.0f, l.any ? l.bestT : 3e38f, t)) continue;
                                    if (l.any) { l.hitTri = (int)id; stop = true; break; }
                                    if (t < l.bestT || (t == l.bestT && (int)id < l.hitTri)) { l.bestT = t; l.hitTri = (int)id; }
                                }
                                l.node = stop ? END : l.stack[--l.sp];
                                mT = std::max(mT, tests);
                            }
                            triIt += mT;
                        } else {
                            for (size_t i = 0; i < cnt; i++) if (L[i].node != END && L[i].node >= 0) {
                                Lane& l = L[i];
                                const BvhNode2& n = sc.bvh->nodes[l.node];
                                float t0, t1;
                                const bool h0 = n.c[0] != -1 || n.c[2] != 0 ? slab(l.r, n, 0, 0.0f, l.bestT, t0) : false;
                                const bool h1 = n.c[1] != -1 || n.c[3] != 0 ? slab(l.r, n, 1, 0.0f, l.bestT, t1) : false;
                                auto ref = [&](int s2) { return n.c[2 + s2] > 0 ? -(1 + ((n.c[s2] << 3) | (n.c[2 + s2] - 1))) : n.c[s2]; };
                                if (h0 && h1) { const bool sw = t1 < t0; l.node = sw ? ref(1) : ref(0); l.stack[l.sp++] = sw ? ref(0) : ref(1); }
                                else if (h0) l.node = ref(0);
                                else if (h1) l.node = ref(1);
                                else l.node = l.stack[--l.sp];
                            }
                            nodeIt++;
                        }
                    }
                }
                printf("vote b%d T=%2d: cost/ray %6.1f (+%.1f with the vote) node-its %.2f tri-its %.2f per ray\n", b, T, (nodeIt * CN + triIt * CT) / rays.size(),
                       its * CB / rays.size(), (double)nodeIt / rays.size(), (double)triIt / rays.size());
            }
        }
    }
    // MSNE_SIM_SUSPEND: model of suspend-and-regroup - when, after an outer iteration, at most T lanes of a warp are still walking
    // (and the warp has run at least MIN node steps), those lanes write their traversal state to a continuation list and the warp
    // takes its next ticket; the continuations are walked 32 at a time in a second pass (same rule), a third pass runs to the end.
    if (getenv("MSNE_SIM_SUSPEND")) {
        const double CS = 80, CR = 60;   // warp instructions of one suspend event (<= T lanes store state + stack) / one resumed warp (loads)
        for (int b = 2; b <= 3; b++) {
            for (int kind = 0; kind < 2; kind++) {
                const std::vector<Ray>& rays = kind == 0 ? byBounce[b] : shadowBy[b];
                for (int T : {0, 2, 4, 6, 8, 12, 16}) {
                    uint64_t warpNode = 0, warpTri = 0, suspends = 0, resumedWarps = 0, suspendedRays = 0;
                    std::vector<Lane> cur, next;
                    cur.resize(rays.size());
                    for (size_t i = 0; i < rays.size(); i++) start(cur[i], rays[i]);
                    for (int pass = 0; pass < 3 && !cur.empty(); pass++) {
                        next.clear();
                        const int Tp = pass == 2 ? 0 : T;
                        for (size_t base = 0; base < cur.size(); base += 32) {
                            const size_t cnt = std::min<size_t>(32, cur.size() - base);
                            Lane* L = &cur[base];
                            if (pass) resumedWarps++;
                            for (;;) {
                                int mN = 0, mT = 0, act = 0;
                                for (size_t i = 0; i < cnt; i++) if (L[i].node != END) mN = std::max(mN, innerPhase(L[i], sc));
                                for (size_t i = 0; i < cnt; i++) if (L[i].node != END) mT = std::max(mT, leafPhase(L[i], sc));
                                warpNode += mN; warpTri += mT;
                                for (size_t i = 0; i < cnt; i++) act += L[i].node != END;
                                if (act == 0) break;
                                if (act <= Tp) {
                                    for (size_t i = 0; i < cnt; i++) if (L[i].node != END) { next.push_back(L[i]); suspendedRays++; }
                                    suspends++;
                                    break;
                                }
                            }
                        }
                        cur.swap(next);
                    }
                    const double cost = warpNode * CN + warpTri * CT + suspends * CS + resumedWarps * CR;
                    printf("suspend %s b%d T=%2d: cost/ray %6.1f (node-steps %.2f tri-steps %.2f per ray; %.1f %% of the rays suspended)\n", kind ? "shadow" : "ext   ", b, T,
                           cost / rays.size(), (double)warpNode / rays.size(), (double)warpTri / rays.size(), 100.0 * suspendedRays / rays.size());
                }
            }
        }
    }
    // MSNE_SIM_REFILL=R: model of dynamic fetch - at the top of an outer iteration, when at least R lanes are idle, the idle
    // lanes take the next rays of the (sorted) stream instead of waiting for the whole ticket (R = 32: today's kernel).
    if (getenv("MSNE_SIM_REFILL")) {
        const double CF = 40;   // warp instructions of one fetch (ticket atomic, ray loads, setup)
        for (int b = 2; b <= 3; b++) {
            const std::vector<Ray>& rays = byBounce[b];
            for (int R : {32, 24, 16, 8, 4, 1}) {
                std::vector<Lane> L(32);
                for (auto& l : L) l.node = END;
                size_t next = 0; uint64_t warpNode = 0, warpTri = 0, fetches = 0;
                for (;;) {
                    int idle = 0; for (auto& l : L) idle += l.node == END;
                    if (idle == 32 && next >= rays.size()) break;
                    if ((idle >= R || idle == 32) && next < rays.size()) {
                        for (auto& l : L) if (l.node == END && next < rays.size()) start(l, rays[next++]);
                        fetches++;
                    }
                    int mN = 0, mT = 0;
                    for (auto& l : L) if (l.node != END) mN = std::max(mN, innerPhase(l, sc));
                    for (auto& l : L) if (l.node != END) mT = std::max(mT, leafPhase(l, sc));
                    warpNode += mN; warpTri += mT;
                }
                printf("refill b%d R=%2d: cost/ray %6.1f (node-steps %.2f, tri-steps %.2f, fetches %.3f per ray)\n", b, R,
                       (warpNode * CN + warpTri * CT + fetches * CF) / rays.size(), (double)warpNode / rays.size(), (double)warpTri / rays.size(), (double)fetches / rays.size());
            }
        }
    }
    return 0;
}
