// Host-side check of the BVH2 -> compressed BVH8 collapse (moonshine_b200/csrc/bvh_build.cpp collapseToBvh8):
// structural invariants + a scalar restatement of the device traversal (traverse.cuh traceRay8, one level) whose
// closest hits must equal brute force over all triangles with the project's watertight test. Test infrastructure.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../moonshine_b200/csrc/bvh_build.h"

using namespace msne;

struct V3 { float x, y, z; };
static V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
struct Tri { V3 a, b, c; };

static bool triTest(const Tri& tr, V3 o, V3 d, float tmin, float tmax, float& t) {
    const V3 A = sub(tr.a, o), B = sub(tr.b, o), C = sub(tr.c, o);
    const float U = dot(d, cross(C, B)), V = dot(d, cross(A, C)), W = dot(d, cross(B, A));
    if ((U < 0 || V < 0 || W < 0) && (U > 0 || V > 0 || W > 0)) return false;
    const float det = (U + V) + W;
    if (det == 0.0f) return false;
    const float rdet = 1.0f / det;
    const float T = (U * dot(A, d) + V * dot(B, d)) + W * dot(C, d);
    t = (T * rdet) * (1.0f / dot(d, d));
    return tmin < t && t < tmax;
}

static uint32_t highBit(uint32_t x) { return 31u - (uint32_t)__builtin_clz(x); }
static float asFloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct Hit { float t; int tri; };

static Hit traverse8(const BuiltBvh8& b, const std::vector<Tri>& tris, V3 o, V3 d, uint64_t& nodesVisited, uint32_t& maxSp) {
    Hit best{1e30f, -1};
    float idir[3], od[3] = {d.x, d.y, d.z}, oo[3] = {o.x, o.y, o.z};
    for (int a = 0; a < 3; a++) idir[a] = std::fabs(od[a]) > 1e-20f ? 1.0f / od[a] : std::copysign(1e20f, od[a]);
    const uint32_t oct = (idir[0] < 0 ? 1u : 0u) | (idir[1] < 0 ? 2u : 0u) | (idir[2] < 0 ? 4u : 0u);
    const uint32_t octInv = 7u - oct;
    struct G { uint32_t x, y; };
    G stack[64]; uint32_t sp = 0;
    G ng{0, 0x80000000u}, tg{0, 0};
    for (;;) {
        if (ng.y > 0x00FFFFFFu) {
            const uint32_t imask = ng.y & 0xFFu, bit = highBit(ng.y);
            ng.y &= ~(1u << bit);
            if (ng.y > 0x00FFFFFFu) { stack[sp++] = ng; if (sp > maxSp) maxSp = sp; }
            const uint32_t slot = (bit - 24u) ^ octInv;
            const uint32_t rel = (uint32_t)__builtin_popcount(imask & ~(0xFFFFFFFFu << slot));
            const BvhNode8& n = b.nodes[ng.x + rel];
            nodesVisited++;
            uint32_t hits = 0;
            float ax[3], bx[3];
            for (int a = 0; a < 3; a++) { ax[a] = asFloat((uint32_t)n.e[a] << 23) * idir[a]; bx[a] = (n.p[a] - oo[a]) * idir[a]; }
            for (int s = 0; s < 8; s++) {
                const uint32_t meta = n.meta[s];
                const bool inner = (meta & 0x18u) == 0x18u && (meta >> 5) == 1u;
                float tn = 0.0f, tf = best.t;
                for (int a = 0; a < 3; a++) {
                    const float q0 = idir[a] < 0 ? n.qhi[a][s] : n.qlo[a][s], q1 = idir[a] < 0 ? n.qlo[a][s] : n.qhi[a][s];
                    tn = std::fmax(tn, std::fmaf(q0, ax[a], bx[a]));
                    tf = std::fmin(tf, std::fmaf(q1, ax[a], bx[a]));
                }
                if (!(tn <= tf * 1.0000004f)) continue;
                const uint32_t bitIndex = ((meta & 0x18u) == 0x18u ? (meta ^ octInv) : meta) & 0x1Fu;
                (void)inner;
                hits |= (meta >> 5) << bitIndex;
            }
            ng.x = n.childBase; ng.y = (hits & 0xFF000000u) | n.imask;
            tg.x = n.itemBase; tg.y = hits & 0x00FFFFFFu;
        } else { tg = ng; ng = G{0, 0}; }
        while (tg.y) {
            const uint32_t bit = highBit(tg.y);
            tg.y &= ~(1u << bit);
            const int ti = (int)b.order[tg.x + bit];
            float t;
            if (triTest(tris[ti], o, d, 0.0f, 1e30f, t) && (t < best.t || (t == best.t && ti < best.tri))) best = Hit{t, ti};
        }
        if (ng.y <= 0x00FFFFFFu) {
            if (sp == 0) return best;
            ng = stack[--sp];
        }
    }
}

int main(int argc, char** argv) {
    const int nTri = argc > 1 ? atoi(argv[1]) : 20000, nRays = argc > 2 ? atoi(argv[2]) : 4000;
    std::mt19937 rng(1234);
    std::uniform_real_distribution<float> U(-1.0f, 1.0f);
    std::vector<Tri> tris(nTri);
    std::vector<BuildBox> boxes(nTri);
    for (int i = 0; i < nTri; i++) {
        // clustered + axis-aligned flat triangles (degenerate boxes) + a few huge ones
        const float s = (i % 97 == 0) ? 1.5f : 0.03f;
        const V3 c{U(rng) * 2.0f, U(rng) * 2.0f, U(rng) * 2.0f};
        Tri t{{c.x + U(rng) * s, c.y + U(rng) * s, c.z + U(rng) * s}, {c.x + U(rng) * s, c.y + U(rng) * s, c.z + U(rng) * s}, {c.x + U(rng) * s, c.y + U(rng) * s, c.z + U(rng) * s}};
        if (i % 3 == 0) { const float z = std::round(c.z * 4.0f) / 4.0f; t.a.z = t.b.z = t.c.z = z; }
        tris[i] = t;
        const float* v[3] = {&t.a.x, &t.b.x, &t.c.x};
        for (int a = 0; a < 3; a++) {
            boxes[i].lo[a] = std::fmin(v[0][a], std::fmin(v[1][a], v[2][a]));
            boxes[i].hi[a] = std::fmax(v[0][a], std::fmax(v[1][a], v[2][a]));
        }
    }
    for (int maxLeaf = 1; maxLeaf <= 3; maxLeaf++) {
        BuiltBvh b2;
        buildBvh2(boxes.data(), (uint32_t)nTri, (uint32_t)maxLeaf, b2);
        BuiltBvh8 b8;
        collapseToBvh8(b2, b8);
        // --- structure: every item once; children indices in range; item masks inside 24 bits
        std::vector<int> seen(nTri, 0);
        if ((int)b8.order.size() != nTri) { printf("FAIL order size %zu\n", b8.order.size()); return 1; }
        for (uint32_t i : b8.order) seen[i]++;
        for (int i = 0; i < nTri; i++) if (seen[i] != 1) { printf("FAIL item %d seen %d times\n", i, seen[i]); return 1; }
        size_t innerRefs = 0;
        for (const BvhNode8& n : b8.nodes) {
            uint32_t items = 0;
            for (int s = 0; s < 8; s++) {
                const uint32_t meta = n.meta[s];
                if (meta == 0) continue;
                if ((n.imask >> s) & 1) {
                    if (meta != (0x20u | (24u + s))) { printf("FAIL inner meta\n"); return 1; }
                    innerRefs++;
                    const uint32_t idx = n.childBase + (uint32_t)__builtin_popcount(n.imask & ((1u << s) - 1u));
                    if (idx >= b8.nodes.size()) { printf("FAIL child index\n"); return 1; }
                } else {
                    const uint32_t cnt = (uint32_t)__builtin_popcount(meta >> 5), off = meta & 31u;
                    if (((meta >> 5) & ((meta >> 5) + 1)) != 0 || off + cnt > 24 || off != items) { printf("FAIL leaf meta %x\n", meta); return 1; }
                    items += cnt;
                    if (n.itemBase + off + cnt > b8.order.size()) { printf("FAIL item range\n"); return 1; }
                    // decoded box must contain every item box of the leaf
                    for (uint32_t k = 0; k < cnt; k++) {
                        const BuildBox& ib = boxes[b8.order[n.itemBase + off + k]];
                        for (int a = 0; a < 3; a++) {
                            const double cell = std::ldexp(1.0, (int)n.e[a] - 127);
                            const double lo = (double)n.p[a] + n.qlo[a][s] * cell, hi = (double)n.p[a] + n.qhi[a][s] * cell;
                            if (lo > ib.lo[a] || hi < ib.hi[a]) { printf("FAIL box containment\n"); return 1; }
                        }
                    }
                }
            }
        }
        if (innerRefs + 1 != b8.nodes.size()) { printf("FAIL node count %zu vs refs %zu\n", b8.nodes.size(), innerRefs); return 1; }
        // --- traversal == brute force
        uint64_t visited = 0; uint32_t maxSp = 0; int nHit = 0;
        for (int r = 0; r < nRays; r++) {
            V3 o{U(rng) * 2.5f, U(rng) * 2.5f, U(rng) * 2.5f}, d{U(rng), U(rng), U(rng)};
            if (r % 5 == 0) d.x = 0.0f;              // axis-parallel components
            if (r % 7 == 0) d.y = -0.0f;
            if (r % 11 == 0) { d.x = 0.0f; d.y = 0.0f; d.z = (r & 1) ? 1.0f : -1.0f; }
            Hit ref{1e30f, -1};
            for (int i = 0; i < nTri; i++) {
                float t;
                if (triTest(tris[i], o, d, 0.0f, 1e30f, t) && (t < ref.t || (t == ref.t && i < ref.tri))) ref = Hit{t, i};
            }
            const Hit got = traverse8(b8, tris, o, d, visited, maxSp);
            if (got.tri != ref.tri || (ref.tri >= 0 && got.t != ref.t)) { printf("FAIL ray %d: tri %d t %g, brute force tri %d t %g\n", r, got.tri, got.t, ref.tri, ref.t); return 1; }
            nHit += ref.tri >= 0;
        }
        if (maxSp > b8.depth) { printf("FAIL stack %u deeper than depth %u\n", maxSp, b8.depth); return 1; }
        printf("maxLeaf %d: %zu bvh2 nodes -> %zu wide nodes, depth %u, %.1f nodes/ray, max stack %u, %d/%d rays hit\n", maxLeaf, b2.nodes.size(), b8.nodes.size(),
               b8.depth, (double)visited / nRays, maxSp, nHit, nRays);
    }
    printf("OK\n");
    return 0;
}
