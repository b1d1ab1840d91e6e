// Binned-SAH BVH2 builder (host). See bvh_build.h.
#include "bvh_build.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace msne {

namespace {

struct Bounds {
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    void grow(const float* l, const float* h) {
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); }
    }
    void growPoint(const float* p) { grow(p, p); }
    float halfArea() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
        return dx * dy + dy * dz + dz * dx;
    }
};

constexpr int kBins = 16;

void setChildBox(BvhNode2& n, int slot, const Bounds& bIn) {
    BuildBox b;
    for (int a = 0; a < 3; a++) { b.lo[a] = bIn.lo[a]; b.hi[a] = bIn.hi[a]; }
    padBox(b);
    float* xy = slot == 0 ? n.n0 : n.n1;
    xy[0] = b.lo[0]; xy[1] = b.hi[0]; xy[2] = b.lo[1]; xy[3] = b.hi[1];
    n.n2[slot * 2 + 0] = b.lo[2]; n.n2[slot * 2 + 1] = b.hi[2];
}

}  // namespace

void buildBvh2(const BuildBox* boxes, uint32_t n, uint32_t maxLeaf, BuiltBvh& out, uint32_t kMaxSahDepth) {
    out.nodes.clear();
    out.depth = 1;
    out.order.resize(n);
    for (uint32_t i = 0; i < n; i++) out.order[i] = i;
    std::vector<float> cent((size_t)n * 3);
    Bounds all;
    for (uint32_t i = 0; i < n; i++) {
        all.grow(boxes[i].lo, boxes[i].hi);
        for (int a = 0; a < 3; a++) cent[(size_t)i * 3 + a] = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);
    }
    for (int a = 0; a < 3; a++) { out.bounds.lo[a] = n ? all.lo[a] : 0.0f; out.bounds.hi[a] = n ? all.hi[a] : 0.0f; }
    padBox(out.bounds);

    // root is always an inner node; an empty child has c = -1, count = 0
    BvhNode2 root;
    std::memset(&root, 0, sizeof(root));
    root.c[0] = root.c[1] = -1;
    out.nodes.push_back(root);
    if (n == 0) return;

    struct Range { uint32_t first, count; int32_t node; uint32_t depth; };  // node whose two children are the split of this range
    std::vector<Range> work;
    if (n <= maxLeaf) {
        Bounds b;
        for (uint32_t i = 0; i < n; i++) b.grow(boxes[i].lo, boxes[i].hi);
        setChildBox(out.nodes[0], 0, b);
        out.nodes[0].c[0] = 0; out.nodes[0].c[2] = (int32_t)n;
        return;
    }
    work.push_back({0, n, 0, 0});
    while (!work.empty()) {
        Range r = work.back(); work.pop_back();
        out.depth = std::max(out.depth, r.depth + 1);
        uint32_t* ord = out.order.data() + r.first;
        // centroid bounds
        Bounds cb;
        for (uint32_t i = 0; i < r.count; i++) cb.growPoint(&cent[(size_t)ord[i] * 3]);
        int bestAxis = -1; int bestBin = -1; float bestCost = 1e38f;
        for (int a = 0; a < 3 && r.depth < kMaxSahDepth; a++) {
            float ext = cb.hi[a] - cb.lo[a];
            if (!(ext > 0.0f)) continue;
            Bounds binB[kBins]; uint32_t binN[kBins] = {0};
            float scale = (float)kBins / ext;
            for (uint32_t i = 0; i < r.count; i++) {
                uint32_t it = ord[i];
                int bi = (int)((cent[(size_t)it * 3 + a] - cb.lo[a]) * scale);
                bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
                binN[bi]++; binB[bi].grow(boxes[it].lo, boxes[it].hi);
            }
            float rightArea[kBins]; uint32_t rightN[kBins];
            Bounds acc; uint32_t cnt = 0;
            for (int b = kBins - 1; b > 0; b--) {
                if (binN[b]) acc.grow(binB[b].lo, binB[b].hi);
                cnt += binN[b];
                rightArea[b] = acc.halfArea(); rightN[b] = cnt;
            }
            Bounds accL; uint32_t cntL = 0;
            for (int b = 0; b < kBins - 1; b++) {
                if (binN[b]) accL.grow(binB[b].lo, binB[b].hi);
                cntL += binN[b];
                if (cntL == 0 || rightN[b + 1] == 0) continue;
                float cost = accL.halfArea() * (float)cntL + rightArea[b + 1] * (float)rightN[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestBin = b; }
            }
        }
        uint32_t mid;
        if (bestAxis >= 0) {
            float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
            float scale = (float)kBins / ext;
            float lo = cb.lo[bestAxis];
            int a = bestAxis; int bb = bestBin;
            uint32_t* m = std::partition(ord, ord + r.count, [&](uint32_t it) {
                int bi = (int)((cent[(size_t)it * 3 + a] - lo) * scale);
                bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
                return bi <= bb;
            });
            mid = (uint32_t)(m - ord);
        } else {
            mid = r.count / 2;  // all centroids coincide: split by index
        }
        if (mid == 0 || mid == r.count) mid = r.count / 2;
        const uint32_t cnt[2] = {mid, r.count - mid};
        const uint32_t fst[2] = {r.first, r.first + mid};
        for (int s = 0; s < 2; s++) {
            Bounds b;
            for (uint32_t i = 0; i < cnt[s]; i++) b.grow(boxes[out.order[fst[s] + i]].lo, boxes[out.order[fst[s] + i]].hi);
            setChildBox(out.nodes[r.node], s, b);
            if (cnt[s] <= maxLeaf) {
                out.nodes[r.node].c[s] = (int32_t)fst[s];
                out.nodes[r.node].c[2 + s] = (int32_t)cnt[s];
            } else {
                BvhNode2 child; std::memset(&child, 0, sizeof(child)); child.c[0] = child.c[1] = -1;
                int32_t idx = (int32_t)out.nodes.size();
                out.nodes.push_back(child);
                out.nodes[r.node].c[s] = idx;
                out.nodes[r.node].c[2 + s] = 0;
                work.push_back({fst[s], cnt[s], idx, r.depth + 1});
            }
        }
    }
}

void reorderTopBfs(BuiltBvh& bvh, uint32_t k) {
    const uint32_t n = (uint32_t)bvh.nodes.size();
    if (k == 0 || n <= 1) return;
    std::vector<uint32_t> newOf(n, 0xFFFFFFFFu), bfs;
    bfs.reserve(std::min(k, n));
    bfs.push_back(0); newOf[0] = 0;
    for (size_t head = 0; head < bfs.size() && bfs.size() < k; head++) {
        const BvhNode2& nd = bvh.nodes[bfs[head]];
        for (int s = 0; s < 2 && bfs.size() < k; s++)
            if (nd.c[2 + s] == 0 && nd.c[s] >= 0) { newOf[(uint32_t)nd.c[s]] = (uint32_t)bfs.size(); bfs.push_back((uint32_t)nd.c[s]); }
    }
    uint32_t next = (uint32_t)bfs.size();
    for (uint32_t i = 0; i < n; i++) if (newOf[i] == 0xFFFFFFFFu) newOf[i] = next++;
    std::vector<BvhNode2> out(n);
    for (uint32_t i = 0; i < n; i++) {
        BvhNode2 nd = bvh.nodes[i];
        for (int s = 0; s < 2; s++) if (nd.c[2 + s] == 0 && nd.c[s] >= 0) nd.c[s] = (int32_t)newOf[(uint32_t)nd.c[s]];
        out[newOf[i]] = nd;
    }
    bvh.nodes.swap(out);
}

}  // namespace msne
