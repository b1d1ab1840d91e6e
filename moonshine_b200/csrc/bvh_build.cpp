// Binned-SAH BVH2 builder (host). See bvh_build.h.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <thread>

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

namespace {

struct Range { uint32_t first, count; int32_t node; uint32_t depth; };  // node whose two children are the split of this range

struct BuildCtx {
    const BuildBox* boxes;
    const float* cent;
    uint32_t* order;
    uint32_t maxLeaf, maxSahDepth;
};

// Splits range `r` (binned SAH, median by index past maxSahDepth), fills the two child boxes of nodes[r.node] and either makes a
// child a leaf or appends a new inner node to `nodes` and pushes its range on `work`. Touches only order[r.first, r.first + r.count).
// runs fn(begin, end, slot) over [0, count) on up to `maxThreads` threads (slot < maxThreads identifies the caller's scratch)
template <class F>
void parallelChunks(uint32_t count, unsigned maxThreads, F fn) {
    const unsigned nT = (unsigned)std::min<uint64_t>(maxThreads, (count + 65535u) / 65536u);
    if (nT <= 1) { fn(0u, count, 0u); return; }
    std::vector<std::thread> pool;
    const uint32_t per = (count + nT - 1) / nT;
    for (unsigned t = 0; t < nT; t++) {
        const uint32_t b0 = std::min(count, t * per), e0 = std::min(count, b0 + per);
        pool.emplace_back([=] { fn(b0, e0, t); });
    }
    for (auto& th : pool) th.join();
}

struct Bins { Bounds box[3][kBins]; uint32_t n[3][kBins]; Bounds cent; };

// Splits range `r` (binned SAH over the three axes in one pass, median by index past maxSahDepth), fills the two child boxes of
// nodes[r.node] and either makes a child a leaf or appends a new inner node to `nodes` and pushes its range on `work`. Touches only
// order[r.first, r.first + r.count). `threads` > 1: the passes over the items of a big range run on that many threads (used for the
// top of the tree); the result is the same for any thread count (min / max / counts are order-independent, the partition is stable
// per chunk and chunks are concatenated in order... except that item order INSIDE the two halves may differ, which never changes
// a split: bins and bounds depend on the set, not the order).
struct Split { uint32_t mid; Bounds box[2]; };
Split splitRange(const BuildCtx& c, const Range& r, unsigned threads = 1, std::vector<uint32_t>* scratch = nullptr) {
    uint32_t* ord = c.order + r.first;
    // centroid bounds
    constexpr unsigned kMaxT = 16;
    threads = std::min(threads, kMaxT);
    Bounds cbT[kMaxT];
    parallelChunks(r.count, threads, [&](uint32_t b0, uint32_t e0, unsigned t) {
        Bounds cb; for (uint32_t i = b0; i < e0; i++) cb.growPoint(&c.cent[(size_t)ord[i] * 3]);
        cbT[t] = cb;
    });
    Bounds cb;
    for (unsigned t = 0; t < kMaxT; t++) if (cbT[t].lo[0] <= cbT[t].hi[0]) cb.grow(cbT[t].lo, cbT[t].hi);
    int bestAxis = -1; int bestBin = -1; float bestCost = 1e38f;
    Bounds bestL, bestR;
    if (r.depth < c.maxSahDepth) {
        float scale[3]; bool use[3];
        for (int a = 0; a < 3; a++) { const float ext = cb.hi[a] - cb.lo[a]; use[a] = ext > 0.0f; scale[a] = use[a] ? (float)kBins / ext : 0.0f; }
        Bins oneBins;                       // ranges handled by one thread (nearly all of them) stay off the heap
        std::vector<Bins> manyBins;
        Bins* binsT = &oneBins;
        const unsigned nBins = std::max(1u, threads);
        if (nBins > 1) { manyBins.resize(nBins); binsT = manyBins.data(); }
        for (unsigned t = 0; t < nBins; t++) std::memset(binsT[t].n, 0, sizeof(binsT[t].n));
        parallelChunks(r.count, threads, [&](uint32_t b0, uint32_t e0, unsigned t) {
            Bins& bt = binsT[t];
            for (uint32_t i = b0; i < e0; i++) {
                const uint32_t it = ord[i];
                for (int a = 0; a < 3; a++) {
                    if (!use[a]) continue;
                    int bi = (int)((c.cent[(size_t)it * 3 + a] - cb.lo[a]) * scale[a]);
                    bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
                    bt.n[a][bi]++; bt.box[a][bi].grow(c.boxes[it].lo, c.boxes[it].hi);
                }
            }
        });
        for (int a = 0; a < 3; a++) {
            if (!use[a]) continue;
            Bounds binB[kBins]; uint32_t binN[kBins] = {0};
            for (unsigned t = 0; t < nBins; t++) {
                const Bins& bt = binsT[t];
                for (int bb = 0; bb < kBins; bb++) if (bt.n[a][bb]) { binN[bb] += bt.n[a][bb]; binB[bb].grow(bt.box[a][bb].lo, bt.box[a][bb].hi); }
            }
            Bounds rightB[kBins]; float rightArea[kBins]; uint32_t rightN[kBins];
            Bounds acc; uint32_t cnt = 0;
            for (int bb = kBins - 1; bb > 0; bb--) {
                if (binN[bb]) acc.grow(binB[bb].lo, binB[bb].hi);
                cnt += binN[bb];
                rightB[bb] = acc; rightArea[bb] = acc.halfArea(); rightN[bb] = cnt;
            }
            Bounds accL; uint32_t cntL = 0;
            for (int bb = 0; bb < kBins - 1; bb++) {
                if (binN[bb]) accL.grow(binB[bb].lo, binB[bb].hi);
                cntL += binN[bb];
                if (cntL == 0 || rightN[bb + 1] == 0) continue;
                float cost = accL.halfArea() * (float)cntL + rightArea[bb + 1] * (float)rightN[bb + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestBin = bb; bestL = accL; bestR = rightB[bb + 1]; }
            }
        }
    }
    uint32_t mid;
    bool boundsKnown = false;
    if (bestAxis >= 0) {
        const float scale = (float)kBins / (cb.hi[bestAxis] - cb.lo[bestAxis]);
        const float lo = cb.lo[bestAxis];
        const int a = bestAxis, bb = bestBin;
        auto left = [&](uint32_t it) {
            int bi = (int)((c.cent[(size_t)it * 3 + a] - lo) * scale);
            bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
            return bi <= bb;
        };
        if (threads > 1 && scratch != nullptr && r.count >= (1u << 18)) {
            // parallel partition: every chunk counts its left items, then scatters into `scratch` at the chunk's offsets
            scratch->resize(r.count);
            uint32_t leftN[kMaxT + 1] = {0}, rightN[kMaxT + 1] = {0}, bounds[kMaxT + 1] = {0};
            unsigned used = 0;
            parallelChunks(r.count, threads, [&](uint32_t b0, uint32_t e0, unsigned t) {
                uint32_t l = 0; for (uint32_t i = b0; i < e0; i++) l += left(ord[i]) ? 1u : 0u;
                leftN[t] = l; rightN[t] = (e0 - b0) - l; bounds[t] = b0;
            });
            for (unsigned t = 0; t < kMaxT; t++) if (leftN[t] + rightN[t]) used = t + 1;
            uint32_t lOff[kMaxT], rOff[kMaxT], totalL = 0;
            for (unsigned t = 0; t < used; t++) { lOff[t] = totalL; totalL += leftN[t]; }
            uint32_t run = totalL;
            for (unsigned t = 0; t < used; t++) { rOff[t] = run; run += rightN[t]; }
            uint32_t* tmp = scratch->data();
            parallelChunks(r.count, threads, [&](uint32_t b0, uint32_t e0, unsigned t) {
                uint32_t l = lOff[t], rr = rOff[t];
                for (uint32_t i = b0; i < e0; i++) { const uint32_t it = ord[i]; if (left(it)) tmp[l++] = it; else tmp[rr++] = it; }
            });
            std::memcpy(ord, tmp, (size_t)r.count * sizeof(uint32_t));
            mid = totalL;
        } else {
            uint32_t* m = std::partition(ord, ord + r.count, left);
            mid = (uint32_t)(m - ord);
        }
        boundsKnown = mid != 0 && mid != r.count;   // the bins hold exactly the two halves' boxes
    } else {
        mid = r.count / 2;  // all centroids coincide (or past the SAH depth cap): split by index
    }
    if (mid == 0 || mid == r.count) mid = r.count / 2;
    Split out;
    out.mid = mid;
    const uint32_t cnt[2] = {mid, r.count - mid};
    const uint32_t fst[2] = {r.first, r.first + mid};
    for (int s = 0; s < 2; s++) {
        Bounds bnd;
        if (boundsKnown) bnd = s == 0 ? bestL : bestR;
        else for (uint32_t i = 0; i < cnt[s]; i++) bnd.grow(c.boxes[c.order[fst[s] + i]].lo, c.boxes[c.order[fst[s] + i]].hi);
        out.box[s] = bnd;
    }
    return out;
}

}  // namespace

// Top-down build on all host cores. Nodes come from one pre-sized array through an atomic cursor; ranges are tasks in a shared
// queue (a worker keeps small children on its own stack and shares big ones); the few ranges at the top that are too big to be
// one task are split on the calling thread with the passes over their items spread over the workers. Which thread builds what is
// timing dependent, so the node array is renumbered depth first at the end: the result (tree AND layout) does not depend on the
// number of threads or on scheduling.
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
    BvhNode2 blank;
    std::memset(&blank, 0, sizeof(blank));
    blank.c[0] = blank.c[1] = -1;
    if (n <= maxLeaf) {
        out.nodes.push_back(blank);
        if (n == 0) return;
        Bounds b;
        for (uint32_t i = 0; i < n; i++) b.grow(boxes[i].lo, boxes[i].hi);
        setChildBox(out.nodes[0], 0, b);
        out.nodes[0].c[0] = 0; out.nodes[0].c[2] = (int32_t)n;
        return;
    }
    const BuildCtx ctx{boxes, cent.data(), out.order.data(), maxLeaf, kMaxSahDepth};
    std::vector<BvhNode2> nodes((size_t)n, blank);       // every inner node has >= 2 items below it: fewer than n nodes
    std::atomic<uint32_t> nextNode{1};
    std::atomic<uint32_t> maxDepth{1};
    // wires the two halves of a split range into nodes[r.node]; inner children get a fresh node and are handed to `emit`
    auto wire = [&](const Range& r, const Split& sp, auto&& emit) {
        const uint32_t cnt[2] = {sp.mid, r.count - sp.mid}, fst[2] = {r.first, r.first + sp.mid};
        for (int s = 0; s < 2; s++) {
            setChildBox(nodes[(size_t)r.node], s, sp.box[s]);
            if (cnt[s] <= maxLeaf) { nodes[(size_t)r.node].c[s] = (int32_t)fst[s]; nodes[(size_t)r.node].c[2 + s] = (int32_t)cnt[s]; }
            else {
                const int32_t idx = (int32_t)nextNode.fetch_add(1);
                nodes[(size_t)r.node].c[s] = idx; nodes[(size_t)r.node].c[2 + s] = 0;
                emit(Range{fst[s], cnt[s], idx, r.depth + 1});
            }
        }
    };
    auto noteDepth = [&](uint32_t d) { uint32_t cur = maxDepth.load(std::memory_order_relaxed); while (d > cur && !maxDepth.compare_exchange_weak(cur, d)) {} };

    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    constexpr uint32_t kBigRange = 1u << 18;             // ranges at least this big: passes over the items run on all threads
    constexpr uint32_t kShare = 1u << 13;                // children at least this big go to the shared queue
    std::vector<Range> queue;
    {   // ---- top of the tree on the calling thread -------------------------------------------------------------------------------
        std::vector<Range> big{{0, n, 0, 0}};
        std::vector<uint32_t> scratch;
        while (!big.empty()) {
            const Range r = big.back(); big.pop_back();
            noteDepth(r.depth + 1);
            const Split sp = splitRange(ctx, r, hw, &scratch);
            wire(r, sp, [&](const Range& k) { (hw > 1 && k.count >= kBigRange ? big : queue).push_back(k); });
        }
    }
    // ---- everything below: a shared queue of ranges -----------------------------------------------------------------------------
    std::mutex mtx;
    std::atomic<int> busy{0};
    auto worker = [&] {
        std::vector<Range> local;
        uint32_t deepest = 1;
        for (;;) {
            Range r;
            {
                std::lock_guard<std::mutex> g(mtx);
                if (queue.empty()) { if (busy.load() == 0) break; else { } }
                if (!queue.empty()) { r = queue.back(); queue.pop_back(); busy.fetch_add(1); local.push_back(r); }
            }
            if (local.empty()) { std::this_thread::yield(); continue; }
            while (!local.empty()) {
                const Range cur = local.back(); local.pop_back();
                deepest = std::max(deepest, cur.depth + 1);
                const Split sp = splitRange(ctx, cur);
                wire(cur, sp, [&](const Range& k) {
                    if (hw > 1 && k.count >= kShare) { std::lock_guard<std::mutex> g(mtx); queue.push_back(k); }
                    else local.push_back(k);
                });
            }
            busy.fetch_sub(1);
        }
        noteDepth(deepest);
    };
    if (hw <= 1 || n < (1u << 15)) worker();
    else {
        std::vector<std::thread> pool;
        for (unsigned w = 1; w < hw; w++) pool.emplace_back(worker);
        worker();
        for (auto& th : pool) th.join();
    }
    // ---- depth-first renumbering (deterministic layout; a node's first child follows it) --------------------------------------------
    const uint32_t used = nextNode.load();
    out.nodes.resize(used);
    std::vector<uint32_t> newOf(used, 0), stack{0};
    uint32_t counter = 0;
    std::vector<uint32_t> orderDfs; orderDfs.reserve(used);
    while (!stack.empty()) {
        const uint32_t k = stack.back(); stack.pop_back();
        newOf[k] = counter++; orderDfs.push_back(k);
        const BvhNode2& nd = nodes[k];
        for (int s = 1; s >= 0; s--) if (nd.c[2 + s] == 0 && nd.c[s] >= 0) stack.push_back((uint32_t)nd.c[s]);
    }
    for (uint32_t i = 0; i < used; i++) {
        BvhNode2 nd = nodes[orderDfs[i]];
        for (int s = 0; s < 2; s++) if (nd.c[2 + s] == 0 && nd.c[s] >= 0) nd.c[s] = (int32_t)newOf[(uint32_t)nd.c[s]];
        out.nodes[i] = nd;
    }
    out.depth = maxDepth.load();
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
