// Host-side BVH construction (replaces the driver-opaque vkCmdBuildAccelerationStructuresKHR calls of
// engine/hrtsystem/Accel.zig:94-184 (BLAS) and :440-489 (TLAS)). Binned-SAH BVH2, emitted directly in
// the 64-byte device node layout of dtypes.h.
#pragma once
#include <cstdint>
#include <vector>
#include "dtypes.h"

namespace msne {

struct BuildBox { float lo[3], hi[3]; };

struct BuiltBvh {
    std::vector<BvhNode2> nodes;   // nodes[0] is the root; child indices are local to this array
    std::vector<uint32_t> order;   // item index for each leaf slot
    BuildBox bounds;               // padded root bounds
    uint32_t depth = 0;            // inner-node levels on the longest root-to-leaf walk (1 = root only)
};

// `maxLeaf` items per leaf at most. Leaf children reference [first, first+count) in `order`.
// Ranges deeper than `maxSahDepth` levels are halved by index instead of split by SAH, which bounds the depth by
// maxSahDepth + ceil(log2 n): the traversal stack (MSNE_STACK entries, one per level) is sized against it (runtime.cpp).
constexpr uint32_t kDefaultMaxSahDepth = 40;
void buildBvh2(const BuildBox* boxes, uint32_t n, uint32_t maxLeaf, BuiltBvh& out, uint32_t maxSahDepth = kDefaultMaxSahDepth);

// Renumbers the nodes so that the first min(k, n) are the top of the tree in breadth-first order (the rest keep their
// relative order). Pure relabelling: boxes, leaves and `order` are untouched.
void reorderTopBfs(BuiltBvh& bvh, uint32_t k);

// outward padding that keeps the FP32 slab test conservative w.r.t. the exact triangle test
inline void padBox(BuildBox& b) {
    float ext = 1e-3f, mag = 0.0f;
    for (int a = 0; a < 3; a++) {
        float e = b.hi[a] - b.lo[a];
        if (e > ext) ext = e;
        float m0 = b.lo[a] < 0 ? -b.lo[a] : b.lo[a], m1 = b.hi[a] < 0 ? -b.hi[a] : b.hi[a];
        if (m0 > mag) mag = m0;
        if (m1 > mag) mag = m1;
    }
    float pad = 1e-5f * ext + 1e-6f * mag;
    for (int a = 0; a < 3; a++) { b.lo[a] -= pad; b.hi[a] += pad; }
}

}  // namespace msne
