// bvh_build.h — host-side builder of the 8-wide quantised BVH ("compressed wide BVH", Ylitie,
// Karras, Laine 2017) that replaces parry3d's per-mesh Qbvh (reference src/shape/mesh.rs:171,
// TriMesh::new) and the bvh-crate item BVH (reference src/scene.rs:1681-1687) on the GPU.
//
// Pipeline: binned-SAH BVH2 (leaves <= 3 primitives) -> greedy collapse to <= 8 children ->
// octant-ordered child slots -> 80-byte nodes (5 x 16 B) with 8-bit quantised child boxes.
// The structure only prunes: every accepted hit is decided by the exact primitive test.
#pragma once
#include <cstdint>
#include <vector>

namespace rtx {

struct Aabb3 { float lo[3], hi[3]; };

// 80-byte node, read on the device as five 128-bit loads.
struct alignas(16) WideNode {
    float p[3];            // quantisation origin
    uint8_t e[3];          // per-axis exponent: child box = p + q * 2^(e-127)
    uint8_t imask;         // bit s set: slot s holds an internal node
    uint32_t child_base;   // index of the first internal child (children are contiguous, slot order)
    uint32_t prim_base;    // index of the first primitive referenced by this node's leaf slots
    uint8_t meta[8];       // internal: 0b001_11sss (sss = slot); leaf: unary count << 5 | offset; empty: 0
    uint8_t qlo[3][8];     // [axis][slot]
    uint8_t qhi[3][8];
};
static_assert(sizeof(WideNode) == 80, "node must be 80 bytes");

struct WideBvh {
    std::vector<WideNode> nodes;       // nodes[0] is the root
    std::vector<uint32_t> prim_order;  // position -> original primitive index (leaf order)
    int max_depth = 0;
};

// Build over `n` primitive boxes.  Leaves hold at most 3 primitives.  depth_limit > 0: a tree deeper than that is rebuilt
// with object-median splits and a size-balanced collapse (depth ~ log8 n).
void build_wide_bvh(const Aabb3* boxes, uint32_t n, WideBvh& out, int depth_limit = 0);

}  // namespace rtx
