// lbvh_build.cuh — wide-BVH construction ON THE DEVICE (north_star: "a binned-SAH BVH built on the host (or an LBVH builder
// kernel)").  Scene::update rebuilds its acceleration structures on every start (reference src/scene.rs:1674-1688)
// and parry builds a Qbvh per TriMesh (src/shape/mesh.rs:171); here a 10 M-triangle mesh is built in tens of milliseconds:
//
//   boxes (n x 6 floats, device)  ->  63-bit Morton codes of the box centres  ->  radix sort (CUB, cold path)
//     ->  binary radix tree, one thread per internal node (Karras 2012)  ->  bottom-up box fit (atomic arrival flags)
//     ->  level-by-level collapse into the SAME 80-byte 8-wide quantised nodes the host builder emits (bvh_build.cpp):
//         greedy largest-area opening to <= 8 children, subtrees of <= 3 primitives become leaf slots, octant-ordered slots,
//         outward-rounded 8-bit planes.
//
// The tree only prunes; accepted hits are decided by the exact primitive tests, so traversal results are identical to those of
// the host-built SAH tree (tests/test_gpu_round2.py).  A Morton tree costs more node visits per ray than the binned-SAH tree,
// which is why the host builder stays the default for static scenes and this one is opt-in (RTX_SCENE_DEVICE_BVH).
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bvh_build.h"

namespace rtx {
namespace lbvh {

struct Box { float lo[3], hi[3]; };

// order-preserving float <-> uint for atomicMin / atomicMax
__device__ __forceinline__ uint32_t f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void bounds_kernel(const Box* __restrict__ boxes, uint32_t n, uint32_t* __restrict__ bounds /* 6: lo xyz, hi xyz (ordered uint) */) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int k = 0; k < 3; k++) {
            const float c = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
            lo[k] = fminf(lo[k], c); hi[k] = fmaxf(hi[k], c);
        }
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(bounds + k, f2ord(lo[k])); atomicMax(bounds + 3 + k, f2ord(hi[k])); }
    }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) {       // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void morton_kernel(const Box* __restrict__ boxes, uint32_t n, const uint32_t* __restrict__ bounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t code = 0;
    for (int k = 0; k < 3; k++) {
        const float lo = ord2f(bounds[k]), hi = ord2f(bounds[3 + k]);
        const float ext = hi - lo;
        const float c = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
        float u = ext > 0.0f ? (c - lo) / ext : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        const uint64_t q = (uint64_t)fminf(u * 2097152.0f, 2097151.0f);
        code |= spread21(q) << (2 - k);
    }
    keys[i] = code; vals[i] = i;
}

// ---- binary radix tree (Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees", HPG 2012) ----
// internal nodes 0 .. n-2, leaves are sorted positions; child references: >= 0 internal index, < 0: ~sorted position
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);       // duplicate codes: fall back to the index
    return __clzll((long long)(a ^ b));
}

__global__ void radix_tree_kernel(const uint64_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right, int* __restrict__ parent /* n-1 internal + n leaves */,
                                  uint2* __restrict__ range /* first, count per internal node */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int lc = (first == gamma) ? ~gamma : gamma;
    const int rc = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc; right[i] = rc;
    range[i] = make_uint2((uint32_t)first, (uint32_t)(last - first + 1));
    if (lc >= 0) parent[lc] = i; else parent[(n - 1) + gamma] = i;
    if (rc >= 0) parent[rc] = i; else parent[(n - 1) + gamma + 1] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void fit_kernel(const Box* __restrict__ boxes, const uint32_t* __restrict__ vals, int n, const int* __restrict__ left, const int* __restrict__ right,
                           const int* __restrict__ parent, Box* __restrict__ node_box, uint32_t* __restrict__ arrived) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int p = parent[(n - 1) + leaf];
    while (p >= 0) {
        if (atomicAdd(arrived + p, 1u) == 0u) return;                // the first child to arrive stops; the second one has both boxes
        __threadfence();
        Box b;
        for (int k = 0; k < 3; k++) { b.lo[k] = INFINITY; b.hi[k] = -INFINITY; }
        const int ch[2] = {left[p], right[p]};
        for (int c = 0; c < 2; c++) {
            // a child box written by another thread is read from L2 (it was published with a fence before that thread's arrival)
            const float* src = ch[c] >= 0 ? node_box[ch[c]].lo : boxes[vals[~ch[c]]].lo;
            for (int k = 0; k < 3; k++) { b.lo[k] = fminf(b.lo[k], __ldcg(src + k)); b.hi[k] = fmaxf(b.hi[k], __ldcg(src + 3 + k)); }
        }
        node_box[p] = b;
        __threadfence();
        p = parent[p];
    }
}

// ---- collapse into 8-wide quantised nodes, one tree level per launch -------------------------------------------------
struct Task { uint32_t wide; int node2; uint32_t depth; };       // node2: >= 0 internal node, < 0: ~sorted position (single primitive)

__device__ __forceinline__ float box_half_area(const Box& b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

__global__ void collapse_kernel(const Task* __restrict__ tasks, uint32_t n_tasks, const int* __restrict__ left, const int* __restrict__ right, const uint2* __restrict__ range,
                                const Box* __restrict__ node_box, const Box* __restrict__ boxes, const uint32_t* __restrict__ vals,
                                WideNode* __restrict__ out_nodes, uint32_t node_cap, uint32_t* __restrict__ prim_order, uint32_t* __restrict__ counters /* [0] nodes, [1] prims, [2] next tasks, [3] max depth, [4] error */,
                                Task* __restrict__ next_tasks) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    const Task task = tasks[t];
    atomicMax(counters + 3, task.depth);
    auto cnt_of = [&](int r) -> uint32_t { return r >= 0 ? range[r].y : 1u; };
    auto first_of = [&](int r) -> uint32_t { return r >= 0 ? range[r].x : (uint32_t)(~r); };
    auto box_of = [&](int r) -> Box { return r >= 0 ? node_box[r] : boxes[vals[~r]]; };
    auto is_leaf = [&](int r) -> bool { return cnt_of(r) <= 3u; };
    int ch[8]; int nch = 0;
    if (is_leaf(task.node2)) ch[nch++] = task.node2;                    // degenerate: the whole tree is one leaf
    else { ch[nch++] = left[task.node2]; ch[nch++] = right[task.node2]; }
    while (nch < 8) {                                                   // greedy: open the largest internal child
        int best = -1; float ba = -1.0f;
        for (int i = 0; i < nch; i++) if (!is_leaf(ch[i])) { const float a = box_half_area(box_of(ch[i])); if (a > ba) { ba = a; best = i; } }
        if (best < 0) break;
        const int c = ch[best];
        ch[best] = left[c]; ch[nch++] = right[c];
    }
    const Box box = box_of(task.node2);
    float cx[3]; for (int k = 0; k < 3; k++) cx[k] = 0.5f * (box.lo[k] + box.hi[k]);
    // octant-ordered slot assignment (greedy on cost[child][slot] = dot(centroid offset, slot direction)), as on the host
    float cost[8][8]; int slot_of[8]; bool slot_used[8], done[8];
    for (int s = 0; s < 8; s++) { slot_used[s] = false; done[s] = false; slot_of[s] = 0; }
    for (int c = 0; c < nch; c++) {
        const Box cb = box_of(ch[c]);
        float d[3]; for (int k = 0; k < 3; k++) d[k] = 0.5f * (cb.lo[k] + cb.hi[k]) - cx[k];
        for (int s = 0; s < 8; s++) cost[c][s] = d[0] * ((s & 4) ? -1.f : 1.f) + d[1] * ((s & 2) ? -1.f : 1.f) + d[2] * ((s & 1) ? -1.f : 1.f);
    }
    for (int it = 0; it < nch; it++) {
        float bc = INFINITY; int bi = -1, bs = -1;
        for (int c = 0; c < nch; c++) if (!done[c]) for (int s = 0; s < 8; s++) if (!slot_used[s] && cost[c][s] < bc) { bc = cost[c][s]; bi = c; bs = s; }
        if (bi < 0) { for (int c = 0; c < nch && bi < 0; c++) if (!done[c]) for (int s = 0; s < 8; s++) if (!slot_used[s]) { bi = c; bs = s; break; } }   // NaN costs
        done[bi] = true; slot_used[bs] = true; slot_of[bi] = bs;
    }
    int child_in_slot[8]; for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
    for (int c = 0; c < nch; c++) child_in_slot[slot_of[c]] = c;

    WideNode w; memset(&w, 0, sizeof(w));
    float scale[3];
    for (int k = 0; k < 3; k++) {                                       // exact-grid quantisation (RTX_DEQUANT == 0 layout of bvh_build.cpp)
        w.p[k] = box.lo[k];
        const float ext = box.hi[k] - box.lo[k];
        int eb = 1;
        if (ext > 0.f) {
            int ex; frexpf(ext / 255.0f, &ex);
            eb = ex + 127;
            while (eb < 254 && w.p[k] + 255.0f * ldexpf(1.0f, eb - 127) < box.hi[k]) eb++;
            eb = max(1, min(254, eb));
        }
        w.e[k] = (uint8_t)eb; scale[k] = ldexpf(1.0f, eb - 127);
    }
    uint32_t n_internal = 0, n_prims = 0;
    for (int c = 0; c < nch; c++) { if (is_leaf(ch[c])) n_prims += cnt_of(ch[c]); else n_internal++; }
    const uint32_t child_base = n_internal ? atomicAdd(counters + 0, n_internal) : 0u;
    const uint32_t prim_base = n_prims ? atomicAdd(counters + 1, n_prims) : 0u;
    if (child_base + n_internal > node_cap) { atomicExch(counters + 4, 1u); return; }
    const uint32_t task_base = n_internal ? atomicAdd(counters + 2, n_internal) : 0u;
    w.child_base = child_base; w.prim_base = prim_base;
    uint32_t next_child = 0, prim_off = 0;
    for (int s = 0; s < 8; s++) {
        for (int k = 0; k < 3; k++) { w.qlo[k][s] = 255; w.qhi[k][s] = 0; }
        const int c = child_in_slot[s];
        if (c < 0) continue;
        const Box cb = box_of(ch[c]);
        for (int k = 0; k < 3; k++) {
            int lo = (int)floorf((cb.lo[k] - w.p[k]) / scale[k]);
            int hi = (int)ceilf((cb.hi[k] - w.p[k]) / scale[k]);
            lo = max(0, min(255, lo)); hi = max(0, min(255, hi));
            while (lo > 0 && w.p[k] + (float)lo * scale[k] > cb.lo[k]) lo--;
            while (hi < 255 && w.p[k] + (float)hi * scale[k] < cb.hi[k]) hi++;
            w.qlo[k][s] = (uint8_t)lo; w.qhi[k][s] = (uint8_t)hi;
        }
        if (!is_leaf(ch[c])) {
            w.imask |= (uint8_t)(1u << s);
            w.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            next_tasks[task_base + next_child] = Task{child_base + next_child, ch[c], task.depth + 1};
            next_child++;
        } else {
            const uint32_t cn = cnt_of(ch[c]), f0 = first_of(ch[c]);
            const uint32_t unary = cn == 1 ? 1u : cn == 2 ? 3u : 7u;
            w.meta[s] = (uint8_t)((unary << 5) | prim_off);
            for (uint32_t i = 0; i < cn; i++) prim_order[prim_base + prim_off + i] = vals[f0 + i];
            prim_off += cn;
        }
    }
    out_nodes[task.wide] = w;
}

// Triangle records (a|face, b|0, c|0 — 48 B) in leaf order, straight from the mesh arrays already on the device.
__global__ void pack_tris_kernel(const uint32_t* __restrict__ prim_order, uint32_t n, const float* __restrict__ verts, const uint32_t* __restrict__ idx,
                                 float4* __restrict__ tris) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t f = prim_order[k];
    const uint32_t ia = idx[3 * (size_t)f], ib = idx[3 * (size_t)f + 1], ic = idx[3 * (size_t)f + 2];
    tris[3 * (size_t)k] = make_float4(verts[3 * (size_t)ia], verts[3 * (size_t)ia + 1], verts[3 * (size_t)ia + 2], __uint_as_float(f));
    tris[3 * (size_t)k + 1] = make_float4(verts[3 * (size_t)ib], verts[3 * (size_t)ib + 1], verts[3 * (size_t)ib + 2], 0.f);
    tris[3 * (size_t)k + 2] = make_float4(verts[3 * (size_t)ic], verts[3 * (size_t)ic + 1], verts[3 * (size_t)ic + 2], 0.f);
}

// Scratch memory of the builder, grown on demand and reused from mesh to mesh (a scene of 64 meshes would otherwise pay 64 x 17
// cudaMalloc / cudaFree pairs, each a device synchronisation).
struct Workspace {
    size_t cap = 0, tmp_bytes = 0;
    uint64_t *keys = nullptr, *keys2 = nullptr; uint32_t *vals = nullptr, *vals2 = nullptr, *bounds = nullptr, *arrived = nullptr, *counters = nullptr, *prim_order = nullptr;
    int *left = nullptr, *right = nullptr, *parent = nullptr; uint2* range = nullptr; Box *node_box = nullptr, *boxes = nullptr; void* tmp = nullptr;
    WideNode* nodes = nullptr; struct TaskT { uint32_t wide; int node2; uint32_t depth; }; void* tq[2] = {nullptr, nullptr};
    void release() {
        cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(bounds); cudaFree(arrived); cudaFree(counters); cudaFree(prim_order);
        cudaFree(left); cudaFree(right); cudaFree(parent); cudaFree(range); cudaFree(node_box); cudaFree(boxes); cudaFree(nodes); cudaFree(tq[0]); cudaFree(tq[1]); cudaFree(tmp);
        *this = Workspace();
    }
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        cudaError_t e;
#define WS(p, bytes) if ((e = cudaMalloc((void**)&p, (bytes))) != cudaSuccess) { release(); return e; }
        WS(keys, n * 8) WS(keys2, n * 8) WS(vals, n * 4) WS(vals2, n * 4) WS(bounds, 6 * 4) WS(arrived, n * 4) WS(counters, 8 * 4) WS(prim_order, n * 4)
        WS(left, n * 4) WS(right, n * 4) WS(parent, 2 * n * 4) WS(range, n * 8) WS(node_box, n * sizeof(Box)) WS(boxes, n * sizeof(Box)) WS(nodes, (n + 8) * sizeof(WideNode))
        WS(tq[0], n * 12) WS(tq[1], n * 12)
        if ((e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, (int)n, 0, 63)) != cudaSuccess) { release(); return e; }
        WS(tmp, tmp_bytes)
#undef WS
        cap = n;
        return cudaSuccess;
    }
};

// Host driver.  d_boxes: n boxes on the current device.  Results are copied into `out` (host): the scene assembly of
// rtx_scene_create (node / primitive index offsets, triangle packing) is shared with the host builder.
// Returns cudaSuccess or the first CUDA error; *deep = 1 when the Morton tree does not fit `depth_limit` levels.
// keep_order_on_device: the leaf order stays in W.prim_order (the caller packs the triangles with a kernel) and out.prim_order is left empty.
inline cudaError_t build_on_device(Workspace& W, const Box* h_boxes, uint32_t n, WideBvh& out, int depth_limit, int* deep, float* device_ms,
                                   bool keep_order_on_device = false) {
    out.nodes.clear(); out.prim_order.clear(); out.max_depth = 0;
    if (deep) *deep = 0;
    if (n == 0) return cudaSuccess;
    cudaError_t e = W.reserve(n);
    if (e != cudaSuccess) return e;
    static_assert(sizeof(Task) == 12, "task size");
    Task* tq[2] = {(Task*)W.tq[0], (Task*)W.tq[1]};
    const uint32_t node_cap = n + 8;                                    // every wide node but the root is an internal binary node: <= n - 1
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define LB(x) do { e = (x); if (e != cudaSuccess) goto done; } while (0)
    LB(cudaMemcpy(W.boxes, h_boxes, (size_t)n * sizeof(Box), cudaMemcpyHostToDevice));
    cudaEventRecord(e0);
    {
        const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        LB(cudaMemcpy(W.bounds, init, sizeof(init), cudaMemcpyHostToDevice));
        bounds_kernel<<<148 * 4, 256>>>(W.boxes, n, W.bounds);
        morton_kernel<<<(n + 255) / 256, 256>>>(W.boxes, n, W.bounds, W.keys, W.vals);
        size_t tb = W.tmp_bytes;
        LB(cub::DeviceRadixSort::SortPairs(W.tmp, tb, W.keys, W.keys2, W.vals, W.vals2, (int)n, 0, 63));
        LB(cudaMemsetAsync(W.arrived, 0, (size_t)n * 4)); LB(cudaMemsetAsync(W.counters, 0, 8 * 4));
        if (n > 1) {
            radix_tree_kernel<<<(n - 1 + 255) / 256, 256>>>(W.keys2, (int)n, W.left, W.right, W.parent, W.range);
            fit_kernel<<<(n + 255) / 256, 256>>>(W.boxes, W.vals2, (int)n, W.left, W.right, W.parent, W.node_box, W.arrived);
        }
        // root task, then one launch per level of the wide tree
        const Task root{0u, n > 1 ? 0 : ~0, 1u};
        LB(cudaMemcpy(tq[0], &root, sizeof(root), cudaMemcpyHostToDevice));
        const uint32_t one = 1u;
        LB(cudaMemcpy(W.counters + 0, &one, 4, cudaMemcpyHostToDevice));  // node 0 = root
        uint32_t n_tasks = 1; int cur = 0;
        for (int level = 0; n_tasks > 0 && level < 4096; level++) {
            LB(cudaMemsetAsync(W.counters + 2, 0, 4));
            collapse_kernel<<<(n_tasks + 127) / 128, 128>>>(tq[cur], n_tasks, W.left, W.right, W.range, W.node_box, W.boxes, W.vals2, W.nodes, node_cap, W.prim_order, W.counters, tq[cur ^ 1]);
            LB(cudaMemcpy(&n_tasks, W.counters + 2, 4, cudaMemcpyDeviceToHost));
            cur ^= 1;
        }
    }
    cudaEventRecord(e1);
    LB(cudaMemcpy(h, W.counters, sizeof(h), cudaMemcpyDeviceToHost));
    if (h[4] || h[1] != n) { e = cudaErrorUnknown; goto done; }
    out.max_depth = (int)h[3];
    if (depth_limit > 0 && out.max_depth > depth_limit) { if (deep) *deep = 1; goto done; }
    out.nodes.resize(h[0]);
    LB(cudaMemcpy(out.nodes.data(), W.nodes, (size_t)h[0] * sizeof(WideNode), cudaMemcpyDeviceToHost));
    if (!keep_order_on_device) { out.prim_order.resize(n); LB(cudaMemcpy(out.prim_order.data(), W.prim_order, (size_t)n * 4, cudaMemcpyDeviceToHost)); }
    if (device_ms) cudaEventElapsedTime(device_ms, e0, e1);
done:
#undef LB
    if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1);
    return e;
}

}  // namespace lbvh
}  // namespace rtx
