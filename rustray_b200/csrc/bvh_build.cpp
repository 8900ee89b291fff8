// bvh_build.cpp — see bvh_build.h.  Host code, runs once per scene (cold path).
#include "bvh_build.h"

#ifndef RTX_DEQUANT
#define RTX_DEQUANT 0     // must match csrc/rtx_device.cuh (0: I2F dequantisation, exact grid)
#endif

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <limits>

namespace rtx {
namespace {

constexpr int kBins = 16;
constexpr uint32_t kMaxLeaf = 3;   // leaf slots hold 1..3 primitives (unary count in meta)

struct Node2 {
    Aabb3 box;
    int32_t left = -1, right = -1;   // children (internal) ...
    uint32_t first = 0, count = 0;   // ... or primitive range (leaf, count > 0)
    uint32_t total = 0;              // primitives below this node (balanced collapse)
};

inline void grow(Aabb3& a, const Aabb3& b) {
    for (int k = 0; k < 3; k++) { a.lo[k] = std::min(a.lo[k], b.lo[k]); a.hi[k] = std::max(a.hi[k], b.hi[k]); }
}
inline Aabb3 empty_box() {
    const float inf = std::numeric_limits<float>::infinity();
    return {{inf, inf, inf}, {-inf, -inf, -inf}};
}
inline float half_area(const Aabb3& b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (!(dx >= 0.f)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
}

struct Builder2 {
    const Aabb3* boxes;
    std::vector<uint32_t> idx;
    std::vector<float> cen;     // 3 per primitive
    std::vector<Node2> nodes;
    bool balanced = false;      // object-median splits: depth <= ceil(log2 n), used when the SAH tree is too deep for the traversal stack

    void build(uint32_t n) {
        idx.resize(n); cen.resize(3 * (size_t)n);
        for (uint32_t i = 0; i < n; i++) {
            idx[i] = i;
            for (int k = 0; k < 3; k++) cen[3 * (size_t)i + k] = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
        }
        nodes.reserve(2 * (size_t)n);
        nodes.emplace_back();
        struct Job { int32_t node; uint32_t first, count; };
        std::vector<Job> stack{{0, 0, n}};
        while (!stack.empty()) {
            Job j = stack.back(); stack.pop_back();
            Aabb3 box = empty_box(), cb = empty_box();
            for (uint32_t i = j.first; i < j.first + j.count; i++) {
                uint32_t p = idx[i];
                grow(box, boxes[p]);
                for (int k = 0; k < 3; k++) { cb.lo[k] = std::min(cb.lo[k], cen[3 * (size_t)p + k]); cb.hi[k] = std::max(cb.hi[k], cen[3 * (size_t)p + k]); }
            }
            nodes[j.node].box = box; nodes[j.node].total = j.count;
            if (j.count == 1) { nodes[j.node].first = j.first; nodes[j.node].count = 1; continue; }

            // binned SAH over the three axes
            float best_cost = std::numeric_limits<float>::infinity(); int best_axis = -1, best_split = -1;
            if (balanced) {
                if (j.count <= kMaxLeaf) { nodes[j.node].first = j.first; nodes[j.node].count = j.count; continue; }
                int ax = 0; for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[ax] - cb.lo[ax]) ax = k;
                const uint32_t mid = j.first + j.count / 2;
                std::nth_element(idx.begin() + j.first, idx.begin() + mid, idx.begin() + j.first + j.count,
                                 [&](uint32_t a, uint32_t b) { return cen[3 * (size_t)a + ax] < cen[3 * (size_t)b + ax]; });
                int32_t l = (int32_t)nodes.size();
                nodes.emplace_back(); nodes.emplace_back();
                nodes[j.node].left = l; nodes[j.node].right = l + 1;
                stack.push_back({l, j.first, mid - j.first});
                stack.push_back({l + 1, mid, j.first + j.count - mid});
                continue;
            }
            for (int ax = 0; ax < 3; ax++) {
                float ext = cb.hi[ax] - cb.lo[ax];
                if (!(ext > 0.f)) continue;
                Aabb3 bb[kBins]; uint32_t bc[kBins];
                for (int b = 0; b < kBins; b++) { bb[b] = empty_box(); bc[b] = 0; }
                float scale = (float)kBins / ext;
                for (uint32_t i = j.first; i < j.first + j.count; i++) {
                    uint32_t p = idx[i];
                    int b = std::min(kBins - 1, (int)((cen[3 * (size_t)p + ax] - cb.lo[ax]) * scale));
                    bc[b]++; grow(bb[b], boxes[p]);
                }
                float right_area[kBins]; uint32_t right_cnt[kBins];
                Aabb3 acc = empty_box(); uint32_t c = 0;
                for (int b = kBins - 1; b > 0; b--) { grow(acc, bb[b]); c += bc[b]; right_area[b] = half_area(acc); right_cnt[b] = c; }
                acc = empty_box(); c = 0;
                for (int b = 0; b < kBins - 1; b++) {
                    grow(acc, bb[b]); c += bc[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    float cost = half_area(acc) * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = b; }
                }
            }
            float leaf_cost = half_area(box) * (float)j.count;
            bool make_leaf = j.count <= kMaxLeaf && (best_axis < 0 || best_cost + 0.5f * half_area(box) >= leaf_cost);
            if (make_leaf) { nodes[j.node].first = j.first; nodes[j.node].count = j.count; continue; }

            uint32_t mid;
            if (best_axis >= 0) {
                float ext = cb.hi[best_axis] - cb.lo[best_axis];
                float scale = (float)kBins / ext;
                auto it = std::partition(idx.begin() + j.first, idx.begin() + j.first + j.count, [&](uint32_t p) {
                    int b = std::min(kBins - 1, (int)((cen[3 * (size_t)p + best_axis] - cb.lo[best_axis]) * scale));
                    return b <= best_split;
                });
                mid = (uint32_t)(it - idx.begin());
            } else {
                mid = j.first + j.count / 2;    // all centroids coincide: split the list in half
            }
            if (mid == j.first || mid == j.first + j.count) mid = j.first + j.count / 2;
            int32_t l = (int32_t)nodes.size();
            nodes.emplace_back(); nodes.emplace_back();
            nodes[j.node].left = l; nodes[j.node].right = l + 1;
            stack.push_back({l, j.first, mid - j.first});
            stack.push_back({l + 1, mid, j.first + j.count - mid});
        }
    }
};

struct WideBuilder {
    const Builder2& b2;
    WideBvh& out;

    // Fill wide node `wi` from BVH2 subtree `n2`; children are appended to `out.nodes`.
    void emit(uint32_t wi, int32_t n2_root, int depth) {
        out.max_depth = std::max(out.max_depth, depth);
        const std::vector<Node2>& N = b2.nodes;
        int32_t ch[8]; int nch = 0;
        if (N[n2_root].count > 0) ch[nch++] = n2_root;     // degenerate: the whole tree is one leaf
        else { ch[nch++] = N[n2_root].left; ch[nch++] = N[n2_root].right; }
        while (nch < 8) {                                   // greedy: open the largest internal child (by area; by size in a balanced tree)
            int best = -1; float ba = -1.f;
            for (int i = 0; i < nch; i++) if (N[ch[i]].count == 0) { float a = b2.balanced ? (float)N[ch[i]].total : half_area(N[ch[i]].box); if (a > ba) { ba = a; best = i; } }
            if (best < 0) break;
            int32_t c = ch[best];
            ch[best] = N[c].left; ch[nch++] = N[c].right;
        }
        const Aabb3& box = N[n2_root].box;
        float cx[3]; for (int k = 0; k < 3; k++) cx[k] = 0.5f * (box.lo[k] + box.hi[k]);

        // octant-ordered slot assignment (greedy on cost[child][slot] = dot(centroid offset, slot dir))
        float cost[8][8]; int slot_of[8]; bool slot_used[8] = {false}; bool done[8] = {false};
        for (int c = 0; c < nch; c++) {
            const Aabb3& cb = N[ch[c]].box;
            float d[3]; for (int k = 0; k < 3; k++) d[k] = 0.5f * (cb.lo[k] + cb.hi[k]) - cx[k];
            for (int s = 0; s < 8; s++)
                cost[c][s] = d[0] * ((s & 4) ? -1.f : 1.f) + d[1] * ((s & 2) ? -1.f : 1.f) + d[2] * ((s & 1) ? -1.f : 1.f);
        }
        for (int it = 0; it < nch; it++) {
            float bc = std::numeric_limits<float>::infinity(); int bi = -1, bs = -1;
            for (int c = 0; c < nch; c++) if (!done[c]) for (int s = 0; s < 8; s++) if (!slot_used[s] && cost[c][s] < bc) { bc = cost[c][s]; bi = c; bs = s; }
            done[bi] = true; slot_used[bs] = true; slot_of[bi] = bs;
        }
        int child_in_slot[8]; for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
        for (int c = 0; c < nch; c++) child_in_slot[slot_of[c]] = c;

        WideNode w; memset(&w, 0, sizeof(w));
        for (int k = 0; k < 3; k++) {
#if RTX_DEQUANT != 0
            // Quantisation grid with a one-step margin on both sides: child planes land in [1, 254] and are then
            // padded outward by one step to [0, 255].  The PRMT dequantisation (byte -> 2^23 + q, 2^23 folded into the
            // FMA addend) costs up to half a step of accuracy; the padding absorbs it.
            float ext = box.hi[k] - box.lo[k];
            int eb = 1;
            if (ext > 0.f) {
                int ex; std::frexp(ext / 253.0f, &ex);         // ext/253 = m * 2^ex, m in [0.5,1)  =>  2^ex * 253 >= ext
                eb = std::max(1, std::min(254, ex + 127));
            }
            for (;;) {
                const float scale = std::ldexp(1.0f, eb - 127);
                w.p[k] = box.lo[k] - scale;
                if (eb >= 254 || (w.p[k] + scale <= box.lo[k] && w.p[k] + 254.0f * scale >= box.hi[k])) break;
                eb++;
            }
#else
            w.p[k] = box.lo[k];
            float ext = box.hi[k] - box.lo[k];
            int eb = 1;
            if (ext > 0.f) {
                int ex; std::frexp(ext / 255.0f, &ex);         // ext/255 = m * 2^ex, m in [0.5,1)  =>  2^ex * 255 >= ext
                eb = ex + 127;
                while (eb < 254 && w.p[k] + 255.0f * std::ldexp(1.0f, eb - 127) < box.hi[k]) eb++;   // cover the extent in float arithmetic
                eb = std::max(1, std::min(254, eb));
            }
#endif
            w.e[k] = (uint8_t)eb;
        }
        uint32_t n_internal = 0;
        for (int s = 0; s < 8; s++) if (child_in_slot[s] >= 0 && N[ch[child_in_slot[s]]].count == 0) n_internal++;
        w.child_base = (uint32_t)out.nodes.size();
        w.prim_base = (uint32_t)out.prim_order.size();
        out.nodes.resize(out.nodes.size() + n_internal);
        uint32_t next_child = w.child_base, prim_off = 0;
        struct Pending { uint32_t wi; int32_t n2; };
        Pending pend[8]; int npend = 0;
        for (int s = 0; s < 8; s++) {
            for (int k = 0; k < 3; k++) { w.qlo[k][s] = 255; w.qhi[k][s] = 0; }
            int c = child_in_slot[s];
            if (c < 0) continue;
            const Node2& cn = N[ch[c]];
            for (int k = 0; k < 3; k++) {
                float scale = std::ldexp(1.0f, (int)w.e[k] - 127);
                int lo = (int)std::floor((cn.box.lo[k] - w.p[k]) / scale);
                int hi = (int)std::ceil((cn.box.hi[k] - w.p[k]) / scale);
#if RTX_DEQUANT != 0
                lo = std::max(1, std::min(254, lo)); hi = std::max(1, std::min(254, hi));
                while (lo > 1 && w.p[k] + (float)lo * scale > cn.box.lo[k]) lo--;
                while (hi < 254 && w.p[k] + (float)hi * scale < cn.box.hi[k]) hi++;
                lo -= 1; hi += 1;                                  // the half-step slack of the PRMT conversion
#else
                lo = std::max(0, std::min(255, lo)); hi = std::max(0, std::min(255, hi));
                while (lo > 0 && w.p[k] + (float)lo * scale > cn.box.lo[k]) lo--;
                while (hi < 255 && w.p[k] + (float)hi * scale < cn.box.hi[k]) hi++;
#endif
                w.qlo[k][s] = (uint8_t)lo; w.qhi[k][s] = (uint8_t)hi;
            }
            if (cn.count == 0) {
                w.imask |= (uint8_t)(1u << s);
                w.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
                pend[npend++] = {next_child++, ch[c]};
            } else {
                assert(cn.count <= kMaxLeaf && prim_off + cn.count <= 24);
                uint32_t unary = cn.count == 1 ? 1u : cn.count == 2 ? 3u : 7u;
                w.meta[s] = (uint8_t)((unary << 5) | prim_off);
                for (uint32_t i = 0; i < cn.count; i++) out.prim_order.push_back(b2.idx[cn.first + i]);
                prim_off += cn.count;
            }
        }
        out.nodes[wi] = w;
        for (int i = 0; i < npend; i++) emit(pend[i].wi, pend[i].n2, depth + 1);
    }
};

}  // namespace

void build_wide_bvh(const Aabb3* boxes, uint32_t n, WideBvh& out, int depth_limit) {
    for (int pass = 0; pass < 2; pass++) {
        out.nodes.clear(); out.prim_order.clear(); out.max_depth = 0;
        if (n == 0) return;
        Builder2 b2; b2.boxes = boxes;
        b2.balanced = pass == 1;
        b2.build(n);
        out.nodes.reserve(n / 2 + 8);
        out.prim_order.reserve(n);
        out.nodes.emplace_back();
        WideBuilder wb{b2, out};
        wb.emit(0, 0, 1);
        // the greedy largest-area collapse does not bound the depth along small-area paths of an unbalanced SAH tree: a tree
        // that would not fit the traversal stack is rebuilt with object-median splits and a size-balanced collapse
        // (depth ~ log8 n) instead of being refused
        if (depth_limit <= 0 || out.max_depth <= depth_limit) return;
    }
}

}  // namespace rtx
