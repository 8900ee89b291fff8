// rtx_kernels.cuh — the wavefront kernels of the ray-casting hot path (sm_100a).
//
//   K1 raygen_kernel        Raytracing::render ray generation        (reference src/raytracing.rs:319-396)
//   K2 closest_kernel       Raytracing::trace, closest hit            (:429-490, called at :726)
//   K4 shade_kernel         get_color_depth_normal_id body            (:720-998), emits shadow + child rays
//   K3 shadow_kernel        Raytracing::trace, stop_on_first_hit      (:883-913) + light accumulation (:917-919)
//   K6 resolve_kernel       mean / clamp / u8 / gamma / normalize     (:406-426) -> the four frame buffers
//   probe_kernel            Raytracing::trace as a test hook (rtx_trace_probe)
//
// The recursion of get_color_depth_normal_id is linear in the two child radiances with scalar
// coefficients, so each queued ray carries one float of throughput:
//   local = (1-f)·ao·a·(1-ρ)·D + ao·f·fog + E ;  w_R = w·(1-f)·ao·a·ρ ;  w_T = w·(1-f)·ao·kt
#pragma once
#include "rtx_device.cuh"

namespace rtx {

constexpr int kTraceBlock = 128;
constexpr int kShadeBlock = 128;

struct WorkCtr { uint32_t* next; };    // one zeroed counter per launch (dynamic warp-granular fetch)

// ---- K1 ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mul4(const float* m, float x, float y, float z, float w, float out[4]) {
#pragma unroll
    for (int r = 0; r < 4; r++) out[r] = xa(xa(xa(xm(m[r], x), xm(m[4 + r], y)), xm(m[8 + r], z)), xm(m[12 + r], w));
}

__device__ __forceinline__ void gen_ray(const FrameDev& F, uint32_t px, uint32_t py, uint32_t x_i, uint32_t y_i, float3& o, float3& d) {
    const float x_f = (float)px, y_f = (float)py, w = (float)F.width, h = (float)F.height;
    const float x_step = xd(2.0f, w), y_step = xd(2.0f, h);
    const float inv_cell = xd(1.0f, (float)F.cell_size);
    float x_trans = xm(xm(x_step, (float)x_i), inv_cell);
    float y_trans = xm(xm(y_step, (float)y_i), inv_cell);
    const bool dof = F.aperture_size > 1.0f && F.focal_length > 1.0f;
    if (dof && F.n_samples > 1) { x_trans = xs(x_trans, xd(x_step, 2.0f)); y_trans = xs(y_trans, xd(y_step, 2.0f)); }
    const float cx = xs(xm(xd(xa(x_f, 0.5f), w), 2.0f), 1.0f), cy = xs(1.0f, xm(xd(xa(y_f, 0.5f), h), 2.0f));
    float pp[4], o4[4], d4[4];
    if (dof) {                                                                   // :338-377
        const float aperture_scale = xd((float)F.width, 800.0f);
        x_trans = xm(x_trans, xm(F.aperture_size, aperture_scale));
        y_trans = xm(y_trans, xm(F.aperture_size, aperture_scale));
        mul4(F.pinv, cx, cy, -1.0f, 1.0f, pp); pp[3] = 1.0f;
        const float rd[3] = {xs(pp[0], 0.0f), xs(pp[1], 0.0f), xs(pp[2], 0.0f)};
        float origin[4], dir[4];
        mul4(F.vinv, 0.0f, 0.0f, 0.0f, 1.0f, origin);
        mul4(F.vinv, rd[0], rd[1], rd[2], 0.0f, dir);
        const float n4 = xsqrt(xa(xa(xm(dir[0], dir[0]), xm(dir[2], dir[2])), xa(xm(dir[1], dir[1]), xm(dir[3], dir[3]))));
        for (int i = 0; i < 4; i++) dir[i] = xd(dir[i], n4);
        const float dist = xnorm(f3(rd[0], rd[1], rd[2]));
        const float f = xd(1.0f, xd(dist, xa(dist, F.focal_length)));
        float p[3]; for (int i = 0; i < 3; i++) p[i] = xa(origin[i], xm(f, dir[i]));
        mul4(F.pinv, xa(cx, x_trans), xa(cy, y_trans), -1.0f, 1.0f, pp); pp[3] = 1.0f;
        mul4(F.vinv, pp[0], pp[1], pp[2], pp[3], o4);
        o = f3(o4[0], o4[1], o4[2]);
        d = f3(xs(p[0], o4[0]), xs(p[1], o4[1]), xs(p[2], o4[2]));
        return;
    }
    mul4(F.pinv, xa(cx, x_trans), xa(cy, y_trans), -1.0f, 1.0f, pp); pp[3] = 1.0f;   // :381-395
    mul4(F.vinv, pp[0], pp[1], pp[2], pp[3], o4);
    mul4(F.vinv, xs(pp[0], 0.0f), xs(pp[1], 0.0f), xs(pp[2], 0.0f), 0.0f, d4);
    o = f3(o4[0], o4[1], o4[2]); d = f3(d4[0], d4[1], d4[2]);
}

// batch = pixels [p0, p0+np) of the owned pixel list x samples [s0, s0+ns); ray i -> (s0 + i/np, p0 + i%np)
__global__ void __launch_bounds__(256) raygen_kernel(FrameDev F, const uint32_t* __restrict__ pixel_list, uint32_t p0, uint32_t np, uint32_t s0,
                                                     uint32_t ns, RayQ q, uint32_t q_base) {
    const uint32_t n = np * ns;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // ray order inside the batch: groups of `sg` samples of one pixel are consecutive (sg = 1: pixel-major per sample)
        const uint32_t sg = F.sample_group;
        const uint32_t blk = i / (np * sg), rem = i % (np * sg);                 // sample block, position inside it
        const uint32_t sb = min(sg, ns - blk * sg);                              // samples in this (possibly short, last) block
        const uint32_t s = s0 + blk * sg + rem % sb, pi = p0 + rem / sb;
        const uint32_t pixel = __ldg(pixel_list + pi);
        const ushort2 xy = F.sample_table[s];
        float3 o, d;
        gen_ray(F, pixel % F.width, pixel / F.width, xy.x, xy.y, o, d);
        d = xnormalize(d);                                                       // :722-723
        const uint32_t flags = (s == F.n_samples - 1) ? RF_ID_OWNER : 0u;
        q.o[q_base + i] = make_float4(o.x, o.y, o.z, 1.0f);
        q.d[q_base + i] = make_float4(d.x, d.y, d.z, __uint_as_float(pixel));
        q.m[q_base + i] = make_uint2((s & 0xffffu) | (1u << 16) | (flags << 24), 1u);
    }
}

// ---- K2 ------------------------------------------------------------------------------------------
// Persistent threads with dynamic ray fetch: every lane owns one ray and one traversal state; when fewer than
// kRefill lanes of a warp still have a ray, the warp leaves the step loop and the idle lanes take new rays
// (one atomicAdd per warp).  All 32 lanes meet at a ballot after every step, so a step is the unit of divergence.
#ifndef RTX_REFILL
#define RTX_REFILL 24
#endif
#ifndef RTX_MIN_BLOCKS
#define RTX_MIN_BLOCKS 7
#endif
#ifndef RTX_LEAF_BATCH
#define RTX_LEAF_BATCH 8
#endif
#ifndef RTX_NODE_REPS
#define RTX_NODE_REPS 1
#endif
constexpr int kNodeReps = RTX_NODE_REPS;   // node phases per loop iteration
constexpr int kRefill = RTX_REFILL;        // refill the warp when fewer lanes than this still own a ray
constexpr int kLeafBatch = RTX_LEAF_BATCH;  // run a leaf phase when at least this many lanes have pending leaf entries
constexpr uint32_t kFull = 0xffffffffu;
#ifndef RTX_SPLIT_LEAF
#define RTX_SPLIT_LEAF 1                    // triangles and TLAS items in separate LEAF rounds, each with its own threshold
#endif
#ifndef RTX_ITEM_BATCH
#define RTX_ITEM_BATCH 4
#endif
constexpr int kItemBatch = RTX_ITEM_BATCH;
#ifndef RTX_LEAF_BATCH_ANY
#define RTX_LEAF_BATCH_ANY RTX_LEAF_BATCH   // any-hit rays end at their first hit: postponing triangle tests costs node visits the hit would have saved
#endif
constexpr int kLeafBatchAny = RTX_LEAF_BATCH_ANY;
#ifndef RTX_TRI_PAIR
#define RTX_TRI_PAIR 0                      // triangle round loads both triangles before testing either (see lane_leaf_tri2)
#endif
#ifndef RTX_TRI_REPS
#define RTX_TRI_REPS 2                      // triangles handled per lane in one triangle round
#endif

template <bool STATS>
__global__ void __launch_bounds__(kTraceBlock, RTX_MIN_BLOCKS) closest_kernel(SceneDev S, RayQ q, uint32_t q_base, uint32_t n_host, const uint32_t* __restrict__ n_ptr,
                                                              HitRec* __restrict__ hits, uint32_t* work, Counters* ctr) {
    // n_ptr != nullptr: the ray count is whatever the previous level's shade kernel appended (sync-free frames: the host never
    // reads it), clamped to the queue capacity n_host
    const uint32_t n = n_ptr ? min(*n_ptr, n_host) : n_host;
    TravStats st{0, 0}; uint32_t n_items = 0, n_sph = 0;
    uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint2 stack[kLaneStack];
    Lane L; bool has = false, exhausted = false; uint32_t my = 0, depth = 1;
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        const uint32_t idle = __ballot_sync(kFull, !has);
        if (idle != 0u && !exhausted) {
            if (STATS) { ph[5] += (lane == 0); ph[6] += !has; }
            const uint32_t leader = __ffs(idle) - 1u, cnt = __popc(idle);
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(work, cnt);
            base = __shfl_sync(kFull, base, leader);
            if (!has) {
                const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < n) {
                    const float4 ro = q.o[q_base + i], rd = q.d[q_base + i];
                    depth = (q.m[q_base + i].x >> 16) & 0xffu;
                    lane_init(L, S, f3(ro.x, ro.y, ro.z), f3(rd.x, rd.y, rd.z), 3.402823466e+38f);
                    has = true; my = i;
                }
            }
            if (base + cnt >= n) exhausted = true;
        }
        if (!__any_sync(kFull, has)) break;
        for (;;) {
            // NODE phase for every lane with a pending node group, then a LEAF round when enough lanes have leaf
            // entries pending (or nobody has node work left)
#pragma unroll
            for (int rep = 0; rep < kNodeReps - 1; rep++) {                    // extra node phases per iteration amortise the votes below
                if (has && L.ng.y > 0x00FFFFFFu) lane_node<UT_CLOSEST, STATS>(L, stack, S, st);
                if (has && L.ng.y <= 0x00FFFFFFu && L.tg.y == 0u && L.sp != 0 && !(L.blas_base >= 0 && L.sp == L.blas_base)) lane_pop(L, stack);
            }
            if (STATS) { ph[0] += (lane == 0); ph[1] += has; ph[2] += (has && L.ng.y > 0x00FFFFFFu); }
            if (has && L.ng.y > 0x00FFFFFFu) lane_node<UT_CLOSEST, STATS>(L, stack, S, st);
#if RTX_SPLIT_LEAF
            {
                const bool want_tri = has && L.tg.y != 0u && L.blas_base >= 0, want_item = has && L.tg.y != 0u && L.blas_base < 0;
                const uint32_t mt = __ballot_sync(kFull, want_tri), mi = __ballot_sync(kFull, want_item);
                const bool no_nodes = !__any_sync(kFull, has && L.ng.y > 0x00FFFFFFu);
                if (mt != 0u && (__popc(mt) >= kLeafBatch || no_nodes)) {
                    if (STATS) { ph[3] += (lane == 0); ph[4] += want_tri; }
                    if (want_tri) {
#if RTX_TRI_PAIR
                        lane_leaf_tri2<UT_CLOSEST, STATS>(L, S, false, depth, st);
#else
                        lane_leaf<UT_CLOSEST, STATS, 1>(L, stack, S, false, depth, st, n_items, n_sph);
#pragma unroll
                        for (int rep = 1; rep < RTX_TRI_REPS; rep++)
                            if (L.tg.y != 0u && L.blas_base >= 0) lane_leaf<UT_CLOSEST, STATS, 1>(L, stack, S, false, depth, st, n_items, n_sph);
#endif
                    }
                }
                if (mi != 0u && (__popc(mi) >= kItemBatch || no_nodes)) {
                    if (STATS) { ph[3] += (lane == 0); ph[4] += want_item; }
                    if (want_item) lane_leaf<UT_CLOSEST, STATS, 2>(L, stack, S, false, depth, st, n_items, n_sph);
                }
            }
#else
            {
                const bool want_leaf = has && L.tg.y != 0u;
                const uint32_t ml = __ballot_sync(kFull, want_leaf);
                if (ml != 0u && (__popc(ml) >= kLeafBatch || !__any_sync(kFull, has && L.ng.y > 0x00FFFFFFu))) {
                    if (STATS) { ph[3] += (lane == 0); ph[4] += want_leaf; }
                    if (want_leaf) lane_leaf<UT_CLOSEST, STATS>(L, stack, S, false, depth, st, n_items, n_sph);
                }
            }
#endif
            if (has && L.ng.y <= 0x00FFFFFFu && L.tg.y == 0u && lane_pop(L, stack)) {
                HitRec h; h.t = L.tmax; h.item = L.bitem; h.prim = L.bprim; h.flags = L.bflags;
                hits[my] = h;
                has = false;
            }
            const uint32_t act = __ballot_sync(kFull, has);
            if (act == 0u || (!exhausted && __popc(act) < kRefill)) break;
        }
    }
    if (STATS) {
        atomicAdd(&ctr->node_visits[0], (unsigned long long)st.nodes); atomicAdd(&ctr->tri_tests[0], (unsigned long long)st.tris);
        atomicAdd(&ctr->item_tests, (unsigned long long)n_items); atomicAdd(&ctr->sphere_tests, (unsigned long long)n_sph);
        for (int k = 0; k < 7; k++) atomicAdd(&ctr->phase[0][k], (unsigned long long)ph[k]);
    }
}

// ---- K3 ------------------------------------------------------------------------------------------
// Shadow rays in two kernels.  K3a answers the order-independent half of the query ("does ANY item that casts
// shadows have a hit at toi <= light distance") with the same persistent any-hit traversal; a ray with no such
// hit is lit and done.  An occluded ray is final when the reference's first-hit order cannot change the result
// (directional light and no alpha-textured material in the scene); the others are compacted into `slow` and
// K3b re-walks them item by item in the reference's bbox-key order (trace_shadow_fast).
template <bool STATS>
__global__ void __launch_bounds__(kTraceBlock, RTX_MIN_BLOCKS) shadow_any_kernel(SceneDev S, FrameDev F, ShadowQ q, const uint32_t* __restrict__ n_ptr, uint32_t n_cap, uint32_t depth,
                                                                 uint32_t* work, uint32_t* __restrict__ slow, uint32_t* slow_count, Counters* ctr) {
    TravStats st{0, 0}; uint32_t n_items = 0, n_sph = 0;
    uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint2 stack[kLaneStack];
    Lane L; bool has = false, exhausted = false, fin = false; uint32_t my = 0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = min(*n_ptr, n_cap);
    for (;;) {
        // write-back of the rays that ended since the last refill, all at once (the lanes that are about to be refilled)
        {
            bool to_slow = false, to_beyond = false;
            if (fin) {
                fin = false;
                const bool occluded = L.bitem != 0xFFFFFFFFu;
                const float okey = __uint_as_float(L.bprim);
                const bool earlier_other = L.bface != 0xFFFFFFFFu && (okey < L.bkey || (okey == L.bkey && L.bface < L.bitem));
                // final unless the reference's first-hit order could change the answer: an alpha-textured material exists, or (finite
                // light distance and) another candidate sorts before the occluder, or the occluder's own key lies beyond the light
                // (then the clipped item walk may not have seen every earlier candidate).  With a merged BLAS the grouped items are
                // not enumerated: whether one of them sorts before the occluder AND is hit beyond the light is asked afterwards, only
                // for occluded rays with a finite light distance (shadow_beyond_kernel).
                if (!occluded || (!S.any_alpha_tex && L.tmax >= 3.402823466e+38f)) to_slow = false;
                else if (S.any_alpha_tex || earlier_other || L.bkey > L.tmax) to_slow = true;
                else if (S.group_root != 0xFFFFFFFFu) to_beyond = true;
                if (!to_slow && !to_beyond) {
                    const float4 rc = q.c[my];
                    const float k = occluded ? 1.0f - rc.w : 1.0f;               // raytracing.rs:898,912
                    atomicAdd(&F.accum_c[__float_as_uint(q.d[my].w)], make_float4(rc.x * k, rc.y * k, rc.z * k, 0.0f));
                    if (q.probe) q.probe[my] = make_uint4(occluded ? 0u : 1u, L.bitem, 0xFFFFFFFFu, 0u);   // rtx_shadow_probe only
                }
            }
            const uint32_t sm = __ballot_sync(kFull, to_slow);
            if (sm != 0u) {
                const uint32_t leader = __ffs(sm) - 1u;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(slow_count, __popc(sm));
                base = __shfl_sync(kFull, base, leader);
                if (to_slow) slow[base + __popc(sm & ((1u << lane) - 1u))] = my;
            }
            const uint32_t bm = __ballot_sync(kFull, to_beyond);
            if (bm != 0u) {
                const uint32_t leader = __ffs(bm) - 1u;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(slow_count + 1, __popc(bm));      // ctr[5]: beyond-light queue
                base = __shfl_sync(kFull, base, leader);
                if (to_beyond) q.beyond[base + __popc(bm & ((1u << lane) - 1u))] = make_uint4(my, __float_as_uint(L.bkey), L.bitem, 0u);
            }
        }
        const uint32_t idle = __ballot_sync(kFull, !has);
        if (idle != 0u && !exhausted) {
            if (STATS) { ph[5] += (lane == 0); ph[6] += !has; }
            const uint32_t leader = __ffs(idle) - 1u, cnt = __popc(idle);
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(work, cnt);
            base = __shfl_sync(kFull, base, leader);
            if (!has) {
                const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < n) {
                    const float4 ro = q.o[i], rd = q.d[i];
                    lane_init(L, S, f3(ro.x, ro.y, ro.z), f3(rd.x, rd.y, rd.z), ro.w);
                    has = true; my = i;
                }
            }
            if (base + cnt >= n) exhausted = true;
        }
        if (!__any_sync(kFull, has)) break;
        for (;;) {
            // NODE phase for every lane with a pending node group, then a LEAF round when enough lanes have leaf
            // entries pending (or nobody has node work left)
#pragma unroll
            for (int rep = 0; rep < kNodeReps - 1; rep++) {                    // extra node phases per iteration amortise the votes below
                if (has && L.ng.y > 0x00FFFFFFu) lane_node<UT_ANY, STATS>(L, stack, S, st);
                if (has && L.ng.y <= 0x00FFFFFFu && L.tg.y == 0u && L.sp != 0 && !(L.blas_base >= 0 && L.sp == L.blas_base)) lane_pop(L, stack);
            }
            if (STATS) { ph[0] += (lane == 0); ph[1] += has; ph[2] += (has && L.ng.y > 0x00FFFFFFu); }
            if (has && L.ng.y > 0x00FFFFFFu) lane_node<UT_ANY, STATS>(L, stack, S, st);
#if RTX_SPLIT_LEAF
            {
                const bool want_tri = has && L.tg.y != 0u && L.blas_base >= 0, want_item = has && L.tg.y != 0u && L.blas_base < 0;
                const uint32_t mt = __ballot_sync(kFull, want_tri), mi = __ballot_sync(kFull, want_item);
                const bool no_nodes = !__any_sync(kFull, has && L.ng.y > 0x00FFFFFFu);
                if (mt != 0u && (__popc(mt) >= kLeafBatchAny || no_nodes)) {
                    if (STATS) { ph[3] += (lane == 0); ph[4] += want_tri; }
                    if (want_tri) {
#if RTX_TRI_PAIR
                        lane_leaf_tri2<UT_ANY, STATS>(L, S, true, depth, st);
#else
                        lane_leaf<UT_ANY, STATS, 1>(L, stack, S, true, depth, st, n_items, n_sph);
#pragma unroll
                        for (int rep = 1; rep < RTX_TRI_REPS; rep++)
                            if (L.tg.y != 0u && L.blas_base >= 0) lane_leaf<UT_ANY, STATS, 1>(L, stack, S, true, depth, st, n_items, n_sph);
#endif
                    }
                }
                if (mi != 0u && (__popc(mi) >= kItemBatch || no_nodes)) {
                    if (STATS) { ph[3] += (lane == 0); ph[4] += want_item; }
                    if (want_item) lane_leaf<UT_ANY, STATS, 2>(L, stack, S, true, depth, st, n_items, n_sph);
                }
            }
#else
            {
                const bool want_leaf = has && L.tg.y != 0u;
                const uint32_t ml = __ballot_sync(kFull, want_leaf);
                if (ml != 0u && (__popc(ml) >= kLeafBatch || !__any_sync(kFull, has && L.ng.y > 0x00FFFFFFu))) {
                    if (STATS) { ph[3] += (lane == 0); ph[4] += want_leaf; }
                    if (want_leaf) lane_leaf<UT_ANY, STATS>(L, stack, S, true, depth, st, n_items, n_sph);
                }
            }
#endif
            if (has && L.ng.y <= 0x00FFFFFFu && L.tg.y == 0u && lane_pop(L, stack)) { has = false; fin = true; }   // result stays in L until the refill
            const uint32_t act = __ballot_sync(kFull, has);
            if (act == 0u || (!exhausted && __popc(act) < kRefill)) break;
        }
    }
    if (STATS) {
        atomicAdd(&ctr->node_visits[1], (unsigned long long)st.nodes); atomicAdd(&ctr->tri_tests[1], (unsigned long long)st.tris);
        atomicAdd(&ctr->item_tests, (unsigned long long)n_items); atomicAdd(&ctr->sphere_tests, (unsigned long long)n_sph);
        for (int k = 0; k < 7; k++) atomicAdd(&ctr->phase[1][k], (unsigned long long)ph[k]);
    }
}

// K3a': occluded rays with a finite light distance, scenes with a merged BLAS (see group_hit_beyond_light).  Not found (the usual
// case: nothing the ray meets behind the light sorts before the occluder) -> the ray is occluded, final; found -> exact walk.
template <bool STATS>
__global__ void __launch_bounds__(kTraceBlock) shadow_beyond_kernel(SceneDev S, FrameDev F, ShadowQ q, const uint32_t* __restrict__ n_ptr, uint32_t n_cap, uint32_t depth,
                                                                    uint32_t* __restrict__ slow, uint32_t* slow_count, Counters* ctr) {
    TravStats st{0, 0}; uint32_t found = 0, seen = 0;
    const uint32_t n = min(*n_ptr, n_cap);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint4 e = q.beyond[j];
        const uint32_t i = e.x;
        const float4 ro = q.o[i], rd = q.d[i];
        seen++;
        if (group_hit_beyond_light<STATS>(S, f3(ro.x, ro.y, ro.z), f3(rd.x, rd.y, rd.z), ro.w, __uint_as_float(e.y), e.z, depth, st)) {
            slow[atomicAdd(slow_count, 1u)] = i; found++;
        } else {
            const float4 rc = q.c[i];
            const float k = 1.0f - rc.w;
            atomicAdd(&F.accum_c[__float_as_uint(rd.w)], make_float4(rc.x * k, rc.y * k, rc.z * k, 0.0f));
            if (q.probe) q.probe[i] = make_uint4(0u, e.z, 0xFFFFFFFFu, 0u);
        }
    }
    if (STATS) {
        atomicAdd(&ctr->node_visits[1], (unsigned long long)st.nodes); atomicAdd(&ctr->tri_tests[1], (unsigned long long)st.tris);
        atomicAdd(&ctr->beyond_rays, (unsigned long long)seen); atomicAdd(&ctr->beyond_found, (unsigned long long)found);
    }
}

// K3b: reference-order walk.  `slow` == nullptr: all rays of the queue (RTX_DEBUG_ORDERED_SHADOW runs the literal
// "closest hit of every item in order" version on everything, without K3a).
template <bool STATS, bool ORDERED>
__global__ void __launch_bounds__(kTraceBlock) shadow_exact_kernel(SceneDev S, FrameDev F, ShadowQ q, const uint32_t* __restrict__ slow,
                                                                   const uint32_t* __restrict__ n_ptr, uint32_t n_cap, uint32_t depth, Counters* ctr) {
    TravStats st{0, 0};
    const uint32_t n = min(*n_ptr, n_cap);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = slow ? slow[j] : j;
        const float4 ro = q.o[i], rd = q.d[i], rc = q.c[i];
        const float3 o = f3(ro.x, ro.y, ro.z), d = f3(rd.x, rd.y, rd.z);
        const float len = ro.w;
        const uint32_t pixel = __float_as_uint(rd.w);
        Best b;
        bool in_light;
        if (ORDERED) {
            trace_shadow_ordered<STATS>(S, o, d, depth, b, st);
            in_light = b.item == 0xFFFFFFFFu || b.t > len;                  // :885-892 (len = +inf for directional)
        } else {
            trace_shadow_fast<STATS>(S, o, d, depth, len, b, st);
            in_light = b.item == 0xFFFFFFFFu;
        }
        float k = 1.0f; uint32_t probe_face = 0;
        if (!in_light) {                                                    // :895-913
            float ssa = rc.w;
            const DItem& occ = S.items[b.item];
            if (occ.flags & IF_ALPHA_TEX) {
                const DItem& recv = S.items[q.r[i]];
                uint32_t face_id = 0;
                if (occ.flags & IF_MESH) {
                    const uint32_t face = __float_as_uint(__ldg(S.tris + (size_t)b.prim * 3).w);
                    face_id = (b.flags & HF_BACK) ? face + occ.n_faces : face;
                }
                probe_face = face_id;
                const float3 shp = o + d * b.t;
                float u, v; item_get_uv(S, recv, shp, face_id, u, v);        // receiver's get_uv (sic, :905)
                float4 tc;
                if (get_tex_color(S, S.mats[occ.material], true, u, v, 4 /*Alpha*/, tc)) ssa *= tc.x;
            }
            k = 1.0f - ssa;
        }
        atomicAdd(&F.accum_c[pixel], make_float4(rc.x * k, rc.y * k, rc.z * k, 0.0f));
        if (q.probe) q.probe[i] = make_uint4(in_light ? 1u : 0u, b.item, (!in_light && (S.items[b.item].flags & IF_ALPHA_TEX)) ? __float_as_uint(b.t) : 0xFFFFFFFFu, probe_face);
    }
    if (STATS) { atomicAdd(&ctr->node_visits[1], (unsigned long long)st.nodes); atomicAdd(&ctr->tri_tests[1], (unsigned long long)st.tris); }
}

// ---- K4 ------------------------------------------------------------------------------------------
struct ShadeOut {
    RayQ child; uint32_t child_cap; uint32_t* child_count;      // level d+1 queue
    ShadowQ shadow; uint32_t shadow_cap; uint32_t* shadow_count;
    uint32_t* overflow;                                         // set to 1 if a queue would overflow (never, by construction)
    uint32_t* skipped;                                          // RTX_OPT_SKIP_ZERO_SHADOW: count of untraced zero-contribution shadow rays
};

#ifndef RTX_SHADE_PREFETCH
#define RTX_SHADE_PREFETCH 0                // measured: see DESIGN.md §6
#endif
#ifndef RTX_SHADE_MIN_BLOCKS
#define RTX_SHADE_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(kShadeBlock, RTX_SHADE_MIN_BLOCKS) shade_kernel(SceneDev S, FrameDev F, RayQ q, uint32_t q_base, uint32_t n_host, const uint32_t* __restrict__ n_ptr,
                                                            const HitRec* __restrict__ hits, ShadeOut out) {
    const float PI = 3.14159265358979323846f;
    const uint32_t n = n_ptr ? min(*n_ptr, n_host) : n_host;
    uint32_t n_enabled = 0;
    for (uint32_t li = 0; li < S.n_lights; li++) n_enabled += S.lights[li].enabled ? 1u : 0u;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const bool active = i < n;
        // --- load ---
        float3 o = f3(0, 0, 0), d = f3(0, 0, 1); float wgt = 0.0f; uint32_t pixel = 0, sample = 0, depth = 1, rflags = 0, path = 1;
        HitRec h; h.item = 0xFFFFFFFFu; h.t = 0; h.prim = 0; h.flags = 0;
        if (active) {
            const float4 ro = q.o[q_base + i], rd = q.d[q_base + i]; const uint2 m = q.m[q_base + i];
            o = f3(ro.x, ro.y, ro.z); d = f3(rd.x, rd.y, rd.z); wgt = ro.w; pixel = __float_as_uint(rd.w);
            sample = m.x & 0xffffu; depth = (m.x >> 16) & 0xffu; rflags = m.x >> 24; path = m.y;
            h = hits[i];
#if RTX_SHADE_PREFETCH
            {   // the next ray of this thread: pull its queue entry and hit record towards L1 while this one is shaded
                const uint32_t ni = i + gridDim.x * blockDim.x;
                if (ni < n) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(q.o + q_base + ni)); asm volatile("prefetch.global.L1 [%0];" ::"l"(q.d + q_base + ni));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(q.m + q_base + ni)); asm volatile("prefetch.global.L1 [%0];" ::"l"(hits + ni));
                }
            }
#endif
        }
        const bool hit = active && h.item != 0xFFFFFFFFu;
        if (active && !hit && (rflags & RF_ID_OWNER)) F.ids[pixel] = 0u;

        bool emit_refl = false, emit_trans = false;
        float3 refl_o, refl_d, trans_o, trans_d; float w_refl = 0.0f, w_trans = 0.0f; uint32_t trans_flags = 0;
        // per-light shadow payloads are produced inside the light loop (warp-aggregated per light)
        float3 hit_point = f3(0, 0, 0), surface_normal = f3(0, 0, 1); float coef_d = 0.0f;
        float4 base_color = make_float4(0, 0, 0, 0), specular_color = make_float4(0, 0, 0, 0);
        float shininess = 0.0f, mat_alpha = 1.0f, shadow_z_lo = 1.0f; bool recv_shadow = false, mat_mc = false;
        float3 direct = f3(0, 0, 0), constant = f3(0, 0, 0); float depth_add = 0.0f;

        // --- queue space: ONE reservation per warp for the shadow rays of all lights (and, before the light loop, one for both
        // child rays).  Every warp of the GPU hits the same two counters, so the atomics' round trips are long: they are issued as
        // early as their counts are known and overlap the normal / texture / light work.  Per light the warp's rays stay
        // contiguous: slot = base + (enabled-light index) * rays-per-light + rank.
        const bool skip_mode = (F.debug_flags & 4u) != 0u;                       // opt-in zero-contribution skipping: per-light appends
        const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
        uint32_t sh_base = 0, sh_cnt = 0, sh_rank = 0;
        {
            bool rs = false;
            if (hit) rs = S.mats[S.items[h.item].material].receive_shadow != 0u;
            const uint32_t m = __ballot_sync(0xffffffffu, rs);
            sh_cnt = __popc(m); sh_rank = __popc(m & lt_mask);
            if (!skip_mode && lane == 0 && sh_cnt != 0u && n_enabled != 0u) sh_base = atomicAdd(out.shadow_count, sh_cnt * n_enabled);
        }
        const uint32_t rng = F.monte_carlo ? mc_base(F.mc_seed, pixel, sample) : 0u;
        float3 view_dir = f3(0, 0, 1);
        if (hit) {
            const DItem& item = S.items[h.item];
            const DMaterial& mat = S.mats[item.material];
            view_dir = norm3(-d);
            uint32_t face_id;
            const float3 normal = hit_normal(S, item, o, d, h.t, h.prim, h.flags, face_id);
            const float hit_dist = h.t;
            if (depth == 1) {                                                    // :400-403 depth / normal sums (the depth rides along with the colour below)
                depth_add = hit_dist;
                atomicAdd(&F.accum_n[pixel], make_float4(normal.x, normal.y, normal.z, 0.0f));
            }
            if (rflags & RF_ID_OWNER) F.ids[pixel] = item.id;
            surface_normal = normal;
            hit_point = xadd(o, xscale(d, hit_dist));
            const bool has_uv = mat.any_texture != 0;
            float u = 0.0f, v = 0.0f;
            if (has_uv) item_get_uv(S, item, hit_point, face_id, u, v);
            // all eight texture lookups in ONE rolled loop (one copy of the nearest / bilinear fetch code instead of eight:
            // the kernel otherwise outgrows the instruction cache)
            float4 texc[8]; uint32_t tex_mask = 0;
#pragma unroll 1
            for (int tt = 0; tt < 8; tt++) {
                float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (get_tex_color(S, mat, has_uv, u, v, tt, c4)) tex_mask |= 1u << tt;
                texc[tt] = c4;
            }
            float4 tc;
            if ((tex_mask >> 3) & 1u) { tc = texc[3];                                  // Normal :757-784
                float3 tangent = xcross(normal, f3(0, 1, 0));
                if (xnorm(tangent) <= 0.0001f) tangent = xcross(normal, f3(0, 0, 1));
                tangent = xnormalize_s(tangent);
                const float3 bitangent = xnormalize_s(xcross(normal, tangent));
                float3 nm = f3(xs(xm(tc.x, 2.0f), 1.0f), xs(xm(tc.y, 2.0f), 1.0f), xs(xm(tc.z, 2.0f), 1.0f));
                nm.x = xm(nm.x, mat.normal_map_strength); nm.y = xm(nm.y, mat.normal_map_strength);
                nm = xnormalize_s(nm);
                surface_normal = xnormalize_s(f3(xa(xa(xm(tangent.x, nm.x), xm(bitangent.x, nm.y)), xm(normal.x, nm.z)),
                                               xa(xa(xm(tangent.y, nm.x), xm(bitangent.y, nm.y)), xm(normal.y, nm.z)),
                                               xa(xa(xm(tangent.z, nm.x), xm(bitangent.z, nm.y)), xm(normal.z, nm.z))));
            }
            const bool has_rough = (tex_mask >> 5) & 1u; tc = texc[5];                 // Roughness :787-798
            if (F.monte_carlo && mat.monte_carlo && (mat.roughness > 0.0f || has_rough)) {
                float roughness = mat.roughness;
                if (has_rough) roughness = (1.0f / PI / 2.0f) * tc.x;
                surface_normal = jitter(surface_normal, has_rough ? jitter_z_lo(roughness) : mat.rough_z_lo, rng, path, 0);
            }
            float4 ambient_color = make_float4(mat.ambient[0], mat.ambient[1], mat.ambient[2], 1.0f);   // :801-803
            if ((tex_mask >> 1) & 1u) { tc = texc[1]; ambient_color.x *= tc.x; ambient_color.y *= tc.y; ambient_color.z *= tc.z; }
            base_color = make_float4(mat.base[0], mat.base[1], mat.base[2], 1.0f);
            if (tex_mask & 1u) { tc = texc[0]; base_color.x *= tc.x; base_color.y *= tc.y; base_color.z *= tc.z; base_color.w *= tc.w; }
            specular_color = make_float4(mat.specular[0], mat.specular[1], mat.specular[2], 1.0f);
            if ((tex_mask >> 2) & 1u) { tc = texc[2]; specular_color.x *= tc.x; specular_color.y *= tc.y; specular_color.z *= tc.z; }
            float alpha = mat.alpha * base_color.w;                                // :806-811
            if ((tex_mask >> 4) & 1u) alpha *= texc[4].x;

            float reflectivity = mat.reflectivity;                                // :928-933
            if ((tex_mask >> 7) & 1u) reflectivity = texc[7].x;
            const bool can_recurse = depth <= F.max_recursion;
            const bool reflect_on = reflectivity > 0.0f && can_recurse;           // :938
            bool trans_exists = false;
            if (alpha < 1.0f && can_recurse)                                      // :948-952
                trans_exists = create_transmission(surface_normal, d, hit_point, mat.refraction_index, trans_o, trans_d);
            const float a = trans_exists ? alpha : ((alpha < 1.0f && !can_recurse) ? alpha : 1.0f);   // :959-975 (TIR keeps 1)
            float kt = 0.0f;
            if (trans_exists) {                                                   // kr (:925) only matters when a transmission ray exists
                const float kr = fresnel(d, surface_normal, mat.refraction_index);
                kt = (kr < 1.0f ? (1.0f - kr) : 1.0f) * (1.0f - alpha);
            }
            const float fog = fminf(F.fog_density * hit_dist, 1.0f);              // :978-982
            float ao = 1.0f;
            if ((tex_mask >> 6) & 1u) ao = texc[6].x;                                  // :985-991
            const float thru = wgt * ao * (1.0f - fog);
            coef_d = thru * a * (1.0f - reflectivity);
            constant = f3(wgt * (ao * fog * F.fog_color[0] + ambient_color.x), wgt * (ao * fog * F.fog_color[1] + ambient_color.y),
                          wgt * (ao * fog * F.fog_color[2] + ambient_color.z));
            if (reflect_on) {                                                     // :492-498
                emit_refl = true; w_refl = thru * a * reflectivity;
                create_reflection(surface_normal, d, hit_point, refl_o, refl_d);
            }
            if (trans_exists) {
                emit_trans = true; w_trans = thru * kt;
                if ((rflags & RF_ID_OWNER) && approx_equal(alpha, 0.0f)) trans_flags = RF_ID_OWNER;   // :966-969
            }
            shininess = mat.shininess; mat_alpha = mat.alpha; shadow_z_lo = mat.shadow_z_lo;
            recv_shadow = mat.receive_shadow != 0; mat_mc = mat.monte_carlo != 0;
        }

        // --- queue space for both child rays: one reservation per warp, issued before the light loop (see the shadow reservation above)
        uint32_t ch_base = 0;
        const uint32_t m_refl = __ballot_sync(0xffffffffu, emit_refl), m_trans = __ballot_sync(0xffffffffu, emit_trans);
        if (lane == 0 && (m_refl | m_trans) != 0u) ch_base = atomicAdd(out.child_count, __popc(m_refl) + __popc(m_trans));

        // --- lights (:814-920) ---
        uint32_t e_idx = 0;
        for (uint32_t li = 0; li < S.n_lights; li++) {
            const DLight& L = S.lights[li];
            if (!L.enabled) continue;                                             // uniform across the warp
            if (e_idx == 0u && !skip_mode) sh_base = __shfl_sync(0xffffffffu, sh_base, 0);
            bool emit = false; float3 sdir = f3(0, 0, 1), c = f3(0, 0, 0); float len = 3.402823466e+38f;
            if (hit) {
                const float3 lpos = f3(L.pos[0], L.pos[1], L.pos[2]), ldir = f3(L.dir[0], L.dir[1], L.dir[2]);
                const float4 dtl4 = xnormalize_len_s(L.type == 0 ? xneg(ldir) : xsub(lpos, hit_point));   // .w = |lpos - hit_point| for point / spot lights
                const float3 dtl = f3(dtl4.x, dtl4.y, dtl4.z);
                const float dot_light = fmaxf(dot3(surface_normal, dtl), 0.0f);
                const float3 mi = -dtl;
                const float3 reflect_dir = mi - (2.0f * dot3(surface_normal, mi)) * surface_normal;
                const float spec_dot = fmaxf(dot3(reflect_dir, view_dir), 0.0f);
                const float light_power = (spec_dot == 0.0f && shininess > 0.0f) ? 0.0f : powf(spec_dot, shininess);   // powf(+0, y > 0) = +0: skips ~40 instructions for every light behind the mirror direction
                float intensity;
                if (L.type == 0) intensity = L.intensity;
                else {
                    const float r2 = dtl4.w;                                      // the same xnorm(xsub(lpos, hit_point)), computed once
                    intensity = L.intensity / (4.0f * PI * r2);
                    len = r2;
                    if (L.type == 2) {
                        const float dl = dot3(-dtl, norm3(ldir));
                        if (acosf(dl) > L.max_angle) intensity = 0.0f;
                    }
                }
                c = f3((L.color[0] * (specular_color.x * light_power + base_color.x * dot_light)) * intensity,
                       (L.color[1] * (specular_color.y * light_power + base_color.y * dot_light)) * intensity,
                       (L.color[2] * (specular_color.z * light_power + base_color.z * dot_light)) * intensity);
                c = c * coef_d;
                const bool zero = (F.debug_flags & 4u) && c.x == 0.0f && c.y == 0.0f && c.z == 0.0f;
                if (recv_shadow && zero) atomicAdd(out.skipped, 1u);            // opt-in: cannot change the frame
                else if (recv_shadow) {
                    emit = true;
                    sdir = dtl;
                    if (F.monte_carlo && mat_mc) sdir = jitter(sdir, shadow_z_lo, rng, path, 2 + 2 * li);
                } else direct = direct + c;
            }
            const uint32_t slot = skip_mode ? queue_append(out.shadow_count, emit) : sh_base + e_idx * sh_cnt + sh_rank;
            e_idx++;
            if (emit) {
                if (slot < out.shadow_cap) {
                    const float3 so = xadd(hit_point, xscale(surface_normal, 0.001f));
                    out.shadow.o[slot] = make_float4(so.x, so.y, so.z, len);
                    out.shadow.d[slot] = make_float4(sdir.x, sdir.y, sdir.z, __uint_as_float(pixel));
                    out.shadow.c[slot] = make_float4(c.x, c.y, c.z, mat_alpha);
                    out.shadow.r[slot] = h.item;
                } else *out.overflow = 1u;
            }
        }
        if (hit) {
            const float3 acc = direct + constant;
            if (acc.x != 0.0f || acc.y != 0.0f || acc.z != 0.0f || depth_add != 0.0f) atomicAdd(&F.accum_c[pixel], make_float4(acc.x, acc.y, acc.z, depth_add));
        }
        // --- child rays ---
        ch_base = __shfl_sync(0xffffffffu, ch_base, 0);
        {
            const uint32_t slot = ch_base + __popc(m_refl & lt_mask);
            if (emit_refl) {
                if (slot < out.child_cap) {
                    const float3 nd = xnormalize_s(refl_d);                         // :722-723 of the recursive call
                    out.child.o[slot] = make_float4(refl_o.x, refl_o.y, refl_o.z, w_refl);
                    out.child.d[slot] = make_float4(nd.x, nd.y, nd.z, __uint_as_float(pixel));
                    out.child.m[slot] = make_uint2(sample | ((depth + 1) << 16), child_path(path, 0u, depth));
                } else *out.overflow = 1u;
            }
        }
        {
            const uint32_t slot = ch_base + __popc(m_refl) + __popc(m_trans & lt_mask);
            if (emit_trans) {
                if (slot < out.child_cap) {
                    const float3 nd = xnormalize_s(trans_d);
                    out.child.o[slot] = make_float4(trans_o.x, trans_o.y, trans_o.z, w_trans);
                    out.child.d[slot] = make_float4(nd.x, nd.y, nd.z, __uint_as_float(pixel));
                    out.child.m[slot] = make_uint2(sample | ((depth + 1) << 16) | (trans_flags << 24), child_path(path, 1u, depth));
                } else *out.overflow = 1u;
            }
        }
    }
}

// ---- K6 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resolve_kernel(FrameDev F, const uint32_t* __restrict__ pixel_list, uint32_t n_pix,
                                                      uchar4* __restrict__ rgba, float* __restrict__ normals, float* __restrict__ depth,
                                                      uint32_t* __restrict__ object_ids) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(pixel_list + i);
        const float4 c = F.accum_c[p], nn = F.accum_n[p];
        const float ns = (float)F.n_samples;
        float r = fminf(xd(c.x, ns), 1.0f), g = fminf(xd(c.y, ns), 1.0f), b = fminf(xd(c.z, ns), 1.0f);   // :406-413
        uchar4 px;
        if (F.gamma) {
            const float ge = 1.0f / 2.2f;
            px = make_uchar4((unsigned char)as_u8(powf(r, ge) * 255.0f), (unsigned char)as_u8(powf(g, ge) * 255.0f),
                             (unsigned char)as_u8(powf(b, ge) * 255.0f), 255);
        } else px = make_uchar4((unsigned char)as_u8(xm(r, 255.0f)), (unsigned char)as_u8(xm(g, 255.0f)), (unsigned char)as_u8(xm(b, 255.0f)), 255);
        if (rgba) rgba[p] = px;
        if (normals) {
            const float3 n = xnormalize(f3(xd(nn.x, ns), xd(nn.y, ns), xd(nn.z, ns)));       // NaN on a miss, like the reference
            normals[3 * (size_t)p] = n.x; normals[3 * (size_t)p + 1] = n.y; normals[3 * (size_t)p + 2] = n.z;
        }
        if (depth) depth[p] = xd(c.w, ns);
        if (object_ids) object_ids[p] = F.ids[p];
    }
}

__global__ void clear_pixels_kernel(FrameDev F, const uint32_t* __restrict__ pixel_list, uint32_t n_pix) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(pixel_list + i);
        F.accum_c[p] = make_float4(0, 0, 0, 0); F.accum_n[p] = make_float4(0, 0, 0, 0); F.ids[p] = 0u;
    }
}

// ---- probe ---------------------------------------------------------------------------------------
struct ProbeRay { float o[3], d[3]; };
struct ProbeHit { float t; float n[3]; uint32_t item_id, face_id; int32_t item_index; uint32_t reserved; };

__global__ void probe_kernel(SceneDev S, const ProbeRay* __restrict__ rays, uint32_t n, int for_shadow, int stop_first, uint32_t depth,
                             ProbeHit* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 o = f3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d = f3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
    Best b; TravStats st{0, 0}; uint32_t a = 0, c = 0;
    if (stop_first && for_shadow) trace_shadow_ordered<false>(S, o, d, depth, b, st);
    else if (stop_first) {
        // stop_on_first_hit without for_shadow is never issued by the reference; same ordered walk, solid rules of a camera ray
        trace_closest<false>(S, o, d, false, depth, b, st, a, c);
    } else trace_closest<false>(S, o, d, for_shadow != 0, depth, b, st, a, c);
    ProbeHit h; h.reserved = 0;
    if (b.item == 0xFFFFFFFFu) { h.t = -1.0f; h.n[0] = h.n[1] = h.n[2] = 0.0f; h.item_id = 0; h.face_id = 0; h.item_index = -1; }
    else {
        const DItem& it = S.items[b.item];
        uint32_t face_id;
        const float3 nn = hit_normal(S, it, o, d, b.t, b.prim, b.flags, face_id);
        h.t = b.t; h.n[0] = nn.x; h.n[1] = nn.y; h.n[2] = nn.z; h.item_id = it.id; h.face_id = face_id; h.item_index = (int32_t)b.item;
    }
    out[i] = h;
}

// probe through the PRODUCTION closest-hit kernel: pack probe rays into a ray queue, run closest_kernel, convert
__global__ void probe_pack_kernel(const ProbeRay* __restrict__ rays, uint32_t n, uint32_t depth, RayQ q) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q.o[i] = make_float4(rays[i].o[0], rays[i].o[1], rays[i].o[2], 1.0f);
    q.d[i] = make_float4(rays[i].d[0], rays[i].d[1], rays[i].d[2], 0.0f);
    q.m[i] = make_uint2((depth & 0xffu) << 16, 1u);
}
__global__ void probe_unpack_kernel(SceneDev S, const ProbeRay* __restrict__ rays, const HitRec* __restrict__ hits, uint32_t n, ProbeHit* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const HitRec b = hits[i];
    ProbeHit h; h.reserved = 0;
    if (b.item == 0xFFFFFFFFu) { h.t = -1.0f; h.n[0] = h.n[1] = h.n[2] = 0.0f; h.item_id = 0; h.face_id = 0; h.item_index = -1; }
    else {
        const DItem& it = S.items[b.item];
        uint32_t face_id;
        const float3 nn = hit_normal(S, it, f3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), f3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), b.t, b.prim, b.flags, face_id);
        h.t = b.t; h.n[0] = nn.x; h.n[1] = nn.y; h.n[2] = nn.z; h.item_id = it.id; h.face_id = face_id; h.item_index = (int32_t)b.item;
    }
    out[i] = h;
}

// rtx_shadow_probe: probe rays through the PRODUCTION shadow kernels.  Each ray is a one-pixel "frame" with a unit light
// contribution, so accum[i].x ends up as the attenuation factor the kernels apply (raytracing.rs:885-914): 1 lit,
// 1 - receiver alpha [* occluder alpha texel] occluded.
struct ShadowProbeOut { float k; int32_t lit; int32_t occluder_index; float t; uint32_t face_id; uint32_t reserved[3]; };
__global__ void shadow_probe_pack_kernel(SceneDev S, const ProbeRay* __restrict__ rays, const float* __restrict__ len, const int32_t* __restrict__ recv,
                                         uint32_t n, ShadowQ q, float4* __restrict__ acc, uint32_t* count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = n;
    if (i >= n) return;
    const int32_t r = recv ? recv[i] : -1;
    const float alpha = r >= 0 ? S.mats[S.items[r].material].alpha : 1.0f;
    q.o[i] = make_float4(rays[i].o[0], rays[i].o[1], rays[i].o[2], len ? len[i] : 3.402823466e+38f);
    q.d[i] = make_float4(rays[i].d[0], rays[i].d[1], rays[i].d[2], __uint_as_float(i));
    q.c[i] = make_float4(1.0f, 1.0f, 1.0f, alpha);
    q.r[i] = r >= 0 ? (uint32_t)r : 0u;
    q.probe[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u);
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void shadow_probe_unpack_kernel(const float4* __restrict__ acc, const uint4* __restrict__ probe, uint32_t n, ShadowProbeOut* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 p = probe[i];
    ShadowProbeOut o; o.k = acc[i].x; o.lit = (int32_t)p.x; o.occluder_index = p.x == 0u ? (int32_t)p.y : -1;
    o.t = p.z == 0xFFFFFFFFu ? -1.0f : __uint_as_float(p.z); o.face_id = p.w; o.reserved[0] = o.reserved[1] = o.reserved[2] = 0;
    out[i] = o;
}

// debug (RTX_VERIFY=1): recompute every closest hit of a wave with the simple per-thread traversal and record mismatches
struct VerifyRec { float o[3], d[3]; uint32_t depth, index; float t_prod; uint32_t item_prod, prim_prod; float t_ref; uint32_t item_ref, prim_ref; };
__global__ void verify_closest_kernel(SceneDev S, RayQ q, uint32_t q_base, uint32_t n, const HitRec* __restrict__ hits, uint32_t* count, VerifyRec* recs, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 ro = q.o[q_base + i], rd = q.d[q_base + i];
    const uint32_t depth = (q.m[q_base + i].x >> 16) & 0xffu;
    Best b; TravStats st{0, 0}; uint32_t a = 0, c = 0;
    trace_closest<false>(S, f3(ro.x, ro.y, ro.z), f3(rd.x, rd.y, rd.z), false, depth, b, st, a, c);
    const HitRec h = hits[i];
    const bool same = (h.item == b.item) && (b.item == 0xFFFFFFFFu || (h.t == b.t && h.prim == b.prim));
    if (!same) {
        const uint32_t k = atomicAdd(count, 1u);
        if (k < cap) {
            VerifyRec r; r.o[0] = ro.x; r.o[1] = ro.y; r.o[2] = ro.z; r.d[0] = rd.x; r.d[1] = rd.y; r.d[2] = rd.z; r.depth = depth; r.index = i;
            r.t_prod = h.t; r.item_prod = h.item; r.prim_prod = h.prim; r.t_ref = b.t; r.item_ref = b.item; r.prim_ref = b.prim;
            recs[k] = r;
        }
    }
}

// ---- shard pack / unpack ---------------------------------------------------------------------------
__global__ void pack_kernel(const uint32_t* __restrict__ pixel_list, uint32_t n, const uchar4* rgba, const float* normals, const float* depth,
                            const uint32_t* ids, uchar4* p_rgba, float* p_normals, float* p_depth, uint32_t* p_ids) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(pixel_list + i);
        p_rgba[i] = rgba[p];
        p_normals[3 * (size_t)i] = normals[3 * (size_t)p]; p_normals[3 * (size_t)i + 1] = normals[3 * (size_t)p + 1]; p_normals[3 * (size_t)i + 2] = normals[3 * (size_t)p + 2];
        p_depth[i] = depth[p]; p_ids[i] = ids[p];
    }
}
__global__ void unpack_kernel(const uint32_t* __restrict__ pixel_list, uint32_t n, const uchar4* p_rgba, const float* p_normals, const float* p_depth,
                              const uint32_t* p_ids, uchar4* rgba, float* normals, float* depth, uint32_t* ids) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(pixel_list + i);
        rgba[p] = p_rgba[i];
        normals[3 * (size_t)p] = p_normals[3 * (size_t)i]; normals[3 * (size_t)p + 1] = p_normals[3 * (size_t)i + 1]; normals[3 * (size_t)p + 2] = p_normals[3 * (size_t)i + 2];
        depth[p] = p_depth[i]; ids[p] = p_ids[i];
    }
}

// all ranks' packed shards (rank r at packed + r * stride, layout of pack_kernel) -> frame buffers, one launch
__global__ void unpack_all_kernel(const uint2* __restrict__ list /* (pixel, rank) */, const uint32_t* __restrict__ start /* per rank: first entry, then count */,
                                  uint32_t world, uint32_t n, const uint8_t* __restrict__ packed, size_t stride,
                                  uchar4* rgba, float* normals, float* depth, uint32_t* ids) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint2 e = __ldg(list + j);
        const uint32_t p = e.x, r = e.y, i = j - __ldg(start + r), nr = __ldg(start + world + r);
        const uint8_t* base = packed + (size_t)r * stride;
        rgba[p] = ((const uchar4*)base)[i];
        const float* pn = (const float*)(base + (size_t)nr * 4);
        normals[3 * (size_t)p] = pn[3 * (size_t)i]; normals[3 * (size_t)p + 1] = pn[3 * (size_t)i + 1]; normals[3 * (size_t)p + 2] = pn[3 * (size_t)i + 2];
        depth[p] = ((const float*)(base + (size_t)nr * 16))[i];
        ids[p] = ((const uint32_t*)(base + (size_t)nr * 20))[i];
    }
}

}  // namespace rtx
