// rtx_api.cu — C ABI (include/rtx.h) of the B200 ray-casting path: scene flattening + BVH upload,
// the host side of the wavefront loop, the probe hook and the shard helpers.
//
// Wavefront scheduling.  Rays live in per-depth queues (depth 1 .. max_recursion+1), each with room
// for 3*CHUNK rays.  One "wave" pops <= CHUNK rays from the tail of one queue and runs
//   closest_kernel -> shade_kernel (appends <= 2 children per ray to the next queue and one shadow
//   ray per enabled light) -> shadow_kernel
// then reads the two appended counts back (one 16-byte D2H).  The queue to pop is the deepest one
// holding >= CHUNK rays, else a fresh batch of primary rays, else the shallowest non-empty queue;
// that keeps every queue below 3*CHUNK without ever dropping or re-queueing a ray, and keeps waves
// fat (a deep queue is only drained early when it is full).
//
// Sync-free frames.  When all primary rays of the (sharded) frame fit one wave the host never reads a
// counter back: level d+1's ray count is the device counter level d's shade kernel appended to, every
// kernel takes its count from device memory, and the whole frame is enqueued in one go (one stream
// sync at the end).  A queue that would overflow raises a device flag and the frame is redone with the
// synchronised schedule.  This is what a launch-bound frame (config 1) and a strong-scaled shard need.
//
// Multi-GPU in ONE process (rtx_scene_create_multi): the scene is replicated, every device renders the
// interleaved tiles it owns on its own host thread, and its resolve kernel stores the finished pixels
// straight into the frame buffers on the first device over NVLink peer memory — resolve and gather are
// one kernel, there is no separate collective.  Across processes the same stores go through CUDA IPC
// (rtx_gbuffer_*).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <condition_variable>
#include <mutex>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/rtx.h"
#include "bvh_build.h"
#include "rtx_kernels.cuh"
#include "lbvh_build.cuh"

using namespace rtx;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            g_err = std::string(#expr) + ": " + cudaGetErrorString(e__);                           \
            return RTX_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

constexpr uint32_t kCtrPool = 1u << 17;   // counters (8 per wave) zeroed in one memset
// deepest BLAS the traversal stacks hold: kStack - 3 levels for the one-thread walks, 2 * depth + 2 * 6 (TLAS) + 4 <= kLaneStack for the lanes
constexpr int kBlasDepthLimit = (kStack - 3) < (kLaneStack - 16) / 2 ? (kStack - 3) : (kLaneStack - 16) / 2;

inline bool approx_equal_h(float a, float b) { return std::trunc(a * 1000000.0f) == std::trunc(b * 1000000.0f); }

template <typename T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    int alloc(size_t count) {
        if (count <= n && p) return RTX_OK;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        if (count == 0) return RTX_OK;
        CU(cudaMalloc(&p, count * sizeof(T)));
        n = count;
        return RTX_OK;
    }
    int upload(const std::vector<T>& v, cudaStream_t s = 0) {
        int rc = alloc(std::max<size_t>(v.size(), 1));
        if (rc) return rc;
        if (!v.empty()) CU(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
        return RTX_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    // device-to-device copy of another device's buffer (NVLink peer copy); the current device must be dst_dev
    int clone_from(const DevBuf<T>& o, int src_dev, int dst_dev) {
        int rc = alloc(std::max<size_t>(o.n, 1));
        if (rc) return rc;
        if (o.n) CU(cudaMemcpyPeer(p, dst_dev, o.p, src_dev, o.n * sizeof(T)));
        return RTX_OK;
    }
};

struct PixelList { DevBuf<uint32_t> d; uint32_t n = 0; };

}  // namespace

struct RtxScene {
    int device = 0; int sm_count = 148;
    // host mirrors needed for updates
    std::vector<RtxItem> src_items; std::vector<DItem> h_items; std::vector<RtxMaterial> src_mats;
    std::vector<uint32_t> mesh_root, mesh_tri_base; std::vector<RtxMesh> mesh_meta;
    uint32_t n_blas_nodes = 0, n_tris = 0, tlas_cap = 0, n_tlas_nodes = 0;
    size_t texture_bytes = 0; float build_ms = 0.f, device_build_ms = 0.f; uint32_t flags = 0;
    // device scene
    DevBuf<float4> nodes, tris; DevBuf<DItem> items; DevBuf<uint32_t> tlas_prims, fast_prims;
    uint32_t group_root = 0xFFFFFFFFu, n_fast_nodes = 0, n_group_tris = 0; std::vector<uint32_t> group_items;   // merged world-space BLAS
    DevBuf<uint4> s_beyond;
    DevBuf<float> verts, uvs, nrms; DevBuf<uint32_t> idx, uv_idx, n_idx;
    DevBuf<DMaterial> mats; DevBuf<DTex> texs; DevBuf<uchar4> texels; DevBuf<DLight> lights;
    SceneDev dev{};
    // frame state
    DevBuf<float4> accum_c, accum_n; DevBuf<uint32_t> ids; size_t frame_pixels = 0;
    std::map<std::tuple<uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t>, PixelList*> pixel_lists;
    DevBuf<ushort2> sample_table; uint32_t table_samples = 0, cell_size = 1;
    // queues
    uint32_t chunk = 0, levels = 0, level_cap = 0, shadow_cap = 0;
    DevBuf<float4> q_o, q_d; DevBuf<uint2> q_m;
    DevBuf<float4> s_o, s_d, s_c; DevBuf<uint32_t> s_r, s_slow;
    DevBuf<HitRec> hits; DevBuf<uint32_t> ctr_pool; DevBuf<uint32_t> overflow; DevBuf<Counters> counters;
    cudaStream_t shadow_stream = nullptr; cudaEvent_t ev_shaded = nullptr, ev_shadow_done[2] = {nullptr, nullptr};   // shadow kernels overlap the next wave
    uint32_t* h_pool = nullptr;             // pinned copy of the counter pool (per-wave exact / beyond counts are summed at the end of the frame)
    uint32_t* h_ctr = nullptr;              // pinned: 4 uint32 per-wave read-back + 8 * 66 level counters of a sync-free frame
    uint32_t wave_cap = 0;                  // rays one wave may hold (hit records, shadow queue / lights)
    uint32_t n_enabled_lights = 0;          // host mirror (rtx_scene_create / rtx_scene_set_lights)
    bool no_sync_free = false;              // a sync-free frame overflowed a queue once: keep the synchronised schedule
    std::mutex api_mu;                      // serialises frames / updates on one handle
    // rtx_trace_probe / rtx_shadow_probe: own stream, queues and counters, so a pick never touches a frame in flight
    struct Probe {
        cudaStream_t st = nullptr; uint32_t cap = 0; std::mutex mu;
        DevBuf<float4> q_o, q_d, s_c, acc; DevBuf<uint2> q_m; DevBuf<uint32_t> s_r, slow, ctr; DevBuf<HitRec> hits; DevBuf<uint4> out;
        DevBuf<float> len; DevBuf<int32_t> recv; DevBuf<ShadowProbeOut> res; DevBuf<uint4> beyond;
    } probe;
    // single-process multi-GPU: replicas[k] renders shard k+1 of `1 + replicas.size()` (this handle renders shard 0)
    std::vector<RtxScene*> replicas; RtxScene* primary = nullptr; cudaStream_t own_stream = nullptr; cudaEvent_t fence = nullptr;
    std::vector<cudaEvent_t> events;
    int blocks_closest = 0, blocks_closest_st = 0, blocks_shadow = 0, blocks_shadow_st = 0;
    // host-output staging for rtx_render_frame
    DevBuf<uchar4> o_rgba; DevBuf<float> o_normals, o_depth; DevBuf<uint32_t> o_ids;
    // probe staging
    DevBuf<ProbeRay> p_rays; DevBuf<ProbeHit> p_hits;
    // async frame (RendererManager::start / stop / is_done)
    std::thread worker; std::atomic<bool> running{false}, cancel{false}; std::atomic<uint64_t> samples_issued{0};
    uint64_t async_pixels = 0; uint32_t async_samples = 1; int async_result = RTX_OK; RtxStats async_stats{}; std::string async_err;
    // progressive preview of the frame in flight (rtx_render_snapshot): host targets of the async frame + request/ack
    uint8_t* snap_rgba = nullptr; float* snap_normals = nullptr; float* snap_depth = nullptr; uint32_t* snap_ids = nullptr;
    std::atomic<uint32_t> snap_req{0}; std::mutex snap_mu; std::condition_variable snap_cv; uint64_t snap_seq = 0;
};

namespace {

// ---- scene flattening -----------------------------------------------------------------------------
void fill_item(const RtxScene& sc, const RtxItem& s, DItem& d) {
    memset(&d, 0, sizeof(d));
    for (int r = 0; r < 3; r++) {
        d.inv[r] = make_float4(s.tran_inverse[0 + r], s.tran_inverse[4 + r], s.tran_inverse[8 + r], s.tran_inverse[12 + r]);
        d.mat[r] = make_float4(s.trans[0 + r], s.trans[4 + r], s.trans[8 + r], s.trans[12 + r]);
    }
    const RtxMaterial& m = sc.src_mats[s.material];
    // material cache = Material::new + apply_diff_without_textures (reference src/shape/mod.rs:182-246,769-772)
    const float c_alpha = approx_equal_h(1.0f, m.alpha) ? 1.0f : m.alpha;
    uint32_t f = 0;
    if (s.shape == RTX_SHAPE_MESH) f |= IF_MESH;
    if (s.visible) f |= IF_VISIBLE;
    if (s.flip_normals) f |= IF_FLIP;
    if (m.cast_shadow) f |= IF_CAST_SHADOW;
    if (m.reflection_only) f |= IF_REFL_ONLY;
    if (m.backface_cullig) f |= IF_BACKFACE;
    if (m.smooth_shading) f |= IF_SMOOTH;
    if (c_alpha > 0.0f) f |= IF_ALPHA_POS;
    if (c_alpha < 1.0f) f |= IF_ALPHA_LT1;
    if (m.texture[RTX_TEX_ALPHA] >= 0) f |= IF_ALPHA_TEX;
    {   // identity + translation inverse (planes, spheres): the per-item ray transform degenerates to one add per axis
        const float* t = s.tran_inverse;
        if (t[0] == 1.f && t[5] == 1.f && t[10] == 1.f && t[15] == 1.f && t[1] == 0.f && t[2] == 0.f && t[4] == 0.f && t[6] == 0.f && t[8] == 0.f && t[9] == 0.f)
            f |= IF_TRANSLATION;
    }
    d.inv_w = s.tran_inverse[15];
    if (d.inv_w != 1.0f) f |= IF_DIV_W;
    d.id = s.id; d.material = (uint32_t)s.material;
    d.lo.w = s.radius; d.hi.w = c_alpha;
    d.flags = f;
}

void item_world_box(const DItem& d, Aabb3& b) {
    const float inf = INFINITY;
    b = {{inf, inf, inf}, {-inf, -inf, -inf}};
    const float lo[3] = {d.lo.x, d.lo.y, d.lo.z}, hi[3] = {d.hi.x, d.hi.y, d.hi.z};
    for (int c = 0; c < 8; c++) {
        const float p[3] = {(c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2]};
        const float4 rows[3] = {d.mat[0], d.mat[1], d.mat[2]};
        for (int r = 0; r < 3; r++) {
            float v = rows[r].x * p[0] + rows[r].y * p[1] + rows[r].z * p[2] + rows[r].w;
            float pad = std::fabs(v) * 1e-6f + 1e-30f;                 // the TLAS is only a conservative filter
            b.lo[r] = std::min(b.lo[r], v - pad); b.hi[r] = std::max(b.hi[r], v + pad);
        }
    }
}

void append_nodes(std::vector<float4>& out, const WideBvh& bvh, uint32_t node_off, uint32_t prim_off) {
    for (const WideNode& wn : bvh.nodes) {
        WideNode n = wn;
        n.child_base += node_off; n.prim_base += prim_off;
        float4 f[5]; memcpy(f, &n, 80);
        for (int k = 0; k < 5; k++) out.push_back(f[k]);
    }
}

int build_tlas(RtxScene& sc, std::vector<float4>& tlas_nodes, std::vector<uint32_t>& tlas_prims) {
    std::vector<Aabb3> boxes(sc.h_items.size());
    for (size_t i = 0; i < sc.h_items.size(); i++) {
        item_world_box(sc.h_items[i], boxes[i]);
        sc.h_items[i].wlo = make_float4(boxes[i].lo[0], boxes[i].lo[1], boxes[i].lo[2], 0.f);
        sc.h_items[i].whi = make_float4(boxes[i].hi[0], boxes[i].hi[1], boxes[i].hi[2], 0.f);
    }
    WideBvh bvh; build_wide_bvh(boxes.data(), (uint32_t)boxes.size(), bvh, 6);
    if (bvh.max_depth > 6) return fail(RTX_E_INVALID, "TLAS too deep for the traversal stack");
    tlas_nodes.clear();
    append_nodes(tlas_nodes, bvh, sc.n_blas_nodes, 0);
    tlas_prims = bvh.prim_order;
    sc.n_tlas_nodes = (uint32_t)bvh.nodes.size();
    return RTX_OK;
}

// Fast TLAS of the persistent kernels: every item that is not in the merged BLAS, plus ONE entry (kGroupPrim) for the merged BLAS.
// Call after build_tlas (which refreshes the items' world boxes).
int build_tlas_fast(RtxScene& sc, std::vector<float4>& nodes, std::vector<uint32_t>& prims) {
    std::vector<Aabb3> boxes; std::vector<uint32_t> ids;
    const float inf = INFINITY;
    Aabb3 gb = {{inf, inf, inf}, {-inf, -inf, -inf}}; bool have_group = false;
    for (size_t i = 0; i < sc.h_items.size(); i++) {
        const DItem& it = sc.h_items[i];
        const Aabb3 b = {{it.wlo.x, it.wlo.y, it.wlo.z}, {it.whi.x, it.whi.y, it.whi.z}};
        if (sc.group_root != 0xFFFFFFFFu && (it.flags & IF_GROUPED)) {
            for (int k = 0; k < 3; k++) { gb.lo[k] = std::min(gb.lo[k], b.lo[k]); gb.hi[k] = std::max(gb.hi[k], b.hi[k]); }
            have_group = true;
        } else { boxes.push_back(b); ids.push_back((uint32_t)i); }
    }
    if (have_group) { boxes.push_back(gb); ids.push_back(kGroupPrim); }
    WideBvh bvh; build_wide_bvh(boxes.data(), (uint32_t)boxes.size(), bvh, 6);
    if (bvh.max_depth > 6) return fail(RTX_E_INVALID, "TLAS too deep for the traversal stack");
    nodes.clear();
    append_nodes(nodes, bvh, sc.n_blas_nodes + sc.tlas_cap, 0);
    prims.resize(bvh.prim_order.size());
    for (size_t k = 0; k < prims.size(); k++) prims[k] = ids[bvh.prim_order[k]];
    sc.n_fast_nodes = (uint32_t)bvh.nodes.size();
    return RTX_OK;
}

inline bool is_identity16(const float* m) {
    for (int i = 0; i < 16; i++) if (m[i] != ((i % 5 == 0) ? 1.0f : 0.0f)) return false;
    return true;
}

void fill_lights(const RtxLight* l, uint32_t n, std::vector<DLight>& out) {
    out.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        DLight d; memset(&d, 0, sizeof(d));
        memcpy(d.pos, l[i].pos, 12); memcpy(d.dir, l[i].dir, 12); memcpy(d.color, l[i].color, 12);
        d.type = l[i].light_type; d.intensity = l[i].intensity; d.max_angle = l[i].max_angle; d.enabled = l[i].enabled;
        out[i] = d;
    }
}

void refresh_dev(RtxScene& sc) {
    SceneDev& D = sc.dev;
    D.nodes = sc.nodes.p; D.tris = sc.tris.p; D.items = sc.items.p; D.tlas_prims = sc.tlas_prims.p;
    D.verts = sc.verts.p; D.idx = sc.idx.p; D.uvs = sc.uvs.p; D.uv_idx = sc.uv_idx.p; D.nrms = sc.nrms.p; D.n_idx = sc.n_idx.p;
    D.mats = sc.mats.p; D.texs = sc.texs.p; D.texels = sc.texels.p; D.lights = sc.lights.p;
    D.n_items = (uint32_t)sc.h_items.size();
    D.tlas_root = sc.n_blas_nodes; D.use_tlas = 1u;
    D.fast_prims = sc.fast_prims.p; D.fast_root = sc.n_blas_nodes + sc.tlas_cap; D.group_root = sc.group_root;
    D.ball_flip_inside = 1u;
    if (!sc.overflow.p) sc.overflow.alloc(4);
    D.dbg = sc.overflow.p + 1;
    D.any_alpha_tex = 0u;
    {
        // few items: walk the item list directly (each item pre-culled by its padded world box) instead of a TLAS node test
        uint32_t lim = 0; if (const char* e = getenv("RTX_FLAT_ITEMS")) lim = (uint32_t)atoi(e);   // off: measured slower than the one-node TLAS
        D.flat_items = (!sc.h_items.empty() && sc.h_items.size() <= std::min(lim, 24u)) ? ((1u << sc.h_items.size()) - 1u) : 0u;
    }
    for (const DItem& it : sc.h_items) if (it.flags & IF_ALPHA_TEX) D.any_alpha_tex = 1u;
}

// Tile ownership: tiles are taken in row-major groups of `world`; inside group g the tile at position j belongs to rank
// (j + rot(g)) % world.  Every rank owns exactly one tile per group (balanced counts) and the per-group rotation keeps a
// rank from always getting the same image columns (plain t % world gave 8 % time imbalance on config 2 at 8 GPUs).
inline uint32_t shard_rot(uint32_t g) { return (g * 0x9E3779B1u) >> 16; }
inline uint32_t shard_tile(uint32_t rank, uint32_t world, uint32_t g) { return g * world + (rank + world - shard_rot(g) % world) % world; }
template <class F> inline void for_each_owned_tile(const RtxShard& sh, uint32_t n_tiles, F f) {
    for (uint32_t g = 0; g * sh.world < n_tiles; g++) { const uint32_t t = shard_tile(sh.rank, sh.world, g); if (t < n_tiles) f(t); }
}

int get_pixel_list(RtxScene& sc, uint32_t w, uint32_t h, const RtxShard* shard, PixelList** out) {
    RtxShard sh = shard ? *shard : RtxShard{0, 1, 8, 4};
    if (sh.world == 0 || sh.rank >= sh.world || sh.tile_w == 0 || sh.tile_h == 0) return fail(RTX_E_INVALID, "bad shard");
    auto key = std::make_tuple(w, h, sh.rank, sh.world, sh.tile_w, sh.tile_h);
    auto it = sc.pixel_lists.find(key);
    if (it != sc.pixel_lists.end()) { *out = it->second; return RTX_OK; }
    std::vector<uint32_t> px;
    const uint32_t tx = (w + sh.tile_w - 1) / sh.tile_w, ty = (h + sh.tile_h - 1) / sh.tile_h;
    for_each_owned_tile(sh, tx * ty, [&](uint32_t t) {
        const uint32_t x0 = (t % tx) * sh.tile_w, y0 = (t / tx) * sh.tile_h;
        for (uint32_t y = y0; y < std::min(y0 + sh.tile_h, h); y++)
            for (uint32_t x = x0; x < std::min(x0 + sh.tile_w, w); x++) px.push_back(y * w + x);
    });
    PixelList* pl = new PixelList();
    pl->n = (uint32_t)px.size();
    int rc = pl->d.upload(px);
    if (rc) { delete pl; return rc; }
    // the list is read by kernels on the caller's stream (possibly cudaStreamNonBlocking): finish the upload first
    if (cudaStreamSynchronize(0) != cudaSuccess) { delete pl; return fail(RTX_E_CUDA, "pixel list upload failed"); }
    sc.pixel_lists[key] = pl;
    *out = pl;
    return RTX_OK;
}

// rand 0.8 StdRng (ChaCha12) + seed_from_u64 + SliceRandom::shuffle — reference src/raytracing.rs:300-313
struct StdRng12 {
    uint32_t key[8]; uint64_t counter = 0; uint32_t buf[16]; int idx = 16;
    static uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    explicit StdRng12(uint64_t state) {
        for (int i = 0; i < 8; i++) {                                 // PCG32 seed expansion
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27), rot = (uint32_t)(state >> 59);
            key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
        }
    }
    void block() {
        uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                          (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
        uint32_t x[16]; memcpy(x, s, 64);
        auto qr = [&](int a, int b, int c, int d) {
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
        };
        for (int i = 0; i < 6; i++) { qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15); qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14); }
        for (int i = 0; i < 16; i++) buf[i] = x[i] + s[i];
        counter++; idx = 0;
    }
    uint32_t next() { if (idx >= 16) block(); return buf[idx++]; }
    uint32_t below(uint32_t range) {                                   // UniformInt<u32>::sample_single
        uint32_t zone = (range << __builtin_clz(range)) - 1u;
        for (;;) { uint64_t m = (uint64_t)next() * range; if ((uint32_t)m <= zone) return (uint32_t)(m >> 32); }
    }
};
void make_sample_table(uint32_t samples, uint32_t& cell, std::vector<ushort2>& out) {
    cell = 1;
    if (samples > 1) { uint32_t v = (samples + 2) & 0xFFFFu, p = 1; while (p < v) p <<= 1; cell = p / 2; }
    std::vector<ushort2> s; s.reserve((size_t)cell * cell);
    for (uint32_t x = 0; x < cell; x++) for (uint32_t y = 0; y < cell; y++) s.push_back(make_ushort2((unsigned short)x, (unsigned short)y));
    StdRng12 rng(0);
    for (size_t i = s.size() - 1; i >= 1; i--) std::swap(s[i], s[rng.below((uint32_t)(i + 1))]);
    if (s.size() > samples) s.resize(samples);
    out = s;
}

int ensure_frame(RtxScene& sc, uint32_t w, uint32_t h) {
    size_t n = (size_t)w * h;
    int rc;
    if ((rc = sc.accum_c.alloc(n)) || (rc = sc.accum_n.alloc(n)) || (rc = sc.ids.alloc(n))) return rc;
    sc.frame_pixels = n;
    return RTX_OK;
}

// Queue sizing for ONE frame.  chunk = rays per wave: 8 Mi, less for a frame that has fewer primary rays (a 800x600x1
// frame does not allocate gigabytes) and less when max_recursion is so high that levels * 3 * chunk * 40 B would pass 12 GB.
// Everything is derived from THIS frame's (levels, chunk); the buffers only ever grow.
int ensure_queues(RtxScene& sc, uint32_t max_recursion, uint32_t n_lights_enabled, uint64_t n_primary) {
    const uint32_t levels = max_recursion + 1;
    if (levels > 64) return fail(RTX_E_INVALID, "max_recursion > 63 is not supported by the wavefront queues");
    uint32_t chunk = 1u << 23;                    // rays per wave: bigger waves = fewer kernel tails and host round trips
    if (const char* e = getenv("RTX_CHUNK")) { long v = atol(e); if (v >= 1024) chunk = (uint32_t)v; }
    while (chunk > (1u << 16) && (uint64_t)(chunk >> 1) >= n_primary) chunk >>= 1;
    while (chunk > (1u << 17) && (size_t)levels * 3 * chunk * 40 > (size_t)12 << 30) chunk >>= 1;   // <= 12 GB of ray queues
    sc.chunk = chunk; sc.levels = levels; sc.level_cap = 3 * chunk; sc.wave_cap = 2 * chunk;
    sc.shadow_cap = sc.wave_cap * std::max(1u, n_lights_enabled);
    const size_t qn = (size_t)levels * sc.level_cap;
    int rc;
    if ((rc = sc.q_o.alloc(qn)) || (rc = sc.q_d.alloc(qn)) || (rc = sc.q_m.alloc(qn))) return rc;
    // the shadow queue is double-buffered: wave k+1's shade kernel fills one half while wave k's shadow kernels still drain the other
    const size_t sq = 2 * (size_t)sc.shadow_cap;
    if ((rc = sc.s_o.alloc(sq)) || (rc = sc.s_d.alloc(sq)) || (rc = sc.s_c.alloc(sq)) || (rc = sc.s_r.alloc(sq)) || (rc = sc.s_slow.alloc(sc.shadow_cap))) return rc;
    if (!sc.shadow_stream) {
        CU(cudaStreamCreateWithFlags(&sc.shadow_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&sc.ev_shaded, cudaEventDisableTiming));
        for (int k = 0; k < 2; k++) CU(cudaEventCreateWithFlags(&sc.ev_shadow_done[k], cudaEventDisableTiming));
        CU(cudaMallocHost(&sc.h_pool, (size_t)kCtrPool * 4));
    }
    if (sc.group_root != 0xFFFFFFFFu && (rc = sc.s_beyond.alloc(sc.shadow_cap))) return rc;
    if ((rc = sc.hits.alloc(sc.wave_cap)) || (rc = sc.ctr_pool.alloc(kCtrPool)) || (rc = sc.overflow.alloc(4)) || (rc = sc.counters.alloc(1))) return rc;
    if (!sc.h_ctr) CU(cudaMallocHost(&sc.h_ctr, (4 + 8 * 66 + 8) * sizeof(uint32_t)));   // [0..3] overflow flags | [4..531] level counters | [532..539] one wave
    return RTX_OK;
}

int occupancy_blocks(RtxScene& sc) {
    if (sc.blocks_closest) return RTX_OK;
    int b = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, closest_kernel<false>, kTraceBlock, 0)); sc.blocks_closest = std::max(1, b) * sc.sm_count;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, closest_kernel<true>, kTraceBlock, 0)); sc.blocks_closest_st = std::max(1, b) * sc.sm_count;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, shadow_any_kernel<false>, kTraceBlock, 0)); sc.blocks_shadow = std::max(1, b) * sc.sm_count;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, shadow_any_kernel<true>, kTraceBlock, 0)); sc.blocks_shadow_st = std::max(1, b) * sc.sm_count;
    return RTX_OK;
}

RayQ level_queue(RtxScene& sc, uint32_t level /*1-based depth*/) {
    size_t off = (size_t)(level - 1) * sc.level_cap;
    return RayQ{sc.q_o.p + off, sc.q_d.p + off, sc.q_m.p + off};
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* rtx_last_error(void) { return g_err.c_str(); }
int rtx_abi_version(void) { return RTX_ABI_VERSION; }
int rtx_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

int rtx_sample_table(uint32_t samples, uint32_t* cell_size, uint16_t* xy) {
    if (samples == 0 || samples > 65535) return fail(RTX_E_INVALID, "samples out of range");
    uint32_t cell; std::vector<ushort2> t; make_sample_table(samples, cell, t);
    if (cell_size) *cell_size = cell;
    if (xy) for (size_t i = 0; i < t.size(); i++) { xy[2 * i] = t[i].x; xy[2 * i + 1] = t[i].y; }
    return RTX_OK;
}

int rtx_scene_create(const RtxSceneDesc* d, int device, RtxScene** out) {
    uint32_t flags = 0;
    if (const char* e = getenv("RTX_DEVICE_BVH")) if (atoi(e) != 0) flags |= RTX_SCENE_DEVICE_BVH;
    return rtx_scene_create_ex(d, device, flags, out);
}

int rtx_scene_create_ex(const RtxSceneDesc* d, int device, uint32_t scene_flags, RtxScene** out) {
    if (!d || !out) return fail(RTX_E_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(RTX_E_NO_DEVICE, "no CUDA device: librtx_b200 has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(RTX_E_INVALID, "bad device ordinal");
    CU(cudaSetDevice(device));
    auto t0 = std::chrono::steady_clock::now();
    const bool trace_build = getenv("RTX_TRACE_BUILD") != nullptr;
    auto phase = [&](const char* what) {
        if (trace_build) fprintf(stderr, "[build] %8.1f ms  %s\n", std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(), what);
    };
    RtxScene* sc = new RtxScene();
    sc->device = device; sc->flags = scene_flags;
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device)); sc->sm_count = prop.multiProcessorCount;
    const bool device_bvh = (scene_flags & RTX_SCENE_DEVICE_BVH) != 0;
    // Morton-order wide BVH built by kernels (lbvh_build.cuh); a tree too deep for the traversal stack falls back to the host builder
    lbvh::Workspace lbvh_ws;
    struct WsGuard { lbvh::Workspace& w; ~WsGuard() { w.release(); } } ws_guard{lbvh_ws};
    auto build_boxes_on_device = [&](const std::vector<Aabb3>& boxes, WideBvh& bvh, bool keep_order = false) -> bool {
        int deep = 0; float ms = 0.f;
        static_assert(sizeof(lbvh::Box) == sizeof(Aabb3), "box layout");
        const cudaError_t e = lbvh::build_on_device(lbvh_ws, reinterpret_cast<const lbvh::Box*>(boxes.data()), (uint32_t)boxes.size(), bvh, kBlasDepthLimit, &deep, &ms, keep_order);
        if (e != cudaSuccess) cudaGetLastError();
        if (e == cudaSuccess && !deep) sc->device_build_ms += ms;
        return e == cudaSuccess && !deep;
    };
    auto bail = [&](int code, const std::string& m) { rtx_scene_destroy(sc); return fail(code, m); };

    // ---- validate + copy ----
    for (uint32_t i = 0; i < d->n_items; i++) {
        const RtxItem& it = d->items[i];
        if (it.material < 0 || (uint32_t)it.material >= d->n_materials) return bail(RTX_E_INVALID, "item material index out of range");
        if (it.shape == RTX_SHAPE_MESH && (it.mesh < 0 || (uint32_t)it.mesh >= d->n_meshes)) return bail(RTX_E_INVALID, "item mesh index out of range");
        if (it.shape != RTX_SHAPE_MESH && it.shape != RTX_SHAPE_SPHERE) return bail(RTX_E_INVALID, "unknown shape kind");
        // Vector3::from_homogeneous(tran_inverse * dir).unwrap() needs w == 0 (src/shape/mod.rs:760)
        if (it.tran_inverse[3] != 0.0f || it.tran_inverse[7] != 0.0f || it.tran_inverse[11] != 0.0f || !(it.tran_inverse[15] > 0.5f && it.tran_inverse[15] < 2.0f))
            return bail(RTX_E_NON_AFFINE, "item transform is not affine");
    }
    for (uint32_t i = 0; i < d->n_materials; i++)
        for (int t = 0; t < RTX_TEX_COUNT; t++)
            if (d->materials[i].texture[t] >= (int32_t)d->n_textures) return bail(RTX_E_INVALID, "texture index out of range");
    sc->src_items.assign(d->items, d->items + d->n_items);
    sc->src_mats.assign(d->materials, d->materials + d->n_materials);

    // ---- meshes: BLAS per mesh ----
    // The shading-side mesh arrays (vertices, indices, uvs, normals and their index lists) go straight from the caller's memory
    // into their place in the device arrays: no host-side concatenation (config 5: 480 MB).
    std::vector<float4> h_nodes, h_tris;
    struct MeshOff { uint32_t vert, idx, uv, uvidx, nrm, nidx; float lo[3], hi[3]; };
    std::vector<MeshOff> moff(d->n_meshes);
    sc->mesh_root.resize(d->n_meshes); sc->mesh_tri_base.resize(d->n_meshes);
    sc->mesh_meta.assign(d->meshes, d->meshes + d->n_meshes);
    {
        uint64_t tv = 0, ti = 0, tu = 0, tui = 0, tn = 0, tni = 0;        // element counts (floats / uint32s)
        for (uint32_t mi = 0; mi < d->n_meshes; mi++) {
            const RtxMesh& m = d->meshes[mi];
            if (m.n_faces == 0) return bail(RTX_E_EMPTY_MESH, "mesh with 0 triangles (parry TriMesh::new panics)");
            if (m.n_normal_faces && m.n_normal_faces < m.n_faces) return bail(RTX_E_INVALID, "normals_indices shorter than indices");
            MeshOff& o = moff[mi];
            o.vert = (uint32_t)tv; o.idx = (uint32_t)ti; o.uv = (uint32_t)tu; o.uvidx = (uint32_t)tui; o.nrm = (uint32_t)tn; o.nidx = (uint32_t)tni;
            tv += 3 * (uint64_t)m.n_vertices; ti += 3 * (uint64_t)m.n_faces; tu += 2 * (uint64_t)m.n_uvs; tui += 3 * (uint64_t)m.n_uv_faces;
            tn += 3 * (uint64_t)m.n_normals; tni += 3 * (uint64_t)m.n_normal_faces;
            if (tv > 0xffffffffull || ti > 0xffffffffull || tu > 0xffffffffull || tui > 0xffffffffull || tn > 0xffffffffull || tni > 0xffffffffull)
                return bail(RTX_E_INVALID, "mesh arrays exceed 2^32 elements");
        }
        int rc0;
        if ((rc0 = sc->verts.alloc(std::max<uint64_t>(tv, 1))) || (rc0 = sc->idx.alloc(std::max<uint64_t>(ti, 1))) || (rc0 = sc->uvs.alloc(std::max<uint64_t>(tu, 1))) ||
            (rc0 = sc->uv_idx.alloc(std::max<uint64_t>(tui, 1))) || (rc0 = sc->nrms.alloc(std::max<uint64_t>(tn, 1))) || (rc0 = sc->n_idx.alloc(std::max<uint64_t>(tni, 1)))) {
            std::string keep = g_err; rtx_scene_destroy(sc); g_err = keep; return rc0;
        }
        for (uint32_t mi = 0; mi < d->n_meshes; mi++) {
            const RtxMesh& m = d->meshes[mi]; const MeshOff& o = moff[mi];
            cudaMemcpyAsync(sc->verts.p + o.vert, m.vertices, 12 * (size_t)m.n_vertices, cudaMemcpyHostToDevice, 0);
            cudaMemcpyAsync(sc->idx.p + o.idx, m.indices, 12 * (size_t)m.n_faces, cudaMemcpyHostToDevice, 0);
            if (m.n_uvs) cudaMemcpyAsync(sc->uvs.p + o.uv, m.uvs, 8 * (size_t)m.n_uvs, cudaMemcpyHostToDevice, 0);
            if (m.n_uv_faces) cudaMemcpyAsync(sc->uv_idx.p + o.uvidx, m.uv_indices, 12 * (size_t)m.n_uv_faces, cudaMemcpyHostToDevice, 0);
            if (m.n_normals) cudaMemcpyAsync(sc->nrms.p + o.nrm, m.normals, 12 * (size_t)m.n_normals, cudaMemcpyHostToDevice, 0);
            if (m.n_normal_faces) cudaMemcpyAsync(sc->n_idx.p + o.nidx, m.normals_indices, 12 * (size_t)m.n_normal_faces, cudaMemcpyHostToDevice, 0);
        }
        if (cudaGetLastError() != cudaSuccess) return bail(RTX_E_CUDA, "mesh array upload failed");
    }
    bool tris_on_device = false;                                          // RTX_SCENE_DEVICE_BVH: sc->tris is filled by kernels, h_tris stays empty
    std::atomic<int> bad_index{0};                                        // 1 vertex, 2 uv, 3 normal index out of range (checked by the worker threads below)
    phase("mesh arrays uploaded");
    // BLAS builds are independent: one host thread per mesh, up to the hardware concurrency (config 5: 64 meshes of 156 k triangles).
    // With RTX_SCENE_DEVICE_BVH the threads only compute the triangle boxes and the trees are built by kernels, mesh after mesh.
    std::vector<WideBvh> bvhs(d->n_meshes);
    {
        std::vector<std::vector<Aabb3>> mesh_boxes(device_bvh ? d->n_meshes : 0);
        std::atomic<uint32_t> next{0};
        auto work = [&]() {
            std::vector<Aabb3> local;
            for (uint32_t mi = next.fetch_add(1); mi < d->n_meshes; mi = next.fetch_add(1)) {
                const RtxMesh& m = d->meshes[mi]; MeshOff& o = moff[mi];
                {
                    bool ok = true;
                    for (size_t k = 0; k < 3 * (size_t)m.n_faces; k++) ok &= m.indices[k] < m.n_vertices;
                    if (!ok) { bad_index.store(1); continue; }
                    for (size_t k = 0; k < 3 * (size_t)m.n_uv_faces; k++) ok &= m.uv_indices[k] < m.n_uvs;
                    if (!ok) { bad_index.store(2); continue; }
                    for (size_t k = 0; k < 3 * (size_t)m.n_normal_faces; k++) ok &= m.normals_indices[k] < m.n_normals;
                    if (!ok) { bad_index.store(3); continue; }
                }
                std::vector<Aabb3>& boxes = device_bvh ? mesh_boxes[mi] : local;
                boxes.resize(m.n_faces);
                for (int k = 0; k < 3; k++) { o.lo[k] = INFINITY; o.hi[k] = -INFINITY; }
                for (uint32_t f = 0; f < m.n_faces; f++) {
                    Aabb3& b = boxes[f];
                    for (int k = 0; k < 3; k++) { b.lo[k] = INFINITY; b.hi[k] = -INFINITY; }
                    for (int c = 0; c < 3; c++) {
                        const float* v = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f + c];
                        for (int k = 0; k < 3; k++) { b.lo[k] = std::min(b.lo[k], v[k]); b.hi[k] = std::max(b.hi[k], v[k]); }
                    }
                    for (int k = 0; k < 3; k++) { o.lo[k] = std::min(o.lo[k], b.lo[k]); o.hi[k] = std::max(o.hi[k], b.hi[k]); }   // TriMesh::aabb
                }
                if (!device_bvh) build_wide_bvh(boxes.data(), m.n_faces, bvhs[mi], kBlasDepthLimit);
            }
        };
        const uint32_t nt = std::min<uint32_t>(d->n_meshes, std::max(1u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (uint32_t t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (std::thread& t : pool) t.join();
        if (bad_index.load()) return bail(RTX_E_INVALID, bad_index.load() == 1 ? "vertex index out of range" : bad_index.load() == 2 ? "uv index out of range" : "normal index out of range");
        phase(device_bvh ? "indices checked, triangle boxes" : "indices checked, triangle boxes + host BLAS builds");
        if (device_bvh) {
            uint32_t max_faces = 0;
            for (uint32_t mi = 0; mi < d->n_meshes; mi++) max_faces = std::max(max_faces, d->meshes[mi].n_faces);
            lbvh_ws.reserve(max_faces);                                   // one allocation for the largest mesh, reused by all
        }
        if (device_bvh) {
            // the triangle array is laid out up front (mesh after mesh, then room for a merged BLAS) and every mesh's records are
            // packed by a kernel from the mesh arrays already on the device: no host packing, no 48 B/triangle upload
            uint64_t total = 0, group_room = 0;
            for (uint32_t mi = 0; mi < d->n_meshes; mi++) total += d->meshes[mi].n_faces;
            for (uint32_t i = 0; i < d->n_items; i++)
                if (d->items[i].shape == RTX_SHAPE_MESH && is_identity16(d->items[i].trans) && is_identity16(d->items[i].tran_inverse)) group_room += d->meshes[d->items[i].mesh].n_faces;
            if (group_room > 4000000) group_room = 0;
            int rc0 = sc->tris.alloc((size_t)(total + group_room) * 3);
            if (rc0) { std::string keep = g_err; rtx_scene_destroy(sc); g_err = keep; return rc0; }
            tris_on_device = true;
            uint64_t toff = 0;
            for (uint32_t mi = 0; mi < d->n_meshes; mi++) {
                const RtxMesh& m = d->meshes[mi]; const MeshOff& o = moff[mi];
                if (build_boxes_on_device(mesh_boxes[mi], bvhs[mi], true)) {
                    lbvh::pack_tris_kernel<<<(m.n_faces + 255) / 256, 256>>>(lbvh_ws.prim_order, m.n_faces, sc->verts.p + o.vert, sc->idx.p + o.idx, sc->tris.p + toff * 3);
                } else {                                                  // Morton tree too deep for the stack: host builder, host packing, one copy
                    build_wide_bvh(mesh_boxes[mi].data(), m.n_faces, bvhs[mi], kBlasDepthLimit);
                    std::vector<float4> t((size_t)m.n_faces * 3);
                    for (size_t k = 0; k < bvhs[mi].prim_order.size(); k++) {
                        const uint32_t f = bvhs[mi].prim_order[k];
                        const float* a = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f];
                        const float* b = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f + 1];
                        const float* c = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f + 2];
                        float fb; memcpy(&fb, &f, 4);
                        t[3 * k] = make_float4(a[0], a[1], a[2], fb); t[3 * k + 1] = make_float4(b[0], b[1], b[2], 0.f); t[3 * k + 2] = make_float4(c[0], c[1], c[2], 0.f);
                    }
                    cudaMemcpy(sc->tris.p + toff * 3, t.data(), t.size() * sizeof(float4), cudaMemcpyHostToDevice);
                }
                toff += m.n_faces;
                std::vector<Aabb3>().swap(mesh_boxes[mi]);
            }
            if (cudaGetLastError() != cudaSuccess) return bail(RTX_E_CUDA, "triangle packing failed");
        }
    }
    if (device_bvh) phase("device BLAS builds");
    uint32_t n_mesh_tris = 0;
    {
        // node / triangle offsets per mesh, then the meshes are packed side by side by a pool of threads (30 M float4 for config 5)
        std::vector<uint32_t> node_off(d->n_meshes + 1, 0), tri_off(d->n_meshes + 1, 0);
        for (uint32_t mi = 0; mi < d->n_meshes; mi++) {
            const WideBvh& bvh = bvhs[mi];
            if (bvh.max_depth >= kStack - 2 || 2 * bvh.max_depth + 2 * 6 + 4 > kLaneStack) return bail(RTX_E_INVALID, "BLAS too deep for the traversal stack");
            node_off[mi + 1] = node_off[mi] + (uint32_t)bvh.nodes.size(); tri_off[mi + 1] = tri_off[mi] + d->meshes[mi].n_faces;
            sc->mesh_root[mi] = node_off[mi]; sc->mesh_tri_base[mi] = tri_off[mi];
        }
        h_nodes.reserve(((size_t)node_off[d->n_meshes] + 2 * (size_t)d->n_items + 64) * 5);    // room for the TLASes (a merged BLAS may still reallocate)
        h_nodes.resize((size_t)node_off[d->n_meshes] * 5); if (!tris_on_device) h_tris.resize((size_t)tri_off[d->n_meshes] * 3);
        n_mesh_tris = tri_off[d->n_meshes];
        std::atomic<uint32_t> next{0};
        auto pack = [&]() {
            for (uint32_t mi = next.fetch_add(1); mi < d->n_meshes; mi = next.fetch_add(1)) {
                const RtxMesh& m = d->meshes[mi]; WideBvh& bvh = bvhs[mi];
                float4* np = h_nodes.data() + (size_t)node_off[mi] * 5;
                for (size_t k = 0; k < bvh.nodes.size(); k++) {
                    WideNode n = bvh.nodes[k];
                    n.child_base += node_off[mi]; n.prim_base += tri_off[mi];
                    memcpy(np + 5 * k, &n, 80);
                }
                float4* tp = tris_on_device ? nullptr : h_tris.data() + (size_t)tri_off[mi] * 3;
                for (size_t k = 0; !tris_on_device && k < bvh.prim_order.size(); k++) {
                    const uint32_t f = bvh.prim_order[k];
                    const float* a = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f];
                    const float* b = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f + 1];
                    const float* c = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f + 2];
                    float fb; memcpy(&fb, &f, 4);
                    tp[3 * k] = make_float4(a[0], a[1], a[2], fb);
                    tp[3 * k + 1] = make_float4(b[0], b[1], b[2], 0.f);
                    tp[3 * k + 2] = make_float4(c[0], c[1], c[2], 0.f);
                }
                bvh = WideBvh();                                              // release as we go
            }
        };
        const uint32_t nt = std::min<uint32_t>(d->n_meshes, std::max(1u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (uint32_t t = 1; t < nt; t++) pool.emplace_back(pack);
        if (d->n_meshes) pack();
        for (std::thread& t : pool) t.join();
    }
    sc->n_blas_nodes = (uint32_t)(h_nodes.size() / 5); sc->n_tris = n_mesh_tris;
    phase("BLAS nodes + triangles assembled");

    // ---- items ----
    sc->h_items.resize(d->n_items);
    for (uint32_t i = 0; i < d->n_items; i++) {
        DItem& di = sc->h_items[i]; const RtxItem& s = d->items[i];
        fill_item(*sc, s, di);
        if (s.shape == RTX_SHAPE_MESH) {
            const RtxMesh& m = d->meshes[s.mesh]; const MeshOff& o = moff[s.mesh];
            di.lo = make_float4(o.lo[0], o.lo[1], o.lo[2], 0.f); di.hi.x = o.hi[0]; di.hi.y = o.hi[1]; di.hi.z = o.hi[2];
            di.root = sc->mesh_root[s.mesh]; di.tri_base = sc->mesh_tri_base[s.mesh];
            if (m.n_faces <= kDirectTris && !getenv("RTX_NO_DIRECT_TRIS")) di.flags |= IF_DIRECT_TRIS;
            di.n_faces = m.n_faces; di.n_uv_faces = m.n_uv_faces; di.n_normal_faces = m.n_normal_faces;
            di.vert_off = o.vert; di.idx_off = o.idx; di.uv_off = o.uv; di.uvidx_off = o.uvidx; di.nrm_off = o.nrm; di.nidx_off = o.nidx;
            if (m.n_normals > 0 && m.n_normal_faces > 0) di.flags |= IF_HAS_NORMALS;
        } else {
            const float r = s.radius;                                     // Ball::aabb
            di.lo.x = -r; di.lo.y = -r; di.lo.z = -r; di.hi.x = r; di.hi.y = r; di.hi.z = r;
        }
    }

    // ---- merged world-space BLAS over the mesh items that carry the identity transform (every primitive of a glTF file does:
    // Scene::load_gltf bakes the node transforms into the vertices, scene.rs:853-891).  Their triangles are ALSO kept in the
    // per-mesh BLASes: the reference-order walk (K3b) and the probes go item by item.
    {
        uint64_t gt = 0; std::vector<uint32_t> cand;
        for (uint32_t i = 0; i < d->n_items; i++) {
            const RtxItem& s = d->items[i];
            if (s.shape == RTX_SHAPE_MESH && is_identity16(s.trans) && is_identity16(s.tran_inverse)) { cand.push_back(i); gt += d->meshes[s.mesh].n_faces; }
        }
        // not for a handful of quads (room walls: entering them directly costs less than a BLAS node plus the per-hit item work:
        // room of spheres 762 ms ungrouped, 862 ms grouped), not for soups so large that a single-threaded build would dominate
        uint64_t max_tris = 4000000, min_tris = 4096;
        if (const char* e = getenv("RTX_GROUP_MAX_TRIS")) max_tris = (uint64_t)atoll(e);
        if (const char* e = getenv("RTX_GROUP_MIN_TRIS")) min_tris = (uint64_t)atoll(e);
        if (cand.size() >= 2 && gt >= min_tris && gt <= max_tris && !getenv("RTX_NO_GROUP")) {
            std::vector<Aabb3> boxes((size_t)gt); std::vector<uint32_t> p_item((size_t)gt), p_face((size_t)gt);
            size_t k = 0;
            for (uint32_t i : cand) {
                const RtxMesh& m = d->meshes[d->items[i].mesh];
                for (uint32_t f = 0; f < m.n_faces; f++, k++) {
                    Aabb3& b = boxes[k];
                    for (int c = 0; c < 3; c++) { b.lo[c] = INFINITY; b.hi[c] = -INFINITY; }
                    for (int c = 0; c < 3; c++) {
                        const float* v = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f + c];
                        for (int q = 0; q < 3; q++) { b.lo[q] = std::min(b.lo[q], v[q]); b.hi[q] = std::max(b.hi[q], v[q]); }
                    }
                    p_item[k] = i; p_face[k] = f;
                }
            }
            WideBvh bvh;
            if (!device_bvh || !build_boxes_on_device(boxes, bvh)) build_wide_bvh(boxes.data(), (uint32_t)gt, bvh, kBlasDepthLimit);
            if (!(bvh.max_depth >= kStack - 2 || 2 * bvh.max_depth + 2 * 6 + 4 > kLaneStack)) {       // too deep: keep the per-item structure only
                const uint32_t node_off = (uint32_t)(h_nodes.size() / 5), tri_off = sc->n_tris;
                append_nodes(h_nodes, bvh, node_off, tri_off);
                std::vector<float4> g_tris;                                  // device mode: copied behind the per-mesh records
                std::vector<float4>& gt_out = tris_on_device ? g_tris : h_tris;
                gt_out.reserve(gt_out.size() + (size_t)gt * 3);
                for (uint32_t pi : bvh.prim_order) {
                    const RtxMesh& m = d->meshes[d->items[p_item[pi]].mesh]; const uint32_t f = p_face[pi];
                    const float* a = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f];
                    const float* b = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f + 1];
                    const float* c = m.vertices + 3 * (size_t)m.indices[3 * (size_t)f + 2];
                    float fb, ib; memcpy(&fb, &f, 4); memcpy(&ib, &p_item[pi], 4);
                    gt_out.push_back(make_float4(a[0], a[1], a[2], fb));
                    gt_out.push_back(make_float4(b[0], b[1], b[2], ib));            // .w = item: what the reference does per item happens per accepted triangle
                    gt_out.push_back(make_float4(c[0], c[1], c[2], 0.f));
                }
                if (tris_on_device) {
                    if ((size_t)(tri_off + gt) * 3 > sc->tris.n) return bail(RTX_E_INVALID, "internal: merged BLAS does not fit the triangle array");
                    CU(cudaMemcpy(sc->tris.p + (size_t)tri_off * 3, g_tris.data(), g_tris.size() * sizeof(float4), cudaMemcpyHostToDevice));
                }
                sc->group_root = node_off; sc->group_items = cand; sc->n_group_tris = (uint32_t)gt;
                for (uint32_t i : cand) sc->h_items[i].flags |= IF_GROUPED;
                sc->n_blas_nodes = (uint32_t)(h_nodes.size() / 5); sc->n_tris = tri_off + (uint32_t)gt;
            }
        }
    }

    phase("merged BLAS");
    // ---- TLASes: the full one (every item) and the fast one (ungrouped items + the merged BLAS); space reserved for updates ----
    std::vector<float4> tlas_nodes, fast_nodes; std::vector<uint32_t> tlas_prims, fast_prims;
    sc->tlas_cap = std::max<uint32_t>(8, d->n_items + 1);
    if (d->n_items > 0) {
        int rc = build_tlas(*sc, tlas_nodes, tlas_prims);
        if (!rc) rc = build_tlas_fast(*sc, fast_nodes, fast_prims);
        if (rc) { rtx_scene_destroy(sc); return rc; }
    }
    h_nodes.insert(h_nodes.end(), tlas_nodes.begin(), tlas_nodes.end());
    h_nodes.resize(((size_t)sc->n_blas_nodes + sc->tlas_cap) * 5, make_float4(0, 0, 0, 0));
    h_nodes.insert(h_nodes.end(), fast_nodes.begin(), fast_nodes.end());
    h_nodes.resize(((size_t)sc->n_blas_nodes + 2 * (size_t)sc->tlas_cap) * 5, make_float4(0, 0, 0, 0));
    tlas_prims.resize(std::max<size_t>(1, d->n_items), 0);
    fast_prims.resize(std::max<size_t>(1, (size_t)d->n_items + 1), 0);

    // ---- materials / textures / lights ----
    std::vector<DTex> h_texs(d->n_textures); std::vector<uchar4> h_texels;
    for (uint32_t i = 0; i < d->n_textures; i++) {
        const RtxTexture& t = d->textures[i];
        size_t off = h_texels.size(), n = (size_t)t.width * t.height;
        h_texs[i] = DTex{(uint32_t)(off & 0xffffffffu), (uint32_t)(off >> 32), t.width, t.height};
        if (n && !t.rgba) return bail(RTX_E_INVALID, "texture without pixels");
        h_texels.resize(off + n);
        if (n) memcpy(h_texels.data() + off, t.rgba, n * 4);
    }
    sc->texture_bytes = h_texels.size() * 4;
    std::vector<DMaterial> h_mats(d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const RtxMaterial& m = d->materials[i]; DMaterial& o = h_mats[i]; memset(&o, 0, sizeof(o));
        memcpy(o.ambient, m.ambient_color, 12); memcpy(o.base, m.base_color, 12); memcpy(o.specular, m.specular_color, 12);
        o.alpha = m.alpha; o.shininess = m.shininess; o.reflectivity = m.reflectivity; o.refraction_index = m.refraction_index;
        o.normal_map_strength = m.normal_map_strength; o.shadow_softness = m.shadow_softness; o.roughness = m.roughness;
        o.nearest = m.texture_filtering_nearest; o.receive_shadow = m.receive_shadow; o.monte_carlo = m.monte_carlo;
        o.shadow_z_lo = m.shadow_softness <= 0.0f ? 1.0f : cosf(m.shadow_softness * 3.14159265358979323846f);   // jitter(): z range
        o.rough_z_lo = m.roughness <= 0.0f ? 1.0f : cosf(m.roughness * 3.14159265358979323846f);
        for (int t = 0; t < 8; t++) {
            o.tex[t] = (m.texture[t] >= 0 && d->textures[m.texture[t]].width > 0) ? m.texture[t] : -1;
            if (o.tex[t] >= 0) o.any_texture = 1;
        }
    }
    std::vector<DLight> h_lights; fill_lights(d->lights, d->n_lights, h_lights);

    phase("TLAS, materials, textures");
    // ---- upload ----
    int rc;
    if ((rc = sc->nodes.upload(h_nodes)) || (!tris_on_device && (rc = sc->tris.upload(h_tris))) || (rc = sc->items.upload(sc->h_items)) ||
        (rc = sc->tlas_prims.upload(tlas_prims)) || (rc = sc->fast_prims.upload(fast_prims)) ||
        (rc = sc->mats.upload(h_mats)) || (rc = sc->texs.upload(h_texs)) || (rc = sc->texels.upload(h_texels)) || (rc = sc->lights.upload(h_lights))) {
        std::string keep = g_err; rtx_scene_destroy(sc); g_err = keep; return rc;
    }
    CU(cudaDeviceSynchronize());
    phase("uploaded");
    refresh_dev(*sc);
    sc->dev.n_lights = d->n_lights;
    sc->n_enabled_lights = 0;
    for (uint32_t i = 0; i < d->n_lights; i++) if (d->lights[i].enabled) sc->n_enabled_lights++;
    { int rc2 = occupancy_blocks(*sc); if (rc2) { std::string keep = g_err; rtx_scene_destroy(sc); g_err = keep; return rc2; } }
    sc->build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = sc;
    return RTX_OK;
}

int rtx_scene_destroy(RtxScene* sc) {
    if (!sc) return RTX_OK;
    sc->cancel.store(true);
    if (sc->worker.joinable()) sc->worker.join();
    for (RtxScene* r : sc->replicas) rtx_scene_destroy(r);
    sc->replicas.clear();
    cudaSetDevice(sc->device);
    cudaDeviceSynchronize();
    {
        RtxScene::Probe& P = sc->probe;
        P.q_o.release(); P.q_d.release(); P.s_c.release(); P.acc.release(); P.q_m.release(); P.s_r.release(); P.slow.release(); P.ctr.release();
        P.hits.release(); P.out.release(); P.len.release(); P.recv.release(); P.res.release(); P.beyond.release();
        if (P.st) cudaStreamDestroy(P.st);
    }
    if (sc->shadow_stream) { cudaStreamDestroy(sc->shadow_stream); cudaEventDestroy(sc->ev_shaded); cudaEventDestroy(sc->ev_shadow_done[0]); cudaEventDestroy(sc->ev_shadow_done[1]); cudaFreeHost(sc->h_pool); }
    if (sc->own_stream) cudaStreamDestroy(sc->own_stream);
    if (sc->fence) cudaEventDestroy(sc->fence);
    sc->nodes.release(); sc->tris.release(); sc->items.release(); sc->tlas_prims.release(); sc->fast_prims.release(); sc->s_beyond.release();
    sc->verts.release(); sc->uvs.release(); sc->nrms.release(); sc->idx.release(); sc->uv_idx.release(); sc->n_idx.release();
    sc->mats.release(); sc->texs.release(); sc->texels.release(); sc->lights.release();
    sc->accum_c.release(); sc->accum_n.release(); sc->ids.release(); sc->sample_table.release();
    sc->q_o.release(); sc->q_d.release(); sc->q_m.release(); sc->s_o.release(); sc->s_d.release(); sc->s_c.release(); sc->s_r.release(); sc->s_slow.release();
    sc->hits.release(); sc->ctr_pool.release(); sc->overflow.release(); sc->counters.release();
    sc->o_rgba.release(); sc->o_normals.release(); sc->o_depth.release(); sc->o_ids.release(); sc->p_rays.release(); sc->p_hits.release();
    for (auto& kv : sc->pixel_lists) { kv.second->d.release(); delete kv.second; }
    for (cudaEvent_t e : sc->events) cudaEventDestroy(e);
    if (sc->h_ctr) cudaFreeHost(sc->h_ctr);
    delete sc;
    return RTX_OK;
}

int rtx_scene_update_items(RtxScene* sc, const RtxItemXform* x, size_t n) {
    if (!sc || (!x && n)) return fail(RTX_E_INVALID, "null argument");
    // the reference takes the scene's RwLock for writing here (run.rs); a frame in flight holds it for reading
    if (sc->running.load()) return fail(RTX_E_BUSY, "a frame is in flight on this handle");
    std::unique_lock<std::mutex> lk(sc->api_mu, std::defer_lock);
    if (!sc->primary) lk.lock();
    std::lock_guard<std::mutex> plk(sc->probe.mu);
    for (RtxScene* r : sc->replicas) { int rc = rtx_scene_update_items(r, x, n); if (rc) return rc; }
    CU(cudaSetDevice(sc->device));
    for (size_t i = 0; i < n; i++) {
        if (x[i].item_index >= sc->src_items.size()) return fail(RTX_E_INVALID, "item index out of range");
        const float* ti = x[i].tran_inverse;
        if (ti[3] != 0.0f || ti[7] != 0.0f || ti[11] != 0.0f || !(ti[15] > 0.5f && ti[15] < 2.0f)) return fail(RTX_E_NON_AFFINE, "item transform is not affine");
    }
    for (size_t i = 0; i < n; i++) {
        RtxItem& s = sc->src_items[x[i].item_index]; DItem& d = sc->h_items[x[i].item_index];
        memcpy(s.trans, x[i].trans, 64); memcpy(s.tran_inverse, x[i].tran_inverse, 64);
        for (int r = 0; r < 3; r++) {
            d.inv[r] = make_float4(s.tran_inverse[0 + r], s.tran_inverse[4 + r], s.tran_inverse[8 + r], s.tran_inverse[12 + r]);
            d.mat[r] = make_float4(s.trans[0 + r], s.trans[4 + r], s.trans[8 + r], s.trans[12 + r]);
        }
        const float* t = s.tran_inverse;
        const bool tr = t[0] == 1.f && t[5] == 1.f && t[10] == 1.f && t[15] == 1.f && t[1] == 0.f && t[2] == 0.f && t[4] == 0.f && t[6] == 0.f && t[8] == 0.f && t[9] == 0.f;
        d.flags = tr ? (d.flags | IF_TRANSLATION) : (d.flags & ~IF_TRANSLATION);
        d.inv_w = t[15];
        d.flags = (t[15] != 1.0f) ? (d.flags | IF_DIV_W) : (d.flags & ~IF_DIV_W);
    }
    CU(cudaDeviceSynchronize());
    if (sc->group_root != 0xFFFFFFFFu) {
        // an item of the merged BLAS that is moved leaves world space == object space: the group is dissolved (its items are
        // walked one by one again, like every other item); a static glTF scene never gets here
        bool intact = true;
        for (uint32_t gi : sc->group_items) if (!is_identity16(sc->src_items[gi].trans) || !is_identity16(sc->src_items[gi].tran_inverse)) intact = false;
        if (!intact) {
            for (uint32_t gi : sc->group_items) sc->h_items[gi].flags &= ~IF_GROUPED;
            sc->group_root = 0xFFFFFFFFu; sc->group_items.clear(); sc->n_group_tris = 0;
            refresh_dev(*sc);
        }
    }
    if (!sc->h_items.empty()) {                                           // Scene::update rebuilds the item BVH (scene.rs:1681-1687)
        std::vector<float4> tn, fn; std::vector<uint32_t> tp, fp;
        int rc = build_tlas(*sc, tn, tp); if (rc) return rc;              // also refreshes the items' world boxes
        if ((rc = build_tlas_fast(*sc, fn, fp))) return rc;
        if (tn.size() / 5 > sc->tlas_cap || fn.size() / 5 > sc->tlas_cap) return fail(RTX_E_INVALID, "TLAS capacity exceeded");
        CU(cudaMemcpy(sc->nodes.p + (size_t)sc->n_blas_nodes * 5, tn.data(), tn.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(sc->nodes.p + ((size_t)sc->n_blas_nodes + sc->tlas_cap) * 5, fn.data(), fn.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(sc->tlas_prims.p, tp.data(), tp.size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(sc->fast_prims.p, fp.data(), fp.size() * 4, cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(sc->items.p, sc->h_items.data(), sc->h_items.size() * sizeof(DItem), cudaMemcpyHostToDevice));
    return RTX_OK;
}

int rtx_scene_set_lights(RtxScene* sc, const RtxLight* l, uint32_t n) {
    if (!sc || (!l && n)) return fail(RTX_E_INVALID, "null argument");
    if (sc->running.load()) return fail(RTX_E_BUSY, "a frame is in flight on this handle");
    std::unique_lock<std::mutex> lk(sc->api_mu, std::defer_lock);
    if (!sc->primary) lk.lock();
    for (RtxScene* r : sc->replicas) { int rc = rtx_scene_set_lights(r, l, n); if (rc) return rc; }
    CU(cudaSetDevice(sc->device));
    std::vector<DLight> h; fill_lights(l, n, h);
    CU(cudaDeviceSynchronize());
    int rc = sc->lights.upload(h); if (rc) return rc;
    CU(cudaDeviceSynchronize());
    refresh_dev(*sc);
    sc->dev.n_lights = n;
    sc->n_enabled_lights = 0;
    for (uint32_t i = 0; i < n; i++) if (l[i].enabled) sc->n_enabled_lights++;
    return RTX_OK;
}

int rtx_scene_bvh_info(const RtxScene* sc, RtxBvhInfo* info) {
    if (!sc || !info) return fail(RTX_E_INVALID, "null argument");
    info->n_nodes = sc->n_blas_nodes; info->n_triangles = sc->n_tris; info->n_items = (uint32_t)sc->h_items.size();
    info->tlas_nodes = sc->n_tlas_nodes; info->node_bytes = (uint64_t)(sc->n_blas_nodes + sc->n_tlas_nodes + sc->n_fast_nodes) * 80;
    info->grouped_items = (uint32_t)sc->group_items.size(); info->grouped_triangles = sc->n_group_tris;
    info->device_build_ms = sc->device_build_ms;
    info->triangle_bytes = (uint64_t)sc->n_tris * 48; info->item_bytes = sc->h_items.size() * sizeof(DItem);
    info->texture_bytes = sc->texture_bytes; info->build_ms = sc->build_ms;
    return RTX_OK;
}

// ---- the frame ------------------------------------------------------------------------------------
}  // extern "C"

namespace {

struct FrameCtx {
    RtxScene* sc; cudaStream_t st; FrameDev F; PixelList* pl; const RtxConfig* cfg; const RtxCamera* cam;
    void *d_rgba, *d_normals, *d_depth, *d_ids;
    bool want_stats, ordered, primary_single = false; uint32_t L; int gs;
    cudaStream_t sst = nullptr; int shadow_buf = 0; bool shadow_pending[2] = {false, false};   // shadow stream, queue half of the next wave
    uint64_t launches = 0, rays_closest = 0, rays_shadow = 0, rays_exact = 0, rays_beyond = 0, primary = 0; uint32_t waves = 0, batches = 0; size_t ev_next = 2;
    cudaEvent_t event(size_t i) {
        while (sc->events.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); sc->events.push_back(e); }
        return sc->events[i];
    }
};

// One wave at depth d: closest -> shade -> shadow.  n_ptr == nullptr: n rays at q_base (count known on the host);
// otherwise the count is read on the device (at most n).  ctr = this wave's 8 counters.
int launch_wave(FrameCtx& X, uint32_t d, uint32_t q_base, uint32_t n, const uint32_t* n_ptr, uint32_t* ctr, uint32_t child_off) {
    RtxScene* sc = X.sc; cudaStream_t st = X.st;
    RayQ Q = level_queue(*sc, d);
    // Shadow kernels run on their own stream: wave k's any-hit / beyond / exact kernels overlap wave k+1's closest-hit and shade
    // kernels (they only add to the accumulators), which fills the tails of the persistent kernels — what a strong-scaled shard with
    // its thin waves needs most.  The queue half a shade kernel writes must have been drained by the shadow kernels of two waves ago.
    const int buf = X.shadow_buf; X.shadow_buf ^= 1;
    const size_t so_off = (size_t)buf * sc->shadow_cap;
    ShadowQ SQ{sc->s_o.p + so_off, sc->s_d.p + so_off, sc->s_c.p + so_off, sc->s_r.p + so_off, nullptr, sc->s_beyond.p};
    cudaStream_t sst = X.sst;
    if (X.shadow_pending[buf] && sst != st) CU(cudaStreamWaitEvent(st, sc->ev_shadow_done[buf], 0));
    cudaEvent_t e0 = X.event(X.ev_next), e1 = X.event(X.ev_next + 1), e2 = X.event(X.ev_next + 2), e3 = X.event(X.ev_next + 3);
    X.ev_next += 4;
    const bool verify = !n_ptr && getenv("RTX_VERIFY");
    if (verify) cudaMemsetAsync(sc->hits.p, 0xEE, (size_t)n * sizeof(HitRec), st);
    CU(cudaEventRecord(e0, st));
    {
        const int maxb = X.want_stats ? sc->blocks_closest_st : sc->blocks_closest;
        const uint32_t warps = (n + 31) / 32;
        const int blocks = n_ptr ? maxb : (int)std::min<uint32_t>((warps + (kTraceBlock / 32) - 1) / (kTraceBlock / 32), (uint32_t)maxb);
        if (X.want_stats) closest_kernel<true><<<blocks, kTraceBlock, 0, st>>>(sc->dev, Q, q_base, n, n_ptr, sc->hits.p, ctr, sc->counters.p);
        else closest_kernel<false><<<blocks, kTraceBlock, 0, st>>>(sc->dev, Q, q_base, n, n_ptr, sc->hits.p, ctr, sc->counters.p);
        X.launches++;
    }
    CU(cudaEventRecord(e1, st));
    if (verify) {
        static uint32_t* d_cnt = nullptr; static VerifyRec* d_rec = nullptr;
        if (!d_cnt) { cudaMalloc(&d_cnt, 4); cudaMalloc(&d_rec, 64 * sizeof(VerifyRec)); }
        cudaMemsetAsync(d_cnt, 0, 4, st);
        verify_closest_kernel<<<(n + 127) / 128, 128, 0, st>>>(sc->dev, Q, q_base, n, sc->hits.p, d_cnt, d_rec, 64);
        uint32_t hc = 0; VerifyRec hr[64];
        cudaMemcpyAsync(&hc, d_cnt, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        if (hc) {
            cudaMemcpy(hr, d_rec, sizeof(hr), cudaMemcpyDeviceToHost);
            uint32_t ovf[4]; cudaMemcpy(ovf, sc->overflow.p, 16, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[RTX_VERIFY] wave %u depth %u n %u: %u mismatching closest hits (queue overflow %u, lane stack overflow %u) err=%s\n", X.waves, d, n, hc, ovf[0], ovf[1], cudaGetErrorString(cudaGetLastError()));
            for (uint32_t k = 0; k < std::min(hc, 40u); k++)
                fprintf(stderr, "   idx %u o %.9g %.9g %.9g d %.9g %.9g %.9g depth %u | prod t %.9g item %x prim %u | ref t %.9g item %d prim %u\n", hr[k].index, hr[k].o[0], hr[k].o[1],
                        hr[k].o[2], hr[k].d[0], hr[k].d[1], hr[k].d[2], hr[k].depth, hr[k].t_prod, (int)hr[k].item_prod, hr[k].prim_prod, hr[k].t_ref,
                        (int)hr[k].item_ref, hr[k].prim_ref);
        }
    }
    {
        ShadeOut so;
        const bool spawn = d + 1 <= X.L;                                  // depth L never spawns children (depth <= max_recursion fails)
        RayQ C = level_queue(*sc, spawn ? d + 1 : d);
        const uint32_t off = spawn ? child_off : 0;
        so.child = RayQ{C.o + off, C.d + off, C.m + off};
        // a sync-free level is consumed by ONE wave, so it may not hold more than a wave
        so.child_cap = spawn ? (n_ptr || X.primary_single ? std::min(sc->level_cap - off, sc->wave_cap) : sc->level_cap - off) : 0;
        so.child_count = ctr + 2;
        so.shadow = SQ; so.shadow_cap = sc->shadow_cap; so.shadow_count = ctr + 3; so.overflow = sc->overflow.p; so.skipped = sc->overflow.p + 2;
        const int blocks = (int)std::min<uint32_t>((n + kShadeBlock - 1) / kShadeBlock, (uint32_t)sc->sm_count * 16);
        shade_kernel<<<blocks, kShadeBlock, 0, st>>>(sc->dev, X.F, Q, q_base, n, n_ptr, sc->hits.p, so);
        X.launches++;
    }
    if (sst != st) { CU(cudaEventRecord(sc->ev_shaded, st)); CU(cudaStreamWaitEvent(sst, sc->ev_shaded, 0)); }
    CU(cudaEventRecord(e2, sst));
    if (sc->n_enabled_lights > 0) {
        const int eb = sc->sm_count * 8;
        if (X.ordered) {
            shadow_exact_kernel<false, true><<<eb, kTraceBlock, 0, sst>>>(sc->dev, X.F, SQ, nullptr, ctr + 3, sc->shadow_cap, d, sc->counters.p);
            X.launches++;
        } else {
            if (X.want_stats) shadow_any_kernel<true><<<sc->blocks_shadow_st, kTraceBlock, 0, sst>>>(sc->dev, X.F, SQ, ctr + 3, sc->shadow_cap, d, ctr + 1, sc->s_slow.p, ctr + 4, sc->counters.p);
            else shadow_any_kernel<false><<<sc->blocks_shadow, kTraceBlock, 0, sst>>>(sc->dev, X.F, SQ, ctr + 3, sc->shadow_cap, d, ctr + 1, sc->s_slow.p, ctr + 4, sc->counters.p);
            if (sc->dev.group_root != 0xFFFFFFFFu) {
                if (X.want_stats) shadow_beyond_kernel<true><<<eb, kTraceBlock, 0, sst>>>(sc->dev, X.F, SQ, ctr + 5, sc->shadow_cap, d, sc->s_slow.p, ctr + 4, sc->counters.p);
                else shadow_beyond_kernel<false><<<eb, kTraceBlock, 0, sst>>>(sc->dev, X.F, SQ, ctr + 5, sc->shadow_cap, d, sc->s_slow.p, ctr + 4, sc->counters.p);
                X.launches++;
            }
            if (X.want_stats) shadow_exact_kernel<true, false><<<eb, kTraceBlock, 0, sst>>>(sc->dev, X.F, SQ, sc->s_slow.p, ctr + 4, sc->shadow_cap, d, sc->counters.p);
            else shadow_exact_kernel<false, false><<<eb, kTraceBlock, 0, sst>>>(sc->dev, X.F, SQ, sc->s_slow.p, ctr + 4, sc->shadow_cap, d, sc->counters.p);
            X.launches += 2;
        }
    }
    CU(cudaEventRecord(e3, sst));
    if (sst != st) { CU(cudaEventRecord(sc->ev_shadow_done[buf], sst)); X.shadow_pending[buf] = true; }
    return RTX_OK;
}

// The frame of ONE device (its shard of the pixels).  Outputs may be peer memory (another device's frame buffers).
int render_frame_single(RtxScene* sc, const RtxCamera* cam, const RtxConfig* cfg, const RtxShard* shard, void* d_rgba, void* d_normals,
                        void* d_depth, void* d_object_ids, cudaStream_t st, RtxStats* stats) {
    CU(cudaSetDevice(sc->device));
    FrameCtx X; X.sc = sc; X.st = st; X.cfg = cfg; X.cam = cam; X.d_rgba = d_rgba; X.d_normals = d_normals; X.d_depth = d_depth; X.d_ids = d_object_ids;
    X.want_stats = cfg->debug_flags & RTX_DEBUG_COLLECT_STATS; X.ordered = cfg->debug_flags & RTX_DEBUG_ORDERED_SHADOW;
    int rc;
    PixelList* pl;
    if ((rc = get_pixel_list(*sc, cam->width, cam->height, shard, &pl))) return rc;
    X.pl = pl;
    if ((rc = ensure_frame(*sc, cam->width, cam->height))) return rc;
    const uint64_t n_primary = (uint64_t)pl->n * cfg->samples;
    if ((rc = ensure_queues(*sc, cfg->max_recursion, sc->n_enabled_lights, n_primary))) return rc;
    if ((rc = occupancy_blocks(*sc))) return rc;
    uint64_t h2d = 0;
    if (sc->table_samples != cfg->samples) {
        std::vector<ushort2> t; make_sample_table(cfg->samples, sc->cell_size, t);
        if ((rc = sc->sample_table.upload(t, st))) return rc;
        sc->table_samples = cfg->samples; h2d += t.size() * 4;
    }
    FrameDev& F = X.F; memset(&F, 0, sizeof(F));
    memcpy(F.pinv, cam->projection_inverse, 64); memcpy(F.vinv, cam->view_inverse, 64);
    F.width = cam->width; F.height = cam->height; F.n_samples = cfg->samples; F.cell_size = sc->cell_size;
    F.monte_carlo = cfg->monte_carlo; F.max_recursion = cfg->max_recursion; F.gamma = cfg->gamma_correction; F.mc_seed = cfg->mc_seed;
    F.focal_length = cfg->focal_length; F.aperture_size = cfg->aperture_size; F.fog_density = cfg->fog_density;
    memcpy(F.fog_color, cfg->fog_color, 12); F.debug_flags = cfg->debug_flags;
    // ray order inside a primary batch: one warp = 32 samples of ONE pixel (sub-pixel footprint: the lanes walk the same nodes,
    // hit the same material, and their shadow rays leave from the same spot) instead of one sample of an 8x4 tile: -4 % frame time
    // (all samples of a pixel, up to 128, back to back: config-4 stand-in at 128 spp 204.3 -> 202.1 ms against groups of 32)
    F.sample_group = 128; if (const char* e = getenv("RTX_SAMPLE_GROUP")) F.sample_group = std::max(1, atoi(e));
    F.sample_table = sc->sample_table.p; F.accum_c = sc->accum_c.p; F.accum_n = sc->accum_n.p; F.ids = sc->ids.p;
    h2d += sizeof(FrameDev);                                             // kernel parameters (camera + config)

    const uint32_t L = cfg->max_recursion + 1;                           // deepest ray depth
    X.L = L; X.gs = sc->sm_count * 8;
    const int gs = X.gs;
    const uint32_t chunk = sc->chunk;
    bool sync_free = n_primary <= chunk && !sc->no_sync_free && !getenv("RTX_FORCE_SYNC") && !getenv("RTX_VERIFY");
    uint32_t ovf[4] = {0, 0, 0, 0};
    X.sst = (getenv("RTX_NO_OVERLAP") || getenv("RTX_VERIFY") || (cfg->debug_flags & RTX_DEBUG_SERIAL_STREAMS)) ? st : sc->shadow_stream;
    struct ShadowStreamGuard { cudaStream_t s; ~ShadowStreamGuard() { cudaStreamSynchronize(s); } } sguard{sc->shadow_stream};   // no exit leaves it running
    // everything the shadow stream did must be in the accumulators before they are resolved (or cleared again)
    auto join_shadow = [&]() -> int {
        for (int k = 0; k < 2; k++) if (X.shadow_pending[k]) { CU(cudaStreamWaitEvent(st, sc->ev_shadow_done[k], 0)); X.shadow_pending[k] = false; }
        return RTX_OK;
    };

    for (int attempt = 0; attempt < 2; attempt++) {
        X.launches = 0; X.rays_closest = X.rays_shadow = X.rays_exact = X.rays_beyond = X.primary = 0; X.waves = X.batches = 0; X.ev_next = 2;
        X.primary_single = sync_free; X.shadow_buf = 0;
        if ((rc = join_shadow())) return rc;
        // events: [0] frame start, [1] frame end, then 4 per wave (closest start/end, shadow start/end)
        CU(cudaEventRecord(X.event(0), st));
        CU(cudaMemsetAsync(sc->ctr_pool.p, 0, kCtrPool * 4, st));
        CU(cudaMemsetAsync(sc->overflow.p, 0, 16, st));
        if (X.want_stats) CU(cudaMemsetAsync(sc->counters.p, 0, sizeof(Counters), st));
        clear_pixels_kernel<<<gs, 256, 0, st>>>(F, pl->d.p, pl->n); X.launches++;

        if (sync_free) {
            // ---- the whole frame in one go: level d+1's count is level d's child counter, never seen by the host ----
            if (sc->cancel.load(std::memory_order_relaxed)) { CU(cudaStreamSynchronize(st)); return fail(RTX_E_CANCELLED, "frame cancelled by rtx_render_stop"); }
            const uint32_t n1 = (uint32_t)n_primary;
            if (n1) { raygen_kernel<<<std::min<uint32_t>((n1 + 255) / 256, gs), 256, 0, st>>>(F, pl->d.p, 0, pl->n, 0, cfg->samples, level_queue(*sc, 1), 0); X.launches++; X.batches++; }
            X.primary = n1;
            sc->samples_issued.store(n1, std::memory_order_relaxed);
            for (uint32_t d = 1; d <= L && n1; d++) {
                uint32_t* ctr = sc->ctr_pool.p + 8 * d;                   // block d: [2] = rays appended to level d+1
                if ((rc = launch_wave(X, d, 0, d == 1 ? n1 : sc->wave_cap, d == 1 ? nullptr : ctr - 8 + 2, ctr, 0))) return rc;
                X.waves++;
            }
            if ((rc = join_shadow())) return rc;
            resolve_kernel<<<std::min<uint32_t>((pl->n + 255) / 256, gs), 256, 0, st>>>(F, pl->d.p, pl->n, (uchar4*)d_rgba, (float*)d_normals, (float*)d_depth,
                                                                                    (uint32_t*)d_object_ids);
            X.launches++;
            CU(cudaEventRecord(X.event(1), st));
            CU(cudaMemcpyAsync(sc->h_ctr + 4, sc->ctr_pool.p, (size_t)8 * (L + 1) * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(sc->h_ctr, sc->overflow.p, 16, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            CU(cudaGetLastError());
            memcpy(ovf, sc->h_ctr, 16);
            if (getenv("RTX_TRACE_SCHED")) {
                fprintf(stderr, "[sched] sync-free frame: primary %u chunk %u wave_cap %u shadow_cap %u overflow %u |", n1, chunk, sc->wave_cap, sc->shadow_cap, ovf[0]);
                for (uint32_t d = 1; d <= L; d++) fprintf(stderr, " L%u children %u shadow %u exact %u;", d, sc->h_ctr[4 + 8 * d + 2], sc->h_ctr[4 + 8 * d + 3], sc->h_ctr[4 + 8 * d + 4]);
                fprintf(stderr, "\n");
            }
            if (ovf[0]) {                                                 // a level outgrew one wave: redo with the synchronised schedule
                sc->no_sync_free = true; sync_free = false;
                continue;
            }
            X.rays_closest = n1;
            for (uint32_t d = 1; d <= L; d++) {
                const uint32_t* c = sc->h_ctr + 4 + 8 * d;
                if (d < L) X.rays_closest += c[2];
                X.rays_shadow += c[3]; X.rays_exact += c[4]; X.rays_beyond += c[5];
            }
            break;
        }

        // ---- synchronised schedule ----
        std::vector<uint32_t> counts(L + 2, 0);
        // primary batches: pixel ranges of <= chunk pixels x sample ranges so that np*ns <= chunk
        // Batch geometry: as many samples of a pixel as fit (all of them up to 128 spp) rather than few samples of every pixel —
        // the rays of one pixel then travel through the waves together (BVH, texture and accumulator locality: -2 % frame time on
        // config 2), and pixels finish in order, which is what the progressive display wants.
        uint32_t np_cap = std::min(chunk, std::max(65536u, chunk / std::max(1u, cfg->samples)));
        if (const char* e = getenv("RTX_BATCH_PIXELS")) np_cap = std::max(1, atoi(e));
        const uint32_t np_full = std::min(pl->n, np_cap);
        const uint32_t ns_full = std::max(1u, chunk / std::max(1u, np_full));
        uint32_t cur_p0 = 0, cur_s0 = 0;
        bool primary_left = pl->n > 0;
        uint32_t ctr_idx = 0;
        // per-wave counts that only the shadow kernels know (rays re-walked exactly / checked behind the light) are summed from the
        // counter pool once the shadow stream has been joined
        auto harvest = [&](uint32_t used) -> int {
            if (used) CU(cudaMemcpyAsync(sc->h_pool, sc->ctr_pool.p, (size_t)used * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (uint32_t k = 0; k + 8 <= used; k += 8) { X.rays_exact += sc->h_pool[k + 4]; X.rays_beyond += sc->h_pool[k + 5]; }
            return RTX_OK;
        };
        for (;;) {
            if (sc->cancel.load(std::memory_order_relaxed)) { CU(cudaStreamSynchronize(st)); return fail(RTX_E_CANCELLED, "frame cancelled by rtx_render_stop"); }
            if (sc->snap_req.load(std::memory_order_relaxed)) {
                if ((rc = join_shadow())) return rc;
                // progressive preview (the stream is idle here): pixels whose samples were all issued are normalised by the sample
                // count, the pixel range in progress by the samples issued so far, the rest is cleared; rays still queued at
                // deeper levels are simply not in the sums yet
                const size_t npx = (size_t)cam->width * cam->height;
                if (d_rgba) CU(cudaMemsetAsync(d_rgba, 0, npx * 4, st));
                if (d_normals) CU(cudaMemsetAsync(d_normals, 0, npx * 12, st));
                if (d_depth) CU(cudaMemsetAsync(d_depth, 0, npx * 4, st));
                if (d_object_ids) CU(cudaMemsetAsync(d_object_ids, 0, npx * 4, st));
                const uint32_t done_px = primary_left ? cur_p0 : pl->n;
                if (done_px) resolve_kernel<<<std::min<uint32_t>((done_px + 255) / 256, gs), 256, 0, st>>>(F, pl->d.p, done_px, (uchar4*)d_rgba, (float*)d_normals,
                                                                                                        (float*)d_depth, (uint32_t*)d_object_ids);
                if (primary_left && cur_s0 > 0) {
                    FrameDev Fp = F; Fp.n_samples = cur_s0;
                    const uint32_t np = std::min(np_full, pl->n - cur_p0);
                    resolve_kernel<<<std::min<uint32_t>((np + 255) / 256, gs), 256, 0, st>>>(Fp, pl->d.p + cur_p0, np, (uchar4*)d_rgba, (float*)d_normals, (float*)d_depth,
                                                                                           (uint32_t*)d_object_ids);
                }
                if (sc->snap_rgba && d_rgba) CU(cudaMemcpyAsync(sc->snap_rgba, d_rgba, npx * 4, cudaMemcpyDeviceToHost, st));
                if (sc->snap_normals && d_normals) CU(cudaMemcpyAsync(sc->snap_normals, d_normals, npx * 12, cudaMemcpyDeviceToHost, st));
                if (sc->snap_depth && d_depth) CU(cudaMemcpyAsync(sc->snap_depth, d_depth, npx * 4, cudaMemcpyDeviceToHost, st));
                if (sc->snap_ids && d_object_ids) CU(cudaMemcpyAsync(sc->snap_ids, d_object_ids, npx * 4, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                { std::lock_guard<std::mutex> lk(sc->snap_mu); sc->snap_seq++; sc->snap_req.store(0); }
                sc->snap_cv.notify_all();
            }
            int level = -1;
            for (int d = (int)L; d >= 1; d--) if (counts[d] >= chunk) { level = d; break; }
            if (level < 0 && primary_left) level = 0;
            if (level < 0) for (uint32_t d = 1; d <= L; d++) if (counts[d] > 0) { level = (int)d; break; }
            if (level < 0) break;
            if (level == 0) {
                const uint32_t np = std::min(np_full, pl->n - cur_p0), ns = std::min(ns_full, cfg->samples - cur_s0);
                const uint32_t n = np * ns;
                raygen_kernel<<<std::min<uint32_t>((n + 255) / 256, gs), 256, 0, st>>>(F, pl->d.p, cur_p0, np, cur_s0, ns, level_queue(*sc, 1), counts[1]);
                X.launches++; X.batches++;
                counts[1] += n; X.primary += n;
                sc->samples_issued.store(X.primary, std::memory_order_relaxed);
                cur_s0 += ns;
                if (cur_s0 >= cfg->samples) { cur_s0 = 0; cur_p0 += np; if (cur_p0 >= pl->n) primary_left = false; }
                continue;
            }
            // ---- one wave at depth `level` ----
            const uint32_t d = (uint32_t)level;
            const uint32_t n = std::min(counts[d], chunk);
            const uint32_t q_base = counts[d] - n;
            counts[d] -= n;
            if (ctr_idx + 8 > kCtrPool) {                                    // counter pool used up (16 Ki waves): harvest it and start over
                if ((rc = join_shadow()) || (rc = harvest(ctr_idx))) return rc;
                CU(cudaMemsetAsync(sc->ctr_pool.p, 0, kCtrPool * 4, st)); ctr_idx = 0;
            }
            uint32_t* ctr = sc->ctr_pool.p + ctr_idx; ctr_idx += 8;          // [0] closest work, [1] shadow work, [2] child count, [3] shadow count, [4] slow count, [5] beyond count
            if ((rc = launch_wave(X, d, q_base, n, nullptr, ctr, d + 1 <= L ? counts[d + 1] : 0))) return rc;
            // the host only needs what the SHADE kernel counted (children, shadow rays) to schedule the next wave; the shadow kernels
            // of this wave keep running on their own stream while the next wave's closest-hit kernel starts
            CU(cudaMemcpyAsync(sc->h_ctr + 532, ctr, 16, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            memcpy(sc->h_ctr, sc->h_ctr + 532, 16);
            if (d + 1 <= L) counts[d + 1] += sc->h_ctr[2];
            X.rays_closest += n; X.rays_shadow += sc->h_ctr[3];
            X.waves++;
            if (d + 1 <= L && counts[d + 1] > sc->level_cap) return fail(RTX_E_INVALID, "internal: ray queue overflow");
            if (sc->h_ctr[3] > sc->shadow_cap) return fail(RTX_E_INVALID, "internal: shadow queue overflow");
        }
        if ((rc = join_shadow())) return rc;
        resolve_kernel<<<std::min<uint32_t>((pl->n + 255) / 256, gs), 256, 0, st>>>(F, pl->d.p, pl->n, (uchar4*)d_rgba, (float*)d_normals, (float*)d_depth,
                                                                                (uint32_t*)d_object_ids);
        X.launches++;
        CU(cudaEventRecord(X.event(1), st));
        CU(cudaMemcpyAsync(sc->h_ctr, sc->overflow.p, 16, cudaMemcpyDeviceToHost, st));
        if ((rc = harvest(ctr_idx))) return rc;                               // (synchronises st)
        CU(cudaGetLastError());
        memcpy(ovf, sc->h_ctr, 16);
        if (ovf[0]) return fail(RTX_E_INVALID, "internal: ray queue overflow");
        break;
    }
    if (ovf[1]) return fail(RTX_E_INVALID, "internal: traversal stack overflow");
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        float ms = 0.f; cudaEventElapsedTime(&ms, sc->events[0], sc->events[1]);
        float tc = 0.f, ts = 0.f;
        for (size_t i = 2; i + 3 < X.ev_next + 0 && i + 3 < sc->events.size(); i += 4) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, sc->events[i], sc->events[i + 1]); cudaEventElapsedTime(&b, sc->events[i + 2], sc->events[i + 3]);
            tc += a; ts += b;
        }
        stats->device_ms = ms; stats->closest_ms = tc; stats->shadow_ms = ts; stats->shade_ms = ms - tc - ts;
        stats->rays_closest = X.rays_closest; stats->rays_shadow = X.rays_shadow; stats->primary_samples = X.primary;
        stats->kernel_launches = X.launches; stats->waves = X.waves; stats->batches = X.batches;
        stats->h2d_bytes = h2d; stats->d2h_bytes = sync_free ? (uint64_t)8 * (L + 1) * 4 + 16 : (uint64_t)X.waves * 16 + 16;
        stats->rays_shadow_skipped = ovf[2]; stats->rays_shadow += ovf[2]; stats->rays_shadow_exact = X.ordered ? 0 : X.rays_exact; stats->rays_shadow_beyond = X.ordered ? 0 : X.rays_beyond;
        stats->host_syncs = sync_free ? 1u : X.waves + 1u;
        if (X.want_stats) {
            Counters c; CU(cudaMemcpy(&c, sc->counters.p, sizeof(c), cudaMemcpyDeviceToHost));
            if (getenv("RTX_PHASE_STATS"))
                for (int k = 0; k < 2; k++) {
                    const unsigned long long* p = c.phase[k];
                    fprintf(stderr, "[phase %s] steps/warp-steps %llu | lanes with a ray %.1f/32 | NODE lanes %.1f/32 | LEAF rounds %.2f per step, lanes %.1f/32 | refills %llu, lanes %.1f\n",
                            k ? "shadow" : "closest", p[0], (double)p[1] / std::max(1ull, p[0]), (double)p[2] / std::max(1ull, p[0]), (double)p[3] / std::max(1ull, p[0]),
                            (double)p[4] / std::max(1ull, p[3]), p[5], (double)p[6] / std::max(1ull, p[5]));
                }
            stats->node_visits[0] = c.node_visits[0]; stats->node_visits[1] = c.node_visits[1];
            stats->tri_tests[0] = c.tri_tests[0]; stats->tri_tests[1] = c.tri_tests[1];
            stats->sphere_tests = c.sphere_tests; stats->item_tests = c.item_tests;
        }
    }
    return RTX_OK;
}

void add_stats(RtxStats& a, const RtxStats& b) {
    a.rays_closest += b.rays_closest; a.rays_shadow += b.rays_shadow; a.primary_samples += b.primary_samples;
    for (int k = 0; k < 2; k++) { a.node_visits[k] += b.node_visits[k]; a.tri_tests[k] += b.tri_tests[k]; }
    a.sphere_tests += b.sphere_tests; a.item_tests += b.item_tests; a.kernel_launches += b.kernel_launches;
    a.waves += b.waves; a.batches += b.batches; a.host_syncs += b.host_syncs;
    // times: the frame takes as long as its slowest device
    if (b.device_ms > a.device_ms) { a.device_ms = b.device_ms; a.closest_ms = b.closest_ms; a.shadow_ms = b.shadow_ms; a.shade_ms = b.shade_ms; }
    a.h2d_bytes += b.h2d_bytes; a.d2h_bytes += b.d2h_bytes; a.rays_shadow_skipped += b.rays_shadow_skipped; a.rays_shadow_exact += b.rays_shadow_exact; a.rays_shadow_beyond += b.rays_shadow_beyond;
}

// Frame of a handle: one device, or (rtx_scene_create_multi, shard == NULL) every device its interleaved tiles, all
// resolving into the SAME output buffers — the caller's, on the first device, written by the others through peer memory.
int render_frame_any(RtxScene* sc, const RtxCamera* cam, const RtxConfig* cfg, const RtxShard* shard, void* d_rgba, void* d_normals,
                     void* d_depth, void* d_object_ids, cudaStream_t st, RtxStats* stats) {
    if (!sc || !cam || !cfg) return fail(RTX_E_INVALID, "null argument");
    if (cam->width == 0 || cam->height == 0 || (uint64_t)cam->width * cam->height > 0x7fffffffull) return fail(RTX_E_INVALID, "bad frame size");
    if (cfg->samples == 0 || cfg->samples > 65535) return fail(RTX_E_INVALID, "samples out of range (u16)");
    if (cfg->max_recursion > 254) return fail(RTX_E_INVALID, "max_recursion > 254 is not supported");
    if (sc->replicas.empty()) return render_frame_single(sc, cam, cfg, shard, d_rgba, d_normals, d_depth, d_object_ids, st, stats);
    if (shard) return fail(RTX_E_INVALID, "a multi-device scene shards the frame itself: pass shard = NULL");
    const uint32_t world = 1 + (uint32_t)sc->replicas.size();
    CU(cudaSetDevice(sc->device));
    CU(cudaEventRecord(sc->fence, st));                                   // the replicas start after what the caller queued before this frame
    std::vector<int> rcs(world, RTX_OK); std::vector<std::string> errs(world); std::vector<RtxStats> sts(world);
    auto work = [&](uint32_t k) {
        RtxScene* r = k == 0 ? sc : sc->replicas[k - 1];
        const RtxShard sh{k, world, 8, 4};
        cudaStream_t s = st;
        if (k) {
            cudaSetDevice(r->device);
            s = r->own_stream;
            cudaStreamWaitEvent(s, sc->fence, 0);
        }
        rcs[k] = render_frame_single(r, cam, cfg, &sh, d_rgba, d_normals, d_depth, d_object_ids, s, &sts[k]);
        if (rcs[k]) errs[k] = g_err;
    };
    std::vector<std::thread> th;
    for (uint32_t k = 1; k < world; k++) th.emplace_back(work, k);
    work(0);
    for (std::thread& t : th) t.join();
    CU(cudaSetDevice(sc->device));
    for (uint32_t k = 0; k < world; k++) if (rcs[k]) return fail(rcs[k], errs[k]);
    if (stats) { *stats = sts[0]; for (uint32_t k = 1; k < world; k++) add_stats(*stats, sts[k]); }
    return RTX_OK;
}

int frame_to_host(RtxScene* sc, const RtxCamera* cam, const RtxConfig* cfg, uint8_t* rgba, float* normals, float* depth, uint32_t* object_ids,
                  RtxStats* stats) {
    if (!sc || !cam || !cfg) return fail(RTX_E_INVALID, "null argument");
    CU(cudaSetDevice(sc->device));
    const size_t n = (size_t)cam->width * cam->height;
    int rc;
    if ((rc = sc->o_rgba.alloc(n)) || (rc = sc->o_normals.alloc(3 * n)) || (rc = sc->o_depth.alloc(n)) || (rc = sc->o_ids.alloc(n))) return rc;
    RtxStats st;
    rc = render_frame_any(sc, cam, cfg, nullptr, sc->o_rgba.p, sc->o_normals.p, sc->o_depth.p, sc->o_ids.p, nullptr, &st);
    if (rc) return rc;
    // the four copies are queued together and waited for once (page-locked destinations run at full rate; pageable ones are
    // staged by the driver)
    if (rgba) { CU(cudaMemcpyAsync(rgba, sc->o_rgba.p, n * 4, cudaMemcpyDeviceToHost, 0)); st.d2h_bytes += n * 4; }
    if (normals) { CU(cudaMemcpyAsync(normals, sc->o_normals.p, n * 12, cudaMemcpyDeviceToHost, 0)); st.d2h_bytes += n * 12; }
    if (depth) { CU(cudaMemcpyAsync(depth, sc->o_depth.p, n * 4, cudaMemcpyDeviceToHost, 0)); st.d2h_bytes += n * 4; }
    if (object_ids) { CU(cudaMemcpyAsync(object_ids, sc->o_ids.p, n * 4, cudaMemcpyDeviceToHost, 0)); st.d2h_bytes += n * 4; }
    CU(cudaStreamSynchronize(0));
    if (stats) *stats = st;
    return RTX_OK;
}

}  // namespace

extern "C" {

int rtx_render_frame_device(RtxScene* sc, const RtxCamera* cam, const RtxConfig* cfg, const RtxShard* shard, void* d_rgba, void* d_normals,
                            void* d_depth, void* d_object_ids, void* cuda_stream, RtxStats* stats) {
    if (!sc) return fail(RTX_E_INVALID, "null argument");
    if (sc->running.load()) return fail(RTX_E_BUSY, "a frame is in flight on this handle");
    std::lock_guard<std::mutex> lk(sc->api_mu);
    return render_frame_any(sc, cam, cfg, shard, d_rgba, d_normals, d_depth, d_object_ids, (cudaStream_t)cuda_stream, stats);
}

int rtx_render_frame(RtxScene* sc, const RtxCamera* cam, const RtxConfig* cfg, uint8_t* rgba, float* normals, float* depth, uint32_t* object_ids,
                     RtxStats* stats) {
    if (!sc) return fail(RTX_E_INVALID, "null argument");
    if (sc->running.load()) return fail(RTX_E_BUSY, "a frame is in flight on this handle");
    std::lock_guard<std::mutex> lk(sc->api_mu);
    return frame_to_host(sc, cam, cfg, rgba, normals, depth, object_ids, stats);
}

// ---- non-blocking frame (reference src/renderer.rs:105-231) ---------------------------------------------
int rtx_render_frame_async(RtxScene* sc, const RtxCamera* cam, const RtxConfig* cfg, uint8_t* rgba, float* normals, float* depth, uint32_t* object_ids) {
    if (!sc || !cam || !cfg) return fail(RTX_E_INVALID, "null argument");
    if (sc->running.load()) return fail(RTX_E_BUSY, "a frame is already in flight on this handle");
    if (sc->worker.joinable()) sc->worker.join();
    sc->cancel.store(false); sc->samples_issued.store(0);
    sc->async_pixels = (uint64_t)cam->width * cam->height; sc->async_samples = std::max(1u, cfg->samples);
    sc->async_result = RTX_OK; sc->running.store(true);
    sc->snap_rgba = rgba; sc->snap_normals = normals; sc->snap_depth = depth; sc->snap_ids = object_ids; sc->snap_req.store(0);
    const RtxCamera c = *cam; const RtxConfig g = *cfg;
    sc->worker = std::thread([=]() {
        int rc;
        { std::lock_guard<std::mutex> lk(sc->api_mu); rc = frame_to_host(sc, &c, &g, rgba, normals, depth, object_ids, &sc->async_stats); }
        sc->async_result = rc;
        if (rc) sc->async_err = g_err;                                    // g_err is thread-local: keep the worker's message
        { std::lock_guard<std::mutex> lk(sc->snap_mu); sc->running.store(false, std::memory_order_release); sc->snap_req.store(0); }
        sc->snap_cv.notify_all();                                         // a snapshot waiting on a finished frame sees the final buffers
    });
    return RTX_OK;
}

int rtx_render_poll(RtxScene* sc, uint64_t* pixels_rendered, int* running, int* done, int* result, RtxStats* stats) {
    if (!sc) return fail(RTX_E_INVALID, "null argument");
    const bool run = sc->running.load(std::memory_order_acquire);
    const bool finished = !run && sc->async_pixels > 0 && sc->async_result == RTX_OK && sc->worker.joinable();
    uint64_t px = std::min<uint64_t>(sc->samples_issued.load() / sc->async_samples, sc->async_pixels);
    if (run && px == sc->async_pixels) px = sc->async_pixels - 1;         // is_done() must only turn true when the buffers are filled
    if (finished) px = sc->async_pixels;
    if (pixels_rendered) *pixels_rendered = px;
    if (running) *running = run ? 1 : 0;
    if (done) *done = finished ? 1 : 0;
    if (result) *result = run ? RTX_OK : sc->async_result;
    if (!run && sc->async_result != RTX_OK) g_err = sc->async_err;
    if (stats && !run) *stats = sc->async_stats;
    return RTX_OK;
}

int rtx_render_snapshot(RtxScene* sc, uint64_t* pixels_rendered) {
    if (!sc) return fail(RTX_E_INVALID, "null argument");
    {
        std::unique_lock<std::mutex> lk(sc->snap_mu);
        if (sc->running.load(std::memory_order_acquire)) {
            const uint64_t seq = sc->snap_seq;
            sc->snap_req.store(1);
            sc->snap_cv.wait(lk, [&]() { return sc->snap_seq != seq || !sc->running.load(std::memory_order_acquire); });
        }
    }
    return rtx_render_poll(sc, pixels_rendered, nullptr, nullptr, nullptr, nullptr);
}

int rtx_render_stop(RtxScene* sc) {
    if (!sc) return fail(RTX_E_INVALID, "null argument");
    sc->cancel.store(true);
    if (sc->worker.joinable()) sc->worker.join();
    sc->cancel.store(false);
    return RTX_OK;
}

// ---- probe ------------------------------------------------------------------------------------------
}  // extern "C"
namespace {
constexpr uint32_t kProbeChunk = 1u << 20;
// Probes run on their own stream with their own queue / hit / counter buffers: a pick while a frame is in flight (the GUI
// picks on every click, reference run.rs:1613) touches nothing the frame uses; the scene arrays are read-only for both.
int probe_prepare(RtxScene* sc, bool shadow) {
    RtxScene::Probe& P = sc->probe;
    if (!P.st) CU(cudaStreamCreateWithFlags(&P.st, cudaStreamNonBlocking));
    const uint32_t cap = kProbeChunk;
    int rc;
    if ((rc = P.q_o.alloc(cap)) || (rc = P.q_d.alloc(cap)) || (rc = P.q_m.alloc(cap)) || (rc = P.hits.alloc(cap)) || (rc = P.ctr.alloc(64))) return rc;
    if (shadow && ((rc = P.s_c.alloc(cap)) || (rc = P.s_r.alloc(cap)) || (rc = P.slow.alloc(cap)) || (rc = P.acc.alloc(cap)) || (rc = P.out.alloc(cap)) ||
                   (rc = P.len.alloc(cap)) || (rc = P.recv.alloc(cap)) || (rc = P.res.alloc(cap)) || (rc = P.beyond.alloc(cap)))) return rc;
    P.cap = cap;
    return RTX_OK;
}
}  // namespace
extern "C" {

int rtx_trace_probe(RtxScene* sc, const RtxRay* rays, size_t n, int for_shadow, int stop_on_first_hit, uint32_t depth, RtxHit* hits) {
    if (!sc || (n && (!rays || !hits))) return fail(RTX_E_INVALID, "null argument");
    if (n == 0) return RTX_OK;
    if (n > 0x7fffffffull) return fail(RTX_E_INVALID, "too many rays");
    static_assert(sizeof(ProbeRay) == sizeof(RtxRay) && sizeof(ProbeHit) == sizeof(RtxHit), "probe layouts");
    std::lock_guard<std::mutex> lk(sc->probe.mu);
    CU(cudaSetDevice(sc->device));
    int rc;
    if ((rc = probe_prepare(sc, false)) || (rc = sc->p_rays.alloc(n)) || (rc = sc->p_hits.alloc(n))) return rc;
    RtxScene::Probe& P = sc->probe;
    CU(cudaMemcpyAsync(sc->p_rays.p, rays, n * sizeof(RtxRay), cudaMemcpyHostToDevice, P.st));
    if (!for_shadow && !stop_on_first_hit) {
        // camera-type rays go through the PRODUCTION kernel (persistent closest_kernel), a chunk at a time
        RayQ Q{P.q_o.p, P.q_d.p, P.q_m.p};
        for (size_t off = 0; off < n; off += P.cap) {
            const uint32_t m = (uint32_t)std::min<size_t>(P.cap, n - off);
            CU(cudaMemsetAsync(P.ctr.p, 0, 32, P.st));
            probe_pack_kernel<<<(m + 127) / 128, 128, 0, P.st>>>(sc->p_rays.p + off, m, depth, Q);
            const int blocks = (int)std::min<uint32_t>((m + kTraceBlock - 1) / kTraceBlock, (uint32_t)sc->blocks_closest);
            closest_kernel<false><<<blocks, kTraceBlock, 0, P.st>>>(sc->dev, Q, 0, m, nullptr, P.hits.p, P.ctr.p, nullptr);
            probe_unpack_kernel<<<(m + 127) / 128, 128, 0, P.st>>>(sc->dev, sc->p_rays.p + off, P.hits.p, m, sc->p_hits.p + off);
        }
    } else {
        // trace(.., stop_on_first_hit / for_shadow) returning (t, normal, item, face): the literal reference-order walk.
        // The production shadow kernels only answer lit / occluded; they are probed by rtx_shadow_probe.
        probe_kernel<<<(unsigned)((n + 127) / 128), 128, 0, P.st>>>(sc->dev, sc->p_rays.p, (uint32_t)n, for_shadow, stop_on_first_hit, depth, sc->p_hits.p);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hits, sc->p_hits.p, n * sizeof(RtxHit), cudaMemcpyDeviceToHost, P.st));
    CU(cudaStreamSynchronize(P.st));
    return RTX_OK;
}

int rtx_shadow_probe(RtxScene* sc, const RtxRay* rays, const float* light_distance, const int32_t* receiver_item, size_t n, uint32_t depth,
                     RtxShadowHit* out) {
    if (!sc || (n && (!rays || !out))) return fail(RTX_E_INVALID, "null argument");
    if (n == 0) return RTX_OK;
    static_assert(sizeof(ShadowProbeOut) == sizeof(RtxShadowHit), "shadow probe layout");
    if (receiver_item) for (size_t i = 0; i < n; i++) if (receiver_item[i] >= (int32_t)sc->h_items.size()) return fail(RTX_E_INVALID, "receiver item out of range");
    std::lock_guard<std::mutex> lk(sc->probe.mu);
    CU(cudaSetDevice(sc->device));
    int rc;
    if ((rc = probe_prepare(sc, true))) return rc;
    RtxScene::Probe& P = sc->probe;
    if ((rc = sc->p_rays.alloc(P.cap))) return rc;
    FrameDev F; memset(&F, 0, sizeof(F)); F.accum_c = P.acc.p;
    ShadowQ SQ{P.q_o.p, P.q_d.p, P.s_c.p, P.s_r.p, P.out.p, P.beyond.p};
    ShadowProbeOut* d_out = P.res.p;
    for (size_t off = 0; off < n; off += P.cap) {
        const uint32_t m = (uint32_t)std::min<size_t>(P.cap, n - off);
        CU(cudaMemcpyAsync(sc->p_rays.p, rays + off, (size_t)m * sizeof(RtxRay), cudaMemcpyHostToDevice, P.st));
        if (light_distance) CU(cudaMemcpyAsync(P.len.p, light_distance + off, (size_t)m * 4, cudaMemcpyHostToDevice, P.st));
        if (receiver_item) CU(cudaMemcpyAsync(P.recv.p, receiver_item + off, (size_t)m * 4, cudaMemcpyHostToDevice, P.st));
        CU(cudaMemsetAsync(P.ctr.p, 0, 32, P.st));
        shadow_probe_pack_kernel<<<(m + 127) / 128, 128, 0, P.st>>>(sc->dev, sc->p_rays.p, light_distance ? P.len.p : nullptr, receiver_item ? P.recv.p : nullptr, m, SQ,
                                                                  P.acc.p, P.ctr.p + 3);
        // exactly what a frame launches after its shade kernel (launch_wave)
        shadow_any_kernel<false><<<sc->blocks_shadow, kTraceBlock, 0, P.st>>>(sc->dev, F, SQ, P.ctr.p + 3, m, depth, P.ctr.p + 1, P.slow.p, P.ctr.p + 4, nullptr);
        if (sc->dev.group_root != 0xFFFFFFFFu)
            shadow_beyond_kernel<false><<<sc->sm_count * 8, kTraceBlock, 0, P.st>>>(sc->dev, F, SQ, P.ctr.p + 5, m, depth, P.slow.p, P.ctr.p + 4, nullptr);
        shadow_exact_kernel<false, false><<<sc->sm_count * 8, kTraceBlock, 0, P.st>>>(sc->dev, F, SQ, P.slow.p, P.ctr.p + 4, m, depth, nullptr);
        shadow_probe_unpack_kernel<<<(m + 127) / 128, 128, 0, P.st>>>(P.acc.p, P.out.p, m, d_out);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out + off, d_out, (size_t)m * sizeof(RtxShadowHit), cudaMemcpyDeviceToHost, P.st));
        CU(cudaStreamSynchronize(P.st));
    }
    return RTX_OK;
}

// ---- shards -------------------------------------------------------------------------------------------
uint64_t rtx_shard_pixel_count(uint32_t w, uint32_t h, const RtxShard* shard) {
    RtxShard sh = shard ? *shard : RtxShard{0, 1, 8, 4};
    if (sh.world == 0 || sh.rank >= sh.world || sh.tile_w == 0 || sh.tile_h == 0) return 0;
    const uint32_t tx = (w + sh.tile_w - 1) / sh.tile_w, ty = (h + sh.tile_h - 1) / sh.tile_h;
    uint64_t n = 0;
    for_each_owned_tile(sh, tx * ty, [&](uint32_t t) {
        const uint32_t x0 = (t % tx) * sh.tile_w, y0 = (t / tx) * sh.tile_h;
        n += (uint64_t)(std::min(x0 + sh.tile_w, w) - x0) * (std::min(y0 + sh.tile_h, h) - y0);
    });
    return n;
}
uint64_t rtx_shard_packed_bytes(uint32_t w, uint32_t h, const RtxShard* shard) { return rtx_shard_pixel_count(w, h, shard) * 24; }

static int shard_list(uint32_t w, uint32_t h, const RtxShard* shard, cudaStream_t st, uint32_t** d_list, uint32_t* n) {
    // pack/unpack are scene-independent: keep one small cache of device pixel lists per process
    static std::map<std::tuple<int, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t>, std::pair<uint32_t*, uint32_t>> cache;
    RtxShard sh = shard ? *shard : RtxShard{0, 1, 8, 4};
    if (sh.world == 0 || sh.rank >= sh.world || sh.tile_w == 0 || sh.tile_h == 0) return fail(RTX_E_INVALID, "bad shard");
    int dev = 0; CU(cudaGetDevice(&dev));
    auto key = std::make_tuple(dev, w, h, sh.rank, sh.world, sh.tile_w, sh.tile_h);
    auto it = cache.find(key);
    if (it == cache.end()) {
        std::vector<uint32_t> px;
        const uint32_t tx = (w + sh.tile_w - 1) / sh.tile_w, ty = (h + sh.tile_h - 1) / sh.tile_h;
        for_each_owned_tile(sh, tx * ty, [&](uint32_t t) {
            const uint32_t x0 = (t % tx) * sh.tile_w, y0 = (t / tx) * sh.tile_h;
            for (uint32_t y = y0; y < std::min(y0 + sh.tile_h, h); y++)
                for (uint32_t x = x0; x < std::min(x0 + sh.tile_w, w); x++) px.push_back(y * w + x);
        });
        uint32_t* p = nullptr;
        CU(cudaMalloc(&p, std::max<size_t>(1, px.size()) * 4));
        if (!px.empty()) CU(cudaMemcpy(p, px.data(), px.size() * 4, cudaMemcpyHostToDevice));
        it = cache.emplace(key, std::make_pair(p, (uint32_t)px.size())).first;
    }
    (void)st;
    *d_list = it->second.first; *n = it->second.second;
    return RTX_OK;
}

int rtx_shard_pack(uint32_t w, uint32_t h, const RtxShard* shard, const void* d_rgba, const void* d_normals, const void* d_depth,
                   const void* d_ids, void* d_packed, void* cuda_stream) {
    if (!d_rgba || !d_normals || !d_depth || !d_ids || !d_packed) return fail(RTX_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    uint32_t* list; uint32_t n; int rc = shard_list(w, h, shard, st, &list, &n); if (rc) return rc;
    if (n == 0) return RTX_OK;
    uint8_t* base = (uint8_t*)d_packed;
    pack_kernel<<<std::min<uint32_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(list, n, (const uchar4*)d_rgba, (const float*)d_normals, (const float*)d_depth,
                                                                          (const uint32_t*)d_ids, (uchar4*)base, (float*)(base + (size_t)n * 4),
                                                                          (float*)(base + (size_t)n * 16), (uint32_t*)(base + (size_t)n * 20));
    CU(cudaGetLastError());
    return RTX_OK;
}

int rtx_shard_unpack(uint32_t w, uint32_t h, const RtxShard* shard, const void* d_packed, void* d_rgba, void* d_normals, void* d_depth,
                     void* d_ids, void* cuda_stream) {
    if (!d_rgba || !d_normals || !d_depth || !d_ids || !d_packed) return fail(RTX_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    uint32_t* list; uint32_t n; int rc = shard_list(w, h, shard, st, &list, &n); if (rc) return rc;
    if (n == 0) return RTX_OK;
    const uint8_t* base = (const uint8_t*)d_packed;
    unpack_kernel<<<std::min<uint32_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(list, n, (const uchar4*)base, (const float*)(base + (size_t)n * 4),
                                                                            (const float*)(base + (size_t)n * 16), (const uint32_t*)(base + (size_t)n * 20),
                                                                            (uchar4*)d_rgba, (float*)d_normals, (float*)d_depth, (uint32_t*)d_ids);
    CU(cudaGetLastError());
    return RTX_OK;
}

// Scatter the packed shards of ranks first_rank .. world-1 (rank r at d_packed_all + r * stride_bytes) with ONE kernel.
int rtx_shard_unpack_all(uint32_t w, uint32_t h, uint32_t world, uint32_t tile_w, uint32_t tile_h, uint32_t first_rank, const void* d_packed_all,
                         uint64_t stride_bytes, void* d_rgba, void* d_normals, void* d_depth, void* d_ids, void* cuda_stream) {
    if (!d_rgba || !d_normals || !d_depth || !d_ids || !d_packed_all) return fail(RTX_E_INVALID, "null argument");
    if (world == 0 || first_rank > world || tile_w == 0 || tile_h == 0) return fail(RTX_E_INVALID, "bad shard");
    struct Entry { uint2* list; uint32_t* start; uint32_t n; };
    static std::map<std::tuple<int, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t>, Entry> cache;
    static std::mutex mu;
    int dev = 0; CU(cudaGetDevice(&dev));
    Entry en;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto key = std::make_tuple(dev, w, h, world, tile_w, tile_h, first_rank);
        auto it = cache.find(key);
        if (it == cache.end()) {
            std::vector<uint2> list; std::vector<uint32_t> start(2 * world, 0);
            const uint32_t tx = (w + tile_w - 1) / tile_w, ty = (h + tile_h - 1) / tile_h;
            for (uint32_t r = 0; r < world; r++) {
                start[r] = (uint32_t)list.size();
                uint32_t cnt = 0;
                const RtxShard sh{r, world, tile_w, tile_h};
                for_each_owned_tile(sh, tx * ty, [&](uint32_t t) {
                    const uint32_t x0 = (t % tx) * tile_w, y0 = (t / tx) * tile_h;
                    for (uint32_t y = y0; y < std::min(y0 + tile_h, h); y++)
                        for (uint32_t x = x0; x < std::min(x0 + tile_w, w); x++) { if (r >= first_rank) list.push_back(make_uint2(y * w + x, r)); cnt++; }
                });
                start[world + r] = cnt;
            }
            Entry e{nullptr, nullptr, (uint32_t)list.size()};
            CU(cudaMalloc(&e.list, std::max<size_t>(1, list.size()) * sizeof(uint2)));
            CU(cudaMalloc(&e.start, start.size() * 4));
            if (!list.empty()) CU(cudaMemcpy(e.list, list.data(), list.size() * sizeof(uint2), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(e.start, start.data(), start.size() * 4, cudaMemcpyHostToDevice));
            it = cache.emplace(key, e).first;
        }
        en = it->second;
    }
    if (en.n == 0) return RTX_OK;
    unpack_all_kernel<<<std::min<uint32_t>((en.n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)cuda_stream>>>(en.list, en.start, world, en.n, (const uint8_t*)d_packed_all,
                                                                                                         (size_t)stride_bytes, (uchar4*)d_rgba, (float*)d_normals,
                                                                                                         (float*)d_depth, (uint32_t*)d_ids);
    CU(cudaGetLastError());
    return RTX_OK;
}

// ---- one process, several GPUs ----------------------------------------------------------------------------
// The reference's seam is ONE RendererManager::start per frame (src/renderer.rs:105-172); a host that owns that call
// cannot run one process per GPU.  rtx_scene_create_multi builds the scene once on devices[0], copies the flattened
// arrays device-to-device to the others, and enables peer access so that every device's resolve kernel can store into
// the frame buffers of devices[0].  All other entry points take the returned handle unchanged.
int rtx_scene_create_multi(const RtxSceneDesc* d, const int* devices, uint32_t n_devices, RtxScene** out) {
    if (!d || !out || !devices || n_devices == 0) return fail(RTX_E_INVALID, "null argument");
    *out = nullptr;
    int ndev = rtx_device_count();
    for (uint32_t k = 0; k < n_devices; k++) {
        if (devices[k] < 0 || devices[k] >= ndev) return fail(ndev ? RTX_E_INVALID : RTX_E_NO_DEVICE, ndev ? "bad device ordinal" : "no CUDA device: librtx_b200 has no CPU fallback");
        for (uint32_t j = 0; j < k; j++) if (devices[j] == devices[k]) return fail(RTX_E_INVALID, "device listed twice");
    }
    RtxScene* sc = nullptr;
    int rc = rtx_scene_create(d, devices[0], &sc);
    if (rc) return rc;
    auto bail = [&](int code, const std::string& m) { rtx_scene_destroy(sc); return fail(code, m); };
    if (n_devices > 1 && cudaEventCreateWithFlags(&sc->fence, cudaEventDisableTiming) != cudaSuccess) return bail(RTX_E_CUDA, "cudaEventCreate failed");
    for (uint32_t k = 1; k < n_devices; k++) {
        const int dev = devices[k];
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dev, devices[0]) != cudaSuccess || !can) return bail(RTX_E_CUDA, "no peer access between the listed devices");
        if (cudaSetDevice(dev) != cudaSuccess) return bail(RTX_E_CUDA, "cudaSetDevice failed");
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bail(RTX_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
        RtxScene* r = new RtxScene();
        sc->replicas.push_back(r);                                        // owned from here on (bail destroys it)
        r->device = dev; r->primary = sc;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return bail(RTX_E_CUDA, "cudaGetDeviceProperties failed");
        r->sm_count = prop.multiProcessorCount;
        r->src_items = sc->src_items; r->h_items = sc->h_items; r->src_mats = sc->src_mats; r->mesh_root = sc->mesh_root;
        r->mesh_tri_base = sc->mesh_tri_base; r->mesh_meta = sc->mesh_meta;
        r->n_blas_nodes = sc->n_blas_nodes; r->n_tris = sc->n_tris; r->tlas_cap = sc->tlas_cap; r->n_tlas_nodes = sc->n_tlas_nodes;
        r->texture_bytes = sc->texture_bytes; r->build_ms = sc->build_ms; r->n_enabled_lights = sc->n_enabled_lights;
        r->device_build_ms = sc->device_build_ms; r->flags = sc->flags;
        r->group_root = sc->group_root; r->group_items = sc->group_items; r->n_group_tris = sc->n_group_tris; r->n_fast_nodes = sc->n_fast_nodes;
        const int s0 = devices[0];
        if ((rc = r->nodes.clone_from(sc->nodes, s0, dev)) || (rc = r->tris.clone_from(sc->tris, s0, dev)) || (rc = r->items.clone_from(sc->items, s0, dev)) ||
            (rc = r->tlas_prims.clone_from(sc->tlas_prims, s0, dev)) || (rc = r->fast_prims.clone_from(sc->fast_prims, s0, dev)) || (rc = r->verts.clone_from(sc->verts, s0, dev)) || (rc = r->idx.clone_from(sc->idx, s0, dev)) ||
            (rc = r->uvs.clone_from(sc->uvs, s0, dev)) || (rc = r->uv_idx.clone_from(sc->uv_idx, s0, dev)) || (rc = r->nrms.clone_from(sc->nrms, s0, dev)) ||
            (rc = r->n_idx.clone_from(sc->n_idx, s0, dev)) || (rc = r->mats.clone_from(sc->mats, s0, dev)) || (rc = r->texs.clone_from(sc->texs, s0, dev)) ||
            (rc = r->texels.clone_from(sc->texels, s0, dev)) || (rc = r->lights.clone_from(sc->lights, s0, dev))) {
            std::string keep = g_err; rtx_scene_destroy(sc); g_err = keep; return rc;
        }
        if (cudaDeviceSynchronize() != cudaSuccess) return bail(RTX_E_CUDA, "scene replication failed");
        refresh_dev(*r);
        r->dev.n_lights = sc->dev.n_lights;
        if ((rc = occupancy_blocks(*r))) { std::string keep = g_err; rtx_scene_destroy(sc); g_err = keep; return rc; }
        if (cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(RTX_E_CUDA, "cudaStreamCreate failed");
    }
    cudaSetDevice(devices[0]);
    *out = sc;
    return RTX_OK;
}

int rtx_scene_device_count(const RtxScene* sc) { return sc ? 1 + (int)sc->replicas.size() : 0; }

// ---- frame buffers another process can write (CUDA IPC) -----------------------------------------------------
// One allocation of 24 bytes per pixel on `device`: rgba8[n] | normals f32[3n] | depth f32[n] | object ids u32[n].  The
// owner exports a 64-byte handle; the other ranks of a multi-process render open it and pass the four pointers as the
// outputs of rtx_render_frame_device, so their resolve kernels store finished pixels straight into the owner's memory
// over NVLink: no pack, no collective, no unpack.
struct RtxGBuffer { int device; uint32_t w, h; uint8_t* base; bool owner; };

int rtx_gbuffer_create(int device, uint32_t w, uint32_t h, RtxGBuffer** out) {
    if (!out || !w || !h) return fail(RTX_E_INVALID, "bad argument");
    CU(cudaSetDevice(device));
    RtxGBuffer* g = new RtxGBuffer{device, w, h, nullptr, true};
    cudaError_t e = cudaMalloc(&g->base, (size_t)w * h * 24);
    if (e != cudaSuccess) { delete g; return fail(RTX_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    cudaMemset(g->base, 0, (size_t)w * h * 24);
    *out = g;
    return RTX_OK;
}
int rtx_gbuffer_export(RtxGBuffer* g, uint8_t handle[64]) {
    if (!g || !handle || !g->owner) return fail(RTX_E_INVALID, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(g->device));
    cudaIpcMemHandle_t hdl; CU(cudaIpcGetMemHandle(&hdl, g->base));
    memcpy(handle, &hdl, 64);
    return RTX_OK;
}
int rtx_gbuffer_open(int device, uint32_t w, uint32_t h, const uint8_t handle[64], RtxGBuffer** out) {
    if (!out || !handle || !w || !h) return fail(RTX_E_INVALID, "bad argument");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t hdl; memcpy(&hdl, handle, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess));
    *out = new RtxGBuffer{device, w, h, (uint8_t*)p, false};
    return RTX_OK;
}
int rtx_gbuffer_pointers(const RtxGBuffer* g, void** rgba, void** normals, void** depth, void** ids) {
    if (!g) return fail(RTX_E_INVALID, "null argument");
    const size_t n = (size_t)g->w * g->h;
    if (rgba) *rgba = g->base; if (normals) *normals = g->base + 4 * n; if (depth) *depth = g->base + 16 * n; if (ids) *ids = g->base + 20 * n;
    return RTX_OK;
}
int rtx_gbuffer_download(const RtxGBuffer* g, uint8_t* rgba, float* normals, float* depth, uint32_t* ids, void* cuda_stream) {
    if (!g) return fail(RTX_E_INVALID, "null argument");
    CU(cudaSetDevice(g->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t n = (size_t)g->w * g->h;
    if (rgba) CU(cudaMemcpyAsync(rgba, g->base, 4 * n, cudaMemcpyDeviceToHost, st));
    if (normals) CU(cudaMemcpyAsync(normals, g->base + 4 * n, 12 * n, cudaMemcpyDeviceToHost, st));
    if (depth) CU(cudaMemcpyAsync(depth, g->base + 16 * n, 4 * n, cudaMemcpyDeviceToHost, st));
    if (ids) CU(cudaMemcpyAsync(ids, g->base + 20 * n, 4 * n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return RTX_OK;
}
int rtx_gbuffer_destroy(RtxGBuffer* g) {
    if (!g) return RTX_OK;
    cudaSetDevice(g->device);
    if (g->owner) cudaFree(g->base); else cudaIpcCloseMemHandle(g->base);
    delete g;
    return RTX_OK;
}

// ---- host BVH builder, exposed for the CPU test suite (no device involved) ---------------------------------------
int rtx_bvh_build_probe(const float* boxes /* n x (lo.xyz, hi.xyz) */, uint32_t n, int depth_limit, uint32_t* max_depth, uint32_t* n_nodes,
                        uint32_t* prim_order /* n entries or NULL */) {
    if (!boxes || !n) return fail(RTX_E_INVALID, "bad argument");
    static_assert(sizeof(Aabb3) == 24, "Aabb3 layout");
    WideBvh bvh; build_wide_bvh(reinterpret_cast<const Aabb3*>(boxes), n, bvh, depth_limit);
    if (max_depth) *max_depth = (uint32_t)bvh.max_depth;
    if (n_nodes) *n_nodes = (uint32_t)bvh.nodes.size();
    if (prim_order) memcpy(prim_order, bvh.prim_order.data(), (size_t)n * 4);
    return RTX_OK;
}

// ---- read-bandwidth micro-benchmark (roofline denominator for L2-resident scenes) -----------
__global__ void __launch_bounds__(256) bw_read_kernel(const uint4* __restrict__ p, size_t n, uint32_t iters, uint4* sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (uint32_t it = 0; it < iters; it++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint4 v = __ldcg(p + i);                                // L2 (not L1) — what a node fetch that misses L1 sees
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) *sink = acc;      // keeps the loads alive
}
int rtx_bandwidth_probe(int device, uint64_t bytes, uint32_t iters, float* gbytes_per_s) {
    if (!gbytes_per_s || bytes < 4096 || iters == 0) return fail(RTX_E_INVALID, "bad argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
    uint4* buf = nullptr; uint4* sink = nullptr;
    const size_t n = bytes / 16;
    CU(cudaMalloc(&buf, n * 16)); CU(cudaMalloc(&sink, 16));
    CU(cudaMemset(buf, 1, n * 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8;
    bw_read_kernel<<<blocks, 256>>>(buf, n, 2, sink);                     // warm the cache
    cudaEventRecord(e0);
    bw_read_kernel<<<blocks, 256>>>(buf, n, iters, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf); cudaFree(sink);
    if (e != cudaSuccess) return fail(RTX_E_CUDA, cudaGetErrorString(e));
    *gbytes_per_s = (float)((double)n * 16.0 * iters / (ms * 1e-3) / 1e9);
    return RTX_OK;
}

// ---- post-processing (reference src/post_processing.rs:77-181) ---------------------------------------
__global__ void post_kernel(uint32_t w, uint32_t h, int cavity, int outline, uchar4* rgba, const float* normals, const uint32_t* ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = w * h;
    if (i >= total) return;
    const int x = (int)(i % w), y = (int)(i / w), wi = (int)w;
    uchar4 px = rgba[i];
    float r = (float)px.x, g = (float)px.y, b = (float)px.z;
    auto id_at = [&](int ox, int oy) -> uint32_t { const long idx = (long)(y + oy) * wi + (x + ox); return (idx < 0 || idx >= (long)total) ? 0u : ids[idx]; };
    auto n_at = [&](int ox, int oy, int comp) -> float { const long idx = (long)(y + oy) * wi + (x + ox); return (idx < 0 || idx >= (long)total) ? 0.0f : normals[3 * idx + comp]; };
    if (outline) {                                                        // calculate_outline :96-121
        const uint32_t c = id_at(0, 0);
        const float eq = ((id_at(0, 1) == c ? 0.25f : 0.0f) + (id_at(0, -1) == c ? 0.25f : 0.0f)) + ((id_at(-1, 0) == c ? 0.25f : 0.0f) + (id_at(1, 0) == c ? 0.25f : 0.0f));
        const float o = 1.0f - eq;
        if (o > 0.0f) { r = o * 255.0f; g = o * 255.0f; b = o * 255.0f; }
    }
    if (cavity) {                                                         // calculate_curvature :77-94 (.xz() of the normal: .y of that is z)
        const float diff = (n_at(0, 1, 2) - n_at(0, -1, 2)) + (n_at(1, 0, 0) - n_at(-1, 0, 0));
        auto soft = [](float c, float control) { return (c < 0.5f / control) ? c * (1.0f - c * control) : 0.25f / control; };
        const float curv = diff < 0.0f ? -2.0f * soft(-diff, 1.0f) : 2.0f * soft(diff, 1.15f);
        r *= curv + 1.0f; g *= curv + 1.0f; b *= curv + 1.0f;
    }
    // f32::clamp keeps NaN; `as u8` maps NaN to 0
    auto cl = [](float v) { return v != v ? v : fminf(fmaxf(v, 0.0f), 255.0f); };
    rgba[i] = make_uchar4((unsigned char)as_u8(cl(r)), (unsigned char)as_u8(cl(g)), (unsigned char)as_u8(cl(b)), 255);
}

int rtx_post_process_device(uint32_t w, uint32_t h, int cavity, int outline, void* d_rgba, const void* d_normals, const void* d_depth,
                            const void* d_ids, void* cuda_stream) {
    (void)d_depth;
    if (!d_rgba || !d_normals || !d_ids) return fail(RTX_E_INVALID, "null argument");
    const uint32_t n = w * h;
    if (n == 0) return RTX_OK;
    post_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(w, h, cavity, outline, (uchar4*)d_rgba, (const float*)d_normals, (const uint32_t*)d_ids);
    CU(cudaGetLastError());
    return RTX_OK;
}

}  // extern "C"
