// rt_oracle.cpp — CPU restatement of rustray's per-pixel ray-casting path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the timed CPU baseline.  Only
// tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may load
// liboracle.so; the product (rustray_b200/, librtx_b200.so) never links, imports or calls it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or KATs for this path, cannot be
// compiled here (no Rust toolchain) and delegates its intersection arithmetic to un-vendored
// crates (parry3d 0.13, bvh 0.7, nalgebra 0.32, rand 0.8 — Cargo.toml:18-36, no Cargo.lock).
// Each function below restates the reference line range it cites; third-party semantics are
// restated from their published algorithms and pinned by the self-derived KATs of SURVEY.md
// §8(c) (tests/test_oracle_kat.py).  All paths in citations are relative to the reference repo.
//
// Arithmetic: IEEE f32, compiled with -ffp-contract=off (rustc/LLVM never fuses mul+add), with
// nalgebra's evaluation order for mat*vec (column axpy), dot (a+b+c) and cross.
//
// Build: see oracle/Makefile.  Exports the same signatures as include/rtx.h with prefix oracle_.
#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../include/rtx.h"

namespace {

thread_local std::string g_err;
constexpr float PI = 3.14159265358979323846f;   // std::f32::consts::PI

// ------------------------------------------------------------------------------------------
// nalgebra-style f32 vectors
// ------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { return {x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a / norm(a); }                    // 0-vector -> NaN like nalgebra
inline float rmin(float a, float b) { return std::fmin(a, b); }     // f32::min: returns the non-NaN operand
inline float rmax(float a, float b) { return std::fmax(a, b); }

struct M4 { float m[16]; float at(int r, int c) const { return m[c * 4 + r]; } };   // column-major
// Matrix4 * Vector4: nalgebra gemv = axpy over columns: ((m0*x + m1*y) + m2*z) + m3*w
inline void mul4(const M4& a, float x, float y, float z, float w, float out[4]) {
    for (int r = 0; r < 4; r++) out[r] = ((a.at(r, 0) * x + a.at(r, 1) * y) + a.at(r, 2) * z) + a.at(r, 3) * w;
}
// tran * point.to_homogeneous() -> Point3::from_homogeneous (divide by w when w != 0)
inline V3 xform_point(const M4& a, V3 p) {
    float o[4]; mul4(a, p.x, p.y, p.z, 1.0f, o);
    if (o[3] != 0.0f) return {o[0] / o[3], o[1] / o[3], o[2] / o[3]};
    return {o[0], o[1], o[2]};
}
// tran * vector.to_homogeneous() (w = 0) .xyz()
inline V3 xform_vec(const M4& a, V3 v) {
    float o[4]; mul4(a, v.x, v.y, v.z, 0.0f, o);
    return {o[0], o[1], o[2]};
}

// Rust `as` casts (saturating, NaN -> 0)
inline uint32_t as_u32(float f) { if (!(f > 0.0f)) return 0; if (f >= 4294967296.0f) return 0xFFFFFFFFu; return (uint32_t)f; }
inline int32_t as_i32(float f) { if (f != f) return 0; if (f >= 2147483648.0f) return INT32_MAX; if (f <= -2147483648.0f) return INT32_MIN; return (int32_t)f; }
inline uint8_t as_u8(float f) { if (!(f > 0.0f)) return 0; if (f >= 255.0f) return 255; return (uint8_t)f; }

// helper.rs:11-20
inline bool approx_equal(float a, float b) {
    float factor = 1000000.0f;   // 10f32.powi(6)
    return std::trunc(a * factor) == std::trunc(b * factor);
}

// ------------------------------------------------------------------------------------------
// rand 0.8: StdRng (ChaCha12) seeded with seed_from_u64 (PCG32 expansion) + SliceRandom::shuffle
// ------------------------------------------------------------------------------------------
inline uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16]; memcpy(x, s, sizeof(x));
#define QR(a, b, c, d) \
    x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    for (int i = 0; i < rounds; i += 2) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

struct StdRng {            // rand::rngs::StdRng == rand_chacha::ChaCha12Rng, stream 0
    uint32_t key[8]; uint64_t counter = 0; uint32_t buf[16]; int idx = 16;
    static StdRng seed_from_u64(uint64_t state) {           // rand_core SeedableRng::seed_from_u64
        StdRng r;
        for (int i = 0; i < 8; i++) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            r.key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        }
        return r;
    }
    uint32_t next_u32() {
        if (idx >= 16) { chacha_block(key, counter++, 0, 12, buf); idx = 0; }
        return buf[idx++];
    }
    // UniformInt<u32>::sample_single(0, ubound): widening multiply with the conservative zone
    uint32_t gen_range_u32(uint32_t range) {
        int lz = __builtin_clz(range);
        uint32_t zone = (range << lz) - 1u;
        for (;;) {
            uint64_t m = (uint64_t)next_u32() * (uint64_t)range;
            if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
        }
    }
};

// raytracing.rs:290-313 — the per-pixel sample sub-grid
uint32_t sample_cell_size(uint32_t samples) {
    uint32_t cell = 1;
    if (samples > 1) {
        uint32_t v = (samples + 2) & 0xFFFFu;                // u16 arithmetic
        uint32_t p = 1; while (p < v) p <<= 1;               // next_power_of_two
        cell = p / 2;
    }
    return cell;
}
void build_sample_table(uint32_t samples, uint32_t* cell_size, std::vector<uint16_t>& xy) {
    uint32_t cell = sample_cell_size(samples);
    std::vector<std::pair<uint16_t, uint16_t>> s;
    s.reserve((size_t)cell * cell);
    for (uint32_t x = 0; x < cell; x++) for (uint32_t y = 0; y < cell; y++) s.push_back({(uint16_t)x, (uint16_t)y});
    StdRng rng = StdRng::seed_from_u64(0);
    for (size_t i = s.size() - 1; i >= 1; i--) {             // SliceRandom::shuffle
        size_t j = rng.gen_range_u32((uint32_t)(i + 1));
        std::swap(s[i], s[j]);
    }
    if (s.size() > samples) s.resize(samples);               // truncate
    *cell_size = cell;
    xy.clear();
    for (auto& p : s) { xy.push_back(p.first); xy.push_back(p.second); }
}

// ------------------------------------------------------------------------------------------
// counter-based RNG for Monte-Carlo jitter (extension: the reference uses thread_rng()).
// Must stay identical to rustray_b200/csrc (both are checked against each other in tests).
// ------------------------------------------------------------------------------------------
inline uint32_t mix32(uint32_t h) { h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16; return h; }
// Id of a child ray in the reflection / refraction tree (the Monte-Carlo stream key): 2 * path + which is unique down to depth
// 31; deeper (max_recursion > 30) the 32-bit id would wrap and streams collide, so it is hashed with the depth instead.
inline uint32_t child_path(uint32_t path, uint32_t which, uint32_t depth) {
    return depth < 31u ? path * 2u + which : mix32(path * 0x9E3779B1u + which + (depth << 8));
}
inline float mc_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t path, uint32_t slot) {
    uint32_t h = mix32(seed ^ 0x9E3779B9u);
    h = mix32(h ^ (pixel * 0x85EBCA6Bu + 0x165667B1u));
    h = mix32(h ^ (sample * 0xC2B2AE35u + 0x27D4EB2Fu));
    h = mix32(h ^ (path * 0x9E3779B1u + slot * 0x632BE5ABu + 0x7F4A7C15u));
    return (float)(h >> 8) * (1.0f / 16777216.0f);            // [0,1)
}
struct McCtx { bool on; uint32_t seed, pixel, sample; };

// ------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------
struct Tex { uint32_t w = 0, h = 0; std::vector<uint8_t> rgba; };
struct Bvh2Node { float lo[3], hi[3]; uint32_t left, count; };   // count>0: leaf (left = first)
struct Mesh {
    std::vector<V3> verts; std::vector<uint32_t> idx;
    std::vector<float> uvs; std::vector<uint32_t> uv_idx;
    std::vector<V3> normals; std::vector<uint32_t> n_idx;
    uint32_t n_faces = 0, n_uv_faces = 0, n_normal_faces = 0;
    std::vector<Bvh2Node> nodes; std::vector<uint32_t> order;     // oracle-side acceleration only
};
struct Item {
    RtxItem d; M4 trans, inv; V3 lo, hi;
    // material cache (shape/mod.rs:33-38,769-772): non-texture fields copied by the diff rule
    float c_alpha; bool c_cast_shadow, c_reflection_only, c_backface, c_smooth;
};
struct Scene {
    std::vector<Item> items; std::vector<Mesh> meshes; std::vector<RtxMaterial> mats;
    std::vector<Tex> texs; std::vector<RtxLight> lights;
    bool ball_normal_flip_inside = true;
    bool brute_force = false;
    // item BVH used above BVH_MIN_ITEMS items (Scene::update, scene.rs:1681-1687; world boxes as shape/mod.rs:48-79)
    std::vector<Bvh2Node> item_nodes; std::vector<uint32_t> item_order;
};
struct Ray { V3 o, d; };
struct Counters { uint64_t closest = 0, shadow = 0; };

// ------------------------------------------------------------------------------------------
// parry3d restatements
// ------------------------------------------------------------------------------------------
// parry3d Aabb::cast_local_ray(ray, max_toi, solid) — call sites sphere.rs:51, mesh.rs:58
bool aabb_cast_local_ray(V3 lo, V3 hi, const Ray& r, float max_toi, bool solid, float* out) {
    float tmin = 0.0f, tmax = max_toi;
    const float o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    const float mn[3] = {lo.x, lo.y, lo.z}, mx[3] = {hi.x, hi.y, hi.z};
    for (int i = 0; i < 3; i++) {
        if (d[i] == 0.0f) {
            if (o[i] < mn[i] || o[i] > mx[i]) return false;
        } else {
            float denom = 1.0f / d[i];
            float n = (mn[i] - o[i]) * denom, f = (mx[i] - o[i]) * denom;
            if (n > f) std::swap(n, f);
            tmin = rmax(tmin, n);
            tmax = rmin(tmax, f);
            if (tmin > tmax) return false;
        }
    }
    *out = (tmin == 0.0f && !solid) ? tmax : tmin;
    return true;
}

// parry3d query::details::ray_toi_with_ball + Ball::cast_local_ray_and_get_normal — sphere.rs:60
bool ball_cast(float radius, const Ray& r, float max_toi, bool solid, bool flip_inside, float* toi, V3* n) {
    V3 dc = r.o;                                  // center = origin
    float a = dot(r.d, r.d), b = dot(dc, r.d), c = dot(dc, dc) - radius * radius;
    bool inside; float t;
    if (a == 0.0f) { if (c > 0.0f) return false; inside = true; t = 0.0f; }
    else if (c > 0.0f && b > 0.0f) return false;
    else {
        float delta = b * b - a * c;
        if (delta < 0.0f) return false;
        t = (-b - std::sqrt(delta)) / a;
        if (t <= 0.0f) { inside = true; t = solid ? 0.0f : (-b + std::sqrt(delta)) / a; }
        else inside = false;
    }
    if (!(t <= max_toi)) return false;
    V3 pos = r.o + r.d * t;
    V3 nn = normalize(pos);
    *n = (inside && flip_inside) ? -nn : nn;
    *toi = t;
    return true;
}

// parry3d query::details::local_ray_intersection_with_triangle (Ericson) — via TriMesh, mesh.rs:67
// returns toi, geometric normal facing the ray origin, fid 0 (front) / 1 (back).  `solid` ignored.
inline bool tri_cast(V3 a, V3 b, V3 c, const Ray& r, float* toi, V3* nrm, int* fid) {
    V3 ab = b - a, ac = c - a;
    V3 n = cross(ab, ac);
    float d = dot(n, r.d);
    if (d == 0.0f) return false;
    V3 ap = r.o - a;
    float t = dot(ap, n);
    if ((t < 0.0f && d < 0.0f) || (t > 0.0f && d > 0.0f)) return false;
    *fid = d < 0.0f ? 0 : 1;
    d = std::fabs(d);
    V3 e = cross(-r.d, ap);
    float v, w;
    if (t < 0.0f) {
        v = -dot(ac, e); if (v < 0.0f || v > d) return false;
        w = dot(ab, e);  if (w < 0.0f || v + w > d) return false;
        float invd = 1.0f / d; *toi = -t * invd; *nrm = -normalize(n);
    } else {
        v = dot(ac, e);  if (v < 0.0f || v > d) return false;
        w = -dot(ab, e); if (w < 0.0f || v + w > d) return false;
        float invd = 1.0f / d; *toi = t * invd; *nrm = normalize(n);
    }
    return true;
}

// oracle-side acceleration: median-split BVH2 over triangles.  It only prunes; the accepted hit is
// min toi with ties resolved to the lowest face index (parry's Qbvh best-first order among equal
// toi is not reproducible; ties are excluded from the id gate — SURVEY.md §7 "Tie-breaking").
void build_bvh2(Mesh& m) {
    uint32_t nf = m.n_faces;
    m.order.resize(nf);
    std::vector<V3> cen(nf), tlo(nf), thi(nf);
    for (uint32_t f = 0; f < nf; f++) {
        m.order[f] = f;
        V3 a = m.verts[m.idx[3 * f]], b = m.verts[m.idx[3 * f + 1]], c = m.verts[m.idx[3 * f + 2]];
        tlo[f] = {rmin(a.x, rmin(b.x, c.x)), rmin(a.y, rmin(b.y, c.y)), rmin(a.z, rmin(b.z, c.z))};
        thi[f] = {rmax(a.x, rmax(b.x, c.x)), rmax(a.y, rmax(b.y, c.y)), rmax(a.z, rmax(b.z, c.z))};
        cen[f] = (tlo[f] + thi[f]) * 0.5f;
    }
    m.nodes.clear(); m.nodes.reserve(2 * nf);
    struct Job { uint32_t node, first, count; };
    std::vector<Job> st;
    m.nodes.push_back({});
    st.push_back({0, 0, nf});
    while (!st.empty()) {
        Job j = st.back(); st.pop_back();
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t i = j.first; i < j.first + j.count; i++) {
            uint32_t f = m.order[i];
            const float l[3] = {tlo[f].x, tlo[f].y, tlo[f].z}, h[3] = {thi[f].x, thi[f].y, thi[f].z}, c[3] = {cen[f].x, cen[f].y, cen[f].z};
            for (int k = 0; k < 3; k++) { lo[k] = rmin(lo[k], l[k]); hi[k] = rmax(hi[k], h[k]); clo[k] = rmin(clo[k], c[k]); chi[k] = rmax(chi[k], c[k]); }
        }
        Bvh2Node nd; memcpy(nd.lo, lo, 12); memcpy(nd.hi, hi, 12);
        int ax = 0; if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1; if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
        if (j.count <= 4 || !(chi[ax] > clo[ax])) { nd.left = j.first; nd.count = j.count; m.nodes[j.node] = nd; continue; }
        uint32_t mid = j.first + j.count / 2;
        std::nth_element(m.order.begin() + j.first, m.order.begin() + mid, m.order.begin() + j.first + j.count,
                         [&](uint32_t p, uint32_t q) { const float* a = &cen[p].x; const float* b = &cen[q].x; return a[ax] < b[ax]; });
        nd.left = (uint32_t)m.nodes.size(); nd.count = 0; m.nodes[j.node] = nd;
        m.nodes.push_back({}); m.nodes.push_back({});
        st.push_back({nd.left, j.first, mid - j.first});
        st.push_back({nd.left + 1, mid, j.first + j.count - mid});
    }
}

// conservative slab test.  The triangle toi (Ericson form) and the slab distances are rounded differently, so a
// box is only skipped when it lies beyond the best toi by more than a 1e-5 relative margin on both ends:
// pruning may never drop a hit the brute-force loop would keep (tests: BVH == brute force, bit for bit).
inline bool box_hit(const Bvh2Node& n, const float o[3], const float inv[3], float tbest) {
    float t0 = 0.0f, t1 = tbest;
    for (int k = 0; k < 3; k++) {
        float a = (n.lo[k] - o[k]) * inv[k], b = (n.hi[k] - o[k]) * inv[k];
        float tn = rmin(a, b), tf = rmax(a, b);
        tn = tn - std::fabs(tn) * 1e-5f; tf = tf + std::fabs(tf) * 1e-5f;
        t0 = rmax(t0, tn); t1 = rmin(t1, tf);
    }
    return t0 <= t1;
}

// parry3d TriMesh::cast_local_ray_and_get_normal — closest triangle, feature = Face(i) / Face(i+n)
bool trimesh_cast(const Mesh& m, const Ray& r, bool brute, float* toi, V3* nrm, uint32_t* feature) {
    float best = std::numeric_limits<float>::max(); uint32_t bf = 0xFFFFFFFFu; V3 bn{0, 0, 0}; int bfid = 0;
    auto test = [&](uint32_t f) {
        float t; V3 n; int fid;
        if (!tri_cast(m.verts[m.idx[3 * f]], m.verts[m.idx[3 * f + 1]], m.verts[m.idx[3 * f + 2]], r, &t, &n, &fid)) return;
        if (!(t <= std::numeric_limits<float>::max())) return;        // inter.toi <= max_toi
        if (bf == 0xFFFFFFFFu || t < best || (t == best && f < bf)) { best = t; bf = f; bn = n; bfid = fid; }
    };
    if (brute || m.nodes.empty()) {
        for (uint32_t f = 0; f < m.n_faces; f++) test(f);
    } else {
        const float o[3] = {r.o.x, r.o.y, r.o.z};
        const float inv[3] = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
        uint32_t stack[64]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Bvh2Node& n = m.nodes[stack[--sp]];
            if (!box_hit(n, o, inv, bf == 0xFFFFFFFFu ? std::numeric_limits<float>::max() : best)) continue;
            if (n.count) { for (uint32_t i = 0; i < n.count; i++) test(m.order[n.left + i]); }
            else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
        }
    }
    if (bf == 0xFFFFFFFFu) return false;
    *toi = best; *nrm = bn; *feature = bfid == 1 ? bf + m.n_faces : bf;
    return true;
}

// ------------------------------------------------------------------------------------------
// Shape trait (shape/mod.rs:19-46) for Sphere (sphere.rs) and Mesh (mesh.rs)
// ------------------------------------------------------------------------------------------
inline Ray get_inverse_ray(const Item& it, const Ray& r) {          // shape/mod.rs:755-761
    return {xform_point(it.inv, r.o), xform_vec(it.inv, r.d)};
}
inline bool item_solid(const Item& it, bool force_not_solid) {      // sphere.rs:49-50, mesh.rs:55-56
    return !(it.c_alpha < 1.0f /* || cache.has_texture(Alpha): cache never holds textures */) && it.c_backface && !force_not_solid;
}
bool intersect_b_box(const Item& it, const Ray& r, bool force_not_solid, float* dist) {   // sphere.rs:45-52, mesh.rs:51-59
    Ray ri = get_inverse_ray(it, r);
    return aabb_cast_local_ray(it.lo, it.hi, ri, std::numeric_limits<float>::max(), item_solid(it, force_not_solid), dist);
}

// Mesh::get_uv / get_normal share the area-ratio weights (mesh.rs:105-161, 204-259)
inline void area_weights(const Scene& sc, const Item& it, V3 hit, uint32_t f_id, float w[3]) {
    const Mesh& m = sc.meshes[it.d.mesh];
    V3 hl = xform_point(it.inv, hit);
    V3 a = m.verts[m.idx[3 * f_id]], b = m.verts[m.idx[3 * f_id + 1]], c = m.verts[m.idx[3 * f_id + 2]];
    V3 f1 = a - hl, f2 = b - hl, f3 = c - hl;
    float area = norm(cross(a - b, a - c));
    w[0] = norm(cross(f2, f3)) / area; w[1] = norm(cross(f3, f1)) / area; w[2] = norm(cross(f1, f2)) / area;
}
V3 mesh_get_normal(const Scene& sc, const Item& it, V3 hit, uint32_t face_id) {             // mesh.rs:204-259
    const Mesh& m = sc.meshes[it.d.mesh];
    uint32_t f_id = face_id % m.n_faces;
    float w[3]; area_weights(sc, it, hit, f_id, w);
    V3 a = m.normals[m.n_idx[3 * f_id]], b = m.normals[m.n_idx[3 * f_id + 1]], c = m.normals[m.n_idx[3 * f_id + 2]];
    V3 p1 = a * w[0], p2 = b * w[1], p3 = c * w[2];
    return {p1.x + p2.x + p3.x, p1.y + p2.y + p3.y, p1.z + p2.z + p3.z};
}
void item_get_uv(const Scene& sc, const Item& it, V3 hit, uint32_t face_id, float uv[2]) {
    if (it.d.shape == RTX_SHAPE_SPHERE) {                                                    // sphere.rs:69-99
        V3 hl = xform_point(it.inv, hit);
        float theta = std::atan2(-hl.z, hl.x);
        float u = (theta + PI) / (2.0f * PI);
        float phi = std::acos((-hl.y) / it.d.radius);
        float v = phi / PI;
        uv[0] = u; uv[1] = -v; return;
    }
    const Mesh& m = sc.meshes[it.d.mesh];                                                    // mesh.rs:105-161
    uint32_t f_id = face_id % m.n_faces;
    if ((int32_t)m.n_uv_faces - 1 < (int32_t)f_id || (int32_t)m.n_faces - 1 < (int32_t)f_id) { uv[0] = 0; uv[1] = 0; return; }
    float w[3]; area_weights(sc, it, hit, f_id, w);
    const float* a = &m.uvs[2 * m.uv_idx[3 * f_id]]; const float* b = &m.uvs[2 * m.uv_idx[3 * f_id + 1]]; const float* c = &m.uvs[2 * m.uv_idx[3 * f_id + 2]];
    float ux = a[0] * w[0] + b[0] * w[1] + c[0] * w[2];
    float uy = a[1] * w[0] + b[1] * w[1] + c[1] * w[2];
    uv[0] = ux; uv[1] = -uy;
}

bool item_intersect(const Scene& sc, const Item& it, const Ray& r, bool force_not_solid, float* toi, V3* normal, uint32_t* face) {
    Ray ri = get_inverse_ray(it, r);
    bool solid = item_solid(it, force_not_solid);
    if (it.d.shape == RTX_SHAPE_SPHERE) {                                                    // sphere.rs:54-67
        V3 n;
        if (!ball_cast(it.d.radius, ri, std::numeric_limits<float>::max(), solid, sc.ball_normal_flip_inside, toi, &n)) return false;
        *normal = normalize(xform_vec(it.trans, n)); *face = 0; return true;
    }
    const Mesh& m = sc.meshes[it.d.mesh];                                                    // mesh.rs:61-103
    V3 n; uint32_t feature;
    if (!trimesh_cast(m, ri, sc.brute_force, toi, &n, &feature)) return false;
    V3 nn;
    if (it.c_smooth && !m.normals.empty() && m.n_normal_faces > 0) {
        V3 hit = r.o + r.d * (*toi);
        nn = mesh_get_normal(sc, it, hit, feature);
        nn = normalize(xform_vec(it.trans, nn));
        if (feature >= m.n_faces) nn = -nn;                                                  // is_backface
    } else nn = normalize(xform_vec(it.trans, n));
    if (it.d.flip_normals) nn = -nn;
    *normal = nn; *face = feature; return true;
}

// ------------------------------------------------------------------------------------------
// Raytracing (raytracing.rs)
// ------------------------------------------------------------------------------------------
struct Hit { float t; V3 n; int item; uint32_t face; };

// Scene::update's item BVH (scene.rs:1681-1687, bvh crate 0.7 BVHNode::build over the Bounded impl of shape/mod.rs:48-79:
// the eight corners of the local box through `trans`).  Only a candidate filter: boxes are padded so that no item that
// passes the exact local-space slab test (intersect_b_box) is ever dropped, and candidates are handed out in scene order,
// so trace() returns exactly what the all-items loop returns — the crate's DFS order (ties between equal bbox
// distances only) is not reproduced.  Median split on the widest centroid axis, <= 2 items per leaf.
constexpr size_t BVH_MIN_ITEMS = 50;                                       // raytracing.rs:23
void build_item_bvh(Scene& sc) {
    sc.item_nodes.clear(); sc.item_order.clear();
    const size_t n = sc.items.size();
    if (n <= BVH_MIN_ITEMS) return;
    std::vector<float> lo(3 * n), hi(3 * n), ce(3 * n);
    for (size_t i = 0; i < n; i++) {
        const Item& it = sc.items[i];
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int c = 0; c < 8; c++) {
            V3 p = {(c & 1) ? it.hi.x : it.lo.x, (c & 2) ? it.hi.y : it.lo.y, (c & 4) ? it.hi.z : it.lo.z};
            float o[4]; mul4(it.trans, p.x, p.y, p.z, 1.0f, o);
            for (int k = 0; k < 3; k++) { mn[k] = rmin(mn[k], o[k]); mx[k] = rmax(mx[k], o[k]); }
        }
        for (int k = 0; k < 3; k++) {
            float pad = 1e-4f * (std::fabs(mn[k]) + std::fabs(mx[k])) + 1e-5f;
            lo[3 * i + k] = mn[k] - pad; hi[3 * i + k] = mx[k] + pad; ce[3 * i + k] = 0.5f * (mn[k] + mx[k]);
        }
    }
    sc.item_order.resize(n);
    for (size_t i = 0; i < n; i++) sc.item_order[i] = (uint32_t)i;
    struct Job { uint32_t node, first, count; };
    std::vector<Job> jobs; sc.item_nodes.push_back({}); jobs.push_back({0, 0, (uint32_t)n});
    while (!jobs.empty()) {
        Job j = jobs.back(); jobs.pop_back();
        Bvh2Node nd; for (int k = 0; k < 3; k++) { nd.lo[k] = INFINITY; nd.hi[k] = -INFINITY; }
        float cl[3] = {INFINITY, INFINITY, INFINITY}, ch[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t q = j.first; q < j.first + j.count; q++) {
            uint32_t i = sc.item_order[q];
            for (int k = 0; k < 3; k++) { nd.lo[k] = rmin(nd.lo[k], lo[3 * i + k]); nd.hi[k] = rmax(nd.hi[k], hi[3 * i + k]); cl[k] = rmin(cl[k], ce[3 * i + k]); ch[k] = rmax(ch[k], ce[3 * i + k]); }
        }
        int ax = 0; for (int k = 1; k < 3; k++) if (ch[k] - cl[k] > ch[ax] - cl[ax]) ax = k;
        if (j.count <= 2 || !(ch[ax] > cl[ax])) { nd.left = j.first; nd.count = j.count; sc.item_nodes[j.node] = nd; continue; }
        uint32_t mid = j.first + j.count / 2;
        std::nth_element(sc.item_order.begin() + j.first, sc.item_order.begin() + mid, sc.item_order.begin() + j.first + j.count,
                         [&](uint32_t a, uint32_t b) { return ce[3 * a + ax] < ce[3 * b + ax]; });
        nd.left = (uint32_t)sc.item_nodes.size(); nd.count = 0;
        sc.item_nodes[j.node] = nd;
        sc.item_nodes.push_back({}); sc.item_nodes.push_back({});
        jobs.push_back({nd.left, j.first, mid - j.first}); jobs.push_back({nd.left + 1, mid, j.first + j.count - mid});
    }
}
// Scene::get_possible_hits_by_ray (scene.rs:1715-1722): candidate items, returned in scene order
inline void item_bvh_candidates(const Scene& sc, const Ray& r, std::vector<uint32_t>& out) {
    out.clear();
    const float o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    uint32_t stack[64]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const Bvh2Node& n = sc.item_nodes[stack[--sp]];
        float t0 = 0.0f, t1 = std::numeric_limits<float>::max(); bool miss = false;
        for (int k = 0; k < 3 && !miss; k++) {
            if (d[k] == 0.0f) { if (o[k] < n.lo[k] || o[k] > n.hi[k]) miss = true; continue; }
            float inv = 1.0f / d[k], a = (n.lo[k] - o[k]) * inv, b = (n.hi[k] - o[k]) * inv;
            if (a > b) std::swap(a, b);
            t0 = rmax(t0, a); t1 = rmin(t1, b);
            if (t0 > t1 * 1.00001f + 1e-6f) miss = true;
        }
        if (miss) continue;
        if (n.count) { for (uint32_t q = 0; q < n.count; q++) out.push_back(sc.item_order[n.left + q]); }
        else if (sp + 2 <= 64) { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
    }
    std::sort(out.begin(), out.end());
}

// Raytracing::trace (raytracing.rs:429-490).  items.len() > BVH_MIN_ITEMS takes its candidates from the item BVH.
bool trace(const Scene& sc, const Ray& r, bool stop_on_first_hit, bool for_shadow, uint32_t depth, Hit* out, Counters* cnt) {
    if (cnt) { if (for_shadow) cnt->shadow++; else cnt->closest++; }
    struct Cand { int item; float dist; };
    Cand small[64]; std::vector<Cand> big; Cand* hits = small; size_t nh = 0;
    const bool use_bvh = !sc.item_nodes.empty() && !sc.brute_force;
    static thread_local std::vector<uint32_t> cand;
    if (use_bvh) item_bvh_candidates(sc, r, cand);
    const size_t n_cand = use_bvh ? cand.size() : sc.items.size();
    if (n_cand > 64) { big.resize(n_cand); hits = big.data(); }
    for (size_t q = 0; q < n_cand; q++) {
        const size_t i = use_bvh ? cand[q] : q;
        const Item& it = sc.items[i];
        float dist;
        if (intersect_b_box(it, r, for_shadow, &dist)) {
            if (it.d.visible && it.c_alpha > 0.0f && (!for_shadow || it.c_cast_shadow) && (!it.c_reflection_only || depth > 1))
                hits[nh++] = {(int)i, dist};
        }
    }
    if (nh == 0) return false;
    for (size_t i = 0; i < nh; i++) assert(hits[i].dist == hits[i].dist && "partial_cmp().unwrap() would abort");
    std::stable_sort(hits, hits + nh, [](const Cand& a, const Cand& b) { return a.dist < b.dist; });   // :466
    bool have = false; Hit best{};
    for (size_t i = 0; i < nh; i++) {
        float t; V3 n; uint32_t f;
        if (item_intersect(sc, sc.items[hits[i].item], r, for_shadow, &t, &n, &f)) {
            if (!have || t < best.t) { best = {t, n, hits[i].item, f}; have = true; }
        }
        if (have && stop_on_first_hit) { *out = best; return true; }
    }
    if (have) *out = best;
    return have;
}

// raytracing.rs:492-563
inline Ray create_reflection(V3 normal, V3 incident, V3 p) { return {p + normal * 0.001f, incident - (2.0f * dot(incident, normal)) * normal}; }
inline bool create_transmission(V3 normal, V3 incident, V3 p, float index, Ray* out) {
    V3 ref_n = normal; float eta_t = index, eta_i = 1.0f; float i_dot_n = dot(incident, normal);
    if (i_dot_n < 0.0f) i_dot_n = -i_dot_n; else { ref_n = -normal; eta_t = 1.0f; eta_i = index; }
    float eta = eta_i / eta_t;
    float k = 1.0f - (eta * eta) * (1.0f - i_dot_n * i_dot_n);
    if (k < 0.0f) return false;
    out->o = p + ref_n * (-0.001f);
    out->d = (incident + i_dot_n * ref_n) * eta - ref_n * std::sqrt(k);
    return true;
}
inline float fresnel(V3 incident, V3 normal, float index) {
    float i_dot_n = dot(incident, normal);
    float eta_i = 1.0f, eta_t = index;
    if (i_dot_n > 0.0f) { eta_i = eta_t; eta_t = 1.0f; }
    float sin_t = eta_i / eta_t * std::sqrt(rmax(1.0f - i_dot_n * i_dot_n, 0.0f));
    if (sin_t > 1.0f) return 1.0f;
    float cos_t = std::sqrt(rmax(1.0f - sin_t * sin_t, 0.0f));
    float cos_i = std::fabs(cos_t);                                   // sic (:558)
    float r_s = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    float r_p = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (r_s * r_s + r_p * r_p) / 2.0f;
}
// raytracing.rs:565-626; the two thread_rng draws are replaced by mc_uniform(.., slot), mc_uniform(.., slot+1)
V3 jitter(V3 dir, float spread, const McCtx& mc, uint32_t path, uint32_t slot) {
    if (spread <= 0.0f) return dir;
    V3 b3 = normalize(dir);
    V3 diff = std::fabs(b3.x) < 0.5f ? v3(1, 0, 0) : v3(0, 1, 0);
    V3 b1 = normalize(cross(b3, diff));
    V3 b2 = cross(b1, b3);
    float z_lo = std::cos(spread * PI);
    if (!(z_lo < 1.0f)) return dir;                                    // z_range.is_empty()
    float z = z_lo + mc_uniform(mc.seed, mc.pixel, mc.sample, path, slot) * (1.0f - z_lo);
    float r = std::sqrt(1.0f - z * z);
    float theta = -PI + mc_uniform(mc.seed, mc.pixel, mc.sample, path, slot + 1) * (PI - (-PI));
    float x = r * std::cos(theta), y = r * std::sin(theta);
    return normalize(x * b1 + y * b2 + z * b3);
}

struct V4 { float x, y, z, w; };
inline V4 texel(const Tex& t, uint32_t x, uint32_t y) {
    const uint8_t* p = &t.rgba[((size_t)y * t.w + x) * 4];
    return {(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f};
}
inline uint32_t wrap(float val, uint32_t bound) {                     // raytracing.rs:629-642
    int32_t sb = (int32_t)bound;
    float fc = val * (float)bound;
    int32_t w = as_i32(fc) % sb;
    return w < 0 ? (uint32_t)(w + sb) : (uint32_t)w;
}
inline float lerp(float a, float b, float f) { return a + f * (b - a); }   // helper.rs:35-38
V4 tex_interpolate(const Tex& t, float xf, float yf) {                // shape/mod.rs:542-629
    float x = xf * (float)t.w, y = yf * (float)t.h;
    if (x < 0.0f) x = x + (float)t.w;
    if (y < 0.0f) y = y + (float)t.h;
    uint32_t x0 = as_u32(std::floor(x)), x1 = as_u32(std::ceil(x));
    uint32_t y0 = as_u32(std::floor(y)), y1 = as_u32(std::ceil(y));
    if (x0 >= t.w) x0 = t.w - 1; if (y0 >= t.h) y0 = t.h - 1;
    if (x1 >= t.w) x1 = t.w - 1; if (y1 >= t.h) y1 = t.h - 1;
    float fx = x - (float)x0, fy = y - (float)y0;
    V4 p0 = texel(t, x0, y0), p1 = texel(t, x1, y0), p2 = texel(t, x0, y1), p3 = texel(t, x1, y1);
    V4 a = {lerp(p0.x, p1.x, fx), lerp(p0.y, p1.y, fx), lerp(p0.z, p1.z, fx), lerp(p0.w, p1.w, fx)};
    V4 b = {lerp(p2.x, p3.x, fx), lerp(p2.y, p3.y, fx), lerp(p2.z, p3.z, fx), lerp(p2.w, p3.w, fx)};
    return {lerp(a.x, b.x, fy), lerp(a.y, b.y, fy), lerp(a.z, b.z, fy), lerp(a.w, b.w, fy)};
}
// Raytracing::get_tex_color (raytracing.rs:651-675)
bool get_tex_color(const Scene& sc, const RtxMaterial& m, const float* uv, int tex_type, V4* out) {
    int ti = m.texture[tex_type];
    if (ti < 0 || !uv) return false;
    const Tex& t = sc.texs[ti];
    if (t.w == 0) return false;
    if (m.texture_filtering_nearest) *out = texel(t, wrap(uv[0], t.w), wrap(uv[1], t.h));
    else *out = tex_interpolate(t, uv[0], uv[1]);
    return true;
}
inline bool has_any_texture(const Scene& sc, const RtxMaterial& m) {
    for (int i = 0; i < RTX_TEX_COUNT; i++) if (m.texture[i] >= 0 && sc.texs[m.texture[i]].w > 0) return true;
    return false;
}
V4 get_item_color(const Scene& sc, const RtxMaterial& m, const float* uv, int which) {   // raytracing.rs:677-712
    const float* c = which == 0 ? m.ambient_color : which == 1 ? m.base_color : m.specular_color;
    int tt = which == 0 ? RTX_TEX_AMBIENT_EMISSIVE : which == 1 ? RTX_TEX_BASE : RTX_TEX_SPECULAR;
    V4 col = {c[0], c[1], c[2], 1.0f}, tc;
    if (get_tex_color(sc, m, uv, tt, &tc)) { col.x *= tc.x; col.y *= tc.y; col.z *= tc.z; col.w *= tc.w; }
    return col;
}

struct Radiance { V3 color; float depth; V3 normal; uint32_t id; };

// Raytracing::get_color_depth_normal_id (raytracing.rs:720-998) — recursive, as written.
Radiance shade(const Scene& sc, const RtxConfig& cfg, Ray ray, uint32_t depth, const McCtx& mc, uint32_t path, Counters* cnt) {
    Ray r = ray; r.d = normalize(r.d);
    Radiance out{{0, 0, 0}, 0.0f, {0, 0, 0}, 0};
    Hit h;
    if (!trace(sc, r, false, false, depth, &h, cnt)) return out;
    const Item& item = sc.items[h.item];
    const RtxMaterial& mat = sc.mats[item.d.material];
    float hit_dist = h.t; V3 normal = h.n; uint32_t face_id = h.face;
    out.depth = hit_dist; out.normal = normal; out.id = item.d.id;
    V3 color{0, 0, 0};
    V3 surface_normal = normal;
    V3 hit_point = r.o + (r.d * hit_dist);

    float uvs[2]; const float* uv = nullptr;
    if (has_any_texture(sc, mat)) { item_get_uv(sc, item, hit_point, face_id, uvs); uv = uvs; }

    V4 ntc;
    if (get_tex_color(sc, mat, uv, RTX_TEX_NORMAL, &ntc)) {                                  // :757-784
        V3 tangent = cross(normal, v3(0, 1, 0));
        if (norm(tangent) <= 0.0001f) tangent = cross(normal, v3(0, 0, 1));
        tangent = normalize(tangent);
        V3 bitangent = normalize(cross(normal, tangent));
        V3 nm = {ntc.x * 2.0f - 1.0f, ntc.y * 2.0f - 1.0f, ntc.z * 2.0f - 1.0f};
        nm.x *= mat.normal_map_strength; nm.y *= mat.normal_map_strength;
        nm = normalize(nm);
        // Matrix3::from_columns([t, b, n]) * nm  (column axpy)
        V3 tn = {(tangent.x * nm.x + bitangent.x * nm.y) + normal.x * nm.z,
                 (tangent.y * nm.x + bitangent.y * nm.y) + normal.y * nm.z,
                 (tangent.z * nm.x + bitangent.z * nm.y) + normal.z * nm.z};
        surface_normal = normalize(tn);
    }
    V4 rtc; bool has_rough_tex = get_tex_color(sc, mat, uv, RTX_TEX_ROUGHNESS, &rtc);        // :787-798
    if (cfg.monte_carlo && mat.monte_carlo && (mat.roughness > 0.0f || has_rough_tex)) {
        float roughness = mat.roughness;
        if (has_rough_tex) roughness = (1.0f / PI / 2.0f) * rtc.x;
        surface_normal = jitter(surface_normal, roughness, mc, path, 0);
    }
    V4 ambient_color = get_item_color(sc, mat, uv, 0);
    V4 base_color = get_item_color(sc, mat, uv, 1);
    V4 specular_color = get_item_color(sc, mat, uv, 2);
    float alpha = mat.alpha * base_color.w;                                                  // :806-811
    V4 atc; if (get_tex_color(sc, mat, uv, RTX_TEX_ALPHA, &atc)) alpha *= atc.x;

    uint32_t li = 0;
    for (const RtxLight& light : sc.lights) {                                                // :814-920
        uint32_t light_slot = 2 + 2 * li; li++;
        if (!light.enabled) continue;
        V3 lpos = {light.pos[0], light.pos[1], light.pos[2]}, ldir = {light.dir[0], light.dir[1], light.dir[2]};
        V3 dtl = light.light_type == RTX_LIGHT_DIRECTIONAL ? normalize(-ldir) : normalize(lpos - hit_point);
        float dot_light = rmax(dot(surface_normal, dtl), 0.0f);
        V3 base = {base_color.x * dot_light, base_color.y * dot_light, base_color.z * dot_light};
        V3 mi = -dtl;
        V3 reflect_dir = mi - 2.0f * dot(surface_normal, mi) * surface_normal;              // reflect() :714-718
        V3 view_dir = normalize(-r.d);
        float spec_dot = rmax(dot(reflect_dir, view_dir), 0.0f);
        float light_power = std::pow(spec_dot, mat.shininess);
        V3 specular = {specular_color.x * light_power, specular_color.y * light_power, specular_color.z * light_power};
        float intensity;
        if (light.light_type == RTX_LIGHT_DIRECTIONAL) intensity = light.intensity;
        else {
            float r2 = norm(lpos - hit_point);
            intensity = light.intensity / (4.0f * PI * r2);
            if (light.light_type == RTX_LIGHT_SPOT) {
                V3 light_dir = normalize(ldir);
                float d = dot(-dtl, light_dir);
                float angle = std::acos(d);
                if (angle > light.max_angle) intensity = 0.0f;
            }
        }
        if (mat.receive_shadow) {                                                            // :872-914
            V3 start = hit_point + (surface_normal * 0.001f);
            V3 sdir = dtl;
            if (cfg.monte_carlo && mat.monte_carlo) sdir = jitter(sdir, mat.shadow_softness, mc, path, light_slot);
            Ray sray{start, sdir};
            Hit sh; bool shit = trace(sc, sray, true, true, depth, &sh, cnt);
            bool in_light = !shit;
            if (!in_light && (light.light_type == RTX_LIGHT_POINT || light.light_type == RTX_LIGHT_SPOT)) {
                float len = norm(lpos - hit_point);
                in_light = sh.t > len;
            }
            if (!in_light) {
                float shadow_source_alpha = mat.alpha;
                const RtxMaterial& smat = sc.mats[sc.items[sh.item].d.material];
                V3 shp = sray.o + (sray.d * sh.t);
                float suv[2]; item_get_uv(sc, item, shp, sh.face, suv);                      // receiver's get_uv (sic, :905)
                V4 satc; if (get_tex_color(sc, smat, suv, RTX_TEX_ALPHA, &satc)) shadow_source_alpha *= satc.x;
                intensity = intensity * (1.0f - shadow_source_alpha);
            }
        }
        color.x = color.x + ((light.color[0] * (specular.x + base.x)) * intensity);
        color.y = color.y + ((light.color[1] * (specular.y + base.y)) * intensity);
        color.z = color.z + ((light.color[2] * (specular.z + base.z)) * intensity);
    }
    float refraction_index = mat.refraction_index;
    float kr = fresnel(r.d, surface_normal, refraction_index);                               // :925
    float reflectivity = mat.reflectivity;
    V4 rftc; if (get_tex_color(sc, mat, uv, RTX_TEX_REFLECTIVITY, &rftc)) reflectivity = rftc.x;
    color = color * (1.0f - reflectivity);
    if (reflectivity > 0.0f && depth <= cfg.max_recursion) {                                 // :938-945
        Ray rr = create_reflection(surface_normal, r.d, hit_point);
        V3 rc = shade(sc, cfg, rr, depth + 1, mc, child_path(path, 0, depth), cnt).color;
        color = color + (rc * reflectivity);
    }
    if (alpha < 1.0f && depth <= cfg.max_recursion) {                                        // :948-975
        Ray tr;
        if (create_transmission(surface_normal, r.d, hit_point, refraction_index, &tr)) {
            Radiance t = shade(sc, cfg, tr, depth + 1, mc, child_path(path, 1, depth), cnt);
            if (kr < 1.0f) color = (color * alpha) + (t.color * (1.0f - kr) * (1.0f - alpha));
            else color = (color * alpha) + (t.color * (1.0f - alpha));
            if (approx_equal(alpha, 0.0f)) out.id = t.id;
        }
    } else if (alpha < 1.0f) color = color * alpha;
    {                                                                                        // fog :978-982
        float fog_amount = rmin(cfg.fog_density * hit_dist, 1.0f);
        V3 fc = {cfg.fog_color[0], cfg.fog_color[1], cfg.fog_color[2]};
        color = ((1.0f - fog_amount) * color) + (fc * fog_amount);
    }
    V4 ao; if (get_tex_color(sc, mat, uv, RTX_TEX_AMBIENT_OCCLUSION, &ao)) { color.x *= ao.x; color.y *= ao.x; color.z *= ao.x; }
    color = color + v3(ambient_color.x, ambient_color.y, ambient_color.z);
    out.color = color;
    return out;
}

// Raytracing::render ray generation (raytracing.rs:319-396)
Ray gen_ray(const RtxCamera& cam, const RtxConfig& cfg, const M4& pinv, const M4& vinv, int x, int y, uint32_t x_i, uint32_t y_i, uint32_t cell_size) {
    float x_f = (float)x, y_f = (float)y, w = (float)cam.width, h = (float)cam.height;
    float x_step = 2.0f / w, y_step = 2.0f / h;
    float x_trans = x_step * (float)x_i * (1.0f / (float)cell_size);
    float y_trans = y_step * (float)y_i * (1.0f / (float)cell_size);
    bool dof = cfg.aperture_size > 1.0f && cfg.focal_length > 1.0f;
    if (dof && cfg.samples > 1) { x_trans -= x_step / 2.0f; y_trans -= y_step / 2.0f; }
    float o4[4], d4[4], pp[4];
    if (dof) {                                                                               // :338-377
        float aperture_scale = (float)cam.width / 800.0f;
        x_trans *= cfg.aperture_size * aperture_scale;
        y_trans *= cfg.aperture_size * aperture_scale;
        float cx = ((x_f + 0.5f) / w) * 2.0f - 1.0f, cy = 1.0f - ((y_f + 0.5f) / h) * 2.0f;
        mul4(pinv, cx, cy, -1.0f, 1.0f, pp); pp[3] = 1.0f;
        float rd[4] = {pp[0] - 0.0f, pp[1] - 0.0f, pp[2] - 0.0f, 0.0f};
        float origin[4]; mul4(vinv, 0.0f, 0.0f, 0.0f, 1.0f, origin);
        float dir[4]; mul4(vinv, rd[0], rd[1], rd[2], rd[3], dir);
        // Vector4::normalize: dot4 = (a + c) + (b + d) in nalgebra's 4-lane form
        float a = dir[0] * dir[0], b = dir[1] * dir[1], c = dir[2] * dir[2], d = dir[3] * dir[3];
        float n4 = std::sqrt((a + c) + (b + d));
        for (int i = 0; i < 4; i++) dir[i] = dir[i] / n4;
        float dist = norm(v3(rd[0], rd[1], rd[2]));
        float f = 1.0f / (dist / (dist + cfg.focal_length));
        float p[4]; for (int i = 0; i < 4; i++) p[i] = origin[i] + f * dir[i];
        float sx = (((x_f + 0.5f) / w) * 2.0f - 1.0f) + x_trans, sy = (1.0f - ((y_f + 0.5f) / h) * 2.0f) + y_trans;
        mul4(pinv, sx, sy, -1.0f, 1.0f, pp); pp[3] = 1.0f;
        mul4(vinv, pp[0], pp[1], pp[2], pp[3], o4);
        return {v3(o4[0], o4[1], o4[2]), v3(p[0] - o4[0], p[1] - o4[1], p[2] - o4[2])};
    }
    float sx = (((x_f + 0.5f) / w) * 2.0f - 1.0f) + x_trans, sy = (1.0f - ((y_f + 0.5f) / h) * 2.0f) + y_trans;   // :381-395
    mul4(pinv, sx, sy, -1.0f, 1.0f, pp); pp[3] = 1.0f;
    float rd[4] = {pp[0] - 0.0f, pp[1] - 0.0f, pp[2] - 0.0f, 0.0f};
    mul4(vinv, pp[0], pp[1], pp[2], pp[3], o4);
    mul4(vinv, rd[0], rd[1], rd[2], rd[3], d4);
    return {v3(o4[0], o4[1], o4[2]), v3(d4[0], d4[1], d4[2])};
}

struct Pixel { uint8_t r, g, b; V3 normal; float depth; uint32_t id; };
// Raytracing::render (raytracing.rs:275-427)
Pixel render_pixel(const Scene& sc, const RtxCamera& cam, const RtxConfig& cfg, const M4& pinv, const M4& vinv,
                   int x, int y, uint32_t cell_size, const uint16_t* table, uint32_t n_samples, Counters* cnt) {
    V3 color{0, 0, 0}, normal{0, 0, 0}; float depth = 0.0f; uint32_t id = 0;
    for (uint32_t s = 0; s < n_samples; s++) {
        Ray ray = gen_ray(cam, cfg, pinv, vinv, x, y, table[2 * s], table[2 * s + 1], cell_size);
        McCtx mc{cfg.monte_carlo != 0, cfg.mc_seed, (uint32_t)(y * (int)cam.width + x), s};
        Radiance res = shade(sc, cfg, ray, 1, mc, 1, cnt);
        color = color + res.color; depth += res.depth; normal = normal + res.normal; id = res.id;
    }
    float n = (float)n_samples;
    color = color / n; depth /= n; normal = normal / n;
    color.x = rmin(color.x, 1.0f); color.y = rmin(color.y, 1.0f); color.z = rmin(color.z, 1.0f);
    Pixel p;
    p.r = as_u8(color.x * 255.0f); p.g = as_u8(color.y * 255.0f); p.b = as_u8(color.z * 255.0f);
    if (cfg.gamma_correction) {
        const float g = 1.0f / 2.2f;
        p.r = as_u8(std::pow(color.x, g) * 255.0f); p.g = as_u8(std::pow(color.y, g) * 255.0f); p.b = as_u8(std::pow(color.z, g) * 255.0f);
    }
    p.normal = normalize(normal); p.depth = depth; p.id = id;
    return p;
}

// material cache = Material::new(0,"") + apply_diff_without_textures (shape/mod.rs:182-246,769-772)
void derive_cache(Item& it, const RtxMaterial& m) {
    it.c_alpha = approx_equal(1.0f, m.alpha) ? 1.0f : m.alpha;
    it.c_cast_shadow = m.cast_shadow != 0;
    it.c_reflection_only = m.reflection_only != 0;
    it.c_backface = m.backface_cullig != 0;
    it.c_smooth = m.smooth_shading != 0;
}
void update_item(Scene& sc, Item& it) {
    memcpy(it.trans.m, it.d.trans, 64); memcpy(it.inv.m, it.d.tran_inverse, 64);
    if (it.d.shape == RTX_SHAPE_SPHERE) { float r = it.d.radius; it.lo = {-r, -r, -r}; it.hi = {r, r, r}; }
    else {
        const Mesh& m = sc.meshes[it.d.mesh];
        V3 lo{INFINITY, INFINITY, INFINITY}, hi{-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t k = 0; k < m.n_faces * 3; k++) {                    // TriMesh::aabb = union of triangle AABBs
            V3 v = m.verts[m.idx[k]];
            lo = {rmin(lo.x, v.x), rmin(lo.y, v.y), rmin(lo.z, v.z)}; hi = {rmax(hi.x, v.x), rmax(hi.y, v.y), rmax(hi.z, v.z)};
        }
        it.lo = lo; it.hi = hi;
    }
    derive_cache(it, sc.mats[it.d.material]);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C exports (same signatures as include/rtx.h, prefix oracle_)
// ------------------------------------------------------------------------------------------
extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }
int oracle_abi_version(void) { return RTX_ABI_VERSION; }

int oracle_scene_create(const RtxSceneDesc* d, int /*device*/, RtxScene** out) {
    if (!d || !out) { g_err = "null argument"; return RTX_E_INVALID; }
    Scene* sc = new Scene();
    sc->mats.assign(d->materials, d->materials + d->n_materials);
    for (uint32_t i = 0; i < d->n_textures; i++) {
        Tex t; t.w = d->textures[i].width; t.h = d->textures[i].height;
        t.rgba.assign(d->textures[i].rgba, d->textures[i].rgba + (size_t)t.w * t.h * 4);
        sc->texs.push_back(std::move(t));
    }
    for (uint32_t i = 0; i < d->n_meshes; i++) {
        const RtxMesh& s = d->meshes[i]; Mesh m;
        if (s.n_faces == 0) { delete sc; g_err = "mesh with 0 triangles"; return RTX_E_EMPTY_MESH; }
        m.n_faces = s.n_faces; m.n_uv_faces = s.n_uv_faces; m.n_normal_faces = s.n_normal_faces;
        for (uint32_t k = 0; k < s.n_vertices; k++) m.verts.push_back({s.vertices[3 * k], s.vertices[3 * k + 1], s.vertices[3 * k + 2]});
        m.idx.assign(s.indices, s.indices + 3 * (size_t)s.n_faces);
        if (s.n_uvs) m.uvs.assign(s.uvs, s.uvs + 2 * (size_t)s.n_uvs);
        if (s.n_uv_faces) m.uv_idx.assign(s.uv_indices, s.uv_indices + 3 * (size_t)s.n_uv_faces);
        for (uint32_t k = 0; k < s.n_normals; k++) m.normals.push_back({s.normals[3 * k], s.normals[3 * k + 1], s.normals[3 * k + 2]});
        if (s.n_normal_faces) m.n_idx.assign(s.normals_indices, s.normals_indices + 3 * (size_t)s.n_normal_faces);
        for (uint32_t v : m.idx) if (v >= s.n_vertices) { delete sc; g_err = "index out of range"; return RTX_E_INVALID; }
        sc->meshes.push_back(std::move(m));
    }
    {   // oracle-side acceleration structures, one thread per mesh (config 5: 64 meshes of 156 k triangles)
        std::atomic<size_t> next{0};
        auto work = [&]() { for (size_t i = next.fetch_add(1); i < sc->meshes.size(); i = next.fetch_add(1)) build_bvh2(sc->meshes[i]); };
        std::vector<std::thread> th;
        const unsigned nt = std::max(1u, std::min<unsigned>((unsigned)sc->meshes.size(), std::thread::hardware_concurrency()));
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    for (uint32_t i = 0; i < d->n_items; i++) {
        Item it; it.d = d->items[i];
        if (it.d.material < 0 || (uint32_t)it.d.material >= d->n_materials) { delete sc; g_err = "bad material index"; return RTX_E_INVALID; }
        if (it.d.shape == RTX_SHAPE_MESH && (it.d.mesh < 0 || (uint32_t)it.d.mesh >= d->n_meshes)) { delete sc; g_err = "bad mesh index"; return RTX_E_INVALID; }
        if (it.d.tran_inverse[3] != 0.0f || it.d.tran_inverse[7] != 0.0f || it.d.tran_inverse[11] != 0.0f || !(it.d.tran_inverse[15] > 0.5f && it.d.tran_inverse[15] < 2.0f)) { delete sc; g_err = "item transform is not affine"; return RTX_E_NON_AFFINE; }
        update_item(*sc, it);
        sc->items.push_back(it);
    }
    sc->lights.assign(d->lights, d->lights + d->n_lights);
    build_item_bvh(*sc);
    *out = reinterpret_cast<RtxScene*>(sc);
    return RTX_OK;
}

int oracle_scene_destroy(RtxScene* s) { delete reinterpret_cast<Scene*>(s); return RTX_OK; }

int oracle_scene_update_items(RtxScene* s, const RtxItemXform* x, size_t n) {
    Scene* sc = reinterpret_cast<Scene*>(s);
    for (size_t i = 0; i < n; i++) {
        if (x[i].item_index >= sc->items.size()) { g_err = "bad item index"; return RTX_E_INVALID; }
        Item& it = sc->items[x[i].item_index];
        memcpy(it.d.trans, x[i].trans, 64); memcpy(it.d.tran_inverse, x[i].tran_inverse, 64);
        update_item(*sc, it);
    }
    build_item_bvh(*sc);                                                  // Scene::update rebuilds it on every start (scene.rs:1674-1688)
    return RTX_OK;
}

int oracle_scene_set_lights(RtxScene* s, const RtxLight* l, uint32_t n) {
    reinterpret_cast<Scene*>(s)->lights.assign(l, l + n); return RTX_OK;
}

// options: bit0 = brute-force triangle loop (validates the oracle's own BVH); bit1 = Ball normal NOT
// negated when the ray starts inside (SURVEY.md §8(c) "Uncertain" switch; default is negated).
int oracle_scene_set_options(RtxScene* s, uint32_t options) {
    Scene* sc = reinterpret_cast<Scene*>(s);
    sc->brute_force = options & 1; sc->ball_normal_flip_inside = !(options & 2);
    return RTX_OK;
}

int oracle_sample_table(uint32_t samples, uint32_t* cell_size, uint16_t* xy) {
    if (samples == 0 || samples > 65535) { g_err = "samples out of range"; return RTX_E_INVALID; }
    std::vector<uint16_t> t; uint32_t c;
    build_sample_table(samples, &c, t);
    if (cell_size) *cell_size = c;
    if (xy) memcpy(xy, t.data(), t.size() * sizeof(uint16_t));
    return RTX_OK;
}

int oracle_trace_probe(RtxScene* s, const RtxRay* rays, size_t n, int for_shadow, int stop_on_first_hit, uint32_t depth, RtxHit* hits) {
    const Scene& sc = *reinterpret_cast<Scene*>(s);
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back([&, t]() {
        for (size_t i = t; i < n; i += nt) {
            Ray r{{rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]}, {rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]}};
            Hit h; RtxHit o; memset(&o, 0, sizeof(o));
            if (trace(sc, r, stop_on_first_hit != 0, for_shadow != 0, depth, &h, nullptr)) {
                o.t = h.t; o.normal[0] = h.n.x; o.normal[1] = h.n.y; o.normal[2] = h.n.z;
                o.item_id = sc.items[h.item].d.id; o.face_id = h.face; o.item_index = h.item;
            } else { o.t = -1.0f; o.item_index = -1; }
            hits[i] = o;
        }
    });
    for (auto& t : th) t.join();
    return RTX_OK;
}

// The shadow query of the shading loop (raytracing.rs:872-914) for a unit light contribution — the CPU statement of what
// rtx_shadow_probe reads back from the production shadow kernels.  occluder_index / t / face_id are ALWAYS the reference's
// first-hit item here (the device reports t / face only for alpha-textured occluders).
int oracle_shadow_probe(RtxScene* s, const RtxRay* rays, const float* light_distance, const int32_t* receiver_item, size_t n, uint32_t depth,
                        RtxShadowHit* out) {
    const Scene& sc = *reinterpret_cast<Scene*>(s);
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back([&, t]() {
        for (size_t i = t; i < n; i += nt) {
            Ray r{{rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]}, {rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]}};
            RtxShadowHit o; memset(&o, 0, sizeof(o));
            Hit sh; bool shit = trace(sc, r, true, true, depth, &sh, nullptr);                  // :883
            bool in_light = !shit;
            if (!in_light && light_distance) in_light = sh.t > light_distance[i];              // :885-892 (point / spot lights only)
            o.lit = in_light ? 1 : 0; o.k = 1.0f; o.occluder_index = -1; o.t = -1.0f;
            if (!in_light) {                                                                   // :895-913
                const int recv = receiver_item ? receiver_item[i] : -1;
                const Item& item = sc.items[recv >= 0 ? recv : 0];
                float shadow_source_alpha = recv >= 0 ? sc.mats[item.d.material].alpha : 1.0f;
                const RtxMaterial& smat = sc.mats[sc.items[sh.item].d.material];
                V3 shp = r.o + (r.d * sh.t);
                float suv[2]; item_get_uv(sc, item, shp, sh.face, suv);                        // receiver's get_uv (sic, :905)
                V4 satc; if (get_tex_color(sc, smat, suv, RTX_TEX_ALPHA, &satc)) shadow_source_alpha *= satc.x;
                o.k = 1.0f * (1.0f - shadow_source_alpha);
                o.occluder_index = sh.item; o.t = sh.t; o.face_id = sh.face;
            }
            out[i] = o;
        }
    });
    for (auto& t : th) t.join();
    return RTX_OK;
}

// Extended render: n_threads workers over 2x2 cells (renderer.rs:17,253-318); cell_step > 1 renders
// only every cell_step-th cell (bounded CPU-baseline sample); faithful != 0 rebuilds and shuffles the
// sample sub-grid for every pixel like raytracing.rs:290-313 does.
int oracle_render_frame_ex(RtxScene* s, const RtxCamera* cam, const RtxConfig* cfg, uint8_t* rgba, float* normals, float* depth,
                           uint32_t* ids, RtxStats* stats, int n_threads, int cell_step, int faithful) {
    const Scene& sc = *reinterpret_cast<Scene*>(s);
    if (!cam || !cfg || cam->width == 0 || cam->height == 0 || cfg->samples == 0) { g_err = "bad camera/config"; return RTX_E_INVALID; }
    M4 pinv, vinv; memcpy(pinv.m, cam->projection_inverse, 64); memcpy(vinv.m, cam->view_inverse, 64);
    uint32_t cell; std::vector<uint16_t> table; build_sample_table(cfg->samples, &cell, table);
    uint32_t ns = (uint32_t)table.size() / 2;
    int w = cam->width, h = cam->height, cw = (w + 1) / 2, ch = (h + 1) / 2;
    size_t n_cells = (size_t)cw * ch;
    if (cell_step < 1) cell_step = 1;
    if (n_threads < 1) n_threads = 1;
    std::atomic<size_t> next{0};
    std::vector<Counters> cnts(n_threads);
    std::atomic<uint64_t> prim{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back([&, t]() {
        std::vector<uint16_t> local; uint64_t np = 0;
        for (;;) {
            size_t c = next.fetch_add(1) * (size_t)cell_step;
            if (c >= n_cells) break;
            int cx = (int)(c % cw) * 2, cy = (int)(c / cw) * 2;
            for (int y = cy; y < std::min(cy + 2, h); y++) for (int x = cx; x < std::min(cx + 2, w); x++) {
                const uint16_t* tb = table.data();
                if (faithful) { uint32_t c2; build_sample_table(cfg->samples, &c2, local); tb = local.data(); }
                Pixel p = render_pixel(sc, *cam, *cfg, pinv, vinv, x, y, cell, tb, ns, &cnts[t]);
                size_t i = (size_t)y * w + x; np += ns;
                if (rgba) { rgba[4 * i] = p.r; rgba[4 * i + 1] = p.g; rgba[4 * i + 2] = p.b; rgba[4 * i + 3] = 255; }
                if (normals) { normals[3 * i] = p.normal.x; normals[3 * i + 1] = p.normal.y; normals[3 * i + 2] = p.normal.z; }
                if (depth) depth[i] = p.depth;
                if (ids) ids[i] = p.id;
            }
        }
        prim += np;
    });
    for (auto& t : th) t.join();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (auto& c : cnts) { stats->rays_closest += c.closest; stats->rays_shadow += c.shadow; }
        stats->primary_samples = prim.load();
        stats->device_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return RTX_OK;
}

int oracle_render_frame(RtxScene* s, const RtxCamera* cam, const RtxConfig* cfg, uint8_t* rgba, float* normals, float* depth,
                        uint32_t* ids, RtxStats* stats) {
    int nt = (int)std::max(1u, std::thread::hardware_concurrency());
    return oracle_render_frame_ex(s, cam, cfg, rgba, normals, depth, ids, stats, nt, 1, 0);
}

// ---- KAT helpers (tests/test_oracle_kat.py) ---------------------------------------------------
void oracle_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) { chacha_block(key, counter, stream, rounds, out); }
void oracle_seed_from_u64(uint64_t seed, uint32_t key_out[8]) { StdRng r = StdRng::seed_from_u64(seed); memcpy(key_out, r.key, 32); }
float oracle_fresnel(const float i[3], const float n[3], float index) { return fresnel({i[0], i[1], i[2]}, {n[0], n[1], n[2]}, index); }
int oracle_approx_equal(float a, float b) { return approx_equal(a, b); }
void oracle_gen_ray(const RtxCamera* cam, const RtxConfig* cfg, int x, int y, uint32_t x_i, uint32_t y_i, uint32_t cell, float o[3], float d[3]) {
    M4 pinv, vinv; memcpy(pinv.m, cam->projection_inverse, 64); memcpy(vinv.m, cam->view_inverse, 64);
    Ray r = gen_ray(*cam, *cfg, pinv, vinv, x, y, x_i, y_i, cell);
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; d[0] = r.d.x; d[1] = r.d.y; d[2] = r.d.z;
}
float oracle_mc_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t path, uint32_t slot) { return mc_uniform(seed, pixel, sample, path, slot); }
int oracle_tex_fetch(uint32_t w, uint32_t h, const uint8_t* rgba, int nearest, float u, float v, float out[4]) {
    Tex t; t.w = w; t.h = h; t.rgba.assign(rgba, rgba + (size_t)w * h * 4);
    V4 c = nearest ? texel(t, wrap(u, w), wrap(v, h)) : tex_interpolate(t, u, v);
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w; return 0;
}
int oracle_tri_cast(const float a[3], const float b[3], const float c[3], const float o[3], const float d[3], float* toi, float n[3], int* fid) {
    V3 nn; Ray r{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}};
    bool hit = tri_cast({a[0], a[1], a[2]}, {b[0], b[1], b[2]}, {c[0], c[1], c[2]}, r, toi, &nn, fid);
    if (hit) { n[0] = nn.x; n[1] = nn.y; n[2] = nn.z; }
    return hit;
}

}  // extern "C"
