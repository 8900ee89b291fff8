// rtx_device.cuh — device-side data layout and math of the B200 ray-casting path.
//
// Data layout in HBM (all 16-byte aligned, read with 128-bit loads):
//   nodes   : 80-byte wide BVH nodes (5 x float4), all per-mesh BLASes + the item TLAS concatenated
//   tris    : 48 bytes per triangle (3 x float4: a|face, b|-, c|-) in BVH leaf order, object space
//   items   : DItem (M^-1 rows, M rows, local AABB, flags, offsets)
//   mesh arrays (verts / indices / uvs / uv indices / normals / normal indices) for shading only
//   materials, texture table + RGBA8 texel pool, lights
//
// Arithmetic that decides a hit (ray -> object space, AABB slab key, ball and triangle tests) uses
// __fmul_rn/__fadd_rn/... so that nvcc cannot contract it into FMAs: the reference is Rust (never
// contracted), and the CPU oracle is built with -ffp-contract=off, so hit distances are bit-equal.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef RTX_SHADE_NOINLINE
#define RTX_SHADE_NOINLINE 0
#endif
#if RTX_SHADE_NOINLINE
#define RTX_SHADE_INLINE __noinline__      // keeps shade_kernel's code small enough for the instruction cache
#else
#define RTX_SHADE_INLINE __forceinline__
#endif

namespace rtx {

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
enum : uint32_t {
    IF_MESH = 1u, IF_VISIBLE = 2u, IF_FLIP = 4u, IF_CAST_SHADOW = 8u, IF_REFL_ONLY = 16u, IF_BACKFACE = 32u,
    IF_SMOOTH = 64u, IF_HAS_NORMALS = 128u, IF_ALPHA_TEX = 256u, IF_ALPHA_POS = 512u, IF_ALPHA_LT1 = 1024u,
    IF_DIV_W = 4096u,          // tran_inverse[3][3] != 1 (rounding of the cofactor inverse): points are divided by it, like from_homogeneous
    IF_TRANSLATION = 2048u,    // tran_inverse is identity + translation: o' = o + t, d' = d (bit-identical to the general product)
    IF_DIRECT_TRIS = 8192u,    // mesh of <= kDirectTris triangles (quads, planes): entering the item queues its triangles, no BLAS node visit
    IF_GROUPED = 16384u        // mesh item with an identity transform whose triangles are (also) in the merged world-space BLAS
};
constexpr uint32_t kGroupPrim = 0xFFFFFFFFu;   // entry of the fast TLAS's primitive list that stands for the merged BLAS
constexpr uint32_t kDirectTris = 4;

struct alignas(16) DItem {
    float4 inv[3];      // rows 0..2 of tran_inverse (affine)
    float4 mat[3];      // rows 0..2 of trans
    float4 lo;          // local AABB min, w = radius
    float4 hi;          // local AABB max, w = cached material alpha
    uint32_t flags, id, material, root;               // root = BLAS root node (mesh)
    uint32_t n_faces, n_uv_faces, n_normal_faces; float inv_w;   // inv_w = tran_inverse[3][3]: Point3::from_homogeneous divides by it
    uint32_t vert_off, idx_off, uv_off, uvidx_off;    // element offsets into the mesh arrays
    uint32_t nrm_off, nidx_off, tri_base, pad2;          // tri_base = first triangle of the mesh in `tris`
    float4 wlo, whi;    // padded world-space AABB (the TLAS leaf box): cheap pre-cull when the item list is walked without a TLAS
};

struct alignas(16) DMaterial {
    float ambient[3]; float alpha;
    float base[3]; float shininess;
    float specular[3]; float reflectivity;
    float refraction_index, normal_map_strength, shadow_softness, roughness;
    int32_t tex[8];
    uint32_t nearest, receive_shadow, monte_carlo, any_texture;
    float shadow_z_lo, rough_z_lo; uint32_t pad[2];     // cos(spread * pi) of jitter() for the two per-material spreads (host libm, like the oracle)
};

struct DTex { uint32_t offset_lo, offset_hi, w, h; };     // texel offset (64-bit) into the pool
struct alignas(16) DLight { float pos[3]; uint32_t type; float dir[3]; float intensity; float color[3]; float max_angle; uint32_t enabled, pad[3]; };

struct SceneDev {
    const float4* nodes; const float4* tris; const DItem* items;
    const uint32_t* tlas_prims;       // FULL item TLAS (every item; one BLAS per mesh): reference-order walks (K3b), probes, RTX_VERIFY
    const uint32_t* fast_prims;       // fast TLAS of the persistent kernels: items that are not grouped + kGroupPrim
    const float* verts; const uint32_t* idx; const float* uvs; const uint32_t* uv_idx; const float* nrms; const uint32_t* n_idx;
    const DMaterial* mats; const DTex* texs; const uchar4* texels; const DLight* lights;
    uint32_t n_items, n_lights, tlas_root, use_tlas;
    uint32_t ball_flip_inside, any_alpha_tex, flat_items /* bit mask of tlas_prims entries when n_items <= 24, else 0 */, pad;
    uint32_t fast_root, group_root /* root of the merged BLAS, ~0u: no group */, pad1, pad2;
    uint32_t* dbg;      // debug counters: [0] lane stack overflow
};

struct FrameDev {
    float pinv[16], vinv[16];
    uint32_t width, height, n_samples, cell_size;
    uint32_t monte_carlo, max_recursion, gamma, mc_seed;
    float focal_length, aperture_size, fog_density, pad0;
    float fog_color[3]; uint32_t debug_flags;
    uint32_t sample_group;      // raygen: consecutive rays = this many samples of one pixel
    const ushort2* sample_table;
    float4* accum_c;      // per frame pixel: rgb sum, depth sum
    float4* accum_n;      // per frame pixel: normal sum
    uint32_t* ids;        // per frame pixel: object id of the last sample
};

// ray queue (SoA): o.xyz + weight | d.xyz + pixel | meta (sample:16 depth:8 flags:8, path)
struct RayQ { float4* o; float4* d; uint2* m; };
// shadow queue (SoA): o.xyz + light distance | d.xyz + pixel | contribution rgb + receiver alpha | receiver item
struct ShadowQ { float4* o; float4* d; float4* c; uint32_t* r; uint4* probe; uint4* beyond; };   // beyond: (ray, occluder key bits, occluder item, -) for shadow_beyond_kernel   // probe: nullptr in frames; rtx_shadow_probe reads (lit, occluder item, toi bits, face) per ray
struct alignas(16) HitRec { float t; uint32_t item; uint32_t prim; uint32_t flags; };   // item = ~0u: miss
enum : uint32_t { HF_BACK = 1u, HF_INSIDE = 2u, HF_NEGN = 4u };   // HF_NEGN: parry returned -normalize(n) (t < 0 branch)
enum : uint32_t { RF_ID_OWNER = 1u };

struct Counters {      // device counters, 64-bit
    unsigned long long node_visits[2], tri_tests[2], sphere_tests, item_tests;   // [0] closest kernel, [1] shadow kernel
    // lane utilisation of the step loop (STATS kernels only; printed with RTX_PHASE_STATS=1): per kernel
    // [0] steps (per warp)  [1] lanes with a ray, summed over steps  [2] lanes in the NODE phase  [3] LEAF rounds  [4] lanes in LEAF rounds
    // [5] refills  [6] lanes refilled
    unsigned long long phase[2][8];
    unsigned long long beyond_rays, beyond_found;   // shadow_beyond_kernel: rays checked / rays sent on to the exact walk
};

// ------------------------------------------------------------------------------------------------
// exact (never contracted) f32 helpers — mirror nalgebra's evaluation order
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float xm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 xsub(float3 a, float3 b) { return f3(xs(a.x, b.x), xs(a.y, b.y), xs(a.z, b.z)); }
__device__ __forceinline__ float3 xadd(float3 a, float3 b) { return f3(xa(a.x, b.x), xa(a.y, b.y), xa(a.z, b.z)); }
__device__ __forceinline__ float3 xscale(float3 a, float s) { return f3(xm(a.x, s), xm(a.y, s), xm(a.z, s)); }
__device__ __forceinline__ float3 xneg(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float xdot(float3 a, float3 b) { return xa(xa(xm(a.x, b.x), xm(a.y, b.y)), xm(a.z, b.z)); }
__device__ __forceinline__ float3 xcross(float3 a, float3 b) {
    return f3(xs(xm(a.y, b.z), xm(a.z, b.y)), xs(xm(a.z, b.x), xm(a.x, b.z)), xs(xm(a.x, b.y), xm(a.y, b.x)));
}
__device__ __forceinline__ float xnorm(float3 a) { return xsqrt(xdot(a, a)); }
__device__ __forceinline__ float3 xnormalize(float3 a) { float n = xnorm(a); return f3(xd(a.x, n), xd(a.y, n), xd(a.z, n)); }
// out-of-line copy for the shade kernel (13 call sites of sqrt + 3 IEEE divisions): keeps its code inside the instruction cache
#ifndef RTX_XNORM_NOINLINE
#define RTX_XNORM_NOINLINE 1
#endif
#if RTX_XNORM_NOINLINE
__device__ __noinline__ float4 xnormalize_len_s(float3 a) { float n = xnorm(a); return make_float4(xd(a.x, n), xd(a.y, n), xd(a.z, n), n); }   // .w = the norm
__device__ __forceinline__ float3 xnormalize_s(float3 a) { const float4 r = xnormalize_len_s(a); return f3(r.x, r.y, r.z); }
#else
__device__ __forceinline__ float3 xnormalize_s(float3 a) { return xnormalize(a); }
__device__ __forceinline__ float4 xnormalize_len_s(float3 a) { float n = xnorm(a); return make_float4(xd(a.x, n), xd(a.y, n), xd(a.z, n), n); }
#endif
// row r of an affine 3x4 (row = m[r][0..3]) times (x,y,z,w): ((m0*x + m1*y) + m2*z) + m3*w
__device__ __forceinline__ float xrow(float4 r, float3 v, float w) { return xa(xa(xa(xm(r.x, v.x), xm(r.y, v.y)), xm(r.z, v.z)), xm(r.w, w)); }
__device__ __forceinline__ float3 xform_point(const float4 m[3], float3 p) { return f3(xrow(m[0], p, 1.0f), xrow(m[1], p, 1.0f), xrow(m[2], p, 1.0f)); }
// tran_inverse * point.to_homogeneous() -> Point3::from_homogeneous: w = m33 (affine), divide when it is not 1
__device__ __forceinline__ float3 xform_point_w(const float4 m[3], float3 p, uint32_t flags, float w) {
    float3 r = xform_point(m, p);
    if (flags & IF_DIV_W) r = f3(xd(r.x, w), xd(r.y, w), xd(r.z, w));
    return r;
}
__device__ __forceinline__ float3 xform_vec(const float4 m[3], float3 v) { return f3(xrow(m[0], v, 0.0f), xrow(m[1], v, 0.0f), xrow(m[2], v, 0.0f)); }

// plain float3 helpers (shading; contraction allowed)
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float len3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 norm3(float3 a) { float n = len3(a); return f3(a.x / n, a.y / n, a.z / n); }
__device__ __forceinline__ float3 norm3_fast(float3 a) { const float s = rsqrtf(dot3(a, a)); return f3(a.x * s, a.y * s, a.z * s); }   // MUFU.RSQ: Monte-Carlo directions only

// Rust `as` casts: saturating, NaN -> 0
__device__ __forceinline__ uint32_t as_u32(float f) { return __float2uint_rz(f); }
__device__ __forceinline__ int32_t as_i32(float f) { return __float2int_rz(f); }
__device__ __forceinline__ uint32_t as_u8(float f) { uint32_t v = __float2uint_rz(f); return v > 255u ? 255u : v; }
__device__ __forceinline__ bool approx_equal(float a, float b) { return truncf(xm(a, 1000000.0f)) == truncf(xm(b, 1000000.0f)); }

// counter-based RNG — identical to oracle/rt_oracle.cpp mc_uniform
__device__ __forceinline__ uint32_t mix32(uint32_t h) { h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16; return h; }
// split in two so that the (seed, pixel, sample) part is hashed once per ray and each draw costs one round
__device__ __forceinline__ uint32_t mc_base(uint32_t seed, uint32_t pixel, uint32_t sample) {
    uint32_t h = mix32(seed ^ 0x9E3779B9u);
    h = mix32(h ^ (pixel * 0x85EBCA6Bu + 0x165667B1u));
    return mix32(h ^ (sample * 0xC2B2AE35u + 0x27D4EB2Fu));
}
__device__ __forceinline__ float mc_draw(uint32_t base, uint32_t path, uint32_t slot) {
    const uint32_t h = mix32(base ^ (path * 0x9E3779B1u + slot * 0x632BE5ABu + 0x7F4A7C15u));
    return __fmul_rn((float)(h >> 8), 1.0f / 16777216.0f);
}
// id of a child ray in the reflection / refraction tree — identical to oracle/rt_oracle.cpp child_path
__device__ __forceinline__ uint32_t child_path(uint32_t path, uint32_t which, uint32_t depth) {
    return depth < 31u ? path * 2u + which : mix32(path * 0x9E3779B1u + which + (depth << 8));
}
__device__ __forceinline__ float mc_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t path, uint32_t slot) {
    return mc_draw(mc_base(seed, pixel, sample), path, slot);
}

// ------------------------------------------------------------------------------------------------
// primitive tests (restating parry3d 0.13; see oracle/rt_oracle.cpp for the citations)
// ------------------------------------------------------------------------------------------------
// Aabb::cast_local_ray(ray, f32::MAX, solid) — reference call sites src/shape/sphere.rs:51, mesh.rs:58
__device__ __forceinline__ bool aabb_cast(float3 lo, float3 hi, float3 o, float3 d, bool solid, float& key, float* t_exit = nullptr) {
    float tmin = 0.0f, tmax = 3.402823466e+38f;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, mn[3] = {lo.x, lo.y, lo.z}, mx[3] = {hi.x, hi.y, hi.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (dd[i] == 0.0f) {
            if (oo[i] < mn[i] || oo[i] > mx[i]) return false;
        } else {
            float denom = __frcp_rn(dd[i]);                              // == 1.0f / d, correctly rounded
            float n = xm(xs(mn[i], oo[i]), denom), f = xm(xs(mx[i], oo[i]), denom);
            if (n > f) { float t = n; n = f; f = t; }
            tmin = fmaxf(tmin, n);
            tmax = fminf(tmax, f);
            if (tmin > tmax) return false;
        }
    }
    key = (tmin == 0.0f && !solid) ? tmax : tmin;
    if (t_exit) *t_exit = tmax;
    return true;
}

// ray_toi_with_ball (center = origin) — reference call site src/shape/sphere.rs:60
__device__ __forceinline__ bool ball_cast(float radius, float3 o, float3 d, bool solid, float& toi, bool& inside) {
    float a = xdot(d, d), b = xdot(o, d), c = xs(xdot(o, o), xm(radius, radius));
    if (a == 0.0f) { if (c > 0.0f) return false; inside = true; toi = 0.0f; return true; }
    if (c > 0.0f && b > 0.0f) return false;
    float delta = xs(xm(b, b), xm(a, c));
    if (delta < 0.0f) return false;
    float sq = xsqrt(delta);
    float t = xd(xs(-b, sq), a);
    if (t <= 0.0f) { inside = true; toi = solid ? 0.0f : xd(xa(-b, sq), a); }
    else { inside = false; toi = t; }
    return toi <= 3.402823466e+38f;
}

// local_ray_intersection_with_triangle (Ericson) — via TriMesh, reference call site src/shape/mesh.rs:67
// `back`: bit0 = FeatureId::Face(1) (backface), bit2 (HF_NEGN) = normal is -normalize(n).
__device__ __forceinline__ bool tri_cast(float3 a, float3 b, float3 c, float3 o, float3 d, float& toi, uint32_t& back) {
    float3 ab = xsub(b, a), ac = xsub(c, a);
    float3 n = xcross(ab, ac);
    float dn = xdot(n, d);
    if (dn == 0.0f) return false;
    float3 ap = xsub(o, a);
    float t = xdot(ap, n);
    if ((t < 0.0f && dn < 0.0f) || (t > 0.0f && dn > 0.0f)) return false;
    back = dn < 0.0f ? 0u : 1u;
    dn = fabsf(dn);
    float3 e = xcross(xneg(d), ap);
    float v, w;
    if (t < 0.0f) {
        v = -xdot(ac, e); if (v < 0.0f || v > dn) return false;
        w = xdot(ab, e);  if (w < 0.0f || xa(v, w) > dn) return false;
        toi = xm(-t, xd(1.0f, dn));
        back |= HF_NEGN;
    } else {
        v = xdot(ac, e);  if (v < 0.0f || v > dn) return false;
        w = -xdot(ab, e); if (w < 0.0f || xa(v, w) > dn) return false;
        toi = xm(t, xd(1.0f, dn));
    }
    return toi <= 3.402823466e+38f;
}

// ------------------------------------------------------------------------------------------------
// wide-BVH traversal (8-wide quantised nodes, octant-ordered, stack of (node group, hit mask))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) { uint32_t r; asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x)); return r; }
__device__ __forceinline__ uint32_t byte_of(uint32_t x, int j) { return (x >> (8 * j)) & 0xffu; }
__device__ __forceinline__ uint32_t bfind(uint32_t x) { return 31u - __clz(x); }
#ifndef RTX_DEQUANT
#define RTX_DEQUANT 0
#endif
__device__ __forceinline__ float byte_as_f23(uint32_t x, int j) { return __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650u + j)); }   // 2^23 + byte j

// Blackwell packed FP32 (SASS FFMA2: two IEEE fused multiply-adds per instruction).  Tried for the six plane distances of a
// child (3 instructions instead of 6, clean codegen with the scalar operand broadcast): measured 1.5 % SLOWER on config 2,
// i.e. FFMA2 does not save issue bandwidth on this kernel.  Kept as a compile-time switch, off.
#ifndef RTX_FFMA2
#define RTX_FFMA2 0
#endif
__device__ __forceinline__ void ffma2(float a0, float a1, float b0, float b1, float c0, float c1, float& r0, float& r1) {
#if RTX_FFMA2
    asm("{ .reg .b64 a, b, c, r;\n\t mov.b64 a, {%2, %3};\n\t mov.b64 b, {%4, %5};\n\t mov.b64 c, {%6, %7};\n\t"
        " fma.rn.f32x2 r, a, b, c;\n\t mov.b64 {%0, %1}, r; }"
        : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
#else
    r0 = fmaf(a0, b0, c0); r1 = fmaf(a1, b1, c1);
#endif
}

struct TravStats { uint32_t nodes, tris; };
struct MeshHit { float t; uint32_t prim, face, back; };
enum { TM_CLOSEST = 0, TM_ANY_LE = 1, TM_ANY_GT = 2, TM_CLASSIFY = 3 };
constexpr int kStack = 28;

struct WideRay {        // per-ray constants of the node test
    float3 o, d, idir; uint32_t octinv4;
};
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ WideRay make_wide_ray(float3 o, float3 d) {
    WideRay r; r.o = o; r.d = d;
    const float eps = 1e-20f;
    // a zero component (either sign) counts as positive, consistently with the near/far selection in node_test;
    // the reciprocal only feeds the conservative (padded) box test, so MUFU.RCP accuracy is enough
    r.idir.x = fast_rcp(fabsf(d.x) > eps ? d.x : (d.x < 0.0f ? -eps : eps));
    r.idir.y = fast_rcp(fabsf(d.y) > eps ? d.y : (d.y < 0.0f ? -eps : eps));
    r.idir.z = fast_rcp(fabsf(d.z) > eps ? d.z : (d.z < 0.0f ? -eps : eps));
    uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
    r.octinv4 = (7u - oct) * 0x01010101u;
    return r;
}

// Intersect the 8 children of one node; returns hit mask: bits 24..31 internal children in
// traversal priority order, bits 0..23 leaf primitives.  [tmin, tmax] is the live ray interval.
__device__ __forceinline__ uint32_t node_test(const float4* __restrict__ nodes, uint32_t node_index, const WideRay& r,
                                              float tmin, float tmax, uint2& ngroup, uint2& tgroup) {
    const float4* np = nodes + (size_t)node_index * 5;
    const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    const uint32_t ew = __float_as_uint(n0.w);
    const float ax = __uint_as_float((ew & 0xffu) << 23) * r.idir.x;
    const float ay = __uint_as_float(((ew >> 8) & 0xffu) << 23) * r.idir.y;
    const float az = __uint_as_float(((ew >> 16) & 0xffu) << 23) * r.idir.z;
    const float bx = (n0.x - r.o.x) * r.idir.x, by = (n0.y - r.o.y) * r.idir.y, bz = (n0.z - r.o.z) * r.idir.z;
    // conservative padding: rounding of q*a + b is bounded by ~ulp(|b| + 255|a|); the margin also covers the few-ulp
    // gap between a triangle's toi (Ericson form) and the slab distance of its own box when tmax is a previous hit
    const float px = fmaf(fabsf(ax), 255.0f, fabsf(bx)) * 2e-6f;
    const float py = fmaf(fabsf(ay), 255.0f, fabsf(by)) * 2e-6f;
    const float pz = fmaf(fabsf(az), 255.0f, fabsf(bz)) * 2e-6f;
    // byte q -> float (2^23 + q) with one PRMT; the 2^23 is folded into the addend: q*a + b = (2^23+q)*a + (b - 2^23*a)
    const float k23 = 8388608.0f;
    const float blx = fmaf(-k23, ax, bx - px), bhx = fmaf(-k23, ax, bx + px);
    const float bly = fmaf(-k23, ay, by - py), bhy = fmaf(-k23, ay, by + py);
    const float blz = fmaf(-k23, az, bz - pz), bhz = fmaf(-k23, az, bz + pz);
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = __float_as_uint(half == 0 ? n1.z : n1.w);
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (r.octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = __float_as_uint(half == 0 ? n2.x : n2.y), qloy = __float_as_uint(half == 0 ? n2.z : n2.w);
        const uint32_t qloz = __float_as_uint(half == 0 ? n3.x : n3.y), qhix = __float_as_uint(half == 0 ? n3.z : n3.w);
        const uint32_t qhiy = __float_as_uint(half == 0 ? n4.x : n4.y), qhiz = __float_as_uint(half == 0 ? n4.z : n4.w);
        const uint32_t xn = r.d.x < 0.0f ? qhix : qlox, xf = r.d.x < 0.0f ? qlox : qhix;
        const uint32_t yn = r.d.y < 0.0f ? qhiy : qloy, yf = r.d.y < 0.0f ? qloy : qhiy;
        const uint32_t zn = r.d.z < 0.0f ? qhiz : qloz, zf = r.d.z < 0.0f ? qloz : qhiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
#if RTX_DEQUANT == 0      // all six planes through I2F.U8 (XU pipe), near/far pairs through one FFMA2 per axis
            float t0x, t1x, t0y, t1y, t0z, t1z;
            ffma2((float)byte_of(xn, j), (float)byte_of(xf, j), ax, ax, bx - px, bx + px, t0x, t1x);
            ffma2((float)byte_of(yn, j), (float)byte_of(yf, j), ay, ay, by - py, by + py, t0y, t1y);
            ffma2((float)byte_of(zn, j), (float)byte_of(zf, j), az, az, bz - pz, bz + pz, t0z, t1z);
#elif RTX_DEQUANT == 1    // all six through PRMT (ALU pipe)
            const float t0x = fmaf(byte_as_f23(xn, j), ax, blx), t1x = fmaf(byte_as_f23(xf, j), ax, bhx);
            const float t0y = fmaf(byte_as_f23(yn, j), ay, bly), t1y = fmaf(byte_as_f23(yf, j), ay, bhy);
            const float t0z = fmaf(byte_as_f23(zn, j), az, blz), t1z = fmaf(byte_as_f23(zf, j), az, bhz);
#else                     // near planes I2F (XU), far planes PRMT (ALU): balances the two pipes
            const float t0x = fmaf((float)byte_of(xn, j), ax, bx - px), t1x = fmaf(byte_as_f23(xf, j), ax, bhx);
            const float t0y = fmaf((float)byte_of(yn, j), ay, by - py), t1y = fmaf(byte_as_f23(yf, j), ay, bhy);
            const float t0z = fmaf((float)byte_of(zn, j), az, bz - pz), t1z = fmaf(byte_as_f23(zf, j), az, bhz);
#endif
            const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
            const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
            if (cmin <= cmax) hitmask |= byte_of(child_bits4, j) << byte_of(bit_index4, j);
        }
    }
    ngroup.x = __float_as_uint(n1.x); tgroup.x = __float_as_uint(n1.y);
    ngroup.y = (hitmask & 0xFF000000u) | (ew >> 24);
    tgroup.y = hitmask & 0x00FFFFFFu;
    return hitmask;
}

// Traverse one mesh BLAS in object space.
//  TM_CLOSEST: mh = closest triangle with toi <= limit (ties: lowest face index); limit prunes.
//  TM_ANY_LE : true as soon as a triangle with toi <= limit is found.
//  TM_ANY_GT : true as soon as a triangle with toi >  limit is found.
//  TM_CLASSIFY: true as soon as a triangle with toi <= limit is found; otherwise mh.back = 1 if any triangle
//               was hit beyond limit (after the first such hit the search is clipped to [0, limit]).
template <int MODE, bool STATS>
__device__ __forceinline__ bool traverse_mesh(const float4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t root,
                                              float3 o, float3 d, float limit, MeshHit& mh, TravStats& st) {
    const WideRay r = make_wide_ray(o, d);
    uint2 stack[kStack]; int sp = 0;
    uint2 ngroup = make_uint2(root, 0x80000000u), tgroup = make_uint2(0u, 0u);
    float tmin = (MODE == TM_ANY_GT) ? limit : 0.0f;
    float tmax = (MODE == TM_ANY_GT || MODE == TM_CLASSIFY) ? 3.402823466e+38f : limit;
    bool found = false;
    if (MODE == TM_CLOSEST) { mh.t = 3.402823466e+38f; mh.face = 0xFFFFFFFFu; }
    if (MODE == TM_CLASSIFY) mh.back = 0u;
    for (;;) {
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t hits = ngroup.y, imask = ngroup.y;
            const uint32_t cbit = bfind(hits);
            const uint32_t base = ngroup.x;
            ngroup.y &= ~(1u << cbit);
            if (ngroup.y > 0x00FFFFFFu) { stack[sp++] = ngroup; }
            const uint32_t slot = (cbit - 24u) ^ (r.octinv4 & 0xffu);
            const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
            if (STATS) st.nodes++;
            node_test(nodes, base + rel, r, tmin, tmax, ngroup, tgroup);
        } else {
            tgroup = ngroup; ngroup = make_uint2(0u, 0u);
        }
        while (tgroup.y != 0u) {
            const uint32_t ti = bfind(tgroup.y);
            tgroup.y &= ~(1u << ti);
            const uint32_t prim = tgroup.x + ti;
            const float4* tp = tris + (size_t)prim * 3;
            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
            if (STATS) st.tris++;
            float toi; uint32_t back;
            if (tri_cast(f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), o, d, toi, back)) {
                if (MODE == TM_CLOSEST) {
                    const uint32_t face = __float_as_uint(v0.w);
                    if (toi < mh.t || (toi == mh.t && face < mh.face)) {
                        if (toi <= limit) { mh.t = toi; mh.prim = prim; mh.face = face; mh.back = back; found = true; tmax = toi; }
                    }
                } else if (MODE == TM_ANY_LE) {
                    if (toi <= limit) return true;
                } else if (MODE == TM_CLASSIFY) {
                    if (toi <= limit) return true;
                    mh.back = 1u; tmax = limit;
                } else {
                    if (toi > limit) return true;
                }
            }
        }
        if (ngroup.y <= 0x00FFFFFFu) {
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    return found;
}

// ------------------------------------------------------------------------------------------------
// item level: Raytracing::trace (reference src/raytracing.rs:429-490)
// ------------------------------------------------------------------------------------------------
struct Best { float t; float key; uint32_t item, prim, flags; };

__device__ __forceinline__ bool item_passes(uint32_t flags, bool for_shadow, uint32_t depth) {
    // raytracing.rs:454: visible && cache.alpha > 0 && (!for_shadow || cast_shadow) && (!reflection_only || depth > 1)
    if (!(flags & IF_VISIBLE) || !(flags & IF_ALPHA_POS)) return false;
    if (for_shadow && !(flags & IF_CAST_SHADOW)) return false;
    if ((flags & IF_REFL_ONLY) && depth <= 1) return false;
    return true;
}
__device__ __forceinline__ bool item_solid(uint32_t flags, bool force_not_solid) {
    // sphere.rs:49-50 / mesh.rs:55-56 (the cache never has an alpha texture)
    return !(flags & IF_ALPHA_LT1) && (flags & IF_BACKFACE) && !force_not_solid;
}

// closest hit of one item, merged into `best` with the reference's order rule: strictly smaller t
// wins; equal t -> smaller (bbox key, item index) = earlier in the stable sort of :466.
template <bool STATS>
__device__ __forceinline__ void visit_item_closest(const SceneDev& S, uint32_t ii, float3 o, float3 d, bool for_shadow, uint32_t depth,
                                                   Best& best, TravStats& st, uint32_t& n_items, uint32_t& n_sph) {
    const DItem* it = S.items + ii;
    const uint32_t flags = it->flags;
    if (!item_passes(flags, for_shadow, depth)) return;
    float4 inv[3] = {__ldg(&it->inv[0]), __ldg(&it->inv[1]), __ldg(&it->inv[2])};
    const float3 lo3 = xform_point_w(inv, o, flags, it->inv_w), ld3 = xform_vec(inv, d);
    const float4 lo = __ldg(&it->lo), hi = __ldg(&it->hi);
    const bool solid = item_solid(flags, for_shadow);
    float key;
    if (STATS) n_items++;
    if (!aabb_cast(f3(lo.x, lo.y, lo.z), f3(hi.x, hi.y, hi.z), lo3, ld3, solid, key)) return;
    float t; uint32_t prim = 0, hf = 0;
    if (flags & IF_MESH) {
        MeshHit mh;
        if (!traverse_mesh<TM_CLOSEST, STATS>(S.nodes, S.tris, it->root, lo3, ld3, best.t, mh, st)) return;
        t = mh.t; prim = mh.prim; hf = mh.back;
    } else {
        bool inside;
        if (STATS) n_sph++;
        if (!ball_cast(lo.w, lo3, ld3, solid, t, inside)) return;
        hf = inside ? HF_INSIDE : 0u;
    }
    if (best.item == 0xFFFFFFFFu || t < best.t || (t == best.t && (key < best.key || (key == best.key && ii < best.item)))) {
        best.t = t; best.key = key; best.item = ii; best.prim = prim; best.flags = hf;
    }
}

// Closest hit over the whole scene (stop_on_first_hit = false).
template <bool STATS>
__device__ __forceinline__ void trace_closest(const SceneDev& S, float3 o, float3 d, bool for_shadow, uint32_t depth, Best& best,
                                              TravStats& st, uint32_t& n_items, uint32_t& n_sph) {
    best.t = 3.402823466e+38f; best.key = 0.0f; best.item = 0xFFFFFFFFu; best.prim = 0; best.flags = 0;
    if (!S.use_tlas) {
        for (uint32_t i = 0; i < S.n_items; i++) visit_item_closest<STATS>(S, i, o, d, for_shadow, depth, best, st, n_items, n_sph);
        return;
    }
    const WideRay r = make_wide_ray(o, d);
    uint2 stack[16]; int sp = 0;
    uint2 ngroup = make_uint2(S.tlas_root, 0x80000000u), tgroup = make_uint2(0u, 0u);
    for (;;) {
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t hits = ngroup.y, imask = ngroup.y;
            const uint32_t cbit = bfind(hits);
            const uint32_t base = ngroup.x;
            ngroup.y &= ~(1u << cbit);
            if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
            const uint32_t slot = (cbit - 24u) ^ (r.octinv4 & 0xffu);
            const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
            if (STATS) st.nodes++;
            node_test(S.nodes, base + rel, r, 0.0f, best.t, ngroup, tgroup);
        } else { tgroup = ngroup; ngroup = make_uint2(0u, 0u); }
        while (tgroup.y != 0u) {
            const uint32_t ti = bfind(tgroup.y);
            tgroup.y &= ~(1u << ti);
            visit_item_closest<STATS>(S, __ldg(S.tlas_prims + tgroup.x + ti), o, d, for_shadow, depth, best, st, n_items, n_sph);
        }
        if (ngroup.y <= 0x00FFFFFFu) { if (sp == 0) break; ngroup = stack[--sp]; }
    }
}

// ---- shadow rays: trace(.., stop_on_first_hit = true, for_shadow = true) ---------------------------
// The reference returns the CLOSEST hit of the FIRST item, in stable bbox-key order, that is hit at
// all (raytracing.rs:466-487).  Candidates are produced kCand at a time in (key, index) order.
constexpr int kCand = 8;
struct CandList { float key[kCand]; uint32_t item[kCand]; int n; bool more; };

__device__ __forceinline__ void cand_consider(const SceneDev& S, uint32_t ii, float3 o, float3 d, uint32_t depth, float ckey, uint32_t citem,
                                              bool have_cursor, CandList& cl) {
    const DItem* it = S.items + ii;
    const uint32_t flags = it->flags;
    if (!item_passes(flags, true, depth)) return;
    float4 inv[3] = {__ldg(&it->inv[0]), __ldg(&it->inv[1]), __ldg(&it->inv[2])};
    const float3 lo3 = xform_point_w(inv, o, flags, it->inv_w), ld3 = xform_vec(inv, d);
    const float4 lo = __ldg(&it->lo), hi = __ldg(&it->hi);
    float key;
    if (!aabb_cast(f3(lo.x, lo.y, lo.z), f3(hi.x, hi.y, hi.z), lo3, ld3, false, key)) return;
    if (have_cursor && (key < ckey || (key == ckey && ii <= citem))) return;     // already processed
    // insert into the sorted bounded list
    int pos = cl.n;
    while (pos > 0 && (key < cl.key[pos - 1] || (key == cl.key[pos - 1] && ii < cl.item[pos - 1]))) pos--;
    if (pos >= kCand) { cl.more = true; return; }
    if (cl.n == kCand) cl.more = true;
    int last = cl.n < kCand ? cl.n : kCand - 1;
    for (int k = last; k > pos; k--) { cl.key[k] = cl.key[k - 1]; cl.item[k] = cl.item[k - 1]; }
    cl.key[pos] = key; cl.item[pos] = ii;
    if (cl.n < kCand) cl.n++;
}

__device__ __forceinline__ void collect_candidates(const SceneDev& S, float3 o, float3 d, uint32_t depth, float ckey, uint32_t citem,
                                                   bool have_cursor, CandList& cl) {
    cl.n = 0; cl.more = false;
    if (!S.use_tlas) {
        for (uint32_t i = 0; i < S.n_items; i++) cand_consider(S, i, o, d, depth, ckey, citem, have_cursor, cl);
        return;
    }
    const WideRay r = make_wide_ray(o, d);
    uint2 stack[16]; int sp = 0;
    uint2 ngroup = make_uint2(S.tlas_root, 0x80000000u), tgroup = make_uint2(0u, 0u);
    for (;;) {
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t hits = ngroup.y, imask = ngroup.y;
            const uint32_t cbit = bfind(hits);
            const uint32_t base = ngroup.x;
            ngroup.y &= ~(1u << cbit);
            if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
            const uint32_t slot = (cbit - 24u) ^ (r.octinv4 & 0xffu);
            const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
            node_test(S.nodes, base + rel, r, 0.0f, 3.402823466e+38f, ngroup, tgroup);
        } else { tgroup = ngroup; ngroup = make_uint2(0u, 0u); }
        while (tgroup.y != 0u) {
            const uint32_t ti = bfind(tgroup.y);
            tgroup.y &= ~(1u << ti);
            cand_consider(S, __ldg(S.tlas_prims + tgroup.x + ti), o, d, depth, ckey, citem, have_cursor, cl);
        }
        if (ngroup.y <= 0x00FFFFFFu) { if (sp == 0) break; ngroup = stack[--sp]; }
    }
}

// one item, shadow semantics (force_not_solid = true)
template <int MODE, bool STATS>
__device__ __forceinline__ bool item_shadow_test(const SceneDev& S, uint32_t ii, float3 o, float3 d, float limit, MeshHit& mh, uint32_t& hf,
                                                 TravStats& st) {
    const DItem* it = S.items + ii;
    float4 inv[3] = {__ldg(&it->inv[0]), __ldg(&it->inv[1]), __ldg(&it->inv[2])};
    const float3 lo3 = xform_point_w(inv, o, it->flags, it->inv_w), ld3 = xform_vec(inv, d);
    if (it->flags & IF_MESH) {
        bool h = traverse_mesh<MODE, STATS>(S.nodes, S.tris, it->root, lo3, ld3, limit, mh, st);
        if (MODE == TM_CLOSEST && h) hf = mh.back;
        return h;
    }
    float t; bool inside;
    if (!ball_cast(__ldg(&it->lo).w, lo3, ld3, false, t, inside)) return false;
    if (MODE == TM_CLOSEST) { mh.t = t; mh.prim = 0; mh.face = 0; mh.back = 0; hf = inside ? HF_INSIDE : 0u; return true; }
    if (MODE == TM_ANY_LE) return t <= limit;
    return t > limit;
}

// Literal restatement: per item in order, full closest hit; return the first item hit.
template <bool STATS>
__device__ __forceinline__ void trace_shadow_ordered(const SceneDev& S, float3 o, float3 d, uint32_t depth, Best& best, TravStats& st) {
    best.item = 0xFFFFFFFFu; best.t = 3.402823466e+38f; best.prim = 0; best.flags = 0; best.key = 0.0f;
    float ckey = 0.0f; uint32_t citem = 0; bool have = false;
    CandList cl;
    for (;;) {
        collect_candidates(S, o, d, depth, ckey, citem, have, cl);
        for (int k = 0; k < cl.n; k++) {
            MeshHit mh; uint32_t hf = 0;
            if (item_shadow_test<TM_CLOSEST, STATS>(S, cl.item[k], o, d, 3.402823466e+38f, mh, hf, st)) {
                best.item = cl.item[k]; best.t = mh.t; best.prim = mh.prim; best.flags = hf; best.key = cl.key[k];
                return;
            }
        }
        if (!cl.more || cl.n == 0) return;
        ckey = cl.key[cl.n - 1]; citem = cl.item[cl.n - 1]; have = true;
    }
}

// Equivalent single-pass walk used by the shadow kernels.  `len` = light distance (+inf for directional).
// Per candidate item, in order, ONE traversal classifies it: a hit at toi <= len (the item's closest hit is
// then <= len: occluder, stop), only hits beyond len (the reference stops here with toi > len: lit), or no
// hit (next item).  When the occluder's material has an alpha texture its closest hit is also returned
// (needed for the attenuation lookup, raytracing.rs:898-912).
template <bool STATS>
__device__ __forceinline__ void trace_shadow_fast(const SceneDev& S, float3 o, float3 d, uint32_t depth, float len, Best& best, TravStats& st) {
    best.item = 0xFFFFFFFFu; best.t = 3.402823466e+38f; best.prim = 0; best.flags = 0; best.key = 0.0f;
    float ckey = 0.0f; uint32_t citem = 0; bool have = false;
    CandList cl;
    for (;;) {
        collect_candidates(S, o, d, depth, ckey, citem, have, cl);
        for (int k = 0; k < cl.n; k++) {
            const uint32_t ii = cl.item[k];
            const DItem* it = S.items + ii;
            float4 inv[3] = {__ldg(&it->inv[0]), __ldg(&it->inv[1]), __ldg(&it->inv[2])};
            const float3 lo3 = xform_point_w(inv, o, it->flags, it->inv_w), ld3 = xform_vec(inv, d);
            int cls;                                                   // 0 none, 1 hit <= len, 2 only beyond len
            MeshHit mh;
            if (it->flags & IF_MESH) {
                const bool le = traverse_mesh<TM_CLASSIFY, STATS>(S.nodes, S.tris, it->root, lo3, ld3, len, mh, st);
                cls = le ? 1 : (mh.back ? 2 : 0);
            } else {
                float t; bool inside;
                cls = ball_cast(__ldg(&it->lo).w, lo3, ld3, false, t, inside) ? (t <= len ? 1 : 2) : 0;
            }
            if (cls == 2) return;                                      // first item hit at all lies beyond the light: lit
            if (cls == 1) {
                best.item = ii; best.key = cl.key[k]; best.t = 0.0f;
                if (it->flags & IF_ALPHA_TEX) {
                    uint32_t hf = 0;
                    item_shadow_test<TM_CLOSEST, STATS>(S, ii, o, d, 3.402823466e+38f, mh, hf, st);
                    best.t = mh.t; best.prim = mh.prim; best.flags = hf;
                }
                return;
            }
        }
        if (!cl.more || cl.n == 0) return;
        ckey = cl.key[cl.n - 1]; citem = cl.item[cl.n - 1]; have = true;
    }
}

// ------------------------------------------------------------------------------------------------
// unified two-level traversal: ONE loop and ONE stack per lane (TLAS over items -> instance push ->
// per-mesh BLAS -> instance pop), written as a resumable step so that a warp can refill finished
// lanes with new rays (persistent threads with dynamic fetch).
//   UT_CLOSEST : Raytracing::trace(.., stop_on_first_hit = false): global closest hit, reference order rule
//   UT_ANY     : "is there any item with a hit at toi <= tmax" (order-independent half of the shadow query)
// ------------------------------------------------------------------------------------------------
enum { UT_CLOSEST = 0, UT_ANY = 1 };
constexpr int kLaneStack = 48;

struct Lane {
    WideRay w;                  // world-space ray (o, d, 1/d, octant)
    WideRay r;                  // ray of the current space (world in the TLAS, object space inside an item)
    float tmax;                 // UT_CLOSEST: best toi so far (prunes); UT_ANY: light distance
    uint2 ng, tg;               // pending node group / leaf group
    int sp, blas_base;          // blas_base < 0: walking the TLAS
    uint32_t cur_item; float cur_key;
    float bkey; uint32_t bitem, bprim, bface, bflags;      // result (bitem = ~0u: none)
};

__device__ __forceinline__ void lane_init(Lane& L, const SceneDev& S, float3 o, float3 d, float tmax) {
    L.w = make_wide_ray(o, d); L.r = L.w; L.tmax = tmax;
    if (S.flat_items) { L.ng = make_uint2(0u, 0u); L.tg = make_uint2(0u, S.flat_items); }      // few items: every item is a leaf entry, no TLAS node
    else { L.ng = make_uint2(S.fast_root, 0x80000000u); L.tg = make_uint2(0u, 0u); }
    L.sp = 0; L.blas_base = -1;
    L.cur_item = 0; L.cur_key = 0.0f; L.bkey = 0.0f; L.bitem = 0xFFFFFFFFu; L.bprim = 0; L.bface = 0xFFFFFFFFu; L.bflags = 0;
}

// Traversal stack exhausted (cannot happen for BVH depths accepted by rtx_scene_create): end this ray and raise the
// device-side error counter; the frame then returns an error instead of corrupting memory or spinning.
__device__ __forceinline__ void lane_abort(Lane& L, const SceneDev& S) {
    atomicAdd(S.dbg, 1u);
    L.ng = make_uint2(0u, 0u); L.tg = make_uint2(0u, 0u); L.sp = 0; L.blas_base = -1;
}

// closest-hit merge with the reference's order rule (strictly smaller toi wins; equal toi: earlier in the
// stable bbox-key sort, i.e. smaller (key, item index); inside one mesh: lowest face index)
__device__ __forceinline__ void lane_accept(Lane& L, float toi, float key, uint32_t item, uint32_t prim, uint32_t face, uint32_t flags) {
    bool take = toi < L.tmax;
    if (!take && toi == L.tmax && L.bitem != 0xFFFFFFFFu)
        take = key < L.bkey || (key == L.bkey && (item < L.bitem || (item == L.bitem && face < L.bface)));
    if (take) { L.tmax = toi; L.bkey = key; L.bitem = item; L.bprim = prim; L.bface = face; L.bflags = flags; }
}

// The traversal is split into warp-synchronous phases (see the kernels): in a NODE phase every lane that has a
// pending node group tests one node; in a LEAF phase every lane that has pending leaf entries handles exactly
// one (a triangle, or an item of the TLAS).  Leaf entries are postponed — parked in `tg`, or on the stack when a
// later node test produces more — until enough lanes of the warp have some, so triangle tests run on mostly
// full warps instead of the 4-5 lanes that happen to reach a leaf in the same step.
template <int MODE, bool STATS>
__device__ __forceinline__ void lane_node(Lane& L, uint2* stack, const SceneDev& S, TravStats& st) {
#ifndef RTX_NO_GUARD
    if (L.sp + 2 > kLaneStack) { lane_abort(L, S); return; }
#endif
    if (L.tg.y != 0u) stack[L.sp++] = L.tg;                               // park postponed leaf entries
    const uint32_t hits = L.ng.y, imask = L.ng.y;
    const uint32_t cbit = bfind(hits);
    const uint32_t base = L.ng.x;
    L.ng.y &= ~(1u << cbit);
    if (L.ng.y > 0x00FFFFFFu) stack[L.sp++] = L.ng;
    const uint32_t slot = (cbit - 24u) ^ (L.r.octinv4 & 0xffu);
    const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
    if (STATS) st.nodes++;
    // UT_ANY: the item level must see every candidate that can matter to the order rule
    // ... until an occluder is known: after that only candidates that sort BEFORE it matter (key < occluder's key), and an
    // item's key is never smaller than the distance at which the ray enters its world box
    // Before an occluder is known the walk is clipped at the light distance: an item the ray enters beyond the light can neither
    // occlude nor sort before an occluder whose own key is <= the light distance (an occluder with a larger key — the ray
    // starts inside its non-solid box — sends the ray to the exact walk, see shadow_any_kernel).
    const float tm = (MODE == UT_ANY && L.blas_base < 0) ? (L.bflags != 0u ? L.bkey : L.tmax) : L.tmax;
    node_test(S.nodes, base + rel, L.r, 0.0f, tm, L.ng, L.tg);
}

// UT_ANY bookkeeping: `bitem/bkey` = the occluder found (an item with a hit at toi <= len), `okey/oitem` (kept in
// bprim/bface as raw bits) = the earliest (key, index) among all OTHER candidate items.  The any-hit answer is final
// unless such a candidate sorts before the occluder (then the reference's first-hit rule may stop at it instead).
__device__ __forceinline__ void lane_note_other(Lane& L, float key, uint32_t item) {
    const float ok = __uint_as_float(L.bprim);
    if (L.bface == 0xFFFFFFFFu || key < ok || (key == ok && item < L.bface)) { L.bprim = __float_as_uint(key); L.bface = item; }
}
__device__ __forceinline__ void lane_any_hit(Lane& L, float key, uint32_t item, const SceneDev& S) {
    L.bitem = item; L.bkey = key; L.bflags = 1u;                          // bflags = 1: enumeration mode (no more BLAS work)
    if (L.blas_base >= 0) { L.sp = L.blas_base; L.blas_base = -1; L.r = L.w; }
    L.ng = make_uint2(0u, 0u); L.tg = make_uint2(0u, 0u);
    // directional light (no finite light distance) and no alpha-textured occluder anywhere: the first-hit order cannot change
    // "occluded" any more (see shadow_any_kernel), the other candidates need not be enumerated
    if (!S.any_alpha_tex && L.tmax >= 3.402823466e+38f) L.sp = 0;
}

// Merged world-space BLAS ("the group").  glTF scenes arrive as dozens of mesh items that all carry the identity transform
// (Scene::load_gltf bakes node transforms into the vertices, reference src/scene.rs:853-891); their boxes overlap heavily, so a
// TLAS over them sends every ray through several instance entries.  The host therefore also builds ONE BVH over all their
// triangles (object space == world space), each triangle tagged with its item.  What the reference does per ITEM is applied
// lazily per accepted triangle: the item filter of raytracing.rs:454 and the exact local-space slab test (`intersect_b_box`,
// which both admits the item and yields the sort key of :466).  cur_item = 0x80000000 | cached item (0x7FFFFFFF: none) marks
// group mode; cur_key = that item's key, NaN when the item is filtered out or its slab test fails.
__device__ __forceinline__ void group_item_key(const SceneDev& S, uint32_t ii, float3 o, float3 d, bool for_shadow, uint32_t depth, float& key) {
    const DItem* it = S.items + ii;
    const uint32_t flags = it->flags;
    key = __int_as_float(0x7fc00000);
    if (!item_passes(flags, for_shadow, depth)) return;
    const float4 lo = __ldg(&it->lo), hi = __ldg(&it->hi);
    float k;
    if (aabb_cast(f3(lo.x, lo.y, lo.z), f3(hi.x, hi.y, hi.z), o, d, item_solid(flags, for_shadow), k)) key = k;
}

// Shadow rays with a finite light distance whose occluder A is known (shadow_any_kernel): the reference stops at the FIRST item,
// in (bbox key, index) order, that is hit at all, and calls the ray lit when that hit lies beyond the light (raytracing.rs:466-487,
// 885-892).  "Occluded" can therefore only be wrong if some item sorted before A is hit beyond the light.  For the grouped items
// that is one any-hit query on the merged BLAS over (len, inf): true = such an item may exist (the ray goes to the exact walk).
// Stage 1 is a point query on the item TLAS: an item B that is entered before A (key_B < key_A <= len) and hit beyond the light
// holds the whole ray segment from key_B to that hit inside its box, in particular the point at the light distance — so only
// items whose box contains o + d * len (conservatively) can matter, and for most light positions there is none.
template <bool STATS>
__device__ __forceinline__ bool group_hit_beyond_light(const SceneDev& S, float3 o, float3 d, float len, float key_a, uint32_t item_a, uint32_t depth, TravStats& st) {
    const WideRay r = make_wide_ray(o, d);
    uint2 stack[kStack]; int sp = 0;
    uint2 ngroup = make_uint2(S.tlas_root, 0x80000000u), tgroup = make_uint2(0u, 0u);
    bool suspect = false;
    const float t_lo = len * 0.9999f - 1e-6f, t_hi = len * 1.0001f + 1e-6f;
    for (;;) {
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t imask = ngroup.y;
            const uint32_t cbit = bfind(imask);
            const uint32_t base = ngroup.x;
            ngroup.y &= ~(1u << cbit);
            if (ngroup.y > 0x00FFFFFFu) { if (sp >= kStack) return true; stack[sp++] = ngroup; }
            const uint32_t slot = (cbit - 24u) ^ (r.octinv4 & 0xffu);
            const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
            if (STATS) st.nodes++;
            node_test(S.nodes, base + rel, r, t_lo, t_hi, ngroup, tgroup);
        } else { tgroup = ngroup; ngroup = make_uint2(0u, 0u); }
        while (tgroup.y != 0u) {
            const uint32_t ti = bfind(tgroup.y);
            tgroup.y &= ~(1u << ti);
            const uint32_t ii = __ldg(S.tlas_prims + tgroup.x + ti);
            if (!(S.items[ii].flags & IF_GROUPED)) continue;
            float kb; group_item_key(S, ii, o, d, true, depth, kb);
            if (kb == kb && (kb < key_a || (kb == key_a && ii < item_a))) suspect = true;
        }
        if (suspect) break;
        if (ngroup.y <= 0x00FFFFFFu) { if (sp == 0) break; ngroup = stack[--sp]; }
    }
    if (!suspect) return false;
    sp = 0;
    ngroup = make_uint2(S.group_root, 0x80000000u); tgroup = make_uint2(0u, 0u);
    uint32_t c_item = 0xFFFFFFFFu; float c_key = 0.0f;
    for (;;) {
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t imask = ngroup.y;
            const uint32_t cbit = bfind(imask);
            const uint32_t base = ngroup.x;
            ngroup.y &= ~(1u << cbit);
            if (ngroup.y > 0x00FFFFFFu) { if (sp >= kStack) return true; stack[sp++] = ngroup; }
            const uint32_t slot = (cbit - 24u) ^ (r.octinv4 & 0xffu);
            const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
            if (STATS) st.nodes++;
            node_test(S.nodes, base + rel, r, len, 3.402823466e+38f, ngroup, tgroup);
        } else { tgroup = ngroup; ngroup = make_uint2(0u, 0u); }
        while (tgroup.y != 0u) {
            const uint32_t ti = bfind(tgroup.y);
            tgroup.y &= ~(1u << ti);
            const float4* tp = S.tris + (size_t)(tgroup.x + ti) * 3;
            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
            if (STATS) st.tris++;
            float toi; uint32_t back;
            if (tri_cast(f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), o, d, toi, back) && toi > len) {
                const uint32_t ii = __float_as_uint(v1.w);
                if (ii != c_item) { group_item_key(S, ii, o, d, true, depth, c_key); c_item = ii; }
                if (c_key == c_key && (c_key < key_a || (c_key == key_a && ii < item_a))) return true;
            }
        }
        if (ngroup.y <= 0x00FFFFFFu) { if (sp == 0) break; ngroup = stack[--sp]; }
    }
    return false;
}

// What a triangle hit means for the lane (closest: merge with the reference's order rule; any-hit: the ray is occluded).  In group
// mode the item work (filter + exact slab key) happens here, once per item change.
template <int MODE>
__device__ __forceinline__ void lane_tri_hit(Lane& L, const SceneDev& S, float toi, uint32_t back, uint32_t prim, uint32_t face, uint32_t tri_item, bool for_shadow, uint32_t depth) {
    if (L.cur_item & 0x80000000u) {                                       // triangle of the merged BLAS: item work only for a hit that would be taken
        const bool cand = (MODE == UT_CLOSEST) ? (toi < L.tmax || (toi == L.tmax && L.bitem != 0xFFFFFFFFu)) : (toi <= L.tmax);
        if (cand) {
            if ((L.cur_item & 0x7FFFFFFFu) != tri_item) { group_item_key(S, tri_item, L.w.o, L.w.d, for_shadow, depth, L.cur_key); L.cur_item = 0x80000000u | tri_item; }
            if (L.cur_key == L.cur_key) {
                if (MODE == UT_CLOSEST) lane_accept(L, toi, L.cur_key, tri_item, prim, face, back);
                else lane_any_hit(L, L.cur_key, tri_item, S);
            }
        }
    } else if (MODE == UT_CLOSEST) lane_accept(L, toi, L.cur_key, L.cur_item, prim, face, back);
    else if (toi <= L.tmax) lane_any_hit(L, L.cur_key, L.cur_item, S);
}

// Triangle round, two triangles at once: both records are loaded before either is tested, so the second load's latency hides
// behind the first test (the triangle rounds wait for their loads, not for issue slots).
template <int MODE, bool STATS>
__device__ __forceinline__ void lane_leaf_tri2(Lane& L, const SceneDev& S, bool for_shadow, uint32_t depth, TravStats& st) {
    const uint32_t t0 = bfind(L.tg.y);
    L.tg.y &= ~(1u << t0);
    const bool two = L.tg.y != 0u;
    const uint32_t t1 = two ? bfind(L.tg.y) : t0;
    L.tg.y &= ~(1u << t1);
    const uint32_t p0 = L.tg.x + t0, p1 = L.tg.x + t1;
    const float4* a = S.tris + (size_t)p0 * 3; const float4* b = S.tris + (size_t)p1 * 3;
    const float4 a0 = __ldg(a), a1 = __ldg(a + 1), a2 = __ldg(a + 2), b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
    if (STATS) st.tris += two ? 2u : 1u;
    float toi0, toi1; uint32_t back0, back1;
    const bool h0 = tri_cast(f3(a0.x, a0.y, a0.z), f3(a1.x, a1.y, a1.z), f3(a2.x, a2.y, a2.z), L.r.o, L.r.d, toi0, back0);
    const bool h1 = two && tri_cast(f3(b0.x, b0.y, b0.z), f3(b1.x, b1.y, b1.z), f3(b2.x, b2.y, b2.z), L.r.o, L.r.d, toi1, back1);
    if (h0) lane_tri_hit<MODE>(L, S, toi0, back0, p0, __float_as_uint(a0.w), __float_as_uint(a1.w), for_shadow, depth);
    // any-hit: the first hit ended the ray (its groups are gone); closest: the merge is order-independent
    if (h1 && !(MODE == UT_ANY && L.bflags != 0u)) lane_tri_hit<MODE>(L, S, toi1, back1, p1, __float_as_uint(b0.w), __float_as_uint(b1.w), for_shadow, depth);
}

// WHICH: 0 = whatever the entry is; 1 = the caller knows it is a triangle (L.blas_base >= 0); 2 = an item of the TLAS.
// The kernels run triangles and items in separate rounds: the two paths share no code, so a mixed round would execute
// both at a fraction of the lanes each.
template <int MODE, bool STATS, int WHICH = 0>
__device__ __forceinline__ void lane_leaf(Lane& L, uint2* stack, const SceneDev& S, bool for_shadow, uint32_t depth,
                                          TravStats& st, uint32_t& n_items, uint32_t& n_sph) {
    const uint32_t ti = bfind(L.tg.y);
    L.tg.y &= ~(1u << ti);
    const uint32_t prim = L.tg.x + ti;
    if (WHICH == 1 || (WHICH == 0 && L.blas_base >= 0)) {                 // triangle of the current item
        const float4* tp = S.tris + (size_t)prim * 3;
        const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
        if (STATS) st.tris++;
        float toi; uint32_t back;
        if (tri_cast(f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), L.r.o, L.r.d, toi, back)) {
            if (L.cur_item & 0x80000000u) {                               // triangle of the merged BLAS: item work only for a hit that would be taken
                const bool cand = (MODE == UT_CLOSEST) ? (toi < L.tmax || (toi == L.tmax && L.bitem != 0xFFFFFFFFu)) : (toi <= L.tmax);
                if (cand) {
                    const uint32_t ii = __float_as_uint(v1.w);
                    if ((L.cur_item & 0x7FFFFFFFu) != ii) { group_item_key(S, ii, L.w.o, L.w.d, for_shadow, depth, L.cur_key); L.cur_item = 0x80000000u | ii; }
                    if (L.cur_key == L.cur_key) {
                        if (MODE == UT_CLOSEST) lane_accept(L, toi, L.cur_key, ii, prim, __float_as_uint(v0.w), back);
                        else lane_any_hit(L, L.cur_key, ii, S);
                    }
                }
            } else if (MODE == UT_CLOSEST) lane_accept(L, toi, L.cur_key, L.cur_item, prim, __float_as_uint(v0.w), back);
            else if (toi <= L.tmax) lane_any_hit(L, L.cur_key, L.cur_item, S);
        }
        return;
    }
    const uint32_t ii = __ldg(S.fast_prims + prim);                       // entry of the fast TLAS
    if (ii == kGroupPrim) {                                               // the merged BLAS: world space, no item work at entry
        if (MODE == UT_ANY && L.bflags != 0u) return;                     // occluder known: grouped items are not enumerated (shadow_beyond_kernel settles their order)
#ifndef RTX_NO_GUARD
        if (L.sp + 2 > kLaneStack) { lane_abort(L, S); return; }
#endif
        if (L.ng.y > 0x00FFFFFFu) stack[L.sp++] = L.ng;
        if (L.tg.y != 0u) stack[L.sp++] = L.tg;
        L.blas_base = L.sp; L.cur_item = 0xFFFFFFFFu; L.cur_key = 0.0f;
        L.r = L.w;
        L.ng = make_uint2(S.group_root, 0x80000000u); L.tg = make_uint2(0u, 0u);
        return;
    }
    const DItem* it = S.items + ii;
    const uint32_t flags = it->flags;
    if (!item_passes(flags, for_shadow, depth)) return;
    if (S.flat_items) {                                                   // no TLAS node above this item: conservative world-box pre-cull
        const float4 wl = __ldg(&it->wlo), wh = __ldg(&it->whi);
        const float ax = (wl.x - L.r.o.x) * L.r.idir.x, bx = (wh.x - L.r.o.x) * L.r.idir.x;
        const float ay = (wl.y - L.r.o.y) * L.r.idir.y, by = (wh.y - L.r.o.y) * L.r.idir.y;
        const float az = (wl.z - L.r.o.z) * L.r.idir.z, bz = (wh.z - L.r.o.z) * L.r.idir.z;
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        // closest: nothing beyond the best hit matters; any-hit: every candidate must be seen (order rule), no clipping
        const float lim = (MODE == UT_CLOSEST) ? L.tmax : 3.402823466e+38f;
        if (!(tn * 0.999999f <= fminf(tf * 1.000001f, lim))) return;
    }
    float3 lo3, ld3;
    if (flags & IF_TRANSLATION) {                                         // ((1*x + 0*y) + 0*z) + t*1 == x + t for finite inputs
        const float4 r0 = __ldg(&it->inv[0]), r1 = __ldg(&it->inv[1]), r2 = __ldg(&it->inv[2]);
        lo3 = f3(xa(L.w.o.x, r0.w), xa(L.w.o.y, r1.w), xa(L.w.o.z, r2.w)); ld3 = L.w.d;
    } else {
        float4 inv[3] = {__ldg(&it->inv[0]), __ldg(&it->inv[1]), __ldg(&it->inv[2])};
        lo3 = xform_point_w(inv, L.w.o, flags, it->inv_w); ld3 = xform_vec(inv, L.w.d);
    }
    const float4 lo = __ldg(&it->lo), hi = __ldg(&it->hi);
    const bool solid = item_solid(flags, for_shadow);
    float key, t_exit;
    if (STATS) n_items++;
    if (!aabb_cast(f3(lo.x, lo.y, lo.z), f3(hi.x, hi.y, hi.z), lo3, ld3, solid, key, &t_exit)) return;
    // UT_ANY: which OTHER candidates can turn "occluded" into the reference's "lit"?  Only an item that sorts before the occluder and
    // whose first hit lies BEYOND the light (raytracing.rs:466-487, 885-892).  Its hits lie inside its box, so a box the ray leaves
    // before the light cannot hold one: such a candidate either has a hit in front of the light (same verdict) or none (not a
    // candidate at all) and need not be noted.  The margin covers the few ulps between a triangle's toi and its box's slab distance.
    const bool reaches_light = MODE == UT_ANY && fmaf(t_exit, 1.0001f, 1e-6f) >= L.tmax;
    if (MODE == UT_ANY && L.bflags != 0u) { if (reaches_light) lane_note_other(L, key, ii); return; }    // occluder known: only enumerate
    if (!(flags & IF_MESH)) {
        float t; bool inside;
        if (STATS) n_sph++;
        const bool h = ball_cast(lo.w, lo3, ld3, solid, t, inside);
        if (MODE == UT_CLOSEST) { if (h) lane_accept(L, t, key, ii, 0u, 0u, inside ? HF_INSIDE : 0u); }
        else if (h && t <= L.tmax) lane_any_hit(L, key, ii, S);
        else if (h) lane_note_other(L, key, ii);                          // hit only beyond the light: the one kind of candidate that matters
        return;
    }
    if (MODE == UT_ANY && reaches_light) lane_note_other(L, key, ii);     // harmless if it becomes the occluder: (key, item) is then not < itself
    // enter the instance: park what is left of the TLAS groups under the BLAS part of the stack
#ifndef RTX_NO_GUARD
    if (L.sp + 2 > kLaneStack) { lane_abort(L, S); return; }
#endif
    if (L.ng.y > 0x00FFFFFFu) stack[L.sp++] = L.ng;
    if (L.tg.y != 0u) stack[L.sp++] = L.tg;
    L.blas_base = L.sp; L.cur_item = ii; L.cur_key = key;
    L.r = make_wide_ray(lo3, ld3);
    if (flags & IF_DIRECT_TRIS) { L.ng = make_uint2(0u, 0u); L.tg = make_uint2(it->tri_base, (1u << it->n_faces) - 1u); }
    else { L.ng = make_uint2(it->root, 0x80000000u); L.tg = make_uint2(0u, 0u); }
}

// Both groups empty: leave the instance if its part of the stack is drained, then pop.  Returns true when the ray is done.
__device__ __forceinline__ bool lane_pop(Lane& L, const uint2* stack) {
    if (L.blas_base >= 0 && L.sp == L.blas_base) { L.blas_base = -1; L.r = L.w; }
    if (L.sp == 0) return true;
    const uint2 e = stack[--L.sp];
    if (e.y > 0x00FFFFFFu) L.ng = e; else L.tg = e;
    return false;
}

// ------------------------------------------------------------------------------------------------
// hit attributes: Sphere::intersect / Mesh::intersect normal, get_uv, get_normal
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 ld3(const float* p, uint32_t i) { return f3(__ldg(p + 3 * (size_t)i), __ldg(p + 3 * (size_t)i + 1), __ldg(p + 3 * (size_t)i + 2)); }

// area-ratio weights of Mesh::get_uv / get_normal (reference src/shape/mesh.rs:105-161, 204-259)
__device__ __forceinline__ void area_weights(const SceneDev& S, const DItem& it, float3 hit, uint32_t f_id, float w[3]) {
    const float3 hl = xform_point_w(it.inv, hit, it.flags, it.inv_w);
    const uint32_t* fi = S.idx + it.idx_off + 3 * (size_t)f_id;
    const float3 a = ld3(S.verts + it.vert_off, __ldg(fi)), b = ld3(S.verts + it.vert_off, __ldg(fi + 1)), c = ld3(S.verts + it.vert_off, __ldg(fi + 2));
    const float3 f1 = xsub(a, hl), f2 = xsub(b, hl), f3_ = xsub(c, hl);
    const float area = xnorm(xcross(xsub(a, b), xsub(a, c)));
    w[0] = xd(xnorm(xcross(f2, f3_)), area); w[1] = xd(xnorm(xcross(f3_, f1)), area); w[2] = xd(xnorm(xcross(f1, f2)), area);
}

__device__ RTX_SHADE_INLINE void item_get_uv(const SceneDev& S, const DItem& it, float3 hit, uint32_t face_id, float& u, float& v) {
    const float PI = 3.14159265358979323846f;
    if (!(it.flags & IF_MESH)) {                                  // sphere.rs:69-99
        const float3 hl = xform_point_w(it.inv, hit, it.flags, it.inv_w);
        const float theta = atan2f(-hl.z, hl.x);
        u = (theta + PI) / (2.0f * PI);
        const float phi = acosf((-hl.y) / it.lo.w);
        v = -(phi / PI);
        return;
    }
    const uint32_t f_id = face_id % it.n_faces;
    if ((int32_t)it.n_uv_faces - 1 < (int32_t)f_id || (int32_t)it.n_faces - 1 < (int32_t)f_id) { u = 0.0f; v = 0.0f; return; }
    float w[3]; area_weights(S, it, hit, f_id, w);
    const uint32_t* ui = S.uv_idx + it.uvidx_off + 3 * (size_t)f_id;
    const float* ua = S.uvs + it.uv_off + 2 * (size_t)__ldg(ui); const float* ub = S.uvs + it.uv_off + 2 * (size_t)__ldg(ui + 1);
    const float* uc = S.uvs + it.uv_off + 2 * (size_t)__ldg(ui + 2);
    u = xa(xa(xm(__ldg(ua), w[0]), xm(__ldg(ub), w[1])), xm(__ldg(uc), w[2]));
    v = -xa(xa(xm(__ldg(ua + 1), w[0]), xm(__ldg(ub + 1), w[1])), xm(__ldg(uc + 1), w[2]));
}

// World-space shading normal returned by Shape::intersect (sphere.rs:54-67, mesh.rs:61-103) and the
// reference face id (parry FeatureId: face, +n_faces on a backface).
__device__ __forceinline__ float3 hit_normal(const SceneDev& S, const DItem& it, float3 o, float3 d, float t, uint32_t prim, uint32_t hflags,
                                             uint32_t& face_id) {
    if (!(it.flags & IF_MESH)) {
        const float3 lo3 = xform_point_w(it.inv, o, it.flags, it.inv_w), ld = xform_vec(it.inv, d);
        float3 n = xnormalize_s(xadd(lo3, xscale(ld, t)));
        if ((hflags & HF_INSIDE) && S.ball_flip_inside) n = xneg(n);
        face_id = 0;
        return xnormalize_s(xform_vec(it.mat, n));
    }
    const float4* tp = S.tris + (size_t)prim * 3;
    const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    const uint32_t face = __float_as_uint(v0.w);
    const bool back = hflags & HF_BACK;
    face_id = back ? face + it.n_faces : face;
    float3 n;
    if ((it.flags & IF_SMOOTH) && (it.flags & IF_HAS_NORMALS)) {
        const float3 hit = xadd(o, xscale(d, t));
        float w[3]; area_weights(S, it, hit, face, w);
        const uint32_t* ni = S.n_idx + it.nidx_off + 3 * (size_t)face;
        const float3 a = ld3(S.nrms + it.nrm_off, __ldg(ni)), b = ld3(S.nrms + it.nrm_off, __ldg(ni + 1)), c = ld3(S.nrms + it.nrm_off, __ldg(ni + 2));
        const float3 p1 = xscale(a, w[0]), p2 = xscale(b, w[1]), p3 = xscale(c, w[2]);
        n = f3(xa(xa(p1.x, p2.x), p3.x), xa(xa(p1.y, p2.y), p3.y), xa(xa(p1.z, p2.z), p3.z));
        n = xnormalize_s(xform_vec(it.mat, n));
        if (back) n = xneg(n);
    } else {
        const float3 a = f3(v0.x, v0.y, v0.z), b = f3(v1.x, v1.y, v1.z), c = f3(v2.x, v2.y, v2.z);
        float3 g = xnormalize_s(xcross(xsub(b, a), xsub(c, a)));
        if (hflags & HF_NEGN) g = xneg(g);      // parry: normal faces the ray origin
        n = xnormalize_s(xform_vec(it.mat, g));
    }
    if (it.flags & IF_FLIP) n = xneg(n);
    return n;
}

// ------------------------------------------------------------------------------------------------
// textures (reference src/raytracing.rs:629-675, src/shape/mod.rs:510-629)
// ------------------------------------------------------------------------------------------------
// (float)p / 255.0f, correctly rounded, in three operations: q = p*r, q += (p - q*255)*r with r = RN(1/255).  Equal to the
// IEEE division for all 256 inputs (checked exhaustively in tests/test_oracle_kat.py); the division costs ~10 instructions
// and there are 16 of them per bilinear fetch.
__device__ __forceinline__ float u8_unit(uint32_t p) {
    const float pf = (float)p, r = 1.0f / 255.0f;
    const float q = __fmul_rn(pf, r);
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, pf), r, q);
}
__device__ __forceinline__ float4 texel_at(const SceneDev& S, const DTex& t, uint32_t x, uint32_t y) {
    const size_t off = ((size_t)t.offset_hi << 32 | t.offset_lo) + (size_t)y * t.w + x;
    const uchar4 p = __ldg(S.texels + off);
    return make_float4(u8_unit(p.x), u8_unit(p.y), u8_unit(p.z), u8_unit(p.w));
}
__device__ __forceinline__ uint32_t tex_wrap(float val, uint32_t bound) {
    const int32_t sb = (int32_t)bound;
    const int32_t i = as_i32(xm(val, (float)bound));
    // power-of-two sizes (every glTF map of the bench scenes): the truncating remainder moved into [0, sb) is the low bits of the
    // two's-complement value — same result as the general path, without the ~20-instruction integer division
    if ((bound & (bound - 1u)) == 0u) return (uint32_t)i & (bound - 1u);
    const int32_t w = i % sb;
    return w < 0 ? (uint32_t)(w + sb) : (uint32_t)w;
}
__device__ __forceinline__ float lerp1(float a, float b, float f) { return xa(a, xm(f, xs(b, a))); }
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float f) { return make_float4(lerp1(a.x, b.x, f), lerp1(a.y, b.y, f), lerp1(a.z, b.z, f), lerp1(a.w, b.w, f)); }
__device__ RTX_SHADE_INLINE float4 tex_interpolate(const SceneDev& S, const DTex& t, float xf, float yf) {
    float x = xm(xf, (float)t.w), y = xm(yf, (float)t.h);
    if (x < 0.0f) x = xa(x, (float)t.w);
    if (y < 0.0f) y = xa(y, (float)t.h);
    uint32_t x0 = as_u32(floorf(x)), x1 = as_u32(ceilf(x)), y0 = as_u32(floorf(y)), y1 = as_u32(ceilf(y));
    if (x0 >= t.w) x0 = t.w - 1; if (y0 >= t.h) y0 = t.h - 1;
    if (x1 >= t.w) x1 = t.w - 1; if (y1 >= t.h) y1 = t.h - 1;
    const float fx = xs(x, (float)x0), fy = xs(y, (float)y0);
    const float4 a = lerp4(texel_at(S, t, x0, y0), texel_at(S, t, x1, y0), fx);
    const float4 b = lerp4(texel_at(S, t, x0, y1), texel_at(S, t, x1, y1), fx);
    return lerp4(a, b, fy);
}
// Raytracing::get_tex_color: false when the material has no such texture (or no uv)
__device__ __forceinline__ bool get_tex_color(const SceneDev& S, const DMaterial& m, bool has_uv, float u, float v, int type, float4& out) {
    const int ti = m.tex[type];
    if (ti < 0 || !has_uv) return false;
    const DTex t = S.texs[ti];
    if (t.w == 0) return false;
    out = m.nearest ? texel_at(S, t, tex_wrap(u, t.w), tex_wrap(v, t.h)) : tex_interpolate(S, t, u, v);
    return true;
}

// ------------------------------------------------------------------------------------------------
// shading helpers (reference src/raytracing.rs:492-626)
// ------------------------------------------------------------------------------------------------
// Ray geometry is computed with the never-contracted helpers: whether a child ray exists (total internal
// reflection) and what it hits must not depend on FMA contraction, or ray totals drift from the reference's.
__device__ __forceinline__ bool create_transmission(float3 normal, float3 incident, float3 p, float index, float3& o, float3& d) {
    float3 ref_n = normal; float eta_t = index, eta_i = 1.0f; float i_dot_n = xdot(incident, normal);
    if (i_dot_n < 0.0f) i_dot_n = -i_dot_n; else { ref_n = xneg(normal); eta_t = 1.0f; eta_i = index; }
    const float eta = xd(eta_i, eta_t);
    const float k = xs(1.0f, xm(xm(eta, eta), xs(1.0f, xm(i_dot_n, i_dot_n))));
    if (k < 0.0f) return false;
    o = xadd(p, xscale(ref_n, -0.001f));
    d = xsub(xscale(xadd(incident, xscale(ref_n, i_dot_n)), eta), xscale(ref_n, xsqrt(k)));
    return true;
}
__device__ __forceinline__ void create_reflection(float3 normal, float3 incident, float3 p, float3& o, float3& d) {
    o = xadd(p, xscale(normal, 0.001f));
    d = xsub(incident, xscale(normal, xm(2.0f, xdot(incident, normal))));
}
__device__ __forceinline__ float fresnel(float3 incident, float3 normal, float index) {
    const float i_dot_n = dot3(incident, normal);
    float eta_i = 1.0f, eta_t = index;
    if (i_dot_n > 0.0f) { eta_i = eta_t; eta_t = 1.0f; }
    const float sin_t = eta_i / eta_t * sqrtf(fmaxf(1.0f - i_dot_n * i_dot_n, 0.0f));
    if (sin_t > 1.0f) return 1.0f;
    const float cos_t = sqrtf(fmaxf(1.0f - sin_t * sin_t, 0.0f));
    const float cos_i = fabsf(cos_t);
    const float r_s = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    const float r_p = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (r_s * r_s + r_p * r_p) / 2.0f;
}
// jitter (reference src/raytracing.rs:565-626).  z_lo = cos(spread * pi).  The sampled direction is Monte-Carlo data
// (the reference draws from thread_rng), so the basis and the trigonometry use the fast units; the uniforms are the
// oracle's bit for bit.
__device__ RTX_SHADE_INLINE float3 jitter(float3 dir, float z_lo, uint32_t rng, uint32_t path, uint32_t slot) {
    const float PI = 3.14159265358979323846f;
    if (!(z_lo < 1.0f)) return dir;                                  // spread <= 0 (host stores 1) or empty z range
    const float3 b3 = norm3_fast(dir);
    const float3 diff = fabsf(b3.x) < 0.5f ? f3(1, 0, 0) : f3(0, 1, 0);
    const float3 b1 = norm3_fast(cross3(b3, diff));
    const float3 b2 = cross3(b1, b3);
    const float z = z_lo + mc_draw(rng, path, slot) * (1.0f - z_lo);
    const float r = sqrtf(1.0f - z * z);
    const float theta = -PI + mc_draw(rng, path, slot + 1) * (PI - (-PI));
    float sn, cs; __sincosf(theta, &sn, &cs);
    return norm3_fast((r * cs) * b1 + (r * sn) * b2 + z * b3);
}
__device__ __forceinline__ float jitter_z_lo(float spread) { return spread <= 0.0f ? 1.0f : cosf(spread * 3.14159265358979323846f); }

// warp-aggregated queue append: returns the slot of this lane (valid only where `emit`)
__device__ __forceinline__ uint32_t queue_append(uint32_t* counter, bool emit) {
    const uint32_t mask = __ballot_sync(__activemask(), emit);
    if (!emit) return 0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t leader = __ffs(mask) - 1u;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

}  // namespace rtx
