/*
 * rtx.h — C ABI of the B200 ray-casting hot path (librtx_b200.so).
 *
 * This is the drop-in boundary for rustray's per-pixel render loop.  The reference has no
 * FFI; the seam it replaces is the Rust call chain
 *     RendererManager::start / start_thread      (reference src/renderer.rs:105-172, 253-318)
 *       -> Raytracing::render(x, y) -> PixelData (reference src/raytracing.rs:275-427)
 *       -> Run::apply_pixels (four frame buffers) (reference src/run.rs:506-545)
 * i.e. "render this frame of this scene with this config into image / normals / depth /
 * objects".  Every entry point below cites the reference interface it stands in for.
 *
 * Conventions
 *  - plain C, no torch / CUDA types in signatures (device pointers are void*, streams void*).
 *  - all 4x4 matrices are COLUMN-MAJOR float[16] (element (r,c) at [c*4+r]) — the in-memory
 *    layout of nalgebra::Matrix4<f32>, so a Rust host passes `m.as_slice().as_ptr()`.
 *  - every function returns 0 on success or a negative RTX_E_* code; the reference's
 *    `unwrap()` panics become error codes (rtx_last_error() gives the text).
 *  - the caller owns all input memory; rtx_scene_create copies what it needs to the device.
 *  - frame buffers are indexed y*width+x (reference src/run.rs:118, 533-536).
 */
#ifndef RTX_B200_H
#define RTX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTX_ABI_VERSION 2

/* error codes */
#define RTX_OK              0
#define RTX_E_INVALID      -1   /* bad argument / inconsistent scene description            */
#define RTX_E_CUDA         -2   /* CUDA runtime error (no device, OOM, launch failure)       */
#define RTX_E_NO_DEVICE    -3   /* no CUDA device: this library has no CPU fallback          */
#define RTX_E_NON_AFFINE   -4   /* item transform is not affine (reference would panic in
                                   Vector3::from_homogeneous, src/shape/mod.rs:760)          */
#define RTX_E_EMPTY_MESH   -5   /* mesh with 0 triangles (parry TriMesh::new panics)         */
#define RTX_E_CANCELLED    -6   /* rtx_render_stop was called while the frame was in flight  */
#define RTX_E_BUSY         -7   /* a frame is in flight on this scene handle (frames, item / light updates; probes are
                                   always allowed: they use their own stream and buffers)     */

/* TextureType order — reference src/shape/mod.rs:633-643 */
enum {
    RTX_TEX_BASE = 0, RTX_TEX_AMBIENT_EMISSIVE = 1, RTX_TEX_SPECULAR = 2, RTX_TEX_NORMAL = 3,
    RTX_TEX_ALPHA = 4, RTX_TEX_ROUGHNESS = 5, RTX_TEX_AMBIENT_OCCLUSION = 6,
    RTX_TEX_REFLECTIVITY = 7, RTX_TEX_COUNT = 8
};

/* LightType — reference src/scene.rs:32-37 */
enum { RTX_LIGHT_DIRECTIONAL = 0, RTX_LIGHT_POINT = 1, RTX_LIGHT_SPOT = 2 };

/* shape kind — reference src/shape/sphere.rs (Sphere), src/shape/mesh.rs (Mesh) */
enum { RTX_SHAPE_SPHERE = 0, RTX_SHAPE_MESH = 1 };

/* A decoded texture: what `DynamicImage::get_pixel(x,y).to_rgba()` returns for every texel
 * (reference src/shape/mod.rs:521-531), row-major, 4 bytes per texel. */
typedef struct RtxTexture {
    uint32_t width, height;
    const uint8_t* rgba;
} RtxTexture;

/* Material — reference src/shape/mod.rs:95-134 (same field names, same defaults :137-180) */
typedef struct RtxMaterial {
    uint32_t id;
    float ambient_color[3];
    float base_color[3];
    float specular_color[3];
    int32_t  texture[RTX_TEX_COUNT];     /* index into RtxSceneDesc.textures, -1 = none (width 0) */
    uint32_t texture_filtering_nearest;
    float alpha, shininess, reflectivity, refraction_index, normal_map_strength;
    uint32_t cast_shadow, receive_shadow;
    float shadow_softness;
    uint32_t monte_carlo;
    float roughness;
    uint32_t smooth_shading, reflection_only, backface_cullig /* sic */;
} RtxMaterial;

/* Mesh buffers — reference src/shape/mesh.rs:10-21.  uv / normal arrays may be empty. */
typedef struct RtxMesh {
    const float*    vertices;         /* n_vertices * 3 */
    const uint32_t* indices;          /* n_faces * 3    */
    const float*    uvs;              /* n_uvs * 2      */
    const uint32_t* uv_indices;       /* n_uv_faces * 3 */
    const float*    normals;          /* n_normals * 3  */
    const uint32_t* normals_indices;  /* n_normal_faces * 3 */
    uint32_t n_vertices, n_faces, n_uvs, n_uv_faces, n_normals, n_normal_faces;
} RtxMesh;

/* Item = ShapeBasics + shape payload — reference src/shape/mod.rs:661-680.
 * `tran_inverse` is what ShapeBasics::calc_inverse stored (:763-767); the local AABB and the
 * material cache (:33-38) are derived by the library from mesh/radius and the material. */
typedef struct RtxItem {
    uint32_t id;
    uint32_t shape;          /* RTX_SHAPE_* */
    uint32_t visible, flip_normals;
    float trans[16];
    float tran_inverse[16];
    int32_t  material;       /* index into materials */
    int32_t  mesh;           /* index into meshes (RTX_SHAPE_MESH)  */
    float    radius;         /* Ball radius (RTX_SHAPE_SPHERE)      */
    uint32_t reserved;
} RtxItem;

/* Light — reference src/scene.rs:40-51 */
typedef struct RtxLight {
    uint32_t enabled, id, light_type;
    float pos[3], dir[3], color[3];
    float intensity, max_angle /* rad */;
} RtxLight;

/* The flattened result of Scene::load* + init + update (reference src/scene.rs:121-157,
 * 1666-1688).  Item order == scene.items order (it decides shadow first-hit ties). */
typedef struct RtxSceneDesc {
    const RtxItem*     items;      uint32_t n_items;
    const RtxMesh*     meshes;     uint32_t n_meshes;
    const RtxMaterial* materials;  uint32_t n_materials;
    const RtxTexture*  textures;   uint32_t n_textures;
    const RtxLight*    lights;     uint32_t n_lights;
} RtxSceneDesc;

/* Camera — the two matrices Raytracing::render reads (reference src/camera.rs:34-38,79-90)
 * plus the frame size (cam.width / cam.height). */
typedef struct RtxCamera {
    float projection_inverse[16];
    float view_inverse[16];
    uint32_t width, height;
} RtxCamera;

/* RaytracingConfig — reference src/raytracing.rs:92-127 (same names, same defaults).
 * `mc_seed` is an extension: the reference draws Monte-Carlo jitter from thread_rng()
 * (:616-618), which is irreproducible; here it is a counter-based generator keyed on
 * (mc_seed, pixel, sample, ray path, slot). */
typedef struct RtxConfig {
    uint32_t monte_carlo;
    uint32_t samples;
    float focal_length, aperture_size;
    float fog_density, fog_color[3];
    uint32_t max_recursion;
    uint32_t gamma_correction;
    uint32_t mc_seed;
    uint32_t debug_flags;               /* RTX_DEBUG_* bits, 0 in production */
} RtxConfig;

#define RTX_DEBUG_COLLECT_STATS  1u   /* count node visits / primitive tests per ray (slower)        */
#define RTX_OPT_SKIP_ZERO_SHADOW 4u   /* opt-in: a shadow ray whose light contribution is exactly (0,0,0) cannot change the
                                         frame and is not traced; still counted in rays_shadow (the reference calls trace
                                         for it) and reported in rays_shadow_skipped.  Off by default: the bench traces
                                         every ray the reference traces.                                              */
#define RTX_DEBUG_SERIAL_STREAMS  8u   /* measurement: shadow kernels on the frame's own stream instead of overlapping the next
                                         wave, so that per-kernel CUDA-event times (RtxStats.closest_ms / shadow_ms) are
                                         exclusive; the frame is a few percent slower                                     */
#define RTX_DEBUG_ORDERED_SHADOW 2u   /* shadow rays: literal "closest hit of every item in bbox
                                         order" walk instead of the equivalent two-phase any-hit     */

/* Interleaved-tile shard of one frame: tile t (row-major over ceil(w/tile_w) x ceil(h/tile_h))
 * belongs to rank (t + rot(t / world)) % world with rot(g) = (g * 0x9E3779B1 mod 2^32) >> 16: one tile per rank in every
 * group of `world` consecutive tiles, rotated per group (balanced and decorrelated from image columns).
 * world = 1 renders everything. */
typedef struct RtxShard {
    uint32_t rank, world, tile_w, tile_h;
} RtxShard;

/* Per-frame counters.  One ray == one Raytracing::trace call (reference
 * src/raytracing.rs:726 closest-hit, :883 shadow). */
typedef struct RtxStats {
    uint64_t rays_closest, rays_shadow;
    uint64_t primary_samples;
    /* traversal work, [0] = closest-hit kernel, [1] = shadow kernel; only with RTX_DEBUG_COLLECT_STATS */
    uint64_t node_visits[2], tri_tests[2];
    uint64_t sphere_tests, item_tests;
    uint64_t kernel_launches;
    uint32_t waves, batches;
    float device_ms;       /* CUDA events around all device work of the frame                      */
    float closest_ms;      /* CUDA events, sum over closest-hit kernel launches                     */
    float shadow_ms;       /* CUDA events, sum over shadow kernel launches                          */
    float shade_ms;        /* device_ms - closest_ms - shadow_ms (raygen, shade, resolve, gaps)     */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t rays_shadow_skipped;   /* RTX_OPT_SKIP_ZERO_SHADOW: shadow rays not traced (included in rays_shadow) */
    uint32_t host_syncs;            /* stream synchronisations inside the frame: 1 for a sync-free frame, waves + 1 otherwise */
    uint32_t reserved;
    uint64_t rays_shadow_exact;     /* shadow rays whose answer could depend on the reference's first-hit order and that were
                                       re-walked item by item (shadow_exact_kernel); included in rays_shadow */
    uint64_t rays_shadow_beyond;    /* occluded shadow rays with a finite light distance that were checked for a hit BEHIND the
                                       light by an item sorting before the occluder (shadow_beyond_kernel, merged BLAS only) */
} RtxStats;

typedef struct RtxRay { float origin[3]; float dir[3]; } RtxRay;
/* Result of Raytracing::trace: Option<(f32, Vector3, &dyn Shape, u32)> (src/raytracing.rs:429) */
typedef struct RtxHit {
    float t;              /* < 0  => None */
    float normal[3];
    uint32_t item_id;     /* ShapeBasics.id, 0 on miss */
    uint32_t face_id;     /* parry FeatureId::Face index (+n_faces on a backface), 0 for spheres */
    int32_t  item_index;  /* position in RtxSceneDesc.items, -1 on miss */
    uint32_t reserved;
} RtxHit;

/* One shadow query answered by the PRODUCTION shadow kernels (any-hit + reference-order walk), i.e. the outcome of
 * `trace(&shadow_ray, true, true, depth)` followed by the light-distance test and the attenuation of
 * reference src/raytracing.rs:883-914 for a unit light contribution:
 *   k              1 when lit, else 1 - receiver.alpha [* occluder alpha-texture texel at the receiver's uv (sic, :905)]
 *   lit            1 = no item hit at all, or the first item hit (bbox-key order) is hit beyond the light distance
 *   occluder_index an item with a hit at toi <= light distance (-1 when lit).  It is the reference's first-hit item
 *                  whenever the order rule can matter (alpha-textured occluder, or a finite light distance with several
 *                  candidates); otherwise any occluder
 *   t, face_id     hit distance / parry face id on the occluder, only when its material has an alpha texture (else -1, 0) */
typedef struct RtxShadowHit {
    float k; int32_t lit; int32_t occluder_index; float t; uint32_t face_id; uint32_t reserved[3];
} RtxShadowHit;

typedef struct RtxBvhInfo {
    uint32_t n_nodes, n_triangles, n_items, tlas_nodes;
    uint64_t node_bytes, triangle_bytes, item_bytes, texture_bytes;
    float build_ms;
    uint32_t grouped_items, grouped_triangles;   /* identity-transform mesh items merged into one world-space BLAS (their triangles
                                                    are counted twice in n_triangles: the per-mesh BLASes stay for the exact walk) */
    float device_build_ms;                        /* RTX_SCENE_DEVICE_BVH: time spent in the builder kernels (part of build_ms) */
} RtxBvhInfo;

typedef struct RtxScene RtxScene;   /* opaque */

/* Scene::load + init + update, flattened (reference src/scene.rs:121-157,1674-1688): copies the
 * description, builds the per-mesh wide BVHs and the item-level structure, uploads to
 * `device` (CUDA ordinal).  Fails with RTX_E_NO_DEVICE when there is no GPU. */
int rtx_scene_create(const RtxSceneDesc* desc, int device, RtxScene** out);

/* Same, with flags.  RTX_SCENE_DEVICE_BVH: the per-mesh wide BVHs (and the merged world-space BLAS) are built by kernels
 * — Morton codes, radix sort, binary radix tree, bottom-up fit, level-by-level collapse into the same quantised 8-wide nodes —
 * instead of the host's binned-SAH builder: tens of milliseconds for 10 M triangles, for scenes that are rebuilt often
 * (Scene::update rebuilds on every start, reference src/scene.rs:1674-1688).  The tree only prunes, so every traversal result
 * is identical; a Morton tree costs more node visits per ray, which is why it is opt-in (also: environment RTX_DEVICE_BVH=1). */
#define RTX_SCENE_DEVICE_BVH 1u
int rtx_scene_create_ex(const RtxSceneDesc* desc, int device, uint32_t flags, RtxScene** out);

/* Same scene on several GPUs of ONE process — the reference has one RendererManager::start per frame
 * (src/renderer.rs:105-172), so a host that owns that call cannot be one process per GPU.  The scene is built once on
 * devices[0] and copied device-to-device to the others; rtx_render_frame / _async / _device (shard = NULL) then split
 * the frame into interleaved tiles, one host thread and one stream per device, and every device's resolve kernel stores
 * its finished pixels directly into the frame buffers on devices[0] through NVLink peer memory (no separate gather).
 * Needs peer access between devices[0] and every other listed device.  All other calls take the handle unchanged. */
int rtx_scene_create_multi(const RtxSceneDesc* desc, const int* devices, uint32_t n_devices, RtxScene** out);
int rtx_scene_device_count(const RtxScene* scene);

/* Scene::apply_frame / ShapeBasics::apply_mat + Scene::update (reference src/scene.rs:1695-1713,
 * src/shape/mod.rs:748-753,1674-1688): replace transforms of n items (by position in items). */
typedef struct RtxItemXform { uint32_t item_index; float trans[16]; float tran_inverse[16]; } RtxItemXform;
int rtx_scene_update_items(RtxScene* scene, const RtxItemXform* xforms, size_t n);

/* Replace lights / one material's scalar fields between frames (GUI editors, src/run.rs). */
int rtx_scene_set_lights(RtxScene* scene, const RtxLight* lights, uint32_t n_lights);

/* RendererManager::start .. is_done + every Run::apply_pixels write (reference
 * src/renderer.rs:105-172,228; src/run.rs:506-545).  Blocking.  HOST output buffers:
 * rgba w*h*4 (a = 255, run.rs:527), normals w*h*3, depth w*h, object_ids w*h.
 * Any output pointer may be NULL.  `stats` may be NULL. */
int rtx_render_frame(RtxScene* scene, const RtxCamera* cam, const RtxConfig* cfg,
                     uint8_t* rgba, float* normals, float* depth, uint32_t* object_ids,
                     RtxStats* stats);

/* Non-blocking variant for the GUI path: RendererManager::start returns at once and the window polls
 * is_running / is_done / get_rendered_pixels and may call stop (reference src/renderer.rs:105-172, 174-231).
 * rtx_render_frame_async starts the frame on a worker thread (same HOST buffers as rtx_render_frame, they must stay
 * valid until done); rtx_render_poll reports progress — pixels_rendered counts pixels whose samples have all been
 * issued, and equals width*height exactly when the frame is complete and the buffers are filled; `result` receives the
 * frame's return code once it is no longer running.  rtx_render_stop cancels between waves and joins the worker. */
int rtx_render_frame_async(RtxScene* scene, const RtxCamera* cam, const RtxConfig* cfg,
                           uint8_t* rgba, float* normals, float* depth, uint32_t* object_ids);
int rtx_render_poll(RtxScene* scene, uint64_t* pixels_rendered, int* running, int* done, int* result, RtxStats* stats);
int rtx_render_stop(RtxScene* scene);
/* Progressive display (the reference streams finished pixels to the window through an mpsc channel, src/renderer.rs:
 * 305-312 -> src/run.rs:506-545): while an async frame is in flight, refresh ITS host buffers with the current state —
 * pixels whose samples have all been issued are normalised by the sample count, the pixel range in progress by the samples
 * issued so far, untouched pixels are zero.  Returns after the buffers were written (at the next wave boundary), or at once
 * when no frame is running (the buffers then hold the finished frame).  *pixels_rendered as in rtx_render_poll. */
int rtx_render_snapshot(RtxScene* scene, uint64_t* pixels_rendered);

/* Same frame, DEVICE output buffers (same layouts) on the scene's device, restricted to the
 * pixels of `shard` (NULL = whole frame); pixels of other ranks are left untouched.
 * Work is enqueued on `cuda_stream` (a cudaStream_t, NULL = default stream) and the call
 * returns after the stream has drained (the wavefront loop reads queue sizes back). */
int rtx_render_frame_device(RtxScene* scene, const RtxCamera* cam, const RtxConfig* cfg,
                            const RtxShard* shard,
                            void* d_rgba, void* d_normals, void* d_depth, void* d_object_ids,
                            void* cuda_stream, RtxStats* stats);

/* Multi-GPU gather helpers: pack the owned pixels of a shard into a compact device buffer
 * (layout: n*4 rgba bytes | n*3 normal floats | n depth floats | n id uint32, n =
 * rtx_shard_pixel_count) and scatter such a buffer (from any rank) into full frame buffers. */
uint64_t rtx_shard_pixel_count(uint32_t width, uint32_t height, const RtxShard* shard);
uint64_t rtx_shard_packed_bytes(uint32_t width, uint32_t height, const RtxShard* shard);
int rtx_shard_pack(uint32_t width, uint32_t height, const RtxShard* shard,
                   const void* d_rgba, const void* d_normals, const void* d_depth,
                   const void* d_object_ids, void* d_packed, void* cuda_stream);
int rtx_shard_unpack(uint32_t width, uint32_t height, const RtxShard* shard,
                     const void* d_packed, void* d_rgba, void* d_normals, void* d_depth,
                     void* d_object_ids, void* cuda_stream);

/* Rank 0 after ONE gather into a contiguous buffer (rank r's packed shard at d_packed_all + r * stride_bytes): scatter
 * the shards of ranks first_rank .. world-1 into the frame buffers with one kernel. */
int rtx_shard_unpack_all(uint32_t width, uint32_t height, uint32_t world, uint32_t tile_w, uint32_t tile_h,
                         uint32_t first_rank, const void* d_packed_all, uint64_t stride_bytes,
                         void* d_rgba, void* d_normals, void* d_depth, void* d_object_ids, void* cuda_stream);

/* Frame buffers that OTHER PROCESSES render into (one process per GPU under torchrun): 24 bytes per pixel in one
 * device allocation, rgba8[n] | normals f32[3n] | depth f32[n] | ids u32[n] (= the four buffers of Run, src/run.rs:117-120).
 * The owner exports a 64-byte CUDA IPC handle, the other ranks open it and pass rtx_gbuffer_pointers() as the outputs of
 * rtx_render_frame_device with their shard: finished pixels travel as NVLink stores from the resolve kernel, and after a
 * barrier the owner holds the whole frame.  rtx_gbuffer_download copies it to (page-locked) host memory. */
typedef struct RtxGBuffer RtxGBuffer;
int rtx_gbuffer_create(int device, uint32_t width, uint32_t height, RtxGBuffer** out);
int rtx_gbuffer_export(RtxGBuffer* g, uint8_t handle[64]);
int rtx_gbuffer_open(int device, uint32_t width, uint32_t height, const uint8_t handle[64], RtxGBuffer** out);
int rtx_gbuffer_pointers(const RtxGBuffer* g, void** d_rgba, void** d_normals, void** d_depth, void** d_object_ids);
int rtx_gbuffer_download(const RtxGBuffer* g, uint8_t* rgba, float* normals, float* depth, uint32_t* object_ids, void* cuda_stream);
int rtx_gbuffer_destroy(RtxGBuffer* g);

/* Raytracing::trace (reference src/raytracing.rs:429-490) and, with for_shadow = 0 on a
 * primary ray, Raytracing::pick (:237-273).  HOST arrays of n rays / n hits.  Ray directions
 * are used as given (the callers in the reference normalise first, :262,:723). */
int rtx_trace_probe(RtxScene* scene, const RtxRay* rays, size_t n, int for_shadow,
                    int stop_on_first_hit, uint32_t depth, RtxHit* hits);

/* Shadow rays through the production kernels (see RtxShadowHit).  light_distance: n floats, NULL = +inf for every ray
 * (directional light, raytracing.rs:887); receiver_item: index of the item the shadow ray leaves from (its material's alpha
 * and its get_uv enter the attenuation, :897,:905), NULL = alpha 1 / item 0. */
int rtx_shadow_probe(RtxScene* scene, const RtxRay* rays, const float* light_distance, const int32_t* receiver_item,
                     size_t n, uint32_t depth, RtxShadowHit* out);

/* The per-pixel sample sub-grid of Raytracing::render (reference src/raytracing.rs:290-313):
 * cell_size, and `samples` (x_i, y_i) pairs after StdRng::seed_from_u64(0) shuffle + truncate.
 * xy must hold 2*samples uint16. */
int rtx_sample_table(uint32_t samples, uint32_t* cell_size, uint16_t* xy);

int rtx_scene_bvh_info(const RtxScene* scene, RtxBvhInfo* info);
int rtx_scene_destroy(RtxScene* scene);
const char* rtx_last_error(void);
int rtx_abi_version(void);
int rtx_device_count(void);

/* The host-side wide-BVH builder on n boxes (6 floats each: lo.xyz, hi.xyz), no device involved: depth and node count of
 * the tree and its leaf order.  depth_limit > 0 = the limit rtx_scene_create applies (a deeper SAH tree is rebuilt balanced). */
int rtx_bvh_build_probe(const float* boxes, uint32_t n, int depth_limit, uint32_t* max_depth, uint32_t* n_nodes,
                        uint32_t* prim_order);

/* Read bandwidth of a `bytes`-sized buffer swept `iters` times with 128-bit loads (ld.global.cg): with bytes well below
 * the 126 MB L2 this is the L2 figure SURVEY.md §8(d) asks to quote next to the HBM roofline for L2-resident scenes. */
int rtx_bandwidth_probe(int device, uint64_t bytes, uint32_t iters, float* gbytes_per_s);

/* Post-processing on the finished G-buffer (reference src/post_processing.rs:77-181), device
 * side, in place on d_rgba.  cavity/outline as in PostProcessingConfig. */
int rtx_post_process_device(uint32_t width, uint32_t height, int cavity, int outline,
                            void* d_rgba, const void* d_normals, const void* d_depth,
                            const void* d_object_ids, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* RTX_B200_H */
