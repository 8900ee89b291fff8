"""ctypes mirror of include/rtx.h plus the flat scene container that feeds it.

`FlatScene` is the "flattened once" scene of the north star: plain arrays in exactly the layout
`RtxSceneDesc` points at.  It can be built from the stand-in loader (`scene_loader.Scene`), saved
to / loaded from a compressed .npz (the committed fixtures under tests/golden/scenes — the
reference's scene files do not exist on the GPU box) and turned into an `RtxSceneDesc`.
"""
from __future__ import annotations

import ctypes as C
import io
import os
from typing import Dict, List, Optional

import numpy as np

c_f16 = C.c_float * 16
c_f3 = C.c_float * 3


class RtxTexture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgba", C.c_void_p)]


class RtxMaterial(C.Structure):
    _fields_ = [("id", C.c_uint32), ("ambient_color", c_f3), ("base_color", c_f3), ("specular_color", c_f3),
                ("texture", C.c_int32 * 8), ("texture_filtering_nearest", C.c_uint32),
                ("alpha", C.c_float), ("shininess", C.c_float), ("reflectivity", C.c_float),
                ("refraction_index", C.c_float), ("normal_map_strength", C.c_float),
                ("cast_shadow", C.c_uint32), ("receive_shadow", C.c_uint32), ("shadow_softness", C.c_float),
                ("monte_carlo", C.c_uint32), ("roughness", C.c_float),
                ("smooth_shading", C.c_uint32), ("reflection_only", C.c_uint32), ("backface_cullig", C.c_uint32)]


class RtxMesh(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("indices", C.c_void_p), ("uvs", C.c_void_p),
                ("uv_indices", C.c_void_p), ("normals", C.c_void_p), ("normals_indices", C.c_void_p),
                ("n_vertices", C.c_uint32), ("n_faces", C.c_uint32), ("n_uvs", C.c_uint32),
                ("n_uv_faces", C.c_uint32), ("n_normals", C.c_uint32), ("n_normal_faces", C.c_uint32)]


class RtxItem(C.Structure):
    _fields_ = [("id", C.c_uint32), ("shape", C.c_uint32), ("visible", C.c_uint32), ("flip_normals", C.c_uint32),
                ("trans", c_f16), ("tran_inverse", c_f16),
                ("material", C.c_int32), ("mesh", C.c_int32), ("radius", C.c_float), ("reserved", C.c_uint32)]


class RtxLight(C.Structure):
    _fields_ = [("enabled", C.c_uint32), ("id", C.c_uint32), ("light_type", C.c_uint32),
                ("pos", c_f3), ("dir", c_f3), ("color", c_f3), ("intensity", C.c_float), ("max_angle", C.c_float)]


class RtxSceneDesc(C.Structure):
    _fields_ = [("items", C.POINTER(RtxItem)), ("n_items", C.c_uint32),
                ("meshes", C.POINTER(RtxMesh)), ("n_meshes", C.c_uint32),
                ("materials", C.POINTER(RtxMaterial)), ("n_materials", C.c_uint32),
                ("textures", C.POINTER(RtxTexture)), ("n_textures", C.c_uint32),
                ("lights", C.POINTER(RtxLight)), ("n_lights", C.c_uint32)]


class RtxCamera(C.Structure):
    _fields_ = [("projection_inverse", c_f16), ("view_inverse", c_f16), ("width", C.c_uint32), ("height", C.c_uint32)]


class RtxConfig(C.Structure):
    _fields_ = [("monte_carlo", C.c_uint32), ("samples", C.c_uint32), ("focal_length", C.c_float),
                ("aperture_size", C.c_float), ("fog_density", C.c_float), ("fog_color", c_f3),
                ("max_recursion", C.c_uint32), ("gamma_correction", C.c_uint32), ("mc_seed", C.c_uint32),
                ("debug_flags", C.c_uint32)]


class RtxShard(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("tile_w", C.c_uint32), ("tile_h", C.c_uint32)]


class RtxStats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("primary_samples", C.c_uint64),
                ("node_visits", C.c_uint64 * 2), ("tri_tests", C.c_uint64 * 2), ("sphere_tests", C.c_uint64),
                ("item_tests", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("waves", C.c_uint32), ("batches", C.c_uint32),
                ("device_ms", C.c_float), ("closest_ms", C.c_float), ("shadow_ms", C.c_float), ("shade_ms", C.c_float),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("rays_shadow_skipped", C.c_uint64),
                ("host_syncs", C.c_uint32), ("reserved", C.c_uint32), ("rays_shadow_exact", C.c_uint64), ("rays_shadow_beyond", C.c_uint64)]

    def as_dict(self) -> dict:
        return {k: (list(getattr(self, k)) if hasattr(getattr(self, k), '__len__') else getattr(self, k)) for k, _ in self._fields_}


class RtxRay(C.Structure):
    _fields_ = [("origin", c_f3), ("dir", c_f3)]


class RtxHit(C.Structure):
    _fields_ = [("t", C.c_float), ("normal", c_f3), ("item_id", C.c_uint32), ("face_id", C.c_uint32),
                ("item_index", C.c_int32), ("reserved", C.c_uint32)]


class RtxShadowHit(C.Structure):
    _fields_ = [("k", C.c_float), ("lit", C.c_int32), ("occluder_index", C.c_int32), ("t", C.c_float),
                ("face_id", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class RtxBvhInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_triangles", C.c_uint32), ("n_items", C.c_uint32),
                ("tlas_nodes", C.c_uint32), ("node_bytes", C.c_uint64), ("triangle_bytes", C.c_uint64),
                ("item_bytes", C.c_uint64), ("texture_bytes", C.c_uint64), ("build_ms", C.c_float),
                ("grouped_items", C.c_uint32), ("grouped_triangles", C.c_uint32), ("device_build_ms", C.c_float)]


class RtxItemXform(C.Structure):
    _fields_ = [("item_index", C.c_uint32), ("trans", c_f16), ("tran_inverse", c_f16)]


HIT_DTYPE = np.dtype([("t", "<f4"), ("normal", "<f4", 3), ("item_id", "<u4"), ("face_id", "<u4"),
                      ("item_index", "<i4"), ("reserved", "<u4")])
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("dir", "<f4", 3)])
SHADOW_HIT_DTYPE = np.dtype([("k", "<f4"), ("lit", "<i4"), ("occluder_index", "<i4"), ("t", "<f4"), ("face_id", "<u4"),
                             ("reserved", "<u4", 3)])
assert SHADOW_HIT_DTYPE.itemsize == C.sizeof(RtxShadowHit)
assert HIT_DTYPE.itemsize == C.sizeof(RtxHit) and RAY_DTYPE.itemsize == C.sizeof(RtxRay)


def colmajor16(m: np.ndarray):
    """4x4 (row, col) numpy matrix -> column-major float[16] (nalgebra layout)."""
    return c_f16(*[float(x) for x in np.asarray(m, dtype=np.float32).T.reshape(16)])


def make_camera(cam) -> RtxCamera:
    return RtxCamera(colmajor16(cam.projection_inverse), colmajor16(cam.view_inverse), int(cam.width), int(cam.height))


def make_config(cfg=None, **over) -> RtxConfig:
    from .scene_loader import Config
    cfg = cfg or Config()
    vals = {k: getattr(cfg, k) for k in ("monte_carlo", "samples", "focal_length", "aperture_size", "fog_density",
                                          "fog_color", "max_recursion", "gamma_correction", "mc_seed",
                                          "debug_flags")}
    vals.update(over)
    return RtxConfig(int(bool(vals["monte_carlo"])), int(vals["samples"]), float(vals["focal_length"]),
                     float(vals["aperture_size"]), float(vals["fog_density"]), c_f3(*[float(x) for x in vals["fog_color"]]),
                     int(vals["max_recursion"]), int(bool(vals["gamma_correction"])), int(vals["mc_seed"]),
                     int(vals["debug_flags"]))


class FlatScene:
    """Plain-array scene in RtxSceneDesc layout."""

    def __init__(self):
        self.items = (RtxItem * 0)()
        self.materials = (RtxMaterial * 0)()
        self.lights = (RtxLight * 0)()
        self.mesh_arrays: List[Dict[str, np.ndarray]] = []
        self.textures: List[np.ndarray] = []           # (h, w, 4) uint8
        self.item_names: List[str] = []
        self.extras: Dict[str, np.ndarray] = {}
        self._keep = []

    # ---- construction ---------------------------------------------------------------------
    @staticmethod
    def from_scene(scene) -> "FlatScene":
        from .scene_loader import mat_inverse, SHAPE_MESH
        fs = FlatScene()
        tex_index: Dict[str, int] = {}
        mats: List = []
        mat_index: Dict[int, int] = {}

        def mat_of(m) -> int:
            if id(m) in mat_index:
                return mat_index[id(m)]
            tex = []
            for t in m.textures:
                if t is None:
                    tex.append(-1)
                else:
                    if t not in tex_index:
                        tex_index[t] = len(fs.textures)
                        fs.textures.append(np.ascontiguousarray(scene.texture_data[t]))
                    tex.append(tex_index[t])
            rm = RtxMaterial(int(m.id), c_f3(*map(float, m.ambient_color)), c_f3(*map(float, m.base_color)),
                             c_f3(*map(float, m.specular_color)), (C.c_int32 * 8)(*tex),
                             int(m.texture_filtering_nearest), float(m.alpha), float(m.shininess),
                             float(m.reflectivity), float(m.refraction_index), float(m.normal_map_strength),
                             int(m.cast_shadow), int(m.receive_shadow), float(m.shadow_softness),
                             int(m.monte_carlo), float(m.roughness), int(m.smooth_shading),
                             int(m.reflection_only), int(m.backface_cullig))
            mat_index[id(m)] = len(mats)
            mats.append(rm)
            return mat_index[id(m)]

        items = []
        for it in scene.items:
            inv = mat_inverse(it.trans)
            if inv is None:
                raise ValueError("item %s: singular transform (reference would panic, shape/mod.rs:766)" % it.name)
            mesh_idx = -1
            if it.shape == SHAPE_MESH:
                mesh_idx = len(fs.mesh_arrays)
                md = it.mesh
                fs.mesh_arrays.append({
                    "vertices": np.ascontiguousarray(md.vertices, dtype=np.float32),
                    "indices": np.ascontiguousarray(md.indices, dtype=np.uint32),
                    "uvs": np.ascontiguousarray(md.uvs, dtype=np.float32),
                    "uv_indices": np.ascontiguousarray(md.uv_indices, dtype=np.uint32),
                    "normals": np.ascontiguousarray(md.normals, dtype=np.float32),
                    "normals_indices": np.ascontiguousarray(md.normals_indices, dtype=np.uint32)})
            items.append(RtxItem(int(it.id), int(it.shape), int(it.visible), int(it.flip_normals),
                                 colmajor16(it.trans), colmajor16(inv), mat_of(it.material), mesh_idx,
                                 float(it.radius), 0))
            fs.item_names.append(it.name)
        fs.items = (RtxItem * len(items))(*items)
        fs.materials = (RtxMaterial * len(mats))(*mats)
        lights = [RtxLight(int(l.enabled), int(l.id), int(l.light_type), c_f3(*map(float, l.pos)),
                           c_f3(*map(float, l.dir)), c_f3(*map(float, l.color)), float(l.intensity),
                           float(l.max_angle)) for l in scene.lights]
        fs.lights = (RtxLight * len(lights))(*lights)
        return fs

    # ---- (de)serialisation ----------------------------------------------------------------
    def save(self, path: str, **extras) -> None:
        """extras: named numpy arrays stored alongside (camera matrices, config, ...), read back in `.extras`."""
        arrs = {"x_" + k: np.asarray(v) for k, v in extras.items()}
        arrs.update({"items": np.frombuffer(bytes(self.items), dtype=np.uint8),
                "materials": np.frombuffer(bytes(self.materials), dtype=np.uint8),
                "lights": np.frombuffer(bytes(self.lights), dtype=np.uint8),
                "n_meshes": np.array([len(self.mesh_arrays)]), "n_textures": np.array([len(self.textures)]),
                "item_names": np.array(self.item_names)})
        for i, m in enumerate(self.mesh_arrays):
            for k, v in m.items():
                arrs["mesh%d_%s" % (i, k)] = v
        for i, t in enumerate(self.textures):
            arrs["tex%d" % i] = t
        np.savez_compressed(path, **arrs)

    @staticmethod
    def load(path: str) -> "FlatScene":
        z = np.load(path, allow_pickle=False)
        fs = FlatScene()

        def arr(cls, raw):
            n = raw.size // C.sizeof(cls)
            a = (cls * n)()
            C.memmove(a, raw.tobytes(), raw.size)
            return a
        fs.items = arr(RtxItem, z["items"])
        fs.materials = arr(RtxMaterial, z["materials"])
        fs.lights = arr(RtxLight, z["lights"])
        fs.item_names = [str(s) for s in z["item_names"]]
        for i in range(int(z["n_meshes"][0])):
            fs.mesh_arrays.append({k: np.ascontiguousarray(z["mesh%d_%s" % (i, k)]) for k in
                                   ("vertices", "indices", "uvs", "uv_indices", "normals", "normals_indices")})
        for i in range(int(z["n_textures"][0])):
            fs.textures.append(np.ascontiguousarray(z["tex%d" % i]))
        fs.extras = {k[2:]: z[k] for k in z.files if k.startswith("x_")}
        return fs

    # ---- C view ---------------------------------------------------------------------------
    def desc(self) -> RtxSceneDesc:
        meshes = (RtxMesh * len(self.mesh_arrays))()
        for i, m in enumerate(self.mesh_arrays):
            def p(a):
                return a.ctypes.data if a.size else None
            meshes[i] = RtxMesh(p(m["vertices"]), p(m["indices"]), p(m["uvs"]), p(m["uv_indices"]), p(m["normals"]),
                                p(m["normals_indices"]), m["vertices"].shape[0], m["indices"].shape[0],
                                m["uvs"].shape[0], m["uv_indices"].shape[0], m["normals"].shape[0],
                                m["normals_indices"].shape[0])
        texs = (RtxTexture * len(self.textures))()
        for i, t in enumerate(self.textures):
            texs[i] = RtxTexture(t.shape[1], t.shape[0], t.ctypes.data)
        self._keep = [meshes, texs]
        return RtxSceneDesc(C.cast(self.items, C.POINTER(RtxItem)), len(self.items),
                            C.cast(meshes, C.POINTER(RtxMesh)), len(meshes),
                            C.cast(self.materials, C.POINTER(RtxMaterial)), len(self.materials),
                            C.cast(texs, C.POINTER(RtxTexture)), len(texs),
                            C.cast(self.lights, C.POINTER(RtxLight)), len(self.lights))

    @property
    def n_triangles(self) -> int:
        return int(sum(m["indices"].shape[0] for m in self.mesh_arrays))


def bind(lib: C.CDLL, prefix: str = "rtx_") -> None:
    """Declare argument / return types of the C ABI on a loaded library.  `prefix` lets the test
    suite bind the CPU oracle (which exports the same signatures as `oracle_*`)."""
    P = C.POINTER
    sig = {
        "scene_create": (C.c_int, [P(RtxSceneDesc), C.c_int, P(C.c_void_p)]),
        "scene_create_ex": (C.c_int, [P(RtxSceneDesc), C.c_int, C.c_uint32, P(C.c_void_p)]),
        "scene_update_items": (C.c_int, [C.c_void_p, P(RtxItemXform), C.c_size_t]),
        "scene_set_lights": (C.c_int, [C.c_void_p, P(RtxLight), C.c_uint32]),
        "render_frame": (C.c_int, [C.c_void_p, P(RtxCamera), P(RtxConfig), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, P(RtxStats)]),
        "render_frame_device": (C.c_int, [C.c_void_p, P(RtxCamera), P(RtxConfig), P(RtxShard), C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, P(RtxStats)]),
        "shard_pixel_count": (C.c_uint64, [C.c_uint32, C.c_uint32, P(RtxShard)]),
        "shard_packed_bytes": (C.c_uint64, [C.c_uint32, C.c_uint32, P(RtxShard)]),
        "shard_pack": (C.c_int, [C.c_uint32, C.c_uint32, P(RtxShard), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
        "shard_unpack": (C.c_int, [C.c_uint32, C.c_uint32, P(RtxShard), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
        "shard_unpack_all": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "bvh_build_probe": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, P(C.c_uint32), P(C.c_uint32), C.c_void_p]),
        "bandwidth_probe": (C.c_int, [C.c_int, C.c_uint64, C.c_uint32, P(C.c_float)]),
        "render_frame_async": (C.c_int, [C.c_void_p, P(RtxCamera), P(RtxConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "render_poll": (C.c_int, [C.c_void_p, P(C.c_uint64), P(C.c_int), P(C.c_int), P(C.c_int), P(RtxStats)]),
        "render_stop": (C.c_int, [C.c_void_p]),
        "render_snapshot": (C.c_int, [C.c_void_p, P(C.c_uint64)]),
        "trace_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_void_p]),
        "shadow_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]),
        "scene_create_multi": (C.c_int, [P(RtxSceneDesc), P(C.c_int), C.c_uint32, P(C.c_void_p)]),
        "scene_device_count": (C.c_int, [C.c_void_p]),
        "gbuffer_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, P(C.c_void_p)]),
        "gbuffer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
        "gbuffer_open": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, P(C.c_void_p)]),
        "gbuffer_pointers": (C.c_int, [C.c_void_p, P(C.c_void_p), P(C.c_void_p), P(C.c_void_p), P(C.c_void_p)]),
        "gbuffer_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "gbuffer_destroy": (C.c_int, [C.c_void_p]),
        "sample_table": (C.c_int, [C.c_uint32, P(C.c_uint32), C.c_void_p]),
        "scene_bvh_info": (C.c_int, [C.c_void_p, P(RtxBvhInfo)]),
        "scene_destroy": (C.c_int, [C.c_void_p]),
        "last_error": (C.c_char_p, []),
        "abi_version": (C.c_int, []),
        "device_count": (C.c_int, []),
        "post_process_device": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, prefix + name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args


ABI_SYMBOLS = ["rtx_scene_create", "rtx_scene_update_items", "rtx_scene_set_lights", "rtx_render_frame",
               "rtx_render_frame_device", "rtx_render_frame_async", "rtx_render_poll", "rtx_render_stop", "rtx_render_snapshot", "rtx_shard_pixel_count", "rtx_shard_packed_bytes", "rtx_shard_pack",
               "rtx_shard_unpack", "rtx_trace_probe", "rtx_sample_table", "rtx_scene_bvh_info",
               "rtx_scene_destroy", "rtx_last_error", "rtx_abi_version", "rtx_device_count",
               "rtx_post_process_device", "rtx_shadow_probe", "rtx_scene_create_multi", "rtx_scene_device_count",
               "rtx_gbuffer_create", "rtx_gbuffer_export", "rtx_gbuffer_open", "rtx_gbuffer_pointers", "rtx_gbuffer_download",
               "rtx_gbuffer_destroy", "rtx_shard_unpack_all", "rtx_bandwidth_probe", "rtx_bvh_build_probe", "rtx_scene_create_ex"]


def fixture_path(name: str) -> str:
    """Committed flat-scene fixture (tests/golden/scenes/<name>.npz, made by tests/golden/make_fixtures.py)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(root, "tests", "golden", "scenes", name + ".npz")


def load_fixture(name: str, **cfg_over):
    """-> (FlatScene, RtxCamera, RtxConfig) of a committed fixture; cfg_over overrides config fields."""
    fs = FlatScene.load(fixture_path(name))
    x = fs.extras
    cam = RtxCamera(colmajor16(x["projection_inverse"]), colmajor16(x["view_inverse"]), int(x["size"][0]), int(x["size"][1]))
    c = x["config"]
    vals = dict(monte_carlo=int(c[0]), samples=int(c[1]), focal_length=float(c[2]), aperture_size=float(c[3]),
                fog_density=float(c[4]), fog_color=(float(c[5]), float(c[6]), float(c[7])), max_recursion=int(c[8]),
                gamma_correction=int(c[9]), mc_seed=0, debug_flags=0)
    vals.update(cfg_over)
    cfg = RtxConfig(int(vals["monte_carlo"]), int(vals["samples"]), vals["focal_length"], vals["aperture_size"],
                    vals["fog_density"], c_f3(*vals["fog_color"]), int(vals["max_recursion"]), int(vals["gamma_correction"]),
                    int(vals["mc_seed"]), int(vals["debug_flags"]))
    return fs, cam, cfg


def resize_camera(cam: RtxCamera, width: int, height: int) -> RtxCamera:
    """Same camera at another resolution (Camera::init with a new size only changes the aspect term
    P^-1[0][0] = aspect * tan(fov/2), reference src/camera.rs:69-90)."""
    out = RtxCamera()
    C.memmove(C.byref(out), C.byref(cam), C.sizeof(RtxCamera))
    tan_half = np.float32(cam.projection_inverse[5])
    out.projection_inverse[0] = float(np.float32(np.float32(width) / np.float32(height)) * tan_half)
    out.width, out.height = int(width), int(height)
    return out
