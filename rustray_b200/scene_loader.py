"""Stand-in host loader: scene JSON / OBJ -> flat scene description for the C ABI.

In production the (unchanged) Rust host keeps doing this work (reference src/scene.rs) and the
`rustray-cuda-sys` binding of INTEGRATION.md walks its `Scene` to fill `RtxSceneDesc`.  There is no
Rust toolchain in this environment, so this module reproduces the *output semantics* of the
reference loaders — id assignment order, material mapping, `apply_diff`, transform composition,
default light, camera matrices — so that tests and the bench can feed the same scenes
(`scene/spheres.json`, `scene/floor.json` + `scene/monkey.json`, ...) through `include/rtx.h`.
It is cold-path host code; nothing here is on the measured path.

Every function cites the reference lines it mirrors (paths relative to the reference repo).
All arithmetic that ends up in the scene description is done in IEEE f32 like the reference.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

F = np.float32
PI = F(math.pi)

TEX_BASE, TEX_AMBIENT, TEX_SPECULAR, TEX_NORMAL, TEX_ALPHA, TEX_ROUGHNESS, TEX_AO, TEX_REFLECTIVITY = range(8)
LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_SPOT = 0, 1, 2
SHAPE_SPHERE, SHAPE_MESH = 0, 1


# --------------------------------------------------------------------------------------------
# f32 linear algebra in nalgebra's evaluation order
# --------------------------------------------------------------------------------------------
def mat_identity() -> np.ndarray:
    return np.eye(4, dtype=F)


def mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """4x4 * 4x4 in f32, accumulating over k = 0..3 per element (nalgebra gemm = column axpy)."""
    out = np.zeros((4, 4), dtype=F)
    for j in range(4):
        for i in range(4):
            acc = F(a[i, 0] * b[0, j])
            for k in range(1, 4):
                acc = F(acc + F(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def mat_vec(a: np.ndarray, v) -> np.ndarray:
    out = np.zeros(4, dtype=F)
    for i in range(4):
        acc = F(a[i, 0] * F(v[0]))
        for k in range(1, 4):
            acc = F(acc + F(a[i, k] * F(v[k])))
        out[i] = acc
    return out


def mat_translation(x, y, z) -> np.ndarray:
    m = mat_identity()
    m[0, 3], m[1, 3], m[2, 3] = F(x), F(y), F(z)
    return m


def mat_scaling(x, y, z) -> np.ndarray:
    m = mat_identity()
    m[0, 0], m[1, 1], m[2, 2] = F(x), F(y), F(z)
    return m


def mat_euler(roll, pitch, yaw) -> np.ndarray:
    """nalgebra Rotation3::from_euler_angles(roll, pitch, yaw).to_homogeneous()."""
    sr, cr = F(np.sin(F(roll))), F(np.cos(F(roll)))
    sp, cp = F(np.sin(F(pitch))), F(np.cos(F(pitch)))
    sy, cy = F(np.sin(F(yaw))), F(np.cos(F(yaw)))
    m = mat_identity()
    m[0, 0] = cy * cp
    m[0, 1] = F(F(cy * sp) * sr) - F(sy * cr)
    m[0, 2] = F(F(cy * sp) * cr) + F(sy * sr)
    m[1, 0] = sy * cp
    m[1, 1] = F(F(sy * sp) * sr) + F(cy * cr)
    m[1, 2] = F(F(sy * sp) * cr) - F(cy * sr)
    m[2, 0] = -sp
    m[2, 1] = cp * sr
    m[2, 2] = cp * cr
    return m


def mat_inverse(mm: np.ndarray) -> Optional[np.ndarray]:
    """nalgebra Matrix4::try_inverse (4x4 cofactor expansion), f32. Used by
    ShapeBasics::calc_inverse (src/shape/mod.rs:763-767) and Camera::init_matrices
    (src/camera.rs:89)."""
    m = np.asarray(mm, dtype=F).T.reshape(16).copy()  # column-major flat, m[c*4+r]
    inv = np.zeros(16, dtype=F)

    def p(*idx):
        r = F(1.0)
        first = True
        for i in idx:
            r = F(m[i]) if first else F(r * m[i])
            first = False
        return r

    def s(terms):
        acc = None
        for sign, t in terms:
            if acc is None:
                acc = t if sign > 0 else F(-t)
            else:
                acc = F(acc + t) if sign > 0 else F(acc - t)
        return acc

    inv[0] = s([(1, p(5, 10, 15)), (-1, p(5, 11, 14)), (-1, p(9, 6, 15)), (1, p(9, 7, 14)), (1, p(13, 6, 11)), (-1, p(13, 7, 10))])
    inv[4] = s([(-1, p(4, 10, 15)), (1, p(4, 11, 14)), (1, p(8, 6, 15)), (-1, p(8, 7, 14)), (-1, p(12, 6, 11)), (1, p(12, 7, 10))])
    inv[8] = s([(1, p(4, 9, 15)), (-1, p(4, 11, 13)), (-1, p(8, 5, 15)), (1, p(8, 7, 13)), (1, p(12, 5, 11)), (-1, p(12, 7, 9))])
    inv[12] = s([(-1, p(4, 9, 14)), (1, p(4, 10, 13)), (1, p(8, 5, 14)), (-1, p(8, 6, 13)), (-1, p(12, 5, 10)), (1, p(12, 6, 9))])
    inv[1] = s([(-1, p(1, 10, 15)), (1, p(1, 11, 14)), (1, p(9, 2, 15)), (-1, p(9, 3, 14)), (-1, p(13, 2, 11)), (1, p(13, 3, 10))])
    inv[5] = s([(1, p(0, 10, 15)), (-1, p(0, 11, 14)), (-1, p(8, 2, 15)), (1, p(8, 3, 14)), (1, p(12, 2, 11)), (-1, p(12, 3, 10))])
    inv[9] = s([(-1, p(0, 9, 15)), (1, p(0, 11, 13)), (1, p(8, 1, 15)), (-1, p(8, 3, 13)), (-1, p(12, 1, 11)), (1, p(12, 3, 9))])
    inv[13] = s([(1, p(0, 9, 14)), (-1, p(0, 10, 13)), (-1, p(8, 1, 14)), (1, p(8, 2, 13)), (1, p(12, 1, 10)), (-1, p(12, 2, 9))])
    inv[2] = s([(1, p(1, 6, 15)), (-1, p(1, 7, 14)), (-1, p(5, 2, 15)), (1, p(5, 3, 14)), (1, p(13, 2, 7)), (-1, p(13, 3, 6))])
    inv[6] = s([(-1, p(0, 6, 15)), (1, p(0, 7, 14)), (1, p(4, 2, 15)), (-1, p(4, 3, 14)), (-1, p(12, 2, 7)), (1, p(12, 3, 6))])
    inv[10] = s([(1, p(0, 5, 15)), (-1, p(0, 7, 13)), (-1, p(4, 1, 15)), (1, p(4, 3, 13)), (1, p(12, 1, 7)), (-1, p(12, 3, 5))])
    inv[14] = s([(-1, p(0, 5, 14)), (1, p(0, 6, 13)), (1, p(4, 1, 14)), (-1, p(4, 2, 13)), (-1, p(12, 1, 6)), (1, p(12, 2, 5))])
    inv[3] = s([(-1, p(1, 6, 11)), (1, p(1, 7, 10)), (1, p(5, 2, 11)), (-1, p(5, 3, 10)), (-1, p(9, 2, 7)), (1, p(9, 3, 6))])
    inv[7] = s([(1, p(0, 6, 11)), (-1, p(0, 7, 10)), (-1, p(4, 2, 11)), (1, p(4, 3, 10)), (1, p(8, 2, 7)), (-1, p(8, 3, 6))])
    inv[11] = s([(-1, p(0, 5, 11)), (1, p(0, 7, 9)), (1, p(4, 1, 11)), (-1, p(4, 3, 9)), (-1, p(8, 1, 7)), (1, p(8, 3, 5))])
    inv[15] = s([(1, p(0, 5, 10)), (-1, p(0, 6, 9)), (-1, p(4, 1, 10)), (1, p(4, 2, 9)), (1, p(8, 1, 6)), (-1, p(8, 2, 5))])

    det = F(F(F(m[0] * inv[0]) + F(m[1] * inv[4])) + F(m[2] * inv[8])) + F(m[3] * inv[12])
    det = F(det)
    if det == 0:
        return None
    inv_det = F(F(1.0) / det)
    inv = (inv * inv_det).astype(F)
    return inv.reshape(4, 4).T.copy()


def approx_equal(a, b) -> bool:
    """reference src/helper.rs:11-20 (6 decimal places, truncating)."""
    fac = F(10.0) ** 6
    with np.errstate(invalid="ignore", over="ignore"):
        return bool(np.trunc(F(F(a) * F(fac))) == np.trunc(F(F(b) * F(fac))))


def to_radians(deg) -> np.float32:
    return F(F(deg) * F(PI / F(180.0)))


# --------------------------------------------------------------------------------------------
# scene objects
# --------------------------------------------------------------------------------------------
@dataclass
class Material:
    """reference src/shape/mod.rs:95-180"""
    id: int = 0
    name: str = ""
    ambient_color: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0], dtype=F))
    base_color: np.ndarray = field(default_factory=lambda: np.array([1, 1, 1], dtype=F))
    specular_color: np.ndarray = field(default_factory=lambda: np.array([0.8, 0.8, 0.8], dtype=F))
    textures: List[Optional[str]] = field(default_factory=lambda: [None] * 8)   # keys into Scene.texture_data
    texture_filtering_nearest: bool = False
    alpha: float = 1.0
    shininess: float = 150.0
    reflectivity: float = 0.0
    refraction_index: float = 1.0
    normal_map_strength: float = 1.0
    cast_shadow: bool = True
    receive_shadow: bool = True
    shadow_softness: float = 0.01
    monte_carlo: bool = True
    roughness: float = 0.0
    smooth_shading: bool = True
    reflection_only: bool = False
    backface_cullig: bool = True

    _SCALARS = ("alpha", "shininess", "reflectivity", "refraction_index", "normal_map_strength",
                "shadow_softness", "roughness")
    _BOOLS = ("texture_filtering_nearest", "cast_shadow", "receive_shadow", "monte_carlo",
              "smooth_shading", "reflection_only", "backface_cullig")

    def apply_diff(self, new: "Material") -> None:
        """Material::apply_diff (src/shape/mod.rs:182-299): copy only fields that differ from the
        DEFAULT material."""
        d = Material()
        for col in ("ambient_color", "base_color", "specular_color"):
            dv, nv = getattr(d, col), getattr(new, col)
            if any(not approx_equal(dv[i], nv[i]) for i in range(3)):
                setattr(self, col, nv.copy())
        for k in self._SCALARS:
            if not approx_equal(getattr(d, k), getattr(new, k)):
                setattr(self, k, getattr(new, k))
        for k in self._BOOLS:
            if getattr(d, k) != getattr(new, k):
                setattr(self, k, getattr(new, k))
        for t in range(8):
            if new.textures[t] is not None:
                self.textures[t] = new.textures[t]


@dataclass
class MeshData:
    """reference src/shape/mesh.rs:10-21"""
    vertices: np.ndarray
    indices: np.ndarray
    uvs: np.ndarray
    uv_indices: np.ndarray
    normals: np.ndarray
    normals_indices: np.ndarray


@dataclass
class Item:
    """ShapeBasics + payload (src/shape/mod.rs:661-680)"""
    id: int
    name: str
    shape: int
    material: Material
    trans: np.ndarray
    visible: bool = True
    flip_normals: bool = False
    radius: float = 0.0
    mesh: Optional[MeshData] = None

    def apply_transformation(self, translation, scale, rotation) -> None:
        """ShapeBasics::get_transformation (src/shape/mod.rs:708-729): trans·T·Rz·Ry·Rx·S."""
        t = self.trans
        t = mat_mul(t, mat_translation(*translation))
        t = mat_mul(t, mat_euler(0.0, 0.0, rotation[2]))
        t = mat_mul(t, mat_euler(0.0, rotation[1], 0.0))
        t = mat_mul(t, mat_euler(rotation[0], 0.0, 0.0))
        t = mat_mul(t, mat_scaling(*scale))
        self.trans = t


@dataclass
class Light:
    """reference src/scene.rs:40-51"""
    id: int
    name: str
    pos: np.ndarray
    dir: np.ndarray
    color: np.ndarray
    intensity: float
    max_angle: float
    light_type: int
    enabled: bool = True


@dataclass
class Config:
    """RaytracingConfig (src/raytracing.rs:92-127)"""
    monte_carlo: bool = False
    samples: int = 1
    focal_length: float = 1.0
    aperture_size: float = 1.0
    fog_density: float = 0.0
    fog_color: Tuple[float, float, float] = (0.4, 0.4, 0.4)
    max_recursion: int = 6
    gamma_correction: bool = False
    mc_seed: int = 0
    debug_flags: int = 0


@dataclass
class Camera:
    """reference src/camera.rs"""
    eye_pos: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0], dtype=F))
    up: np.ndarray = field(default_factory=lambda: np.array([0, 1, 0], dtype=F))
    dir: np.ndarray = field(default_factory=lambda: np.array([0, 0, -1], dtype=F))
    fov: np.float32 = to_radians(90.0)
    clipping_near: float = 0.001
    clipping_far: float = 1000.0
    width: int = 0
    height: int = 0
    projection: np.ndarray = field(default_factory=mat_identity)
    view: np.ndarray = field(default_factory=mat_identity)
    projection_inverse: np.ndarray = field(default_factory=mat_identity)
    view_inverse: np.ndarray = field(default_factory=mat_identity)

    def init(self, width: int, height: int) -> None:
        self.width, self.height = int(width), int(height)
        self.init_matrices()

    def init_matrices(self) -> None:
        """Camera::init_matrices (src/camera.rs:79-90): Perspective3::new + analytic inverse,
        Isometry3::look_at_rh + Matrix4::try_inverse."""
        aspect = F(F(self.width) / F(self.height))
        m11 = F(F(1.0) / F(np.tan(F(self.fov / F(2.0)))))
        m00 = F(m11 / aspect)
        zn, zf = F(self.clipping_near), F(self.clipping_far)
        m22 = F(F(zf + zn) / F(zn - zf))
        m23 = F(F(F(zf * zn) * F(2.0)) / F(zn - zf))
        p = np.zeros((4, 4), dtype=F)
        p[0, 0], p[1, 1], p[2, 2], p[2, 3], p[3, 2] = m00, m11, m22, m23, F(-1.0)
        self.projection = p
        pi = np.zeros((4, 4), dtype=F)
        pi[0, 0] = F(F(1.0) / m00)
        pi[1, 1] = F(F(1.0) / m11)
        pi[2, 3] = F(F(1.0) / F(-1.0))
        pi[3, 2] = F(F(1.0) / m23)
        pi[3, 3] = F(F(-m22) / F(m23 * F(-1.0)))
        self.projection_inverse = pi
        # look_at_rh: camera looks along -z
        d = self.dir.astype(F)
        z = -d / F(np.sqrt(F(np.dot(d, d))))
        x = np.cross(self.up.astype(F), z).astype(F)
        x = (x / F(np.sqrt(F(np.dot(x, x))))).astype(F)
        y = np.cross(z, x).astype(F)
        v = mat_identity()
        v[0, :3], v[1, :3], v[2, :3] = x, y, z
        e = self.eye_pos.astype(F)
        v[0, 3], v[1, 3], v[2, 3] = -F(np.dot(x, e)), -F(np.dot(y, e)), -F(np.dot(z, e))
        v[np.abs(v) == 0] = 0  # drop -0.0
        self.view = v
        self.view_inverse = mat_inverse(v)

    def is_default_cam(self) -> bool:
        """Camera::is_default_cam (src/camera.rs:92-123)"""
        return (all(approx_equal(self.eye_pos[i], 0.0) for i in range(3))
                and all(approx_equal(self.dir[i], v) for i, v in enumerate((0.0, 0.0, -1.0)))
                and all(approx_equal(self.up[i], v) for i, v in enumerate((0.0, 1.0, 0.0)))
                and approx_equal(self.fov, to_radians(90.0))
                and approx_equal(self.clipping_near, 0.001) and approx_equal(self.clipping_far, 1000.0))

    def is_point_in_frustum(self, pt) -> bool:
        pv = mat_mul(self.projection, self.view)
        c = mat_vec(pv, [pt[0], pt[1], pt[2], 1.0])
        return bool(abs(c[0]) <= c[3] and abs(c[1]) <= c[3] and abs(c[2]) <= c[3])


def _decode_texture(path: str) -> np.ndarray:
    """image::open(path) + get_pixel().to_rgba() for every texel (src/shape/mod.rs:382,521-531).
    PIL stands in for the `image` crate; JPEG IDCT output can differ from it by a level or two,
    which is why oracle and GPU are always fed these same decoded bytes."""
    from PIL import Image
    im = Image.open(path)
    im = im.convert("RGBA")
    return np.ascontiguousarray(np.asarray(im, dtype=np.uint8))


class Scene:
    """Mirror of reference `Scene` (src/scene.rs:68-83) restricted to what the hot path reads."""

    def __init__(self, asset_root: str = "."):
        self.asset_root = asset_root
        self.item_id = 0
        self.cam = Camera()
        self.items: List[Item] = []
        self.lights: List[Light] = []
        self.materials: List[Material] = []
        self.config = Config()
        self.post = {"cavity": False, "outline": False}
        self.texture_data: Dict[str, np.ndarray] = {}
        self.animation = None

    # -- helpers ------------------------------------------------------------------------------
    def _path(self, p: str) -> str:
        return p if os.path.isabs(p) else os.path.join(self.asset_root, p)

    def get_next_id(self) -> int:
        self.item_id += 1
        return self.item_id

    def load_texture(self, mat: Material, path: str, tex_type: int) -> None:
        key = os.path.normpath(path)
        if key not in self.texture_data:
            self.texture_data[key] = _decode_texture(self._path(path))
        mat.textures[tex_type] = key

    @staticmethod
    def _xyz(obj, key, default, names=("x", "y", "z")):
        """get_{point,vec,color}_from_json_object (src/scene.rs:643-702)"""
        v = np.array(default, dtype=F)
        if obj is None or not isinstance(obj, dict):
            return v
        o = obj.get(key)
        if isinstance(o, dict) and all(o.get(n) is not None for n in names):
            v = np.array([o[n] for n in names], dtype=F)
        return v

    # -- loaders ------------------------------------------------------------------------------
    def load(self, path: str) -> List[int]:
        """Scene::load (src/scene.rs:121-157)"""
        ext = os.path.splitext(path)[1].lower()
        if ext == ".json":
            return self.load_json(path)
        if ext == ".obj":
            return self.load_wavefront(path)
        if ext in (".gltf", ".glb"):
            from .gltf_loader import load_gltf
            return load_gltf(self, path)
        raise ValueError("can not load %s" % path)

    def load_json(self, path: str) -> List[int]:
        """Scene::load_json (src/scene.rs:159-641)"""
        loaded: List[int] = []
        with open(self._path(path), "r") as fh:
            data = json.load(fh)

        cfg = data.get("config")
        if cfg is not None:                                          # :179-198 (JSON beats CLI)
            c = self.config
            if cfg.get("monte_carlo") is not None: c.monte_carlo = bool(cfg["monte_carlo"])
            if cfg.get("samples") is not None: c.samples = int(cfg["samples"]) & 0xFFFF
            if cfg.get("focal_length") is not None: c.focal_length = float(F(cfg["focal_length"]))
            if cfg.get("aperture_size") is not None: c.aperture_size = float(F(cfg["aperture_size"]))
            if cfg.get("fog_density") is not None: c.fog_density = float(F(cfg["fog_density"]))
            if cfg.get("fog_color") is not None:
                fc = cfg["fog_color"]
                c.fog_color = (float(F(fc["r"])), float(F(fc["g"])), float(F(fc["b"])))
            if cfg.get("max_recursion") is not None: c.max_recursion = int(cfg["max_recursion"]) & 0xFFFF
            if cfg.get("gamma_correction") is not None: c.gamma_correction = bool(cfg["gamma_correction"])
        post = data.get("post")
        if post is not None:                                         # :200-205
            for k in ("cavity", "outline"):
                if post.get(k) is not None:
                    self.post[k] = bool(post[k])

        cam = data.get("camera")
        if cam is not None:                                          # :207-239
            self.cam.eye_pos = self._xyz(cam, "pos", (0, 0, 0))
            self.cam.up = self._xyz(cam, "up", (0, 1, 0))
            self.cam.dir = self._xyz(cam, "dir", (0, 0, -1))
            if isinstance(cam.get("fov"), (int, float)):
                self.cam.fov = F(math.radians(float(cam["fov"])))   # f64 to_radians then `as f32`
            if isinstance(cam.get("z_near"), (int, float)): self.cam.clipping_near = float(F(cam["z_near"]))
            if isinstance(cam.get("z_far"), (int, float)): self.cam.clipping_far = float(F(cam["z_far"]))

        for light in data.get("lights") or []:                       # :241-291
            max_angle = F(PI / F(2.0))
            if light.get("max_angle") is not None:
                max_angle = to_radians(F(light["max_angle"]))
            lt = {"point": LIGHT_POINT, "directional": LIGHT_DIRECTIONAL, "spot": LIGHT_SPOT}.get(
                light["light_type"], LIGHT_POINT)
            self.lights.append(Light(
                id=self.get_next_id(), name="light",
                pos=self._xyz(light, "pos", (0, 0, 0)), dir=self._xyz(light, "dir", (0, -1, 0)),
                color=self._xyz(light, "color", (0, 0, 0), ("r", "g", "b")),
                intensity=float(F(light["intensity"])), max_angle=float(max_angle), light_type=lt))

        for obj in data.get("objects") or []:                        # :293-560
            mat = Material(id=self.get_next_id(), name="material")
            item_type = obj["type"]
            name = obj["name"] if obj.get("name") is not None else "unknown"

            colors = obj.get("color")
            if colors is not None:                                   # :311-333
                mat.base_color = self._xyz(colors, "base", mat.base_color, ("r", "g", "b"))
                mat.specular_color = self._xyz(colors, "specular", mat.specular_color, ("r", "g", "b"))
                sp = colors.get("specular")
                if isinstance(sp, dict) and isinstance(sp.get("factor"), float):
                    mat.specular_color = (mat.base_color * F(sp["factor"])).astype(F)
                mat.ambient_color = self._xyz(colors, "ambient", mat.ambient_color, ("r", "g", "b"))
                am = colors.get("ambient")
                if isinstance(am, dict) and isinstance(am.get("factor"), float):
                    mat.ambient_color = (mat.base_color * F(am["factor"])).astype(F)

            for k in Material._SCALARS:                              # :336-349
                if obj.get(k) is not None: setattr(mat, k, float(F(obj[k])))
            for k in Material._BOOLS:
                if obj.get(k) is not None: setattr(mat, k, bool(obj[k]))

            tex = obj.get("texture")
            if tex is not None:                                      # :351-397 (no "reflectivity" key)
                for key, tt in (("base", TEX_BASE), ("ambient", TEX_AMBIENT), ("specular", TEX_SPECULAR),
                                ("normal", TEX_NORMAL), ("alpha", TEX_ALPHA), ("roughness", TEX_ROUGHNESS),
                                ("ambient_occlusion", TEX_AO)):
                    if isinstance(tex.get(key), str):
                        self.load_texture(mat, tex[key], tt)

            visible = bool(obj["visible"]) if obj.get("visible") is not None else True
            flip_normals = bool(obj["flip_normals"]) if obj.get("flip_normals") is not None else False

            rotation = np.zeros(3, dtype=F)
            scale = np.ones(3, dtype=F)
            translation = np.zeros(3, dtype=F)
            tr = obj.get("transformation")
            if tr is not None:                                       # :412-422
                scale = self._xyz(tr, "scale", scale)
                translation = self._xyz(tr, "translation", translation)
                rotation = self._xyz(tr, "rotation", rotation)
                rotation = np.array([to_radians(r) for r in rotation], dtype=F)

            shape: Optional[Item] = None
            if item_type == "sphere":                                # :427-443, sphere.rs:104-118
                pos = self._xyz(obj, "pos", (0, 0, 0))
                radius = float(F(obj["radius"])) if obj.get("radius") is not None else 0.0
                shape = Item(id=self.get_next_id(), name=name, shape=SHAPE_SPHERE, material=mat,
                             trans=mat_translation(*pos), radius=radius)
                loaded.append(shape.id)
            elif item_type == "plane":                               # :445-465, mesh.rs:186-202
                vs = obj["vertices"]
                verts = np.array([[v["x"], v["y"], v["z"]] for v in vs[:4]], dtype=F)
                idx = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
                mesh = MeshData(vertices=verts, indices=idx,
                                uvs=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=F), uv_indices=idx.copy(),
                                normals=np.zeros((0, 3), dtype=F), normals_indices=np.zeros((0, 3), dtype=np.uint32))
                shape = Item(id=self.get_next_id(), name=name, shape=SHAPE_MESH, material=mat,
                             trans=mat_identity(), mesh=mesh)
                loaded.append(shape.id)
            elif item_type in ("wavefront", "json", "gltf"):         # :467-530
                sub = obj["path"]
                if not os.path.exists(self._path(sub)):
                    raise FileNotFoundError(
                        "%s is not staged (the reference downloads it from %s)" % (sub, obj.get("url")))
                ids = self.load(sub) if item_type != "wavefront" else self.load_wavefront(sub)
                for it in self.items:
                    if it.id in ids:                                 # stale ids of nested spheres/planes never match
                        if obj.get("name") is not None:
                            it.name = name
                        it.material.apply_diff(mat)
                        it.visible = visible
                        it.flip_normals = flip_normals
                        it.apply_transformation(translation, scale, rotation)
                loaded.extend(ids)

            if shape is not None:                                    # :532-545
                shape.visible = visible
                shape.flip_normals = flip_normals
                shape.apply_transformation(translation, scale, rotation)
                shape.id = self.get_next_id()
                self.items.append(shape)
                self.materials.append(mat)

        anim = data.get("animation")
        if anim is not None:
            self.animation = anim
        return loaded

    def load_wavefront(self, path: str) -> List[int]:
        """Scene::load_wavefront (src/scene.rs:1126-1367) over a tobj-4-style parse
        (triangulate + single_index)."""
        from .obj_loader import load_obj
        loaded: List[int] = []
        models, mtls = load_obj(self._path(path))
        seen: Dict[int, Material] = {}
        for m in models:
            if len(m["positions"]) == 0:
                continue
            if m["material_id"] is not None:
                mid = m["material_id"]
                if mid in seen:
                    mat = seen[mid]
                else:
                    mat = Material(id=self.get_next_id(), name="")
                    wm = mtls[mid]
                    mat.name = wm["name"]
                    if wm.get("Ns") is not None: mat.shininess = float(F(wm["Ns"]))
                    if wm.get("Ka") is not None: mat.ambient_color = np.array(wm["Ka"], dtype=F)
                    if wm.get("Ks") is not None: mat.specular_color = np.array(wm["Ks"], dtype=F)
                    if wm.get("Kd") is not None: mat.base_color = np.array(wm["Kd"], dtype=F)
                    if wm.get("Ni") is not None: mat.refraction_index = float(F(wm["Ni"]))
                    if wm.get("d") is not None: mat.alpha = float(F(wm["d"]))
                    mat.ambient_color = (mat.base_color * F(0.01)).astype(F)      # :1284
                    if wm.get("illum") is not None and wm["illum"] > 2:
                        mat.reflectivity = 0.5
                    obj_dir = os.path.dirname(path)
                    for key, tt in (("map_Kd", TEX_BASE), ("map_Bump", TEX_NORMAL), ("map_Ka", TEX_AMBIENT),
                                    ("map_Ks", TEX_SPECULAR), ("map_d", TEX_ALPHA)):
                        if wm.get(key):
                            tp = wm[key]
                            if not os.path.isabs(tp):
                                tp = os.path.join(obj_dir, tp)                   # get_texture_path :1643-1658
                            self.load_texture(mat, tp, tt)
                    self.materials.append(mat)
                    seen[mid] = mat
            else:
                mat = Material(id=self.get_next_id(), name="")
            verts = np.asarray(m["positions"], dtype=F).reshape(-1, 3)
            idx = np.asarray(m["indices"], dtype=np.uint32).reshape(-1, 3)
            uvs = np.asarray(m["texcoords"], dtype=F).reshape(-1, 2)
            nrm = np.asarray(m["normals"], dtype=F).reshape(-1, 3)
            uv_idx = idx.copy() if len(uvs) > 0 else np.zeros((0, 3), dtype=np.uint32)   # :1346-1355
            n_idx = idx.copy() if len(nrm) > 0 else np.zeros((0, 3), dtype=np.uint32)
            mesh = MeshData(verts, idx, uvs, uv_idx, nrm, n_idx)
            item = Item(id=self.get_next_id(), name=m["name"], shape=SHAPE_MESH, material=mat,
                        trans=mat_identity(), mesh=mesh)
            loaded.append(item.id)
            self.items.append(item)
        return loaded

    # -- environment defaults -----------------------------------------------------------------
    def _bbox_points(self) -> List[np.ndarray]:
        pts = []
        for it in self.items:
            lo, hi = item_local_aabb(it)
            # parry Aabb::vertices order
            for c in ((lo[0], lo[1], lo[2]), (hi[0], lo[1], lo[2]), (hi[0], hi[1], lo[2]), (lo[0], hi[1], lo[2]),
                      (lo[0], lo[1], hi[2]), (hi[0], lo[1], hi[2]), (hi[0], hi[1], hi[2]), (lo[0], hi[1], hi[2])):
                pts.append(mat_vec(it.trans, [c[0], c[1], c[2], 1.0]))
        return pts

    def find_optimal_camera_pos(self) -> None:
        """Scene::find_optimal_camera_pos (src/scene.rs:1426-1547)"""
        pts = self._bbox_points()
        if not pts:
            return
        arr = np.array(pts, dtype=F)[:, :3]
        mn, mx = arr.min(axis=0), arr.max(axis=0)
        delta = np.abs(mx - mn).astype(F)
        center = (mn + delta / F(2.0)).astype(F)
        ob = np.array([-0.5, 0.5, 1.0], dtype=F)
        direction = (ob / F(np.sqrt(F(np.dot(ob, ob))))).astype(F)
        factor, inc = F(0.0), F(0.01)

        def all_in():
            pv = mat_mul(self.cam.projection, self.cam.view).astype(np.float64)
            hom = np.concatenate([arr.astype(np.float64), np.ones((len(arr), 1))], axis=1)
            c = (hom @ pv.T).astype(F)
            return bool(np.all((np.abs(c[:, 0]) <= c[:, 3]) & (np.abs(c[:, 1]) <= c[:, 3]) & (np.abs(c[:, 2]) <= c[:, 3])))

        self.cam.eye_pos = center.copy()
        while factor < F(1000.0):
            self.cam.eye_pos = (center + direction * factor).astype(F)
            self.cam.dir = (-direction).astype(F)
            self.cam.init_matrices()
            if all_in():
                self.cam.eye_pos = (self.cam.eye_pos + direction * F(1.001)).astype(F)
                break
            factor = F(factor + inc)
        fov = F(0.0)
        while fov < F(90.0):
            self.cam.fov = to_radians(fov)
            self.cam.init_matrices()
            if all_in():
                self.cam.fov = F(self.cam.fov * F(1.1))
                break
            fov = F(fov + inc)
        self.cam.init_matrices()

    def add_default_light(self) -> None:
        """Scene::add_default_light (src/scene.rs:1386-1401)"""
        self.lights.append(Light(id=self.get_next_id(), name="default",
                                 pos=np.array([-2.0, 10.0, 5.0], dtype=F), dir=np.array([0, -1, 0], dtype=F),
                                 color=np.array([1, 1, 1], dtype=F), intensity=200.0,
                                 max_angle=float(F(PI / F(2.0))), light_type=LIGHT_POINT))

    def find_and_set_default_env_if_needed(self) -> None:
        """src/scene.rs:1549-1562"""
        if self.cam.is_default_cam():
            self.find_optimal_camera_pos()
        if len(self.lights) == 0:
            self.add_default_light()


def item_local_aabb(it: Item) -> Tuple[np.ndarray, np.ndarray]:
    """calc_bbox: Ball::aabb / TriMesh::aabb with identity (sphere.rs:39-43, mesh.rs:45-49)"""
    if it.shape == SHAPE_SPHERE:
        r = F(it.radius)
        return np.array([-r, -r, -r], dtype=F), np.array([r, r, r], dtype=F)
    v = it.mesh.vertices
    return v.min(axis=0).astype(F), v.max(axis=0).astype(F)


def load_scene(paths, width: int, height: int, asset_root: str = ".", samples: Optional[int] = None,
               monte_carlo: Optional[bool] = None) -> Scene:
    """main.rs:79-83 + Run::init_scene (src/run.rs:196-245): CLI samples / monte_carlo are written
    into the scene config BEFORE loading, the scene files' "config" blocks then override them;
    all files are loaded into ONE scene in CLI order with ids counting on."""
    sc = Scene(asset_root)
    if monte_carlo is not None: sc.config.monte_carlo = bool(monte_carlo)
    if samples is not None: sc.config.samples = int(samples)
    if isinstance(paths, str):
        paths = [paths]
    for p in paths:
        sc.load(p)
    sc.cam.init(width, height)
    sc.find_and_set_default_env_if_needed()
    return sc
