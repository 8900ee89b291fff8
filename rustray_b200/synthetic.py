"""Synthetic scenes: BASELINE.json config 5 (triangle soup + spheres) and small
feature-coverage scenes for the parity tests.  Everything is seeded; nothing here reads the
reference tree."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from .scene_loader import (Camera, Config, Item, Light, Material, MeshData, Scene, LIGHT_DIRECTIONAL, LIGHT_POINT,
                           LIGHT_SPOT, SHAPE_MESH, SHAPE_SPHERE, mat_identity, mat_translation, to_radians)

F = np.float32


def _mesh(verts, idx, uvs=None, normals=None) -> MeshData:
    verts = np.ascontiguousarray(verts, dtype=F).reshape(-1, 3)
    idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
    z3 = np.zeros((0, 3), dtype=np.uint32)
    return MeshData(verts, idx,
                    np.ascontiguousarray(uvs, dtype=F).reshape(-1, 2) if uvs is not None else np.zeros((0, 2), dtype=F),
                    idx.copy() if uvs is not None else z3,
                    np.ascontiguousarray(normals, dtype=F).reshape(-1, 3) if normals is not None else np.zeros((0, 3), dtype=F),
                    idx.copy() if normals is not None else z3)


def soup_scene(n_triangles: int = 10_000_000, n_spheres: int = 1000, cells: int = 4, seed: int = 0x5EED,
               width: int = 3840, height: int = 2160, extent: float = 50.0) -> Scene:
    """Config 5: `n_triangles` random triangles in [-extent, extent]^3 (edge length U(0.05, 0.5), random
    orientation) split into cells^3 mesh items by spatial cell, `n_spheres` spheres r ~ U(0.2, 1.0)
    (20 % reflectivity 0.5, 10 % alpha 0.5 / ior 1.5), 2 point lights + 1 directional, camera outside
    looking at the centre, fov 60."""
    rng = np.random.default_rng(seed)
    sc = Scene(".")
    c = rng.uniform(-extent, extent, size=(n_triangles, 3)).astype(F)
    # random orientation: two random unit vectors scaled by the edge length
    def unit(n):
        v = rng.normal(size=(n, 3)).astype(F)
        return v / np.linalg.norm(v, axis=1, keepdims=True).astype(F)
    e1 = unit(n_triangles) * rng.uniform(0.05, 0.5, size=(n_triangles, 1)).astype(F)
    e2 = unit(n_triangles) * rng.uniform(0.05, 0.5, size=(n_triangles, 1)).astype(F)
    tri = np.stack([c, c + e1, c + e2], axis=1).astype(F)                  # (n, 3, 3)
    cell = np.clip(((c + extent) / (2 * extent) * cells).astype(np.int64), 0, cells - 1)
    cell_id = (cell[:, 0] * cells + cell[:, 1]) * cells + cell[:, 2]
    order = np.argsort(cell_id, kind="stable")
    tri, cell_id = tri[order], cell_id[order]
    bounds = np.searchsorted(cell_id, np.arange(cells ** 3 + 1))
    palette = rng.uniform(0.2, 1.0, size=(cells ** 3, 3)).astype(F)
    for k in range(cells ** 3):
        a, b = int(bounds[k]), int(bounds[k + 1])
        if b <= a:
            continue
        m = Material(id=sc.get_next_id(), name="soup%d" % k)
        m.base_color = palette[k]
        m.specular_color = (palette[k] * F(0.8)).astype(F)
        verts = tri[a:b].reshape(-1, 3)
        idx = np.arange(3 * (b - a), dtype=np.uint32).reshape(-1, 3)
        sc.items.append(Item(id=sc.get_next_id(), name="soup%d" % k, shape=SHAPE_MESH, material=m, trans=mat_identity(),
                             mesh=_mesh(verts, idx)))
        sc.materials.append(m)
    sp = rng.uniform(-extent, extent, size=(n_spheres, 3)).astype(F)
    sr = rng.uniform(0.2, 1.0, size=n_spheres).astype(F)
    kind = rng.uniform(size=n_spheres)
    for i in range(n_spheres):
        m = Material(id=sc.get_next_id(), name="sphere")
        m.base_color = rng.uniform(0.2, 1.0, size=3).astype(F)
        if kind[i] < 0.2:
            m.reflectivity = 0.5
        elif kind[i] < 0.3:
            m.alpha, m.refraction_index = 0.5, 1.5
        sc.items.append(Item(id=sc.get_next_id(), name="sphere%d" % i, shape=SHAPE_SPHERE, material=m,
                             trans=mat_translation(*sp[i]), radius=float(sr[i])))
        sc.materials.append(m)
    e = extent
    sc.lights.append(Light(sc.get_next_id(), "p1", np.array([e * 1.5, e * 1.5, e * 1.5], dtype=F), np.array([0, -1, 0], dtype=F),
                           np.array([1, 1, 1], dtype=F), 3000.0, math.pi / 2, LIGHT_POINT))
    sc.lights.append(Light(sc.get_next_id(), "p2", np.array([-e * 1.5, e, e * 2.0], dtype=F), np.array([0, -1, 0], dtype=F),
                           np.array([1, 0.9, 0.8], dtype=F), 3000.0, math.pi / 2, LIGHT_POINT))
    sc.lights.append(Light(sc.get_next_id(), "d", np.zeros(3, dtype=F), np.array([0.3, -1.0, -0.5], dtype=F),
                           np.array([1, 1, 1], dtype=F), 0.5, math.pi / 2, LIGHT_DIRECTIONAL))
    sc.cam = Camera()
    sc.cam.eye_pos = np.array([0.0, e * 0.6, e * 3.2], dtype=F)
    d = -sc.cam.eye_pos
    sc.cam.dir = (d / np.linalg.norm(d)).astype(F)
    sc.cam.fov = to_radians(60.0)
    sc.cam.clipping_near, sc.cam.clipping_far = 0.1, 1000.0
    sc.cam.init(width, height)
    sc.config = Config(monte_carlo=True, samples=256)
    return sc


def checker_texture(n: int = 64, squares: int = 8, seed: int = 3, alpha_holes: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:n]
    t = np.zeros((n, n, 4), dtype=np.uint8)
    on = ((x * squares // n + y * squares // n) % 2).astype(bool)
    t[..., :3] = np.where(on[..., None], rng.integers(120, 255, size=3), rng.integers(0, 100, size=3))
    t[..., :3] = np.clip(t[..., :3].astype(np.int32) + rng.integers(-20, 20, size=(n, n, 3)), 0, 255).astype(np.uint8)
    t[..., 3] = 255
    if alpha_holes:
        t[..., 3] = np.where(on, 255, 40)
    return t


def noise_texture(n: int = 32, seed: int = 5, channels_equal: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 255, size=(n, n, 4)).astype(np.uint8)
    if channels_equal:
        t[..., 1] = t[..., 0]; t[..., 2] = t[..., 0]
    t[..., 3] = 255
    return t


def normal_texture(n: int = 32, seed: int = 7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, n, 3)) * 0.25 + np.array([0, 0, 1.0])
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    t = np.zeros((n, n, 4), dtype=np.uint8)
    t[..., :3] = np.clip((v * 0.5 + 0.5) * 255, 0, 255).astype(np.uint8)
    t[..., 3] = 255
    return t


def icosphere_mesh(subdiv: int = 2, radius: float = 1.0, smooth: bool = True, with_uv: bool = True) -> MeshData:
    t = (1.0 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}
        nf = []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    vv = np.array(v, dtype=np.float64)
    uv = np.stack([np.arctan2(vv[:, 2], vv[:, 0]) / (2 * np.pi) + 0.5, np.arccos(np.clip(vv[:, 1], -1, 1)) / np.pi], axis=1)
    return _mesh((vv * radius).astype(F), np.array(f, dtype=np.uint32), uvs=uv.astype(F) if with_uv else None,
                 normals=vv.astype(F) if smooth else None)


def quad_mesh(p0, p1, p2, p3) -> MeshData:
    return _mesh([p0, p1, p2, p3], [[0, 1, 2], [0, 2, 3]], uvs=[[0, 0], [1, 0], [1, 1], [0, 1]])


def feature_scene(width: int = 192, height: int = 128, n_extra_spheres: int = 0, seed: int = 11, spot: bool = True,
                  nearest: bool = False, fog: float = 0.0) -> Scene:
    """Small scene touching every shading feature: textured / normal-mapped / alpha-mapped / AO / roughness /
    reflectivity-mapped materials, a glass sphere, a mirror, a flat-shaded and a smooth-shaded mesh with a
    non-uniform transform, flip_normals, an invisible item, a reflection-only environment sphere, cast/receive
    shadow switches, directional + point + spot lights.  `n_extra_spheres` > 0 adds small spheres so the item
    count crosses the TLAS (16) and the reference's BVH_MIN_ITEMS (50) thresholds."""
    rng = np.random.default_rng(seed)
    sc = Scene(".")
    tex = {"checker": checker_texture(64, 8, 3), "alpha": checker_texture(32, 4, 9, alpha_holes=True),
           "noise": noise_texture(32, 5), "grey": noise_texture(16, 6, channels_equal=True), "normal": normal_texture(32, 7)}
    sc.texture_data.update(tex)

    def mat(**kw) -> Material:
        m = Material(id=sc.get_next_id(), name="m")
        for k, v in kw.items():
            if k == "tex":
                for tt, name in v.items():
                    m.textures[tt] = name
            elif k.endswith("_color"):
                setattr(m, k, np.array(v, dtype=F))
            else:
                setattr(m, k, v)
        sc.materials.append(m)
        return m

    def add(name, shape, material, trans=None, **kw) -> Item:
        it = Item(id=sc.get_next_id(), name=name, shape=shape, material=material, trans=mat_identity() if trans is None else trans, **kw)
        sc.items.append(it)
        return it

    floor = add("floor", SHAPE_MESH, mat(base_color=(0.8, 0.8, 0.9), reflectivity=0.3, tex={0: "checker", 6: "grey"},
                                          texture_filtering_nearest=nearest),
                mesh=quad_mesh((-12, -3, 4), (12, -3, 4), (12, -3, -30), (-12, -3, -30)))
    wall = add("wall", SHAPE_MESH, mat(base_color=(0.9, 0.7, 0.6), tex={0: "noise", 3: "normal"}, normal_map_strength=2.0,
                                        texture_filtering_nearest=nearest),
               mesh=quad_mesh((-12, -3, -30), (12, -3, -30), (12, 12, -30), (-12, 12, -30)))
    add("glass", SHAPE_SPHERE, mat(base_color=(0.9, 0.95, 1.0), alpha=0.15, refraction_index=1.45, reflectivity=0.2),
        trans=mat_translation(-3.0, -1.0, -9.0), radius=2.0)
    add("mirror", SHAPE_SPHERE, mat(base_color=(1, 1, 1), reflectivity=0.9, tex={7: "grey"}), trans=mat_translation(3.5, -1.2, -11.0), radius=1.8)
    add("rough", SHAPE_SPHERE, mat(base_color=(0.9, 0.5, 0.2), roughness=0.1, tex={0: "checker", 2: "noise", 5: "grey"}, reflectivity=0.25),
        trans=mat_translation(0.5, -1.9, -6.0), radius=1.1)
    ico = add("ico_smooth", SHAPE_MESH, mat(base_color=(0.3, 0.9, 0.4), alpha=0.6, refraction_index=1.3, reflectivity=0.3, tex={1: "noise"},
                                             ambient_color=(0.05, 0.05, 0.05)),
              mesh=icosphere_mesh(2, 1.0, smooth=True))
    ico.apply_transformation((-6.5, 0.5, -14.0), (1.5, 2.2, 1.5), (to_radians(20.0), to_radians(35.0), to_radians(-10.0)))
    ico2 = add("ico_flat", SHAPE_MESH, mat(base_color=(0.9, 0.9, 0.2), smooth_shading=False, shininess=40.0),
               mesh=icosphere_mesh(1, 1.3, smooth=True))
    ico2.apply_transformation((6.5, 2.0, -16.0), (1.0, 1.0, 1.0), (0.0, to_radians(15.0), 0.0))
    leaf = add("alpha_card", SHAPE_MESH, mat(base_color=(0.2, 0.8, 0.9), tex={0: "checker", 4: "alpha"}, backface_cullig=False),
               mesh=quad_mesh((-1.5, -1, 0), (1.5, -1, 0), (1.5, 2, 0), (-1.5, 2, 0)))
    leaf.apply_transformation((1.0, 2.5, -8.0), (1.0, 1.0, 1.0), (to_radians(60.0), 0.0, 0.0))
    flipped = add("flipped", SHAPE_MESH, mat(base_color=(0.8, 0.3, 0.8), receive_shadow=False),
                  mesh=quad_mesh((-11.9, -3, 4), (-11.9, -3, -30), (-11.9, 12, -30), (-11.9, 12, 4)), flip_normals=True)
    add("ghost", SHAPE_SPHERE, mat(base_color=(1, 0, 0)), trans=mat_translation(0.0, 3.0, -5.0), radius=1.0, visible=False)
    add("no_shadow", SHAPE_SPHERE, mat(base_color=(0.2, 0.2, 1.0), cast_shadow=False), trans=mat_translation(-1.0, 4.0, -10.0), radius=0.8)
    add("zero_alpha", SHAPE_SPHERE, mat(base_color=(1, 1, 0), alpha=0.0), trans=mat_translation(2.0, 0.0, -4.0), radius=0.7)
    add("env", SHAPE_SPHERE, mat(base_color=(0, 0, 0), ambient_color=(1, 1, 1), tex={1: "checker"}, reflection_only=True, backface_cullig=False),
        radius=90.0)
    for i in range(n_extra_spheres):
        p = rng.uniform([-10, -2.5, -28], [10, 8, -4]).astype(F)
        m = mat(base_color=rng.uniform(0.2, 1, size=3), reflectivity=float(rng.choice([0.0, 0.0, 0.4])),
                alpha=float(rng.choice([1.0, 1.0, 0.5])), refraction_index=1.4)
        add("extra%d" % i, SHAPE_SPHERE, m, trans=mat_translation(*p), radius=float(rng.uniform(0.15, 0.5)))

    def light(**kw):
        sc.lights.append(Light(id=sc.get_next_id(), name="l", pos=np.array(kw.get("pos", (0, 0, 0)), dtype=F),
                               dir=np.array(kw.get("dir", (0, -1, 0)), dtype=F), color=np.array(kw.get("color", (1, 1, 1)), dtype=F),
                               intensity=kw.get("intensity", 1.0), max_angle=kw.get("max_angle", math.pi / 2), light_type=kw["t"]))
    light(t=LIGHT_DIRECTIONAL, dir=(0.4, -1.0, -0.6), intensity=0.6)
    light(t=LIGHT_POINT, pos=(-5.0, 8.0, -4.0), intensity=300.0, color=(1.0, 0.95, 0.9))
    if spot:
        light(t=LIGHT_SPOT, pos=(4.0, 9.0, -8.0), dir=(-0.2, -1.0, -0.3), intensity=500.0, max_angle=float(to_radians(30.0)), color=(0.6, 0.8, 1.0))
    sc.lights.append(Light(id=sc.get_next_id(), name="off", pos=np.array([0, 5, 0], dtype=F), dir=np.array([0, -1, 0], dtype=F),
                           color=np.array([1, 1, 1], dtype=F), intensity=999.0, max_angle=1.0, light_type=LIGHT_POINT, enabled=False))
    sc.cam = Camera()
    sc.cam.eye_pos = np.array([0.5, 1.5, 3.0], dtype=F)
    d = np.array([-0.05, -0.12, -1.0]); sc.cam.dir = (d / np.linalg.norm(d)).astype(F)
    sc.cam.fov = to_radians(70.0)
    sc.cam.clipping_near, sc.cam.clipping_far = 0.1, 500.0
    sc.cam.init(width, height)
    sc.config = Config(fog_density=fog, fog_color=(0.5, 0.55, 0.6))
    return sc


# ---------------------------------------------------------------------------------------------------------
# Stand-ins for BASELINE.json configs[2] / configs[3].  scene/helmet.json and scene/sponza.json download
# DamagedHelmet.glb / Sponza_fixed.glb at load time (reference src/scene.rs:473-493); neither file is in the reference
# tree and there is no network, so these generators build scenes of the same SHAPE — item / triangle / texture counts,
# material mapping of Scene::load_gltf (src/scene.rs:895-960: base, normal, metallic -> Reflectivity, roughness,
# emissive -> AmbientEmissive, occlusion; specular = base * 0.8; roughness = factor / 2 pi; one de-indexed Mesh item per
# primitive with uv.y := 1 - v), the environment sphere of scene/environment.json, the camera / light / config blocks of
# the JSON files.  They are labelled stand-ins everywhere they are reported.
# ---------------------------------------------------------------------------------------------------------
def _surface(fn, nu: int, nv: int, uv_rep=(1.0, 1.0), flip: bool = False):
    """Tessellate P(u, v), u, v in [0, 1], into 2*nu*nv de-indexed triangles with smooth normals and uvs.
    -> (verts (T*3, 3), normals (T*3, 3), uvs (T*3, 2)) as glTF + easy-gltf would deliver them."""
    u, v = np.meshgrid(np.linspace(0.0, 1.0, nu + 1), np.linspace(0.0, 1.0, nv + 1), indexing="ij")
    p = fn(u, v)                                                        # (nu+1, nv+1, 3)
    e = 1e-3
    n = np.cross(fn(np.clip(u + e, 0, 1), v) - fn(np.clip(u - e, 0, 1), v), fn(u, np.clip(v + e, 0, 1)) - fn(u, np.clip(v - e, 0, 1)))
    ln = np.linalg.norm(n, axis=-1, keepdims=True)
    n = np.where(ln > 1e-12, n / np.maximum(ln, 1e-12), np.array([0.0, 1.0, 0.0]))
    if flip:
        n = -n
    t = np.stack([u * uv_rep[0], 1.0 - v * uv_rep[1]], axis=-1)         # the loader's uv.y := 1 - v
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    i, j = i.reshape(-1), j.reshape(-1)
    quad = [(i, j), (i + 1, j), (i + 1, j + 1), (i, j + 1)]
    order = (0, 2, 1, 0, 3, 2) if flip else (0, 1, 2, 0, 2, 3)
    def take(a):
        return np.stack([a[quad[k][0], quad[k][1]] for k in order], axis=1).reshape(-1, a.shape[-1])
    return take(p).astype(F), take(n).astype(F), take(t).astype(F)


def _deindexed_mesh(verts, normals, uvs) -> MeshData:
    n = verts.shape[0]
    idx = np.arange(n, dtype=np.uint32).reshape(-1, 3)
    return MeshData(np.ascontiguousarray(verts, dtype=F), idx, np.ascontiguousarray(uvs, dtype=F), idx.copy(),
                    np.ascontiguousarray(normals, dtype=F), idx.copy())


def _fbm(n: int, rng, octaves=(4, 16, 64), weights=(0.6, 0.3, 0.1)) -> np.ndarray:
    """cheap band-limited noise in [0, 1]: random grids upsampled by pixel repetition + a box blur"""
    out = np.zeros((n, n), dtype=np.float32)
    for o, w in zip(octaves, weights):
        o = min(o, n)
        g = rng.random((o, o), dtype=np.float32)
        out += w * np.kron(g, np.ones((n // o, n // o), dtype=np.float32))
    k = max(1, n // 128)
    if k > 1:
        c = np.cumsum(np.pad(out, ((k, 0), (0, 0)), mode="wrap"), axis=0); out = (c[k:] - c[:-k]) / k
        c = np.cumsum(np.pad(out, ((0, 0), (k, 0)), mode="wrap"), axis=1); out = (c[:, k:] - c[:, :-k]) / k
    return out


def pbr_texture_set(n: int, rng, tint, metal: float = 0.0, rough=(0.3, 0.9), bricks: int = 8, cutout: bool = False):
    """One glTF material's images, decoded as Scene::get_dyn_image_from_gltf_material (src/scene.rs:980-1124) hands them
    to the shading code: base RGBA, normal RGB(a=255), metallic (B of the metallic-roughness image) and roughness (G) as grey
    RGBA, occlusion (R * strength) as grey RGBA, emissive RGB(a=255)."""
    h = _fbm(n, rng)
    y, x = np.mgrid[0:n, 0:n]
    if bricks:
        bw = max(2, n // bricks)
        hb = max(1, bw // 2)
        row = y // hb
        mortar = ((y % hb) < max(1, bw // 24)) | (((x + (row % 2) * hb) % bw) < max(1, bw // 24))
        h = np.where(mortar, h * 0.4, h)
    base = np.zeros((n, n, 4), dtype=np.uint8)
    base[..., :3] = np.clip((0.35 + 0.65 * h)[..., None] * np.asarray(tint, dtype=np.float32) * 255.0, 0, 255).astype(np.uint8)
    base[..., 3] = 255
    if cutout:                                                            # leaves / chains: alpha-masked base colour
        base[..., 3] = np.where(_fbm(n, rng, octaves=(8, 32), weights=(0.7, 0.3)) > 0.5, 255, 0)
    gx = np.roll(h, -1, axis=1) - np.roll(h, 1, axis=1)
    gy = np.roll(h, -1, axis=0) - np.roll(h, 1, axis=0)
    nv = np.stack([-gx * 4.0, -gy * 4.0, np.ones_like(h)], axis=-1)
    nv /= np.linalg.norm(nv, axis=-1, keepdims=True)
    normal = np.zeros((n, n, 4), dtype=np.uint8)
    normal[..., :3] = np.clip((nv * 0.5 + 0.5) * 255.0, 0, 255).astype(np.uint8)
    normal[..., 3] = 255
    m = np.clip(metal * (0.5 + _fbm(n, rng, octaves=(8, 32), weights=(0.7, 0.3))) * 255.0, 0, 255).astype(np.uint8)
    r = np.clip((rough[0] + (rough[1] - rough[0]) * (1.0 - h)) * 255.0, 0, 255).astype(np.uint8)
    grey = lambda c: np.ascontiguousarray(np.repeat(c[..., None], 4, axis=-1))
    ao = np.clip((0.55 + 0.45 * h) * 255.0, 0, 255).astype(np.uint8)
    emis = np.zeros((n, n, 4), dtype=np.uint8)
    glow = _fbm(n, rng, octaves=(8,), weights=(1.0,)) > 0.8
    emis[..., 0] = np.where(glow, 40, 0); emis[..., 1] = np.where(glow, 90, 0); emis[..., 2] = np.where(glow, 160, 0); emis[..., 3] = 255
    return {"base": base, "normal": normal, "metallic": grey(m), "roughness": grey(r), "occlusion": grey(ao), "emissive": emis}


def _gltf_material(sc: Scene, name: str, tex: dict, kinds, nearest: bool, metallic_factor: float = 1.0, roughness_factor: float = 1.0,
                   emissive_factor=None) -> Material:
    """Material as Scene::load_gltf fills it (src/scene.rs:895-960)."""
    m = Material(id=sc.get_next_id(), name=name)
    m.base_color = np.array([1, 1, 1], dtype=F)
    m.specular_color = (m.base_color * F(0.8)).astype(F)
    m.alpha = 1.0
    m.reflectivity = float(F(metallic_factor) * F(0.5))
    m.roughness = float(F(F(F(1.0) / F(math.pi)) / F(2.0)) * F(roughness_factor))
    slot = {"base": 0, "emissive": 1, "normal": 3, "roughness": 5, "occlusion": 6, "metallic": 7}
    for k in kinds:
        key = "%s/%s" % (name, k)
        sc.texture_data[key] = tex[k]
        m.textures[slot[k]] = key
    if "emissive" in kinds:
        m.ambient_color = np.array(emissive_factor if emissive_factor is not None else (1, 1, 1), dtype=F)
    m.texture_filtering_nearest = nearest
    sc.materials.append(m)
    return m


def _environment(sc: Scene, rng, w: int = 1024, h: int = 512) -> None:
    """scene/environment.json: r = 100 reflection-only sphere, black base, white ambient with a 1024x512 ambient texture."""
    y, x = np.mgrid[0:h, 0:w]
    sky = np.zeros((h, w, 4), dtype=np.uint8)
    t = (y / (h - 1.0))[..., None]
    sky[..., :3] = np.clip(((1 - t) * np.array([90, 140, 230]) + t * np.array([200, 190, 170])) + 25 * _fbm(1024, rng)[:h, :w, None], 0, 255).astype(np.uint8)
    sky[..., 3] = 255
    sc.texture_data["environment/ambient"] = sky
    m = Material(id=sc.get_next_id(), name="environment")
    m.base_color = np.array([0, 0, 0], dtype=F); m.ambient_color = np.array([1, 1, 1], dtype=F)
    m.textures[1] = "environment/ambient"
    m.reflection_only = True; m.backface_cullig = False
    sc.materials.append(m)
    sc.items.append(Item(id=sc.get_next_id(), name="environment", shape=SHAPE_SPHERE, material=m, trans=mat_identity(), radius=100.0))


def atrium_scene(width: int = 1280, height: int = 720, detail: float = 1.0, tex_size: int = 1024, seed: int = 0xC4,
                 samples: int = 128, monte_carlo: bool = True) -> Scene:
    """STAND-IN for BASELINE.json configs[3] (scene/sponza.json, 1280x720, samples=128, monte_carlo=1): a two-storey colonnaded
    atrium of >= 100 mesh items / ~262 k de-indexed triangles at detail = 1 (25 PBR materials with base + normal + metallic +
    roughness maps, nearest filtering as the JSON asks, `backface_cullig` left at its default because the JSON spells the key
    `backface_culling`, src/scene.rs:349), the environment sphere of scene/environment.json, no light in the file -> the default
    point light (src/scene.rs:1386-1401), an interior camera as a GLB camera node would give.  More than 50 items, so the
    reference's item-BVH path (BVH_MIN_ITEMS, src/raytracing.rs:23,434) is the one this shape exercises."""
    rng = np.random.default_rng(seed)
    sc = Scene(".")
    _environment(sc, rng)
    kinds = ("base", "normal", "metallic", "roughness")
    q = 1.1 * max(0.001, float(detail)) ** 0.5                            # detail = 1: ~262 k triangles

    def seg(n):
        return max(2, int(round(n * q)))
    palette = [(0.75, 0.7, 0.62), (0.6, 0.55, 0.5), (0.8, 0.78, 0.7), (0.7, 0.3, 0.25), (0.25, 0.4, 0.7), (0.3, 0.6, 0.3),
               (0.85, 0.8, 0.6), (0.5, 0.5, 0.55)]
    mats = []
    for k in range(25):
        metal = (0.0, 0.0, 0.25, 0.0, 0.6)[k % 5]                          # most of Sponza is dielectric; vases, chains, lion heads are not
        ts = tex_size if k < 8 else max(64, tex_size // 2)
        tex = pbr_texture_set(ts, rng, palette[k % len(palette)], metal=metal, bricks=(8, 4, 0, 16)[k % 4], cutout=(k in (11, 17)))
        mats.append(_gltf_material(sc, "atrium_mat%02d" % k, tex, kinds, nearest=True))

    def add(name, mat_index, verts, normals, uvs):
        sc.items.append(Item(id=sc.get_next_id(), name=name, shape=SHAPE_MESH, material=mats[mat_index % 25], trans=mat_identity(),
                             mesh=_deindexed_mesh(verts, normals, uvs)))
    LX, LZ, H1, H2 = 18.0, 8.0, 5.0, 10.0                                 # half length, half width, storey heights
    # floor (gently uneven flagstones), upper gallery floors, outer walls, end walls, roof rim
    add("floor", 0, *_surface(lambda u, v: np.stack([(u * 2 - 1) * LX, 0.03 * np.sin(u * 90) * np.sin(v * 50), (v * 2 - 1) * LZ], -1), seg(96), seg(48), (12, 6)))
    for s, zc in ((-1, -LZ + 1.5), (1, LZ - 1.5)):
        add("gallery%+d" % s, 1, *_surface(lambda u, v, zc=zc: np.stack([(u * 2 - 1) * LX, H1 + 0 * u, zc + (v - 0.5) * 3.0], -1), seg(64), seg(8), (12, 1)))
        add("gallery_under%+d" % s, 1, *_surface(lambda u, v, zc=zc: np.stack([(u * 2 - 1) * LX, H1 - 0.3 + 0 * u, zc + (v - 0.5) * 3.0], -1), seg(64), seg(8), (12, 1), flip=True))
        add("wall%+d" % s, 2, *_surface(lambda u, v, s=s: np.stack([(u * 2 - 1) * LX, v * (H2 + 2), s * LZ + 0.05 * np.sin(u * 40) * np.sin(v * 30)], -1), seg(96), seg(40), (10, 4), flip=(s > 0)))
    for s in (-1, 1):
        add("endwall%+d" % s, 3, *_surface(lambda u, v, s=s: np.stack([s * LX + 0 * u, v * (H2 + 2), (u * 2 - 1) * LZ], -1), seg(48), seg(40), (5, 4), flip=(s < 0)))
    # colonnades: 2 storeys x 2 sides x 9 columns, each a fluted shaft + a capital, arches between neighbours
    n_col = 9
    xs = np.linspace(-LX + 2.0, LX - 2.0, n_col)
    for storey, (y0, hh) in enumerate(((0.0, H1 - 0.3), (H1, H2 - H1))):
        for s in (-1, 1):
            zc = s * (LZ - 3.0)
            for ci, xc in enumerate(xs):
                def shaft(u, v, xc=xc, zc=zc, y0=y0, hh=hh):
                    a = u * 2 * np.pi
                    r = 0.38 * (1.0 + 0.05 * np.cos(a * 12)) * (1.0 - 0.12 * v)
                    return np.stack([xc + r * np.cos(a), y0 + v * hh * 0.86, zc + r * np.sin(a)], -1)
                add("column_s%d_%+d_%d" % (storey, s, ci), 4 + (ci % 2), *_surface(shaft, seg(40), seg(22), (3, 4)))
                def capital(u, v, xc=xc, zc=zc, y0=y0, hh=hh):
                    a = u * 2 * np.pi
                    r = 0.34 + 0.3 * v ** 2 + 0.03 * np.cos(a * 8) * v
                    return np.stack([xc + r * np.cos(a), y0 + hh * (0.86 + 0.14 * v), zc + r * np.sin(a)], -1)
                add("capital_s%d_%+d_%d" % (storey, s, ci), 6, *_surface(capital, seg(32), seg(8), (2, 1)))
            for ci in range(n_col - 1):
                xa, xb = xs[ci], xs[ci + 1]
                def arch(u, v, xa=xa, xb=xb, zc=zc, y0=y0, hh=hh):
                    t = u * np.pi
                    xm, rad = 0.5 * (xa + xb), 0.5 * (xb - xa) - 0.3
                    return np.stack([xm - rad * np.cos(t), y0 + hh * 0.86 + 0.55 * rad * np.sin(t) * 0.5, zc + (v - 0.5) * 0.7], -1)
                if storey == 0 or ci % 2 == 0:
                    add("arch_s%d_%+d_%d" % (storey, s, ci), 7 + (ci % 3), *_surface(arch, seg(36), seg(6), (2, 1), flip=True))
    # hanging curtains (wavy sheets), banners, plants with alpha-cutout leaves, vases, lion-head bosses
    for k in range(10):
        xc, s = xs[(k * 2) % n_col] + 1.0, (-1 if k % 2 else 1)
        def sheet(u, v, xc=xc, s=s, k=k):
            return np.stack([xc + (u - 0.5) * 2.6, H1 + 0.2 + (1 - v) * 3.6, s * (LZ - 3.4) + 0.22 * np.sin(u * (14 + k)) * (0.3 + v)], -1)
        add("curtain%d" % k, 10 + (k % 3), *_surface(sheet, seg(44), seg(36), (1, 1)))
    for k in range(6):
        xc = -LX + 5.0 + k * 5.2
        def leafball(u, v, xc=xc, k=k):
            a, b = u * 2 * np.pi, v * np.pi
            r = 0.9 + 0.35 * np.sin(a * 5 + k) * np.sin(b * 4)
            return np.stack([xc + r * np.sin(b) * np.cos(a), 1.3 + r * np.cos(b), r * np.sin(b) * np.sin(a)], -1)
        add("plant%d" % k, (11, 17)[k % 2], *_surface(leafball, seg(40), seg(28), (3, 3)))
        def vase(u, v, xc=xc):
            a = u * 2 * np.pi
            r = 0.28 + 0.22 * np.sin(v * np.pi) ** 2 + 0.1 * (1 - v)
            return np.stack([xc + r * np.cos(a), v * 0.9, r * np.sin(a)], -1)
        add("vase%d" % k, (14, 19)[k % 2], *_surface(vase, seg(36), seg(20), (2, 1)))
    for k in range(8):
        xc, s = xs[k + (k >= 4)] , (-1 if k % 2 else 1)
        def boss(u, v, xc=xc, s=s, k=k):
            a, b = u * 2 * np.pi, v * np.pi
            r = 0.45 * (1.0 + 0.18 * np.sin(a * 3 + k) * np.sin(b * 5) + 0.08 * np.cos(a * 9))
            return np.stack([xc + r * np.sin(b) * np.cos(a), 3.4 + r * np.cos(b), s * (LZ - 0.15) + 0.6 * r * np.sin(b) * np.sin(a)], -1)
        add("lion%d" % k, (12, 24)[k % 2], *_surface(boss, seg(56), seg(40), (1, 1)))
    sc.cam = Camera()
    sc.cam.eye_pos = np.array([-LX + 2.5, 2.2, 0.6], dtype=F)
    d = np.array([1.0, 0.12, -0.05]); sc.cam.dir = (d / np.linalg.norm(d)).astype(F)
    sc.cam.fov = to_radians(60.0)
    sc.cam.clipping_near, sc.cam.clipping_far = 0.1, 1000.0
    sc.cam.init(width, height)
    sc.add_default_light()
    sc.config = Config(monte_carlo=monte_carlo, samples=samples)
    return sc


def helmet_scene(width: int = 1280, height: int = 720, detail: float = 1.0, tex_size: int = 2048, seed: int = 0xC3,
                 samples: int = 32, monte_carlo: bool = False) -> Scene:
    """STAND-IN for BASELINE.json configs[2] (scene/helmet.json): one glTF primitive of ~70 k de-indexed triangles with the five
    2048x2048 maps of DamagedHelmet (base colour, normal, metallic-roughness -> two textures, occlusion, emissive), the
    environment sphere, the camera / light / transformation / config blocks of the JSON (fov 23, one point light of intensity 100,
    rotation (-25, 15, 0), scale 1.25, 32 spp with monte_carlo off — the file's own config beats the CLI)."""
    rng = np.random.default_rng(seed)
    sc = Scene(".")
    _environment(sc, rng)
    tex = pbr_texture_set(tex_size, rng, (0.8, 0.75, 0.7), metal=0.9, rough=(0.15, 0.8), bricks=0)
    m = _gltf_material(sc, "helmet", tex, ("base", "normal", "metallic", "emissive", "roughness", "occlusion"), nearest=False)
    q = max(0.02, float(detail)) ** 0.5
    nu, nv = max(8, int(round(264 * q))), max(6, int(round(132 * q)))     # 2 * 264 * 132 = 69 696 triangles

    def shell(u, v):
        a, b = u * 2 * np.pi, v * np.pi
        r = 1.0 + 0.12 * np.sin(3 * a) * np.sin(2 * b) ** 2 + 0.25 * np.exp(-((a - np.pi) ** 2 + (b - 1.7) ** 2) * 3.0) + 0.04 * np.cos(11 * a) * np.sin(7 * b)
        return np.stack([r * np.sin(b) * np.cos(a), 0.95 * r * np.cos(b), r * np.sin(b) * np.sin(a)], -1)
    it = Item(id=sc.get_next_id(), name="helmet", shape=SHAPE_MESH, material=m, trans=mat_identity(), mesh=_deindexed_mesh(*_surface(shell, nu, nv, (1, 1))))
    it.apply_transformation((0.3, 0.2, 0.0), (1.25, 1.25, 1.25), (to_radians(-25.0), to_radians(15.0), to_radians(0.0)))
    sc.items.append(it)
    sc.lights.append(Light(id=sc.get_next_id(), name="light point", pos=np.array([6.8627195, 3.287831, 1.4585655], dtype=F), dir=np.array([0, -1, 0], dtype=F),
                           color=np.array([1, 1, 1], dtype=F), intensity=100.0, max_angle=float(F(math.pi / 2)), light_type=LIGHT_POINT))
    sc.cam = Camera()
    sc.cam.eye_pos = np.array([4.2011, 2.7027438, 3.71161], dtype=F)
    sc.cam.up = np.array([-0.32401347, 0.8953957, -0.30542085], dtype=F)
    sc.cam.dir = np.array([-0.6515582, -0.4452714, -0.61417043], dtype=F)
    sc.cam.fov = to_radians(23.0)
    sc.cam.clipping_near, sc.cam.clipping_far = 0.1, 100.0
    sc.cam.init(width, height)
    sc.config = Config(monte_carlo=monte_carlo, samples=samples, focal_length=20.0, aperture_size=1.0, fog_density=0.0, fog_color=(0.4, 0.4, 0.4),
                       max_recursion=6, gamma_correction=False)
    return sc
