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
