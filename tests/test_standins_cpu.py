"""CPU suite, part 5: the labelled stand-in scenes of BASELINE.json configs[2] / configs[3] (rustray_b200/synthetic.py), the
glTF texture mapping on a generated .glb (reference src/scene.rs:895-960, 980-1124), and the oracle's item BVH / shadow query."""
import os

import numpy as np

from rustray_b200 import abi, synthetic
from rustray_b200.scene_loader import load_scene, TEX_AMBIENT, TEX_AO, TEX_BASE, TEX_NORMAL, TEX_REFLECTIVITY, TEX_ROUGHNESS
from oracle.oracle import OracleRenderer
from tests.util import random_rays, scene_to_abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_gltf_texture_mapping_matches_scene_rs():
    """Every branch of get_dyn_image_from_gltf_material on the generated textured_pbr.glb."""
    img = np.load(os.path.join(GOLDEN, "textured_pbr_images.npz"))
    sc = load_scene([os.path.join("tests", "golden", "textured_pbr.glb")], 300, 200, asset_root=ROOT)
    assert [it.name for it in sc.items] == ["patch", "tetra"]
    # ids: light first (scene.rs:732-787), then per primitive object id, then its new material (:893)
    assert [l.id for l in sc.lights] == [1] and [(it.id, it.material.id) for it in sc.items] == [(2, 3), (4, 5)]
    assert sc.lights[0].intensity == np.float32(900.0) / np.float32(10.0)                          # :747
    m = sc.items[0].material
    tex = {t: sc.texture_data[m.textures[t]] for t in (TEX_BASE, TEX_AMBIENT, TEX_NORMAL, TEX_ROUGHNESS, TEX_AO, TEX_REFLECTIVITY)}
    assert np.array_equal(tex[TEX_BASE], img["base"])                                              # RGBA kept (:990-1003)
    assert np.array_equal(tex[TEX_NORMAL][..., :3], img["normal"]) and (tex[TEX_NORMAL][..., 3] == 255).all()   # :1014-1024
    for c in range(4):
        assert np.array_equal(tex[TEX_ROUGHNESS][..., c], img["mr"][..., 1])                       # G -> grey (:1036-1047)
        assert np.array_equal(tex[TEX_REFLECTIVITY][..., c], img["mr"][..., 2])                    # B -> grey (:1083-1095)
        assert np.array_equal(tex[TEX_AO][..., c], np.trunc(img["occ"][..., 0].astype(np.float32) * np.float32(0.75)).astype(np.uint8))   # :1060-1070
    assert np.array_equal(tex[TEX_AMBIENT][..., :3], img["emis"]) and (tex[TEX_AMBIENT][..., 3] == 255).all()   # :1107-1118
    assert np.allclose(m.ambient_color, [1.0, 0.5, 0.25]) and np.allclose(m.base_color, [0.9, 0.8, 0.7])
    assert np.allclose(m.specular_color, np.float32([0.9, 0.8, 0.7]) * np.float32(0.8))
    assert m.reflectivity == np.float32(0.6) * np.float32(0.5) and abs(m.roughness - 0.5 / (2 * np.pi)) < 1e-7
    g = sc.items[1].material
    assert g.textures == [None] * 8 and abs(g.alpha - 0.4) < 1e-7 and g.roughness == 0.0
    # de-indexed, uv.y := 1 - v, node transforms baked (:853-891)
    me = sc.items[0].mesh
    assert me.vertices.shape == (216, 3) and me.uvs.shape == (216, 2) and me.normals.shape == (216, 3)
    assert np.array_equal(me.indices.reshape(-1), np.arange(216)) and sc.items[1].mesh.uvs.shape[0] == 0
    assert me.uvs[:, 1].min() < 0.0 and abs(me.vertices[:, 2].mean() + 6.0) < 0.5                  # translated to z = -6
    assert abs(sc.cam.fov - 0.9) < 1e-6 and np.allclose(sc.cam.eye_pos, [0, 2, 1])
    fs = abi.FlatScene.from_scene(sc)
    assert len(fs.textures) == 6 and fs.n_triangles == 76


def test_atrium_standin_has_the_shape_of_config_4():
    sc = synthetic.atrium_scene(64, 36, detail=1.0, tex_size=64)
    fs = abi.FlatScene.from_scene(sc)
    meshes = [it for it in sc.items if it.mesh is not None]
    assert len(sc.items) > 100 and len(meshes) >= 100                       # > BVH_MIN_ITEMS (raytracing.rs:23)
    assert 240_000 <= fs.n_triangles <= 290_000                             # Sponza: ~262 k triangles
    assert sc.items[0].name == "environment" and sc.items[0].material.reflection_only and not sc.items[0].material.backface_cullig
    assert all(it.material.texture_filtering_nearest and it.material.backface_cullig for it in meshes)   # sponza.json: nearest; the misspelt key is ignored
    assert all(it.material.textures[TEX_BASE] and it.material.textures[TEX_NORMAL] and it.material.textures[TEX_ROUGHNESS]
               and it.material.textures[TEX_REFLECTIVITY] for it in meshes)
    assert len(sc.lights) == 1 and sc.lights[0].intensity == 200.0 and (sc.config.samples, sc.config.monte_carlo) == (128, True)
    assert all(np.array_equal(it.mesh.indices.reshape(-1), np.arange(it.mesh.vertices.shape[0])) for it in meshes)   # de-indexed like load_gltf
    a = abi.FlatScene.from_scene(synthetic.atrium_scene(64, 36, detail=0.02, tex_size=32))
    b = abi.FlatScene.from_scene(synthetic.atrium_scene(64, 36, detail=0.02, tex_size=32))
    assert bytes(a.items) == bytes(b.items) and all(np.array_equal(x["vertices"], y["vertices"]) for x, y in zip(a.mesh_arrays, b.mesh_arrays))   # seeded


def test_helmet_standin_has_the_shape_of_config_3():
    sc = synthetic.helmet_scene(64, 36, tex_size=64)
    fs = abi.FlatScene.from_scene(sc)
    assert [it.name for it in sc.items] == ["environment", "helmet"] and 65_000 <= fs.n_triangles <= 75_000
    m = sc.items[1].material
    assert all(m.textures[t] for t in (TEX_BASE, TEX_AMBIENT, TEX_NORMAL, TEX_ROUGHNESS, TEX_AO, TEX_REFLECTIVITY)) and not m.texture_filtering_nearest
    assert (sc.config.samples, sc.config.monte_carlo, sc.config.aperture_size) == (32, False, 1.0)   # helmet.json's own config block
    assert len(sc.lights) == 1 and sc.lights[0].intensity == 100.0 and abs(float(sc.cam.fov) - np.deg2rad(23.0)) < 1e-6
    full = synthetic.helmet_scene(64, 36, tex_size=2048)
    assert sum(t.nbytes for t in abi.FlatScene.from_scene(full).textures) >= 6 * 2048 * 2048 * 4       # five 2048^2 maps (metallic-roughness counts twice) + environment


def test_oracle_item_bvh_equals_the_all_items_loop():
    """> 50 items: candidates come from the item BVH (scene.rs:1715-1722); it must not change any answer."""
    for sc in (synthetic.atrium_scene(96, 54, detail=0.02, tex_size=32, samples=1, monte_carlo=False),
               synthetic.soup_scene(n_triangles=20_000, n_spheres=150, cells=2, width=96, height=54)):
        fs, cam, cfg = scene_to_abi(sc, samples=1, monte_carlo=0)
        assert len(fs.items) > 50
        a, b = OracleRenderer(fs), OracleRenderer(fs)
        b.set_options(brute_force=True)
        o, d = random_rays(3000, 11, center=(0, 2, 0), radius=30.0)
        for kw in (dict(), dict(depth=2), dict(for_shadow=True, stop_on_first_hit=True)):
            ha, hb = a.trace(o, d, **kw), b.trace(o, d, **kw)
            assert (ha["t"] >= 0).sum() > 50 and ha.tobytes() == hb.tobytes()
        fa, fb = a.render(cam, cfg), b.render(cam, cfg)
        assert np.array_equal(fa.image, fb.image) and np.array_equal(fa.objects, fb.objects)
        assert (fa.stats.rays_closest, fa.stats.rays_shadow) == (fb.stats.rays_closest, fb.stats.rays_shadow)


def test_oracle_shadow_probe_is_the_shading_loops_query():
    """oracle_shadow_probe restates raytracing.rs:883-914; it must agree with trace(.., true, true) + the distance rule."""
    sc = synthetic.feature_scene(96, 64)
    fs, cam, cfg = scene_to_abi(sc)
    c = OracleRenderer(fs)
    o, d = random_rays(4000, 21, center=(0, 1, -12), radius=14.0)
    first = c.trace(o, d, for_shadow=True, stop_on_first_hit=True)
    rng = np.random.default_rng(2)
    length = rng.uniform(2.0, 30.0, o.shape[0]).astype(np.float32)
    recv = rng.integers(0, len(fs.items), o.shape[0]).astype(np.int32)
    for ld in (None, length):
        s = c.shadow_probe(o, d, light_distance=ld, receiver_item=recv)
        hit = first["t"] >= 0
        lit = ~hit if ld is None else (~hit | (first["t"] > ld))
        assert np.array_equal(s["lit"] == 1, lit) and 100 < lit.sum() < lit.size - 100
        assert (s["k"][lit] == 1.0).all() and np.array_equal(s["occluder_index"][~lit], first["item_index"][~lit])
        # (a sphere receiver's get_uv of a point that is not on it is NaN — acos of |y/r| > 1 — and the bilinear fetch keeps it: NaN k)
        assert np.array_equal(s["t"][~lit], first["t"][~lit]) and (np.isnan(s["k"][~lit]) | ((s["k"][~lit] >= 0) & (s["k"][~lit] <= 1))).all()
    # an opaque receiver in front of an opaque occluder: k = 1 - receiver alpha
    alphas = np.array([fs.materials[it.material].alpha for it in fs.items], dtype=np.float32)
    tex_alpha = np.array([fs.materials[it.material].texture[4] >= 0 for it in fs.items])
    s = c.shadow_probe(o, d, receiver_item=recv)
    plain = (s["lit"] == 0) & ~tex_alpha[np.maximum(s["occluder_index"], 0)]
    assert plain.sum() > 100 and np.array_equal(s["k"][plain], (np.float32(1.0) - alphas[recv[plain]]).astype(np.float32))
