"""CPU suite, part 3: the stand-in scene loader (rustray_b200/scene_loader.py, obj_loader.py) against the
semantics of reference src/scene.rs, and the fixture round trip.  Loader tests need the reference's scene
files and are skipped where /root/reference is not mounted (the GPU box)."""
import numpy as np
import pytest

from rustray_b200 import abi
from rustray_b200.scene_loader import (Material, load_scene, mat_inverse, mat_mul, mat_translation, mat_euler, mat_scaling,
                                       approx_equal)
from tests.conftest import REFERENCE, needs_reference


def test_matrix_helpers():
    m = mat_mul(mat_mul(mat_translation(1, 2, 3), mat_euler(0.3, 0.0, 0.0)), mat_scaling(2, 3, 4))
    inv = mat_inverse(m)
    assert np.allclose(mat_mul(m, inv), np.eye(4), atol=1e-6)
    assert mat_inverse(np.zeros((4, 4), dtype=np.float32)) is None
    # Rotation3::from_euler_angles(0, pitch, 0) rotates about +Y
    r = mat_euler(0.0, np.pi / 2, 0.0)
    assert np.allclose(r[:3, :3] @ np.array([0, 0, 1.0]), [1, 0, 0], atol=1e-6)


def test_apply_diff_copies_only_non_default_fields():
    """Material::apply_diff (shape/mod.rs:182-299)"""
    base = Material(id=1)
    base.shininess, base.base_color = 324.0, np.array([0.1, 0.4, 0.8], dtype=np.float32)
    wrap = Material(id=2)
    wrap.alpha, wrap.reflectivity = 0.5, 0.5                       # non-default -> copied
    wrap.shininess = 150.0                                         # default -> NOT copied
    base.apply_diff(wrap)
    assert (base.alpha, base.reflectivity, base.shininess) == (0.5, 0.5, 324.0)
    assert np.allclose(base.base_color, [0.1, 0.4, 0.8])
    assert approx_equal(1.0, 1.0000004) and not approx_equal(1.0, 1.00001)


def test_fixture_round_trip(tmp_path):
    fs, cam, cfg = abi.load_fixture("c2_floor_monkey")
    p = str(tmp_path / "x.npz")
    fs.save(p, size=np.array([cam.width, cam.height]))
    fs2 = abi.FlatScene.load(p)
    assert bytes(fs.items) == bytes(fs2.items) and bytes(fs.materials) == bytes(fs2.materials) and bytes(fs.lights) == bytes(fs2.lights)
    assert fs2.n_triangles == 15746 and (cfg.samples, cfg.monte_carlo) == (32, 1)
    for a, b in zip(fs.mesh_arrays, fs2.mesh_arrays):
        for k in a:
            assert np.array_equal(a[k], b[k])
    d = fs2.desc()
    assert (d.n_items, d.n_meshes, d.n_materials, d.n_textures, d.n_lights) == (2, 2, 2, 1, 4)


@needs_reference
def test_fixtures_are_what_the_loader_produces_today():
    """The committed .npz fixtures equal a fresh load of the reference's scene files."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(__file__), "golden", "make_fixtures.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    for name in ("c1_spheres", "c2_floor_monkey"):
        files, w, h, samples, mc = mk.SCENES[name]
        sc = load_scene(files, w, h, asset_root=REFERENCE, samples=samples, monte_carlo=mc)
        fresh = abi.FlatScene.from_scene(sc)
        fix, cam, cfg = abi.load_fixture(name)
        assert bytes(fresh.items) == bytes(fix.items) and bytes(fresh.materials) == bytes(fix.materials)
        assert np.allclose(np.array(cam.projection_inverse), sc.cam.projection_inverse.T.reshape(-1))


@needs_reference
def test_monkey_obj_single_index_triangulation():
    """tobj triangulate + single_index: 7872 quads -> 15744 triangles, one vertex per distinct v/vt/vn triple,
    fan order (0,1,2),(0,2,3) (scene.rs:1130-1137)."""
    from rustray_b200.obj_loader import load_obj
    models, mtls = load_obj(REFERENCE + "/scene/models/monkey/monkey.obj")
    assert len(models) == 1 and models[0]["name"] == "Suzanne"
    m = models[0]
    idx = np.array(m["indices"]).reshape(-1, 3)
    assert idx.shape[0] == 15744 and idx.max() + 1 == len(m["positions"]) // 3 == len(m["normals"]) // 3 == len(m["texcoords"]) // 2
    assert (idx[0::2, 0] == idx[1::2, 0]).all() and (idx[0::2, 2] == idx[1::2, 1]).all()
    assert idx[0].tolist() == [0, 1, 2] and idx[1].tolist() == [0, 2, 3]
    assert mtls[0]["name"] == "Material.001" and mtls[0]["Ns"] == pytest.approx(323.999994) and mtls[0]["illum"] == 2


@needs_reference
def test_json_config_beats_cli_and_nested_scenes():
    """SURVEY.md fact 6 (main.rs:79-83 -> run.rs:216 -> scene.rs:179-198) and nested `json` objects
    (scene.rs:467-530): ids keep counting, nested spheres/planes are not touched by the wrapper."""
    sc = load_scene(["scene/spheres_in_room.json"], 640, 360, asset_root=REFERENCE, samples=64, monte_carlo=True)
    assert sc.config.samples == 64 and sc.config.monte_carlo      # neither file has a config block
    ids = [it.id for it in sc.items]
    assert ids == sorted(ids) and len(set(ids)) == len(ids)
    assert len(sc.items) == 6 + 8                                  # room planes + spheres
    assert not sc.cam.is_default_cam()


@needs_reference
def test_gltf_loader_semantics():
    """Scene::load_gltf (scene.rs:722-978): one de-indexed Mesh item per primitive, object id before material id,
    file camera and KHR_lights_punctual lights (point intensity / 10), PBR factor mapping; .glb and .gltf agree."""
    a = load_scene(["scene/models/monkey/monkey.gltf"], 320, 180, asset_root=REFERENCE)
    b = load_scene(["scene/models/monkey/monkey.glb"], 320, 180, asset_root=REFERENCE)
    it = a.items[0]
    assert it.name == "Suzanne" and it.mesh.indices.shape == (15744, 3) and it.mesh.vertices.shape == (47232, 3)
    assert np.array_equal(it.mesh.indices.reshape(-1), np.arange(47232))            # fully de-indexed
    assert it.mesh.normals.shape == (47232, 3) and it.mesh.uvs.shape == (47232, 2)
    assert [l.id for l in a.lights] == [1, 2] and (it.id, it.material.id) == (3, 4)  # lights, object id, then material id
    m = it.material
    assert m.reflectivity == 0.0 and m.roughness == pytest.approx(0.4 / (2 * np.pi), rel=1e-3) or m.roughness > 0
    assert np.allclose(m.specular_color, m.base_color * np.float32(0.8))
    assert not a.cam.is_default_cam() and a.cam.fov == pytest.approx(0.3996, abs=1e-3)
    assert np.allclose(b.items[0].mesh.vertices, it.mesh.vertices, atol=1e-6)


def test_animation_keyframe_interpolation():
    """reference src/animation.rs: frame count, keyframe pick, linear interpolation, trans = T·Rz·Ry·Rx·S from identity
    (the helmet.json turntable: rotation y 15 -> 375 degrees over 6 s at 25 fps)."""
    from rustray_b200.animation import Animation
    spec = {"fps": 25, "enabled": True, "keyframes": [
        {"time": 0, "objects": [{"name": "helmet", "transformation": {"rotation": {"x": -25.0, "y": 15.0, "z": 0.0},
                                                                        "scale": {"x": 1.25, "y": 1.25, "z": 1.25}, "translation": {"x": 0.3, "y": 0.2, "z": 0.0}}}]},
        {"time": 6000, "objects": [{"name": "helmet", "transformation": {"rotation": {"x": -25.0, "y": 375.0, "z": 0.0},
                                                                           "scale": {"x": 1.25, "y": 1.25, "z": 1.25}, "translation": {"x": 0.3, "y": 0.2, "z": 0.0}}}]}]}
    an = Animation(spec)
    assert an.has_animation() and an.frames_to_render() == 150
    assert an.trans_for_frame(10, "nobody") is None
    m0, m75 = an.trans_for_frame(0, "helmet"), an.trans_for_frame(75, "helmet")
    assert np.allclose(m0[:3, 3], [0.3, 0.2, 0.0]) and np.allclose(np.linalg.det(m0[:3, :3].astype(np.float64)), 1.25 ** 3, rtol=1e-5)
    # frame 75 = 3000 ms: rotation y = 15 + 180 degrees -> the rotation part is m0's with x and z columns mirrored about Y
    ry = lambda d: np.array([[np.cos(np.radians(d)), 0, np.sin(np.radians(d))], [0, 1, 0], [-np.sin(np.radians(d)), 0, np.cos(np.radians(d))]])
    rx = lambda d: np.array([[1, 0, 0], [0, np.cos(np.radians(d)), -np.sin(np.radians(d))], [0, np.sin(np.radians(d)), np.cos(np.radians(d))]])
    assert np.allclose(m75[:3, :3], ry(195.0) @ rx(-25.0) * 1.25, atol=1e-5)
    assert not Animation({"enabled": False}).has_animation() and not Animation(None).has_animation()
