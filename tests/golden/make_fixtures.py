"""Generate the committed flat-scene fixtures from the reference's scene files.

Run HERE (the container that has /root/reference); the GPU box has no reference tree, so tests and the
bench read these .npz files instead.  Each file holds the flattened scene (rustray_b200.abi.FlatScene),
the camera matrices of Camera::init_matrices and the effective RaytracingConfig after the reference's
"JSON config beats CLI" rule.

    python tests/golden/make_fixtures.py [/root/reference]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from rustray_b200.abi import FlatScene  # noqa: E402
from rustray_b200.scene_loader import load_scene  # noqa: E402

SCENES = {
    # name: (scene files in CLI order, width, height, cli samples, cli monte_carlo)
    "c1_spheres": (["scene/spheres.json"], 800, 600, 1, False),
    "c2_floor_monkey": (["scene/floor.json", "scene/monkey.json"], 1280, 720, 32, True),
    "room_spheres": (["scene/room-no-textures.json", "scene/spheres.json"], 1280, 720, 128, True),
    "kbert": (["scene/floor.json", "scene/kbert.json"], 1280, 720, 64, True),
    # glTF path (easy-gltf semantics: de-indexed mesh, file camera + KHR lights, PBR material with roughness jitter)
    "monkey_gltf": (["scene/models/monkey/monkey.gltf"], 1280, 720, 16, True),
    # nested scene files (room.json + kbert.json through "type": "json" objects), three base-colour textures (png, gif), OBJ + MTL
    "kbert_in_room": (["scene/kbert_in_room.json"], 1280, 720, 32, True),
    # a sphere with base, specular and normal maps (sphere uv + tangent frame, three 2048x1024 JPEGs) in the textured room
    "earth_in_room": (["scene/earth_in_room.json"], 1280, 720, 16, True),
}


def main(ref_root: str, only=None) -> None:
    for name, (files, w, h, samples, mc) in SCENES.items():
        if only and name not in only:
            continue
        sc = load_scene(files, w, h, asset_root=ref_root, samples=samples, monte_carlo=mc)
        fs = FlatScene.from_scene(sc)
        c = sc.config
        cfg = np.array([int(c.monte_carlo), c.samples, c.focal_length, c.aperture_size, c.fog_density, *c.fog_color,
                        c.max_recursion, int(c.gamma_correction)], dtype=np.float64)
        out = os.path.join(HERE, "scenes", name + ".npz")
        fs.save(out, projection_inverse=sc.cam.projection_inverse, view_inverse=sc.cam.view_inverse,
                size=np.array([w, h]), config=cfg, files=np.array(files))
        print("%-18s items=%d tris=%d textures=%d lights=%d -> %.1f KiB" % (
            name, len(fs.items), fs.n_triangles, len(fs.textures), len(fs.lights), os.path.getsize(out) / 1024))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", only=sys.argv[2:] or None)   # [ref root] [fixture names...]
