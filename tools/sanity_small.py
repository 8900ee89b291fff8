"""A small tour of every kernel (grouped scene, device-built BVH, frames on both schedules, probes, shadow probes, shard pack /
unpack, post-processing) — meant to be run under `compute-sanitizer --tool memcheck` on a B200:
    compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from rustray_b200 import abi, synthetic  # noqa: E402
from rustray_b200.renderer import RendererManager  # noqa: E402
from tests.util import random_rays, scene_to_abi  # noqa: E402

for device_bvh in (False, True):
    sc = synthetic.atrium_scene(96, 54, detail=0.03, tex_size=16, samples=2, monte_carlo=True)
    fs, cam, cfg = scene_to_abi(sc)
    g = RendererManager(96, 54, fs, device_bvh=device_bvh)
    i = g.bvh_info()
    f = g.start(cam, cfg)
    os.environ["RTX_FORCE_SYNC"] = "1"
    f2 = g.start(cam, cfg)
    del os.environ["RTX_FORCE_SYNC"]
    assert np.array_equal(f.objects, f2.objects)
    o, d = random_rays(2000, 4, center=(0, 4, 0), radius=9.0)
    g.trace(o, d); g.trace(o, d, for_shadow=True, stop_on_first_hit=True)
    s = g.shadow_probe(o, d, 9.0, np.zeros(2000, dtype=np.int32))
    print("atrium device_bvh=%s: grouped %d items / %d tris, frame %d+%d rays, %d lit of %d probes" % (
        device_bvh, i.grouped_items, i.grouped_triangles, f.stats.rays_closest, f.stats.rays_shadow, int((s["lit"] == 1).sum()), s.size))
    g.close()
sc = synthetic.soup_scene(20_000, 60, cells=2, width=64, height=36)
fs, cam, cfg = scene_to_abi(sc, samples=2, monte_carlo=1)
g = RendererManager(64, 36, fs, device_bvh=True)
f = g.start(cam, cfg)
print("soup:", f.stats.rays_closest, f.stats.rays_shadow, g.bvh_info().grouped_items)
fs, cam, cfg = abi.load_fixture("c2_floor_monkey", samples=2, monte_carlo=1)
cam = abi.resize_camera(cam, 96, 54)
g2 = RendererManager(96, 54, fs)
f = g2.start(cam, cfg)
print("c2 small:", f.stats.rays_closest, f.stats.rays_shadow, f.stats.host_syncs)
print("SANITY_OK")
