"""Frame time of one committed fixture at its own config:  python tools/one_fixture.py room_spheres [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager
name = sys.argv[1]; frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fs, cam, cfg = abi.load_fixture(name)
g = RendererManager(cam.width, cam.height, fs)
i = g.bvh_info()
g.start(cam, cfg)
for _ in range(frames):
    s = g.start(cam, cfg).stats
print("%s: %.2f ms | %d closest + %d shadow rays (%d beyond checks, %d exact) | grouped items %d tris %d | waves %d" % (
    name, s.device_ms, s.rays_closest, s.rays_shadow, s.rays_shadow_beyond, s.rays_shadow_exact, i.grouped_items, i.grouped_triangles, s.waves))
