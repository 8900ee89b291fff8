import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager
from oracle.oracle import OracleRenderer
from tests.util import clone_cfg
W, H = int(sys.argv[1]), int(sys.argv[2])
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
name = sys.argv[4] if len(sys.argv) > 4 else "room_spheres"
fs, cam, cfg = abi.load_fixture(name, samples=1, monte_carlo=0)
cam = abi.resize_camera(cam, W, H)
cfg = clone_cfg(cfg, debug_flags=flags)
r = OracleRenderer(fs).render(cam, cfg)
print("cpu", r.stats.rays_closest, r.stats.rays_shadow)
bad = 0
for h in range(2):
    g = RendererManager(W, H, fs)
    for k in range(4):
        a = g.start(cam, cfg)
        d = np.abs(a.image.astype(int) - r.image.astype(int)).max(-1)
        bad += (a.stats.rays_closest, a.stats.rays_shadow) != (r.stats.rays_closest, r.stats.rays_shadow)
        if os.environ.get("QUIET") is None: print("handle", h, "frame", k, a.stats.rays_closest, a.stats.rays_shadow, "px>1LSB", int((d > 1).sum()), "ids_eq", bool((a.objects == r.objects).all()))
    g.close()
print(name, os.environ.get("RTX_LIB", "default")[-20:], "frames with wrong ray totals:", bad, "of 8")
