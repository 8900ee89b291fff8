import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager, primary_ray
from oracle.oracle import OracleRenderer
from tests.util import random_rays
name = sys.argv[1] if len(sys.argv) > 1 else "c1_spheres"
fs, cam, cfg = abi.load_fixture(name)
g = RendererManager(cam.width, cam.height, fs); c = OracleRenderer(fs)
rng = np.random.default_rng(7)
xs, ys = rng.integers(0, cam.width, 6000), rng.integers(0, cam.height, 6000)
rays = [primary_ray(cam, int(x), int(y)) for x, y in zip(xs, ys)]
o = np.array([r[0] for r in rays]); d = np.array([r[1] for r in rays])
o2, d2 = random_rays(6000, 3)
for (oo, dd, tag) in ((o, d, "primary"), (o2, d2, "random")):
  for kw in (dict(depth=1), dict(depth=2), dict(for_shadow=True), dict(for_shadow=True, stop_on_first_hit=True, depth=2)):
    hg, hc = g.trace(oo, dd, **kw), c.trace(oo, dd, **kw)
    bad = np.nonzero((hg["face_id"] != hc["face_id"]) | (hg["t"] != hc["t"]) | (hg["item_index"] != hc["item_index"]))[0]
    print(tag, kw, "mismatches", bad.size)
    for i in bad[:6]:
        print("  ray", i, "o", oo[i], "d", dd[i], "gpu", hg[i], "cpu", hc[i])
