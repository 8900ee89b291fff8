import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager
from oracle.oracle import OracleRenderer
from tests.util import random_rays
fs, cam, cfg = abi.load_fixture("room_spheres")
g = RendererManager(cam.width, cam.height, fs); c = OracleRenderer(fs)
rng = np.random.default_rng(5)
n = 400000
o = rng.uniform([-8, -4, -18], [8, 6, 0], size=(n, 3)).astype(np.float32)
d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
hc = c.trace(o, d, depth=2)
for k in range(4):
    hg = g.trace(o, d, depth=2)
    bad = np.nonzero((hg["t"] != hc["t"]) | (hg["item_index"] != hc["item_index"]) | (hg["face_id"] != hc["face_id"]))[0]
    print("run", k, "mismatches vs oracle", bad.size, bad[:8])
    for i in bad[:4]:
        print("   o", o[i].tolist(), "d", d[i].tolist(), "\n   gpu", hg[i], "\n   cpu", hc[i])
hs = g.trace(o, d, for_shadow=True, depth=2)
print("old-path (for_shadow) vs oracle mismatches", (hs.tobytes() != c.trace(o, d, for_shadow=True, depth=2).tobytes()))
