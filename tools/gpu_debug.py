import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager
from oracle.oracle import OracleRenderer
from tests.util import random_rays
fs, cam, cfg = abi.load_fixture("kbert")
g = RendererManager(cam.width, cam.height, fs); c = OracleRenderer(fs)
o, d = random_rays(6000, 3)
for kw in (dict(depth=1), dict(depth=2), dict(for_shadow=True), dict(for_shadow=True, stop_on_first_hit=True, depth=2)):
    hg, hc = g.trace(o, d, **kw), c.trace(o, d, **kw)
    bad = np.nonzero((hg["face_id"] != hc["face_id"]) | (hg["t"] != hc["t"]) | (hg["item_index"] != hc["item_index"]))[0]
    print(kw, "mismatches", bad.size)
    for i in bad[:10]:
        print("  ray", i, "o", o[i], "d", d[i], "gpu", hg[i], "cpu", hc[i])
        m = fs.mesh_arrays[fs.items[int(hc[i]["item_index"])].mesh]
        nf = m["indices"].shape[0]
        for f in (int(hg[i]["face_id"]) % nf, int(hc[i]["face_id"]) % nf):
            print("     face", f, m["vertices"][m["indices"][f]].tolist())
c.set_options(brute_force=True)
hb = c.trace(o, d)
hc2 = OracleRenderer(fs).trace(o, d)
print("oracle bvh vs brute mismatches", (hb.tobytes() != hc2.tobytes()))
