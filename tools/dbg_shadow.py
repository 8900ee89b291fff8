import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from rustray_b200 import abi, synthetic
from rustray_b200.renderer import RendererManager
from oracle.oracle import OracleRenderer
from tests.util import scene_to_abi
from tests.test_gpu_round2 import _shadow_rays_of_a_frame
sc = synthetic.feature_scene(224, 144)
fs, cam, cfg = scene_to_abi(sc)
g, c = RendererManager(224, 144, fs), OracleRenderer(fs)
o, d, ln, recv = _shadow_rays_of_a_frame(fs, cam, c, n=8000)
sg, sc_ = g.shadow_probe(o, d, ln, recv, 1), c.shadow_probe(o, d, ln, recv, 1)
bad = ~((sg["k"] == sc_["k"]) | (np.isnan(sg["k"]) & np.isnan(sc_["k"])))
print("mismatch", bad.sum(), "of", bad.size)
names = fs.item_names
for i in np.nonzero(bad)[0][:12]:
    print(i, "gpu", sg[i], "cpu", sc_[i], "recv", recv[i], names[recv[i]], "occ", names[sc_["occluder_index"][i]] if sc_["occluder_index"][i] >= 0 else None, "len", ln[i])
