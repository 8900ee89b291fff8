"""Per-rank frame time of the sharded render measured on ONE GPU (rank r of N rendered one after the other):
shows the tile imbalance of the interleaved assignment.   python tools/shard_balance.py N [spp_per_gpu]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager
from rustray_b200.distributed import ShardedRenderer
n = int(sys.argv[1]); spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
fs, cam, cfg = abi.load_fixture("c2_floor_monkey", samples=spp * n, monte_carlo=1)
rm = RendererManager(cam.width, cam.height, fs)
ms = []
for r in range(n):
    sr = ShardedRenderer(rm, cam.width, cam.height, r, n)
    sr.render_local(cam, cfg)
    st = sr.render_local(cam, cfg)
    ms.append(st.device_ms)
    print("rank %d: %.2f ms, %d rays" % (r, st.device_ms, st.rays_closest + st.rays_shadow))
print("max %.2f mean %.2f imbalance %.1f%%" % (max(ms), sum(ms) / n, 100 * (max(ms) / (sum(ms) / n) - 1)))
