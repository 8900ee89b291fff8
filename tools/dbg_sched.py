import os, sys
sys.path.insert(0, os.getcwd())
os.environ["RTX_TRACE_SCHED"]="1"
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager
fs, cam, cfg = abi.load_fixture("c1_spheres", samples=1, monte_carlo=0)
g = RendererManager(800, 600, fs)
a = g.start(cam, cfg)
print(a.stats.host_syncs, a.stats.waves, a.stats.rays_closest, a.stats.rays_shadow)
