"""Config 5 (synthetic triangle soup + spheres): build + render timing, traversal statistics.
   python tools/bench_c5.py N_TRIANGLES N_SPHERES WIDTH HEIGHT SPP [cells]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi, synthetic
from rustray_b200.renderer import RendererManager
nt, ns, w, h, spp = (int(x) for x in sys.argv[1:6])
cells = int(sys.argv[6]) if len(sys.argv) > 6 else 4
t = time.time(); sc = synthetic.soup_scene(nt, ns, cells=cells, width=w, height=h); print("generate %.1fs" % (time.time() - t))
t = time.time(); fs = abi.FlatScene.from_scene(sc); print("flatten %.1fs" % (time.time() - t))
cam = abi.make_camera(sc.cam); cfg = abi.make_config(sc.config, samples=spp, monte_carlo=1)
t = time.time(); g = RendererManager(w, h, fs); print("scene_create %.1fs" % (time.time() - t))
info = g.bvh_info(); print("nodes %d (%.0f MB) tris %d (%.0f MB) items %d tlas nodes %d" % (info.n_nodes, info.node_bytes / 1e6, info.n_triangles, info.triangle_bytes / 1e6, info.n_items, info.tlas_nodes))
import ctypes as C
cs = abi.RtxConfig(); C.memmove(C.byref(cs), C.byref(cfg), C.sizeof(cfg)); cs.debug_flags = 1; cs.samples = min(spp, 4)
s = g.start(cam, cs).stats
print("per ray: closest nodes %.1f tris %.1f | shadow nodes %.1f tris %.1f | items/ray %.2f" % (
    s.node_visits[0] / s.rays_closest, s.tri_tests[0] / s.rays_closest, s.node_visits[1] / max(1, s.rays_shadow), s.tri_tests[1] / max(1, s.rays_shadow),
    s.item_tests / (s.rays_closest + s.rays_shadow)))
bc = (s.node_visits[0] * 80 + s.tri_tests[0] * 48) / s.rays_closest; bs = (s.node_visits[1] * 80 + s.tri_tests[1] * 48) / max(1, s.rays_shadow)
for i in range(int(os.environ.get("C5_FRAMES", "3"))):
    t = time.time(); f = g.start(cam, cfg); dt = time.time() - t; s = f.stats
    rays = s.rays_closest + s.rays_shadow
    print("frame %d: wall %.2fs device %.1f ms closest %.1f ms shadow %.1f ms | rays %.1fM+%.1fM -> %.0f Mrays/s | closest %.0f GB/s shadow %.0f GB/s algorithmic | waves %d" % (
        i, dt, s.device_ms, s.closest_ms, s.shadow_ms, s.rays_closest / 1e6, s.rays_shadow / 1e6, rays / s.device_ms / 1e3,
        s.rays_closest * bc / s.closest_ms / 1e6, s.rays_shadow * bs / max(1e-9, s.shadow_ms) / 1e6, s.waves))
from PIL import Image
Image.fromarray(f.image).save("gpurun_out/c5.png")
