"""Frame time / Mrays/s of the other committed workloads (not the headline): config 1 (spheres 800x600 1 spp, deterministic),
the README scenes room+spheres (128 spp MC) and floor+kbert (64 spp MC), the glTF monkey (16 spp MC).
   python tools/bench_workloads.py [frames]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 5
out = {}
for name in ("c1_spheres", "c2_floor_monkey", "room_spheres", "kbert", "monkey_gltf"):
    fs, cam, cfg = abi.load_fixture(name)
    g = RendererManager(cam.width, cam.height, fs)
    for _ in range(2):
        g.start(cam, cfg)
    ms, rays = 0.0, 0
    for _ in range(frames):
        s = g.start(cam, cfg).stats
        ms += s.device_ms; rays = s.rays_closest + s.rays_shadow
    ms /= frames
    out[name] = {"size": [cam.width, cam.height], "samples": cfg.samples, "monte_carlo": cfg.monte_carlo, "items": len(fs.items), "triangles": fs.n_triangles,
                 "rays_per_frame": rays, "ms_per_frame": round(ms, 2), "Mrays_per_s": round(rays / ms / 1e3, 1)}
    print(name, out[name])
    g.close()
json.dump(out, open("gpurun_out/workloads.json", "w"), indent=1)
