"""One bench workload on one GPU, a few frames, with the library's own per-frame statistics (and, with RTX_PHASE_STATS=1, the
lane-utilisation counters of the step loop).  Meant for ncu launch lists / captures of the non-headline workloads:
    python tools/run_workload.py c4_standin [frames] [stats]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rustray_b200 import abi  # noqa: E402
from rustray_b200.renderer import RendererManager  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4_standin"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
want_stats = len(sys.argv) > 3 and sys.argv[3] == "stats"
fs, cam, cfg, desc = bench.build_workload(name)
g = RendererManager(cam.width, cam.height, fs)
info = g.bvh_info()
desc.update({"scene_build_ms": info.build_ms, "grouped_items": info.grouped_items, "grouped_triangles": info.grouped_triangles, "bvh_nodes": info.n_nodes})
print(json.dumps(desc))
if want_stats:
    c = abi.RtxConfig(); C.memmove(C.byref(c), C.byref(cfg), C.sizeof(cfg)); c.debug_flags = 1
    s = g.start(cam, c).stats
    n = s.rays_closest + s.rays_shadow
    print("stats frame: rays closest %d shadow %d | node visits/ray closest %.2f shadow %.2f | tri tests/ray %.2f %.2f | item tests/ray %.2f | sphere tests/ray %.3f" % (
        s.rays_closest, s.rays_shadow, s.node_visits[0] / max(1, s.rays_closest), s.node_visits[1] / max(1, s.rays_shadow),
        s.tri_tests[0] / max(1, s.rays_closest), s.tri_tests[1] / max(1, s.rays_shadow), s.item_tests / max(1, n), s.sphere_tests / max(1, n)))
g.start(cam, cfg)
for _ in range(frames):
    s = g.start(cam, cfg).stats
    print("frame: %.2f ms device (closest %.2f, shadow %.2f, shade+rest %.2f) | %d closest + %d shadow rays (%d beyond-light checks, %d via the exact walk) | %.1f Mrays/s | waves %d, syncs %d, launches %d" % (
        s.device_ms, s.closest_ms, s.shadow_ms, s.shade_ms, s.rays_closest, s.rays_shadow, s.rays_shadow_beyond, s.rays_shadow_exact, (s.rays_closest + s.rays_shadow) / s.device_ms / 1e3,
        s.waves, s.host_syncs, s.kernel_launches))
