"""Scene build time of a bench workload with the host (binned SAH) and the device (Morton) BVH builder, phase by phase
(RTX_TRACE_BUILD=1 prints the phases of rtx_scene_create):  python tools/build_time.py c5"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RTX_TRACE_BUILD"] = "1"
import bench  # noqa: E402
from rustray_b200.renderer import RendererManager  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
fs, cam, cfg, desc = bench.build_workload(name)
for device_bvh in (False, True, False, True):
    t0 = time.perf_counter()
    g = RendererManager(cam.width, cam.height, fs, device_bvh=device_bvh)
    wall = (time.perf_counter() - t0) * 1e3
    i = g.bvh_info()
    print("%s builder: rtx_scene_create %.0f ms (wall incl. ctypes %.0f ms), builder kernels %.1f ms, %d nodes, %d triangles" % (
        "device" if device_bvh else "host", i.build_ms, wall, i.device_build_ms, i.n_nodes, i.n_triangles), flush=True)
    g.close()
