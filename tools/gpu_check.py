"""Quick GPU diagnostics (not a test): probe + image parity vs the oracle and a timing of C2."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustray_b200 import abi
from rustray_b200.renderer import RendererManager, primary_ray
from oracle.oracle import OracleRenderer

def cmp_frames(a, b, tag):
    diff = np.abs(a.image.astype(np.int32) - b.image.astype(np.int32)).max(axis=-1)
    print("%s: <=1LSB %.5f  exact %.5f  maxdiff %d  ids_eq %.5f  depth_rel %.3g  rays gpu (%d,%d) cpu (%d,%d)" % (
        tag, (diff <= 1).mean(), (diff == 0).mean(), diff.max(), (a.objects == b.objects).mean(),
        np.nanmax(np.abs(a.depth - b.depth) / np.maximum(np.abs(b.depth), 1e-6)),
        a.stats.rays_closest, a.stats.rays_shadow, b.stats.rays_closest, b.stats.rays_shadow))

for name in ("c1_spheres", "c2_floor_monkey", "room_spheres", "kbert"):
    fs, cam, cfg = abi.load_fixture(name, samples=1, monte_carlo=0)
    cam = abi.resize_camera(cam, 320, 180)
    t = time.time(); rm = RendererManager(320, 180, fs); print(name, "scene_create %.2fs" % (time.time() - t))
    orc = OracleRenderer(fs)
    # probes
    rng = np.random.default_rng(1)
    xs = rng.integers(0, 320, 4000); ys = rng.integers(0, 180, 4000)
    rays = [primary_ray(cam, int(x), int(y)) for x, y in zip(xs, ys)]
    o = np.array([r[0] for r in rays]); d = np.array([r[1] for r in rays])
    hg = rm.trace(o, d); hc = orc.trace(o, d)
    hitm = hc["t"] >= 0
    print("  probe: item_eq %.5f face_eq %.5f t_biteq %.5f n_maxdiff %.3g hits %d" % (
        (hg["item_index"] == hc["item_index"]).mean(), (hg["face_id"] == hc["face_id"]).mean(),
        (hg["t"][hitm] == hc["t"][hitm]).mean() if hitm.any() else 1.0,
        np.abs(hg["normal"][hitm] - hc["normal"][hitm]).max() if hitm.any() else 0.0, hitm.sum()))
    hg = rm.trace(o, d, for_shadow=True, stop_on_first_hit=True); hc = orc.trace(o, d, for_shadow=True, stop_on_first_hit=True)
    print("  shadow probe: item_eq %.5f t_biteq %.5f" % ((hg["item_index"] == hc["item_index"]).mean(), (hg["t"] == hc["t"]).mean()))
    f = rm.start(cam, cfg); ref = orc.render(cam, cfg)
    cmp_frames(f, ref, "  det 1spp")
    cfg2 = abi.RtxConfig(); import ctypes; ctypes.memmove(ctypes.byref(cfg2), ctypes.byref(cfg), ctypes.sizeof(cfg))
    cfg2.debug_flags = 2
    f2 = rm.start(cam, cfg2)
    print("  ordered-shadow vs fast: identical image %s" % bool((f2.image == f.image).all()))
    cfg2.debug_flags = 0; cfg2.samples = 4; cfg2.monte_carlo = 1
    f = rm.start(cam, cfg2); ref = orc.render(cam, cfg2)
    cmp_frames(f, ref, "  mc 4spp")
    rm.close()

fs, cam, cfg = abi.load_fixture("c2_floor_monkey")
rm = RendererManager(cam.width, cam.height, fs)
for i in range(3):
    t = time.time(); f = rm.start(cam, cfg); dt = time.time() - t
    s = f.stats
    print("C2 full: wall %.3fs device %.1f ms trace %.1f ms rays %d+%d => %.1f Mrays/s (device), waves %d launches %d" % (
        dt, s.device_ms, s.closest_ms + s.shadow_ms, s.rays_closest, s.rays_shadow, (s.rays_closest + s.rays_shadow) / s.device_ms / 1e3, s.waves, s.kernel_launches))
from PIL import Image
Image.fromarray(f.image).save("gpurun_out/c2_gpu.png")
