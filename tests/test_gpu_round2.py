"""GPU suite, round 2: the parity holes VERDICT r01 named, the stand-ins of configs[2] / configs[3], sync-free frames,
probes during a frame in flight, the single-process multi-GPU handle and the peer-memory frame buffers.

  * shadow queries go through the PRODUCTION kernels (shadow_any_kernel + shadow_exact_kernel, rtx_shadow_probe) and are compared
    with the oracle's restatement of raytracing.rs:883-914, finite and infinite light distances, on every fixture;
  * the benchmarked frame itself (configs[1], 1280x720x32 Monte Carlo, shared counter RNG) GPU vs oracle, and both against the
    author's own full-resolution rendering of that config;
  * the 40 dB Monte-Carlo gate against an ORACLE high-spp render of a cropped region.
"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np
import pytest

from rustray_b200 import abi, synthetic
from rustray_b200.renderer import RendererManager, RtxError, primary_ray
from oracle.oracle import OracleRenderer
from tests.util import clone_cfg, lsb_stats, psnr, random_rays, scene_to_abi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = ["c1_spheres", "c2_floor_monkey", "room_spheres", "kbert", "monkey_gltf", "kbert_in_room", "earth_in_room"]


def _n_gpus() -> int:
    import torch
    return torch.cuda.device_count()


def _shadow_rays_of_a_frame(fs, cam, c, n=5000, seed=5):
    """Shadow rays exactly as the shading loop builds them (raytracing.rs:872-883): from primary hit points, offset along the
    normal by SHADOW_BIAS, towards every enabled light; light distance for point / spot lights."""
    rng = np.random.default_rng(seed)
    xs, ys = rng.integers(0, cam.width, n), rng.integers(0, cam.height, n)
    rays = [primary_ray(cam, int(x), int(y)) for x, y in zip(xs, ys)]
    o = np.array([r[0] for r in rays]); d = np.array([r[1] for r in rays])
    h = c.trace(o, d)
    ok = h["t"] >= 0
    p = (o[ok] + d[ok] * h["t"][ok, None]).astype(np.float32)
    nn = h["normal"][ok]
    recv = h["item_index"][ok].astype(np.int32)
    so, sd, sl, sr, inf = [], [], [], [], []
    for l in fs.lights:
        if not l.enabled:
            continue
        if l.light_type == 0:
            dirs = np.tile(-np.array(l.dir, dtype=np.float32) / np.linalg.norm(np.array(l.dir, dtype=np.float32)), (p.shape[0], 1))
            ln = np.full(p.shape[0], np.float32(3.402823466e+38))
        else:
            v = np.array(l.pos, dtype=np.float32) - p
            ln = np.linalg.norm(v, axis=1).astype(np.float32)
            dirs = v / ln[:, None]
        so.append((p + nn * np.float32(0.001)).astype(np.float32)); sd.append(dirs.astype(np.float32)); sl.append(ln); sr.append(recv)
    return np.concatenate(so), np.concatenate(sd), np.concatenate(sl), np.concatenate(sr)


def _check_shadow(g, c, fs, o, d, ld, recv, depth=1, min_each=20):
    sg, sc_ = g.shadow_probe(o, d, ld, recv, depth), c.shadow_probe(o, d, ld, recv, depth)
    assert np.array_equal(sg["lit"], sc_["lit"])
    lit = sc_["lit"] == 1
    assert lit.sum() >= min_each and (~lit).sum() >= min_each
    tex_alpha = np.array([fs.materials[it.material].texture[4] >= 0 for it in fs.items])
    at = ~lit & tex_alpha[np.maximum(sc_["occluder_index"], 0)]
    # the factor the kernels apply: bit for bit where it is 1 - receiver alpha; where the occluder's alpha texel enters (:905-909) it is
    # colour arithmetic (1 - alpha * texel may be contracted into one FMA on the device) — and a SPHERE receiver's uv goes through
    # atan2f / acosf, where glibc and CUDA differ in the last ulps, so the texel itself may be a neighbouring blend
    sphere_recv = np.zeros(lit.shape, dtype=bool) if recv is None else np.array([it.shape == 0 for it in fs.items])[np.maximum(recv, 0)]
    assert np.array_equal(sg["k"][~at], sc_["k"][~at], equal_nan=True)
    assert np.allclose(sg["k"][at & ~sphere_recv], sc_["k"][at & ~sphere_recv], rtol=0, atol=1e-5, equal_nan=True)
    assert np.allclose(sg["k"][at & sphere_recv], sc_["k"][at & sphere_recv], rtol=0, atol=2e-3, equal_nan=True)
    assert (sg["occluder_index"][lit] == -1).all() and (sg["occluder_index"][~lit] >= 0).all()
    # alpha-textured occluder: the order rule decides which item attenuates, and its hit point feeds the texture lookup
    assert np.array_equal(sg["occluder_index"][at], sc_["occluder_index"][at]) and np.array_equal(sg["t"][at], sc_["t"][at])
    assert np.array_equal(sg["face_id"][at], sc_["face_id"][at])
    return int(at.sum())


@pytest.mark.parametrize("name", FIXTURES)
def test_shadow_queries_through_the_production_kernels(name):
    fs, cam, cfg = abi.load_fixture(name)
    g, c = RendererManager(cam.width, cam.height, fs), OracleRenderer(fs)
    o, d, ln, recv = _shadow_rays_of_a_frame(fs, cam, c)
    _check_shadow(g, c, fs, o, d, ln, recv)                                                       # the frame's own shadow rays
    _check_shadow(g, c, fs, o, d, None, recv, depth=2, min_each=0)                                # as if every light were directional (a closed room: none lit)
    ro, rd = random_rays(6000, 13)
    rng = np.random.default_rng(1)
    for ld in (None, rng.uniform(1.0, 25.0, ro.shape[0]).astype(np.float32), np.full(ro.shape[0], 0.05, dtype=np.float32)):
        sg, sc_ = g.shadow_probe(ro, rd, ld, None), c.shadow_probe(ro, rd, ld, None)
        assert np.array_equal(sg["lit"], sc_["lit"]) and np.array_equal(sg["k"], sc_["k"], equal_nan=True)


def test_shadow_queries_alpha_textured_occluders_and_many_items():
    """The exact-walk half (alpha texture present -> every occluded ray is re-walked in the reference's order) and the > 50 item TLAS."""
    for kw in (dict(), dict(n_extra_spheres=70)):
        sc = synthetic.feature_scene(224, 144, **kw)
        fs, cam, cfg = scene_to_abi(sc)
        g, c = RendererManager(224, 144, fs), OracleRenderer(fs)
        o, d, ln, recv = _shadow_rays_of_a_frame(fs, cam, c, n=8000)
        n_at = _check_shadow(g, c, fs, o, d, ln, recv)
        ro, rd = random_rays(8000, 17, center=(0, 1, -12), radius=14.0)
        rng = np.random.default_rng(3)
        rr = rng.integers(0, len(fs.items), ro.shape[0]).astype(np.int32)
        rr[np.array([it.shape == 0 for it in fs.items])[rr]] = 0                                   # mesh receivers only: a sphere's uv of a foreign point is NaN
        n_at += _check_shadow(g, c, fs, ro, rd, rng.uniform(3.0, 40.0, ro.shape[0]).astype(np.float32), rr, depth=2)
        n_at += _check_shadow(g, c, fs, ro, rd, None, rr, depth=2, min_each=0)                     # depth 2: the environment sphere occludes every ray
        assert n_at > 50                                                                           # the alpha card did occlude


def test_shadow_probe_edge_cases():
    fs, cam, cfg = abi.load_fixture("c2_floor_monkey")
    g = RendererManager(8, 8, fs)
    assert g.shadow_probe(np.zeros((0, 3)), np.zeros((0, 3))).size == 0
    with pytest.raises(RtxError):
        g.shadow_probe(np.zeros((1, 3)), np.array([[0, 0, -1.0]]), None, np.array([99], dtype=np.int32))
    # > one probe chunk (1 Mi rays): every chunk is answered
    o = np.tile(np.array([[0.0, 5.0, -10.0]], dtype=np.float32), (1_100_000, 1)); d = np.tile(np.array([[0.0, -1.0, 0.0]], dtype=np.float32), (1_100_000, 1))
    s = g.shadow_probe(o, d)
    assert (s["lit"] == 0).all() and (s["k"] == 0.0).all()


# ---- the benchmarked frame ---------------------------------------------------------------------------------------
def test_full_size_config2_frame_against_the_oracle_and_the_authors_rendering():
    """configs[1] as bench.py renders it — 1280x720, 32 spp, Monte Carlo — GPU vs the oracle with the shared counter RNG, and
    both against the author's own rendering of that scene (a real output of the reference, made with thread_rng)."""
    from PIL import Image
    fs, cam, cfg = abi.load_fixture("c2_floor_monkey", samples=32, monte_carlo=1)
    g, c = RendererManager(cam.width, cam.height, fs), OracleRenderer(fs)
    fg = g.start(cam, cfg)
    fc = c.render_ex(cam, cfg)
    within1, exact, mx = lsb_stats(fg.image, fc.image)
    assert within1 >= 0.995, (within1, exact, mx)
    assert np.array_equal(fg.objects, fc.objects) and np.array_equal(fg.depth, fc.depth)
    assert (fg.stats.rays_closest, fg.stats.rays_shadow) == (fc.stats.rays_closest, fc.stats.rays_shadow)
    assert fg.stats.primary_samples == 1280 * 720 * 32
    ref = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "ref_render_c2_1280x720.png")).convert("RGB"))
    pg, pc = psnr(fg.image[..., :3], ref), psnr(fc.image[..., :3], ref)
    print("PSNR vs the author's 1280x720 rendering: GPU %.2f dB, oracle %.2f dB" % (pg, pc))
    assert pg >= 28.0 and pc >= 28.0 and abs(pg - pc) < 0.1


def _crop_camera(cam, x0, y0, w, h):
    """Camera whose w x h frame is the [x0, x0+w) x [y0, y0+h) window of `cam`'s frame: P^-1' = P^-1 . A with A the affine map from
    the window's normalised device coordinates to the full frame's (column-major)."""
    W, H = cam.width, cam.height
    a = np.eye(4, dtype=np.float64)
    a[0, 0] = w / W; a[0, 3] = (2.0 * x0 + w) / W - 1.0
    a[1, 1] = h / H; a[1, 3] = 1.0 - (2.0 * y0 + h) / H
    pinv = np.array(cam.projection_inverse, dtype=np.float64).reshape(4, 4).T
    out = abi.RtxCamera()
    C.memmove(C.byref(out), C.byref(cam), C.sizeof(abi.RtxCamera))
    m = (pinv @ a).astype(np.float32)
    for k, v in enumerate(m.T.reshape(16)):
        out.projection_inverse[k] = float(v)
    out.width, out.height = w, h
    return out


def test_monte_carlo_psnr_gate_against_an_oracle_high_spp_crop():
    """north_star: Monte-Carlo renders reach PSNR >= 40 dB against a high-spp reference render.  The reference here is the
    ORACLE at 1024 spp (its own seed) on a 160x96 window of config 2 that holds the monkey's chin, its soft shadow on the
    checkerboard and the refractions; the GPU renders the same window at 256 spp with another seed."""
    fs, cam, cfg = abi.load_fixture("c2_floor_monkey", monte_carlo=1)
    crop = _crop_camera(cam, 520, 400, 160, 96)
    g, c = RendererManager(160, 96, fs), OracleRenderer(fs)
    hi = c.render_ex(crop, clone_cfg(cfg, samples=1024, mc_seed=1234))
    assert (hi.objects != 0).mean() > 0.3 and len(np.unique(hi.objects)) >= 2
    p256 = psnr(g.start(crop, clone_cfg(cfg, samples=256, mc_seed=7)).image[..., :3], hi.image[..., :3])
    p32 = psnr(g.start(crop, clone_cfg(cfg, samples=32, mc_seed=7)).image[..., :3], hi.image[..., :3])
    print("PSNR vs oracle 1024 spp: GPU 256 spp %.2f dB, GPU 32 spp %.2f dB" % (p256, p32))
    assert p256 >= 40.0 and p32 >= 34.0 and p256 > p32
    # the crop camera is the same camera: its deterministic frame equals the window of the full deterministic frame
    det = clone_cfg(cfg, samples=1, monte_carlo=0)
    full = RendererManager(cam.width, cam.height, fs).start(cam, det)
    win = g.start(crop, det)
    assert lsb_stats(win.image, full.image[400:496, 520:680])[0] >= 0.99 and (win.objects == full.objects[400:496, 520:680]).mean() >= 0.995


# ---- stand-ins of configs[2] and configs[3] -------------------------------------------------------------------------
def _standin_parity(sc, mc_cfg, w, h, uv_exact=True):
    fs, cam, cfg = scene_to_abi(sc, samples=1, monte_carlo=0)
    g, c = RendererManager(w, h, fs), OracleRenderer(fs)
    rng = np.random.default_rng(9)
    xs, ys = rng.integers(0, w, 5000), rng.integers(0, h, 5000)
    rays = [primary_ray(cam, int(x), int(y)) for x, y in zip(xs, ys)]
    o = np.array([r[0] for r in rays]); d = np.array([r[1] for r in rays])
    for depth in (1, 2):
        hg, hc = g.trace(o, d, depth=depth), c.trace(o, d, depth=depth)
        assert (hc["t"] >= 0).sum() > 1000 and hg.tobytes()[:0] == b""
        assert np.array_equal(hg["item_index"], hc["item_index"]) and np.array_equal(hg["face_id"], hc["face_id"]) and np.array_equal(hg["t"], hc["t"])
        hit = hc["t"] >= 0
        assert np.allclose(hg["normal"][hit], hc["normal"][hit], rtol=1e-4, atol=1e-6)
    so, sd, sl, sr = _shadow_rays_of_a_frame(fs, cam, c, n=4000)
    _check_shadow(g, c, fs, so, sd, sl, sr)
    fg, fc = g.start(cam, cfg), c.render(cam, cfg)                                                # deterministic, 1 spp
    within1, exact, mx = lsb_stats(fg.image, fc.image)
    assert within1 >= 0.999, (within1, exact, mx)
    assert np.array_equal(fg.objects, fc.objects) and np.array_equal(fg.depth, fc.depth)
    assert (fg.stats.rays_closest, fg.stats.rays_shadow) == (fc.stats.rays_closest, fc.stats.rays_shadow)
    fg, fc = g.start(cam, mc_cfg(cfg)), c.render(cam, mc_cfg(cfg))                                # the config's own sampling, shared RNG
    within1, exact, mx = lsb_stats(fg.image, fc.image)
    assert within1 >= 0.995, (within1, exact, mx)
    assert np.array_equal(fg.objects, fc.objects)
    for a, b in ((fg.stats.rays_closest, fc.stats.rays_closest), (fg.stats.rays_shadow, fc.stats.rays_shadow)):
        assert abs(int(a) - int(b)) <= 1e-4 * b + 1
    return fg


def test_config4_standin_parity():
    """Atrium (configs[3] stand-in) at test size: > 100 textured mesh items (TLAS path), nearest filtering, normal / roughness /
    metallic maps, alpha cut-outs, environment sphere; 8 spp Monte Carlo as the config runs it."""
    sc = synthetic.atrium_scene(320, 180, detail=0.04, tex_size=64)
    fg = _standin_parity(sc, lambda cfg: clone_cfg(cfg, samples=8, monte_carlo=1), 320, 180)
    assert fg.stats.rays_closest > 320 * 180 * 8                                                  # secondary rays exist


def test_config4_standin_full_size_probes():
    sc = synthetic.atrium_scene(1280, 720, detail=1.0, tex_size=32)
    fs, cam, cfg = scene_to_abi(sc)
    assert fs.n_triangles > 240_000 and len(fs.items) > 100
    g, c = RendererManager(1280, 720, fs), OracleRenderer(fs)
    rng = np.random.default_rng(4)
    xs, ys = rng.integers(0, 1280, 20000), rng.integers(0, 720, 20000)
    rays = [primary_ray(cam, int(x), int(y)) for x, y in zip(xs, ys)]
    o = np.array([r[0] for r in rays]); d = np.array([r[1] for r in rays])
    ro, rd = random_rays(20000, 6, center=(0, 4, 0), radius=9.0)
    for oo, dd, depth in ((o, d, 1), (ro, rd, 2)):
        hg, hc = g.trace(oo, dd, depth=depth), c.trace(oo, dd, depth=depth)
        assert (hc["t"] >= 0).sum() > 5000
        assert np.array_equal(hg["item_index"], hc["item_index"]) and np.array_equal(hg["face_id"], hc["face_id"]) and np.array_equal(hg["t"], hc["t"])
    so, sd, sl, sr = _shadow_rays_of_a_frame(fs, cam, c, n=20000)
    _check_shadow(g, c, fs, so, sd, sl, sr)


def test_config3_standin_parity():
    """Helmet (configs[2] stand-in) at test size: one glTF-style primitive with base / normal / metallic / roughness / occlusion /
    emissive maps and bilinear filtering, transformed item, 4 spp deterministic as the file's config block runs it."""
    sc = synthetic.helmet_scene(320, 180, detail=0.08, tex_size=128)
    _standin_parity(sc, lambda cfg: clone_cfg(cfg, samples=4, monte_carlo=0), 320, 180)
    # "as stated" in BASELINE.json (monte_carlo=1 forced over the file): roughness-map jitter on every hit
    fs, cam, cfg = scene_to_abi(sc, samples=4, monte_carlo=1)
    g, c = RendererManager(320, 180, fs), OracleRenderer(fs)
    fg, fc = g.start(cam, cfg), c.render(cam, cfg)
    assert lsb_stats(fg.image, fc.image)[0] >= 0.99 and np.array_equal(fg.objects, fc.objects)


def test_generated_gltf_with_all_texture_kinds():
    """tests/golden/textured_pbr.glb through gltf_loader -> FlatScene -> GPU vs oracle (scene.rs:980-1124)."""
    from rustray_b200.scene_loader import load_scene
    sc = load_scene([os.path.join("tests", "golden", "textured_pbr.glb")], 300, 200, asset_root=ROOT)
    for mc, spp in ((0, 1), (1, 4)):
        fs, cam, cfg = scene_to_abi(sc, samples=spp, monte_carlo=mc)
        g, c = RendererManager(300, 200, fs), OracleRenderer(fs)
        fg, fc = g.start(cam, cfg), c.render(cam, cfg)
        within1, exact, mx = lsb_stats(fg.image, fc.image)
        assert within1 >= (0.999 if not mc else 0.99), (within1, exact, mx)
        assert np.array_equal(fg.objects, fc.objects) and (fc.objects != 0).mean() > 0.1
        if not mc:
            assert np.array_equal(fg.depth, fc.depth)
            assert (fg.stats.rays_closest, fg.stats.rays_shadow) == (fc.stats.rays_closest, fc.stats.rays_shadow)


# ---- scheduling: sync-free frames ---------------------------------------------------------------------------------------
def test_sync_free_frame_equals_the_synchronised_schedule():
    """A frame whose primary rays fit one wave is enqueued without a single counter read-back; it must be the same frame."""
    for name, w, h, spp, mc in (("c1_spheres", 800, 600, 1, 0), ("room_spheres", 320, 180, 16, 1), ("kbert_in_room", 400, 225, 4, 1)):
        fs, cam, cfg = abi.load_fixture(name, samples=spp, monte_carlo=mc)
        cam = abi.resize_camera(cam, w, h)
        g = RendererManager(w, h, fs)
        a = g.start(cam, cfg)
        # (the room of mirrors doubles its rays per level: a level may outgrow one wave, then the frame was redone synchronised)
        assert a.stats.host_syncs == 1 or name == "room_spheres"
        img, ids, depth, st = a.image.copy(), a.objects.copy(), a.depth.copy(), (a.stats.rays_closest, a.stats.rays_shadow)
        os.environ["RTX_FORCE_SYNC"] = "1"
        try:
            b = g.start(cam, cfg)
        finally:
            del os.environ["RTX_FORCE_SYNC"]
        assert b.stats.host_syncs == b.stats.waves + 1 and b.stats.waves >= 1
        assert np.array_equal(ids, b.objects) and np.array_equal(depth, b.depth)
        assert lsb_stats(img, b.image)[0] >= 0.9999                                              # float atomics: summation order
        assert st == (b.stats.rays_closest, b.stats.rays_shadow)


def test_sync_free_overflow_falls_back_to_the_synchronised_schedule():
    """A hall of mirrors: every hit spawns two children, level sizes double, a level outgrows one wave and the frame is redone."""
    sc = synthetic.feature_scene(96, 64)
    for m in sc.materials:
        m.reflectivity, m.alpha, m.refraction_index = 0.5, 0.5, 1.0
    fs, cam, cfg = scene_to_abi(sc, samples=8, monte_carlo=0, max_recursion=12)
    os.environ["RTX_CHUNK"] = "65536"
    try:
        g, c = RendererManager(96, 64, fs), OracleRenderer(fs)
        fg = g.start(cam, cfg)
    finally:
        del os.environ["RTX_CHUNK"]
    fc = c.render(cam, cfg)
    assert fg.stats.rays_closest > 2 * 65536 * 8 and fg.stats.host_syncs > 1                       # a level passed the wave capacity
    assert np.array_equal(fg.objects, fc.objects) and lsb_stats(fg.image, fc.image)[0] >= 0.99
    for a, b in ((fg.stats.rays_closest, fc.stats.rays_closest), (fg.stats.rays_shadow, fc.stats.rays_shadow)):
        assert abs(int(a) - int(b)) <= 1e-4 * b + 1


# ---- a frame in flight (probes and updates while rtx_render_frame_async runs) ------------------------------------
def test_pick_during_an_async_frame_and_busy_updates():
    fs, cam, cfg = abi.load_fixture("c2_floor_monkey", samples=64, monte_carlo=1)
    g = RendererManager(cam.width, cam.height, fs)
    ref = g.start(cam, cfg)
    img, ids = ref.image.copy(), ref.objects.copy()
    idle_pick = g.pick(cam, 640, 300)
    g.start_async(cam, cfg)
    picks, busy = 0, 0
    while g.is_running():
        assert g.pick(cam, 640, 300) == idle_pick                                                 # own stream and buffers: allowed, same answer
        picks += 1
        try:
            g.update_items([(0, np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32))])
        except RtxError as e:
            assert "-7" in str(e) or "in flight" in str(e)
            busy += 1
    deadline = time.time() + 60
    while not g.is_done() and time.time() < deadline:
        time.sleep(0.01)
    assert g.is_done() and picks >= 1 and busy >= 1
    assert np.array_equal(g.frame.objects, ids) and lsb_stats(g.frame.image, img)[0] >= 0.9999


# ---- several GPUs --------------------------------------------------------------------------------------------------------
@pytest.mark.skipif("_n_gpus() < 2")
def test_one_process_two_gpus_equals_one_gpu():
    """rtx_scene_create_multi: scene copied device-to-device, interleaved tiles, resolve kernels of device 1 store into device 0's
    frame buffers through peer memory."""
    for name, w, h, spp, mc in (("c2_floor_monkey", 640, 360, 4, 1), ("kbert_in_room", 400, 225, 1, 0)):
        fs, cam, cfg = abi.load_fixture(name, samples=spp, monte_carlo=mc)
        cam = abi.resize_camera(cam, w, h)
        one, two = RendererManager(w, h, fs, device=0), RendererManager(w, h, fs, devices=[0, 1])
        a, b = one.start(cam, cfg), two.start(cam, cfg)
        assert np.array_equal(a.objects, b.objects) and np.array_equal(a.depth, b.depth)
        assert lsb_stats(a.image, b.image)[0] >= 0.9999
        assert (a.stats.rays_closest, a.stats.rays_shadow, a.stats.primary_samples) == (b.stats.rays_closest, b.stats.rays_shadow, b.stats.primary_samples)
        hit = a.objects != 0
        assert np.array_equal(a.normals[hit], b.normals[hit])
        two.update_items([(1, np.array(fs.items[1].trans).reshape(4, 4).T, np.array(fs.items[1].tran_inverse).reshape(4, 4).T)])   # replicas follow
        c = two.start(cam, cfg)
        assert np.array_equal(c.objects, a.objects)
        assert two.pick(cam, w // 2, h // 2) == one.pick(cam, w // 2, h // 2)


@pytest.mark.skipif("_n_gpus() < 2")
def test_one_process_two_gpus_with_a_merged_blas():
    """The replicas of a multi-GPU handle carry the merged world-space BLAS and the fast TLAS too."""
    sc = synthetic.atrium_scene(320, 180, detail=0.04, tex_size=32, samples=2, monte_carlo=True)
    fs, cam, cfg = scene_to_abi(sc)
    one, two = RendererManager(320, 180, fs, device=0), RendererManager(320, 180, fs, devices=[0, 1])
    assert one.bvh_info().grouped_items > 100
    a, b = one.start(cam, cfg), two.start(cam, cfg)
    assert np.array_equal(a.objects, b.objects) and np.array_equal(a.depth, b.depth) and lsb_stats(a.image, b.image)[0] >= 0.9999
    assert (a.stats.rays_closest, a.stats.rays_shadow) == (b.stats.rays_closest, b.stats.rays_shadow)


def test_moving_an_item_of_the_merged_blas_dissolves_the_group():
    """rtx_scene_update_items on a grouped item: world space is no longer its object space, the group is dissolved and the
    frame must still be the reference's (Scene::apply_frame + update, scene.rs:1695-1713)."""
    from rustray_b200.scene_loader import mat_inverse, mat_translation
    sc = synthetic.atrium_scene(240, 135, detail=0.04, tex_size=32, samples=1, monte_carlo=False)
    fs, cam, cfg = scene_to_abi(sc)
    g, c = RendererManager(240, 135, fs), OracleRenderer(fs)
    assert g.bvh_info().grouped_items > 100
    names = fs.item_names
    k = names.index("plant2")
    t = mat_translation(0.0, 0.8, 0.5)
    for r in (g, c):
        r.update_items([(k, t, mat_inverse(t))])
    assert g.bvh_info().grouped_items == 0
    fg, fc = g.start(cam, cfg), c.render(cam, cfg)
    assert lsb_stats(fg.image, fc.image)[0] >= 0.999 and np.array_equal(fg.objects, fc.objects) and np.array_equal(fg.depth, fc.depth)
    assert (fg.stats.rays_closest, fg.stats.rays_shadow) == (fc.stats.rays_closest, fc.stats.rays_shadow)
    o, d = random_rays(3000, 8, center=(0, 4, 0), radius=9.0)
    sg, sc_ = g.shadow_probe(o, d, 12.0), c.shadow_probe(o, d, 12.0)
    assert np.array_equal(sg["lit"], sc_["lit"])


def test_one_process_multi_handle_rejects_bad_device_lists():
    fs, cam, cfg = abi.load_fixture("c1_spheres")
    with pytest.raises(RtxError):
        RendererManager(64, 64, fs, devices=[0, 0])
    with pytest.raises(RtxError):
        RendererManager(64, 64, fs, devices=[0, 99])


def test_peer_frame_single_rank_round_trip():
    """rtx_gbuffer_*: the owner's 24 B/pixel allocation, rendered into and downloaded (world = 1: no IPC needed)."""
    from rustray_b200.distributed import PeerFrame
    from rustray_b200.renderer import Frame
    fs, cam, cfg = abi.load_fixture("c1_spheres", samples=1, monte_carlo=0)
    cam = abi.resize_camera(cam, 320, 240)
    g = RendererManager(320, 240, fs)
    ref = g.start(cam, cfg)
    pf = PeerFrame(g._lib, 320, 240, 0, 1, 0)
    p = pf.pointers()
    st = abi.RtxStats()
    g._check(g._lib.rtx_render_frame_device(g._h, C.byref(cam), C.byref(cfg), None, p[0], p[1], p[2], p[3], None, C.byref(st)))
    out = Frame(320, 240)
    pf.download(out)
    assert np.array_equal(out.image, ref.image) and np.array_equal(out.objects, ref.objects) and np.array_equal(out.depth, ref.depth)
    pf.close()


@pytest.mark.skipif("_n_gpus() < 2")
def test_two_ranks_render_into_rank0_over_cuda_ipc():
    """One process per GPU (torchrun), rank 1's resolve kernel stores into rank 0's frame buffers; also the NCCL gather path."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "peer_frame_check.py")], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "PEER_FRAME_OK" in r.stdout and "NCCL_GATHER_OK" in r.stdout


# ---- memory appetite -----------------------------------------------------------------------------
def test_small_frames_allocate_small_queues_and_two_scenes_share_a_device():
    """Queues are sized by the frame (a 800x600x1 frame must not reserve gigabytes), and two scene handles on one device render
    interleaved without disturbing each other."""
    import torch
    fs1, cam1, cfg1 = abi.load_fixture("c1_spheres", samples=1, monte_carlo=0)
    fs2, cam2, cfg2 = abi.load_fixture("kbert", samples=2, monte_carlo=1)
    cam2 = abi.resize_camera(cam2, 480, 270)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info(0)[0]
    a = RendererManager(800, 600, fs1)
    fa = a.start(cam1, cfg1)
    used = free0 - torch.cuda.mem_get_info(0)[0]
    assert used < 700 << 20, "config 1 (0.48 M primary rays) took %d MiB of device memory" % (used >> 20)
    ia, ida = fa.image.copy(), fa.objects.copy()
    b = RendererManager(480, 270, fs2)
    fb = b.start(cam2, cfg2)
    ib, idb = fb.image.copy(), fb.objects.copy()
    for _ in range(2):                                                                           # interleave frames of the two handles
        fa = a.start(cam1, cfg1); fb = b.start(cam2, cfg2)
        assert np.array_equal(fa.objects, ida) and lsb_stats(fa.image, ia)[0] >= 0.9999
        assert np.array_equal(fb.objects, idb) and lsb_stats(fb.image, ib)[0] >= 0.9999
    # a bigger frame on the same handle grows the queues; a smaller one afterwards still works
    big = a.render(abi.resize_camera(cam1, 1600, 1200), clone_cfg(cfg1, samples=4))               # (render() allocates a frame of the camera's size)
    assert big.stats.primary_samples == 1600 * 1200 * 4
    small = a.render(abi.resize_camera(cam1, 200, 150), cfg1)
    assert small.stats.primary_samples == 200 * 150 and small.stats.host_syncs == 1
    a.close(); b.close()


# ---- BVH built on the device -------------------------------------------------------------
def _same_answers(fs, cam, cfg, w, h, rays):
    host, dev = RendererManager(w, h, fs), RendererManager(w, h, fs, device_bvh=True)
    assert dev.bvh_info().device_build_ms > 0.0 and host.bvh_info().device_build_ms == 0.0
    o, d = rays
    for kw in (dict(), dict(depth=2), dict(for_shadow=True), dict(for_shadow=True, stop_on_first_hit=True)):
        assert host.trace(o, d, **kw).tobytes() == dev.trace(o, d, **kw).tobytes()              # the tree only prunes: identical hits, bit for bit
    sa, sb = host.shadow_probe(o, d, 25.0), dev.shadow_probe(o, d, 25.0)
    assert np.array_equal(sa["lit"], sb["lit"]) and np.array_equal(sa["k"], sb["k"], equal_nan=True)
    fa, fb = host.start(cam, cfg), dev.start(cam, cfg)
    assert np.array_equal(fa.objects, fb.objects) and np.array_equal(fa.depth, fb.depth) and lsb_stats(fa.image, fb.image)[0] >= 0.9999
    assert (fa.stats.rays_closest, fa.stats.rays_shadow) == (fb.stats.rays_closest, fb.stats.rays_shadow)
    return host, dev


def test_device_built_bvh_gives_identical_results():
    sc = synthetic.soup_scene(200_000, 150, cells=3, width=256, height=144)
    fs, cam, cfg = scene_to_abi(sc, samples=1, monte_carlo=0)
    host, dev = _same_answers(fs, cam, cfg, 256, 144, random_rays(6000, 9, center=(0, 0, 0), radius=90.0))
    assert dev.bvh_info().grouped_triangles == 200_000                                            # the merged BLAS was built on the device too
    oc = OracleRenderer(fs)
    o, d = random_rays(3000, 10, center=(0, 0, 0), radius=90.0)
    hg, hc = dev.trace(o, d), oc.trace(o, d)
    assert np.array_equal(hg["item_index"], hc["item_index"]) and np.array_equal(hg["t"], hc["t"]) and np.array_equal(hg["face_id"], hc["face_id"])
    fs, cam, cfg = abi.load_fixture("c2_floor_monkey", samples=2, monte_carlo=1)                 # transformed instance + a 2-triangle mesh
    cam = abi.resize_camera(cam, 320, 180)
    _same_answers(fs, cam, cfg, 320, 180, random_rays(4000, 3))
    sc = synthetic.atrium_scene(160, 90, detail=0.03, tex_size=32, samples=1, monte_carlo=False)  # > 100 meshes of a few hundred triangles each
    fs, cam, cfg = scene_to_abi(sc)
    _same_answers(fs, cam, cfg, 160, 90, random_rays(4000, 6, center=(0, 4, 0), radius=9.0))


def test_device_built_bvh_degenerate_meshes():
    """1 / 2 / 3 / 4 / 9 triangles, coincident centroids (identical Morton codes), a flat mesh (zero extent on one axis)."""
    from rustray_b200.scene_loader import Scene, Item, Material, Light, Camera, Config, SHAPE_MESH, LIGHT_POINT, mat_identity, mat_translation, to_radians
    sc = Scene(".")
    rng = np.random.default_rng(5)

    def add(verts, idx, pos):
        m = Material(id=sc.get_next_id(), name="m")
        sc.materials.append(m)
        sc.items.append(Item(id=sc.get_next_id(), name="it%d" % len(sc.items), shape=SHAPE_MESH, material=m, trans=mat_translation(*pos),
                             mesh=synthetic._mesh(np.asarray(verts, dtype=np.float32), np.asarray(idx, dtype=np.uint32))))
    tri = [[-1, -1, 0], [1, -1, 0], [0, 1, 0]]
    for k, n in enumerate((1, 2, 3, 4, 9)):                                                        # n copies of one triangle: coincident centroids
        v = np.concatenate([np.asarray(tri, dtype=np.float32) + np.float32([0, 0, 0.0]) for _ in range(n)])
        add(v, np.arange(3 * n).reshape(-1, 3), (-8 + 4 * k, 0, -12))
    g = 5                                                                                          # flat 5x5 grid (zero extent in y)
    u, w_ = np.meshgrid(np.linspace(-2, 2, g + 1), np.linspace(-2, 2, g + 1), indexing="ij")
    gv = np.stack([u, np.zeros_like(u), w_], -1).reshape(-1, 3)
    gi = [[i * (g + 1) + j, (i + 1) * (g + 1) + j, (i + 1) * (g + 1) + j + 1] for i in range(g) for j in range(g)] + \
         [[i * (g + 1) + j, (i + 1) * (g + 1) + j + 1, i * (g + 1) + j + 1] for i in range(g) for j in range(g)]
    add(gv, gi, (0, -2.5, -12))
    rv = rng.uniform(-1, 1, (300, 3)); add(rv, np.arange(300).reshape(-1, 3), (0, 4, -14))      # 100 random triangles
    sc.lights.append(Light(sc.get_next_id(), "l", np.float32([0, 10, -4]), np.float32([0, -1, 0]), np.float32([1, 1, 1]), 300.0, 1.5, LIGHT_POINT))
    sc.cam = Camera(); sc.cam.fov = to_radians(70.0); sc.cam.init(192, 128)
    sc.config = Config()
    fs, cam, cfg = scene_to_abi(sc)
    _same_answers(fs, cam, cfg, 192, 128, random_rays(5000, 2, center=(0, 0, -12), radius=14.0))
