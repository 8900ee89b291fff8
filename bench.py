#!/usr/bin/env python
"""bench.py — Mrays/s (closest-hit + shadow) and ms/frame of the ray-casting hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo, N B200s (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path (oracle port) on the host cores

A step = ONE frame of BASELINE.json configs[1]: scene/floor.json + scene/monkey.json, 1280x720, samples=32,
monte_carlo=1 — the SAME frame at every N (strong scaling): the frame's interleaved 8x4 tiles are dealt to the N ranks,
every rank renders and resolves its tiles and its resolve kernel stores the finished pixels straight into rank 0's
frame buffers over NVLink peer memory (CUDA IPC); a barrier ends the frame.  One ray = one Raytracing::trace call
(closest-hit or shadow), BASELINE.md §2.

`value`     rays of the frame / device time (CUDA events, max over ranks), output buffers resident in HBM.
`e2e`       same metric through the public host API (RendererManager.start at N = 1; render + PeerFrame.download at
            N > 1) with HOST frame buffers: camera / config H2D and the 24 B/pixel G-buffer D2H inside the timed region.
`roofline`  dominant traversal kernel: achieved = algorithmic bytes (counted node visits * 80 B + triangle tests * 48 B
            + sphere tests * 64 B) / summed CUDA-event time of its launches, against the measured HBM
            copy bandwidth; `bounds` adds what north_star defines — the lesser of bytes/ray over bandwidth and flops/ray
            over the FP32 peak — with the L2 read bandwidth measured live for scenes whose BVH is L2-resident.
`workloads` the other configs of BASELINE.json on the same N GPUs, each with ms/frame, Mrays/s, roofline and (N = 1) a
            CPU baseline: c4_standin / c3_standin (labelled stand-ins of scene/sponza.json / scene/helmet.json, whose .glb
            files are not in the reference tree: rustray_b200/synthetic.py) and c5 (the 10 M-triangle soup at 3840x2160x256).
`cpu_baseline` the C++ oracle (port of the reference's path; the Rust reference cannot be built here) on a bounded
            sample of the same frame, all host threads; `value` rebuilds the sample set per pixel like the reference
            does, `hoisted.value` computes it once per frame (SURVEY.md §8(d) asks for both).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mrays/s (closest-hit+shadow)"
UNIT = "Mrays/s"
SCENE = "c2_floor_monkey"
SPP = 32
S_NODE, S_TRI, S_SPH = 80, 48, 64          # bytes per node visit / triangle test / sphere test
F_NODE, F_TRI, F_SPH = 200, 60, 40         # flops per node visit (8 boxes) / triangle test / sphere test
ALL_WORKLOADS = ("c4_standin", "c3_standin", "c5")


def workload_config(n_gpus):
    return {"workload": "configs[1]: scene/floor.json + scene/monkey.json 1280x720 samples=%d monte_carlo=1 (the same frame at every N)" % SPP,
            "width": 1280, "height": 720, "samples": SPP, "monte_carlo": 1, "triangles": 15746, "items": 2, "lights": 4,
            "sharding": "interleaved 8x4 tiles, one per rank in every group of N tiles (rotated per group); resolve kernels store into rank 0's frame buffers over NVLink peer memory",
            "l2": "256 MiB buffer written between timed steps (L2 flush)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def source_sha():
    """Identity of the kernel sources (the frame's kernels, the device functions they inline and the host BVH builder that shapes
    the trees they walk): profiles/*_traffic.json is only trusted when it was captured from the same code."""
    h = hashlib.sha256()
    for f in ("rtx_kernels.cuh", "rtx_device.cuh", "bvh_build.cpp", "bvh_build.h"):
        h.update(open(os.path.join(ROOT, "rustray_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------
def build_workload(name):
    """-> (FlatScene, RtxCamera, RtxConfig, description dict).  Everything is generated / loaded from files of THIS repo."""
    from rustray_b200 import abi, synthetic
    t0 = time.perf_counter()
    if name == "c2":
        fs, cam, cfg = abi.load_fixture(SCENE, samples=SPP, monte_carlo=1)
        desc = workload_config(1)
    elif name == "c4_standin":
        sc = synthetic.atrium_scene(1280, 720, samples=128, monte_carlo=True)
        fs, cam, cfg = abi.FlatScene.from_scene(sc), abi.make_camera(sc.cam), abi.make_config(sc.config)
        desc = {"workload": "configs[3] STAND-IN: synthetic.atrium_scene in place of scene/sponza.json (Sponza_fixed.glb is not in the reference tree) 1280x720 samples=128 monte_carlo=1",
                "width": 1280, "height": 720, "samples": 128, "monte_carlo": 1, "nearest_filtering": True}
    elif name == "c3_standin":
        sc = synthetic.helmet_scene(1280, 720, samples=32, monte_carlo=False)
        fs, cam, cfg = abi.FlatScene.from_scene(sc), abi.make_camera(sc.cam), abi.make_config(sc.config)
        desc = {"workload": "configs[2] STAND-IN: synthetic.helmet_scene in place of scene/helmet.json (DamagedHelmet.glb is not in the reference tree) 1280x720, the file's own config: samples=32 monte_carlo=0",
                "width": 1280, "height": 720, "samples": 32, "monte_carlo": 0}
    elif name == "c5":
        sc = synthetic.soup_scene()
        fs, cam, cfg = abi.FlatScene.from_scene(sc), abi.make_camera(sc.cam), abi.make_config(sc.config)
        desc = {"workload": "configs[4]: synthetic 10M-triangle soup (64 mesh items) + 1000 spheres 3840x2160 samples=256 monte_carlo=1, seed 0x5EED",
                "width": 3840, "height": 2160, "samples": 256, "monte_carlo": 1}
    else:
        raise SystemExit("unknown workload %r" % name)
    desc.update({"triangles": fs.n_triangles, "items": len(fs.items), "lights": len(fs.lights), "textures": len(fs.textures),
                 "scene_generate_s": round(time.perf_counter() - t0, 2)})
    return fs, cam, cfg, desc


def cpu_oracle_sample(fs, cam, cfg, target_s=6.0, faithful=True, cell_step=None):
    """The reference's CPU implementation of the path, as ported in oracle/ (kind = "port"): all host threads, the reference's
    scheduling shape (2x2 cells pulled by worker threads, renderer.rs:17,253-318) and its per-pixel sample-set rebuild
    (raytracing.rs:290-313, `faithful`), on every `cell_step`-th 2x2 cell of the frame.  cell_step is chosen from a short
    calibration pass so that the sample takes about `target_s` seconds."""
    from oracle.oracle import OracleRenderer
    orc = OracleRenderer(fs)
    cores = os.cpu_count() or 1
    n_cells = ((cam.width + 1) // 2) * ((cam.height + 1) // 2)
    if cell_step is None:
        probe_step = max(1, n_cells // 512)
        t0 = time.perf_counter()
        orc.render_ex(cam, cfg, threads=cores, cell_step=probe_step, faithful=faithful)
        dt = max(1e-3, time.perf_counter() - t0)
        cell_step = max(1, int(round(probe_step * dt / target_s)))
    t0 = time.perf_counter()
    f = orc.render_ex(cam, cfg, threads=cores, cell_step=cell_step, faithful=faithful)
    dt = time.perf_counter() - t0
    rays = f.stats.rays_closest + f.stats.rays_shadow
    orc.close()
    sample = "every %d-th 2x2 cell of the %dx%dx%dspp frame (%d rays in %.1f s), %s" % (
        cell_step, cam.width, cam.height, cfg.samples, rays, dt,
        "per-pixel sample-set rebuild as in the reference" if faithful else "sample set hoisted out of the pixel loop")
    return {"value": rays / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_step": dt * 1e3, "rays": rays,
            "cell_step": cell_step}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fs, cam, cfg, _ = build_workload("c2")
    cores = os.cpu_count() or 1
    cell_step = max(1, int(round(128 / max(1, cores))))                  # ~9 Mrays/s on 16 cores for this scene: a few seconds per step
    vals = []
    for i in range(max(0, args.warmup) + args.steps):
        r = cpu_oracle_sample(fs, cam, cfg, cell_step=cell_step)
        if i >= max(0, args.warmup):
            vals.append(r)
    ms = sum(v["ms_per_step"] for v in vals) / len(vals)
    val = sum(v["rays"] for v in vals) / (ms * 1e-3 * len(vals)) / 1e6
    r = vals[-1]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(max(1, args.gpus)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU path of the reference as ported in oracle/rt_oracle.cpp (the Rust reference cannot be compiled in this image: no cargo/rustc)"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
class Harness:
    """One rank of the GPU arm: process group, device, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1")); self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (args.gpus, args.gpus))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allsum(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev); self.dist.all_reduce(t); return float(t.item())

    def allmax(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev); self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX); return float(t.item())


def roofline_block(st0, closest_ms, shadow_ms, tot_ms, steps, waves, rays_per_step, peak, peak_src, l2_gbs, sm_mhz, sm_count, scene_bytes, traffic):
    """Roofline of the dominant traversal kernel + the two bounds of north_star for the traversal as a whole."""
    bytes_closest = st0.node_visits[0] * S_NODE + st0.tri_tests[0] * S_TRI + st0.sphere_tests * S_SPH
    bytes_shadow = st0.node_visits[1] * S_NODE + st0.tri_tests[1] * S_TRI
    flops_closest = st0.node_visits[0] * F_NODE + st0.tri_tests[0] * F_TRI + st0.sphere_tests * F_SPH
    flops_shadow = st0.node_visits[1] * F_NODE + st0.tri_tests[1] * F_TRI
    dom = "closest_kernel" if closest_ms >= shadow_ms else "shadow_any_kernel"
    k_ms = closest_ms if dom == "closest_kernel" else shadow_ms
    k_bytes = bytes_closest if dom == "closest_kernel" else bytes_shadow
    achieved = (k_bytes * steps) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    n_rays = max(1, st0.rays_closest + st0.rays_shadow)
    b_ray = (bytes_closest + bytes_shadow) / n_rays
    f_ray = (flops_closest + flops_shadow) / n_rays
    fp32_peak = sm_count * 128 * 2 * (sm_mhz or 1965.0) * 1e6            # FMA = 2 flops, at the SM clock sampled under load
    l2_resident = scene_bytes < 100e6
    bw = l2_gbs if (l2_resident and l2_gbs) else peak
    bound_bw = bw * 1e9 / max(1.0, b_ray)
    bound_fp32 = fp32_peak / max(1.0, f_ray)
    trav_ms = (closest_ms + shadow_ms) / steps
    trav_rays_per_s = rays_per_step / (trav_ms * 1e-3) if trav_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": k_bytes / max(1.0, waves / steps),
            "avg_launch_ms": k_ms / max(1, waves), "launches_per_step": waves / steps,
            "bytes_per_ray": {"closest": bytes_closest / max(1, st0.rays_closest), "shadow": bytes_shadow / max(1, st0.rays_shadow)},
            "node_visits_per_ray": {"closest": st0.node_visits[0] / max(1, st0.rays_closest), "shadow": st0.node_visits[1] / max(1, st0.rays_shadow)},
            "tri_tests_per_ray": {"closest": st0.tri_tests[0] / max(1, st0.rays_closest), "shadow": st0.tri_tests[1] / max(1, st0.rays_shadow)},
            "kernel_share_of_step": {"closest_kernel": closest_ms / tot_ms, "shadow_any_kernel": shadow_ms / tot_ms},
            "bounds": {"bytes_per_ray": b_ray, "flops_per_ray": f_ray,
                       "bandwidth_GBps": bw, "bandwidth_kind": "L2 read bandwidth measured live (rtx_bandwidth_probe, 32 MiB)" if (l2_resident and l2_gbs) else "HBM copy bandwidth",
                       "l2_read_GBps_measured": l2_gbs, "hbm_GBps": peak, "scene_bvh_bytes": scene_bytes, "l2_resident": l2_resident,
                       "fp32_peak_TFLOPs": fp32_peak / 1e12, "fp32_clock_mhz": sm_mhz,
                       "bound_bw_Grays_per_s": bound_bw / 1e9, "bound_fp32_Grays_per_s": bound_fp32 / 1e9,
                       "binding": "bandwidth" if bound_bw <= bound_fp32 else "fp32",
                       "traversal_Grays_per_s": trav_rays_per_s / 1e9, "frac_of_bound": trav_rays_per_s / min(bound_bw, bound_fp32),
                       "frac_of_hbm_bound": trav_rays_per_s / (peak * 1e9 / max(1.0, b_ray))},
            "note": "achieved = counted node+triangle+sphere bytes of the dominant kernel / its CUDA-event time, against the HBM copy peak as the bench contract asks; `bounds` is north_star's definition (lesser of bytes/ray over bandwidth and flops/ray over FP32 peak) for both traversal kernels together"}


def measure_workload(H, name, steps, warmup, with_cpu, peak, peak_src, l2_gbs, traffic_db, main=False):
    """Render `steps` timed frames of one workload on all ranks (strong scaling) -> result dict (rank 0) or None."""
    torch = H.torch
    from rustray_b200 import abi
    from rustray_b200.distributed import PeerFrame
    from rustray_b200.renderer import RendererManager
    fs, cam, cfg, desc = build_workload(name)
    w, h = cam.width, cam.height
    rm = RendererManager(w, h, fs, device=H.local)
    info = rm.bvh_info()
    shard = abi.RtxShard(H.rank, H.world, 8, 4)
    pf = PeerFrame(rm._lib, w, h, H.rank, H.world, H.local)
    stream = torch.cuda.current_stream(H.dev).cuda_stream
    sr = None
    if pf.ok and os.environ.get("RTX_BENCH_NO_IPC") == "1":            # exercise the fallback on a box that does have peer memory
        pf.close(); pf.ok = False
    if pf.ok:
        ptrs = pf.pointers()
        gather_mode = "resolve kernels store into rank 0's frame buffers over NVLink peer memory (CUDA IPC)"
    else:
        # no peer memory between these GPUs: every rank resolves locally, ONE NCCL gather of the packed shards, one scatter kernel
        from rustray_b200.distributed import ShardedRenderer

        class _NcclFrame:
            def __init__(self, s): self.s = s
            def finish(self): self.s.gather()
            def download(self, frame, stream_ptr=0):
                torch.cuda.synchronize()
                for src, dst in zip((self.s.rgba, self.s.normals, self.s.depth, self.s.ids), (frame.image, frame.normals, frame.depth, frame.objects)):
                    dst.reshape(-1)[:] = src.cpu().numpy().view(dst.dtype).reshape(-1)
            def close(self): pass
        sr = ShardedRenderer(rm, w, h, H.rank, H.world, device=H.dev)
        ptrs = [sr.rgba.data_ptr(), sr.normals.data_ptr(), sr.depth.data_ptr(), sr.ids.data_ptr()]
        pf = _NcclFrame(sr)
        gather_mode = "CUDA IPC unavailable: one NCCL gather of the packed G-buffer + one scatter kernel"

    def render(c):
        st = abi.RtxStats()
        rc = rm._lib.rtx_render_frame_device(rm._h, C.byref(cam), C.byref(c), C.byref(shard), ptrs[0], ptrs[1], ptrs[2], ptrs[3], C.c_void_p(stream), C.byref(st))
        rm._check(rc)
        return st

    # ---- stats frame (untimed): counted traversal work for the roofline ----
    cfg_stats = abi.RtxConfig(); C.memmove(C.byref(cfg_stats), C.byref(cfg), C.sizeof(cfg)); cfg_stats.debug_flags = 1 | 8
    st0 = render(cfg_stats)
    for _ in range(warmup):
        render(cfg); pf.finish()
    clocks = ClockSampler(H.local)
    if H.rank == 0:
        clocks.start()
    H.barrier()
    tot_ms, rays, launches, closest_ms, shadow_ms, waves, syncs, rank_ms = 0.0, 0, 0, 0.0, 0.0, 0, 0, 0.0
    for _ in range(steps):
        H.flush.fill_(1)
        H.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = render(cfg)
        pf.finish()                                                      # barrier: every rank's pixels are in rank 0's buffers
        e1.record()
        torch.cuda.synchronize()
        tot_ms += H.allmax(e0.elapsed_time(e1))
        rank_ms += st.device_ms
        rays += st.rays_closest + st.rays_shadow
        launches += st.kernel_launches
        closest_ms += st.closest_ms; shadow_ms += st.shadow_ms; waves += st.waves; syncs += st.host_syncs
    H.barrier()
    clk = clocks.stop() if H.rank == 0 else None
    # ---- per-kernel durations for the roofline: in the product schedule the shadow kernels of wave k overlap the closest-hit / shade
    # kernels of wave k+1 on a second stream, so their CUDA-event times are not exclusive.  The same frames are therefore run once
    # more with the two streams serialised (RTX_DEBUG_SERIAL_STREAMS); `value` / `e2e` above and below use the overlapped schedule.
    cfg_serial = abi.RtxConfig(); C.memmove(C.byref(cfg_serial), C.byref(cfg), C.sizeof(cfg)); cfg_serial.debug_flags = 8
    overlapped = {"closest_ms_per_frame": closest_ms / steps, "shadow_ms_per_frame": shadow_ms / steps}
    closest_ms, shadow_ms, waves, serial_ms = 0.0, 0.0, 0, 0.0
    n_serial = max(1, min(steps, 5))
    for _ in range(n_serial):
        st = render(cfg_serial); pf.finish()
        closest_ms += st.closest_ms; shadow_ms += st.shadow_ms; waves += st.waves; serial_ms += st.device_ms
    rays_all = H.allsum(rays)
    launches_all = H.allsum(launches)
    slowest_rank_ms = H.allmax(rank_ms / steps)
    ms_per_step = tot_ms / steps
    value = rays_all / (tot_ms * 1e-3) / 1e6

    # ---- end to end through the host API: host frame buffers, H2D params + D2H G-buffer inside the timed region ----
    H.barrier()
    e2e_s, e2e_rays, h2d, d2h = 0.0, 0, 0, 0
    n_e2e = max(1, min(steps, 5))
    for i in range(1 + n_e2e):
        H.barrier()
        t0 = time.perf_counter()
        if H.world == 1:
            f = rm.start(cam, cfg)                                        # public API of a single-GPU host: RendererManager::start
            st = f.stats
        else:
            st = render(cfg)
            pf.finish()
            if H.rank == 0:
                pf.download(rm.frame, stream)
                st.d2h_bytes += w * h * 24
        H.barrier()
        dt = H.allmax(time.perf_counter() - t0)
        if i >= 1:
            e2e_s += dt; e2e_rays += st.rays_closest + st.rays_shadow; h2d = st.h2d_bytes; d2h = st.d2h_bytes
    e2e_rays = H.allsum(e2e_rays)
    e2e_val = e2e_rays / e2e_s / 1e6
    d2h = H.allmax(d2h)

    out = None
    if H.rank == 0:
        prop = torch.cuda.get_device_properties(H.dev)
        dom_traffic = None
        t = traffic_db.get(name)
        if t and traffic_db.get("source_sha") == source_sha():
            dom_k = "closest_kernel" if closest_ms >= shadow_ms else "shadow_any_kernel"
            dom_traffic = t.get("kernels", {}).get(dom_k)
        rf = roofline_block(st0, closest_ms, shadow_ms, serial_ms, n_serial, waves, rays / steps, peak, peak_src, l2_gbs,
                            clk["sm_mhz"] if clk else None, prop.multi_processor_count, info.node_bytes + info.triangle_bytes,
                            (dom_traffic or {}).get("dram_bytes_per_launch"))
        rf["kernel_times"] = {"source": "%d frames with the shadow stream serialised (RTX_DEBUG_SERIAL_STREAMS), run right after the timed region" % n_serial,
                              "serialised_ms_per_frame": serial_ms / n_serial, "overlapped_schedule": overlapped}
        if dom_traffic is None and traffic_db.get(name):
            rf["traffic_note"] = "profiles traffic file was captured from other kernel sources (sha mismatch): not reported"
        out = {"config": desc, "n_gpus": H.world, "gather_mode": gather_mode, "steps": steps, "warmup": warmup, "ms_per_frame": ms_per_step, "value": value, "unit": UNIT,
               "rays_per_frame": rays_all / steps, "slowest_rank_device_ms": slowest_rank_ms,
               "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / n_e2e},
               "gpu_launches": int(launches_all), "host_syncs_per_frame_rank0": syncs / steps, "waves_per_frame_rank0": waves / n_serial,
               "scene_build_ms": info.build_ms, "bvh_bytes": int(info.node_bytes + info.triangle_bytes), "texture_bytes": int(info.texture_bytes),
               "roofline": rf, "clocks": clk}
        if with_cpu and H.world == 1:
            r = cpu_oracle_sample(fs, cam, cfg, target_s=6.0 if not main else 8.0)
            out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
            if main:
                # SURVEY.md 8(d) asks for both variants: the reference rebuilds its shuffled sample set per pixel (faithful, above);
                # "hoisted" computes it once per frame like any sane port would
                rh = cpu_oracle_sample(fs, cam, cfg, faithful=False, cell_step=r["cell_step"])
                out["cpu_baseline"]["hoisted"] = {"value": rh["value"], "unit": UNIT, "sample": rh["sample"]}
            out["speedup_e2e_vs_cpu"] = e2e_val / r["value"]
        if name == "c5" and H.world == 1:
            # the same scene with its BVHs built by kernels (RTX_SCENE_DEVICE_BVH): build time, and what the Morton tree costs per frame
            t0 = time.perf_counter()
            rd = RendererManager(w, h, fs, device=H.local, device_bvh=True)
            wall = (time.perf_counter() - t0) * 1e3
            di = rd.bvh_info()
            st_d = abi.RtxStats()
            rd._check(rd._lib.rtx_render_frame_device(rd._h, C.byref(cam), C.byref(cfg), None, ptrs[0], ptrs[1], ptrs[2], ptrs[3], C.c_void_p(stream), C.byref(st_d)))
            out["device_bvh"] = {"scene_build_ms": di.build_ms, "builder_kernels_ms": di.device_build_ms, "create_wall_ms": wall, "ms_per_frame": st_d.device_ms,
                                 "note": "rtx_scene_create_ex(RTX_SCENE_DEVICE_BVH): Morton-order wide BVH built by kernels; identical traversal results, more node visits per ray than the host's binned-SAH tree (the default)"}
            rd.close()
    extra = None
    if main:
        extra = (rm, render, pf, cam, cfg)
    else:
        pf.close(); rm.close()
    return out, extra


def run_ours(args):
    H = Harness(args)
    torch = H.torch
    from rustray_b200 import abi
    from rustray_b200.renderer import load_library
    lib = load_library()
    peak, peak_src = measured_peaks()
    l2 = C.c_float(0.0)
    l2_gbs = float(l2.value) if lib.rtx_bandwidth_probe(H.local, 32 << 20, 200, C.byref(l2)) == 0 else None
    traffic_db = {}
    for f in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
        if f.endswith("_traffic.json") and f.startswith("r02"):
            try:
                traffic_db = json.load(open(os.path.join(ROOT, "profiles", f)))
            except Exception:
                pass
    names = [n for n in (args.workloads.split(",") if args.workloads else []) if n]
    for n in names:
        if n not in ALL_WORKLOADS:
            raise SystemExit("unknown workload %r (choose from %s)" % (n, ",".join(ALL_WORKLOADS)))

    warm = max(3, args.warmup)
    main, extra = measure_workload(H, "c2", args.steps, warm, not args.no_cpu_baseline, peak, peak_src, l2_gbs, traffic_db, main=True)
    rm, render, pf, cam, cfg = extra

    # ---- informational: opt-in RTX_OPT_SKIP_ZERO_SHADOW (not the headline: the headline traces every reference ray) ----
    cfg_skip = abi.RtxConfig(); C.memmove(C.byref(cfg_skip), C.byref(cfg), C.sizeof(cfg)); cfg_skip.debug_flags = 4
    skip_ms, skip_st = 0.0, None
    for i in range(4):
        H.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); skip_st = render(cfg_skip); pf.finish(); e1.record(); torch.cuda.synchronize()
        if i > 0:
            skip_ms += H.allmax(e0.elapsed_time(e1)) / 3
    # ---- informational: the same frame gathered with ONE NCCL gather + one scatter kernel instead of peer stores ----
    nccl_ms = None
    if H.world > 1:
        from rustray_b200.distributed import ShardedRenderer
        sr = ShardedRenderer(rm, cam.width, cam.height, H.rank, H.world, device=H.dev)
        nccl_ms = 0.0
        for i in range(4):
            H.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sr.render_local(cam, cfg); sr.gather(); e1.record(); torch.cuda.synchronize()
            if i > 0:
                nccl_ms += H.allmax(e0.elapsed_time(e1)) / 3
    pf.close(); rm.close()

    others = {}
    for n in names:
        k = {"c4_standin": 3, "c3_standin": 5, "c5": 2}[n]
        w_ = {"c4_standin": 1, "c3_standin": 2, "c5": 1}[n]
        res, _ = measure_workload(H, n, k, w_, not args.no_cpu_baseline, peak, peak_src, l2_gbs, traffic_db)
        if H.rank == 0:
            others[n] = res
        H.barrier()

    if H.rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": H.world, "steps": args.steps, "warmup": warm,
                "ms_per_step": main["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(H.world), "clocks": main["clocks"],
                "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "roofline": main["roofline"], "rays_per_step": main["rays_per_frame"],
                "ms_per_frame": main["ms_per_frame"], "slowest_rank_device_ms": main["slowest_rank_device_ms"], "gather_mode": main["gather_mode"],
                "host_syncs_per_frame_rank0": main["host_syncs_per_frame_rank0"], "scene_build_ms": main["scene_build_ms"],
                "opt_skip_zero_shadow": {"ms_per_frame": skip_ms, "rays_skipped_rank0": int(skip_st.rays_shadow_skipped),
                                         "note": "opt-in flag, image-identical; informational, not part of value/e2e"},
                "source_sha": source_sha()}
        if nccl_ms is not None:
            line["gather"] = {"peer_store_ms_per_frame": main["ms_per_frame"], "nccl_gather_ms_per_frame": nccl_ms,
                              "note": "same frame: resolve kernels storing into rank 0's buffers over NVLink (product path) vs ONE NCCL gather of the packed G-buffer + one scatter kernel"}
        if "cpu_baseline" in main:
            line["cpu_baseline"] = main["cpu_baseline"]
        if others:
            line["workloads"] = others
        print(json.dumps(line), flush=True)
    if H.world > 1:
        H.dist.barrier()
        H.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workloads", default=",".join(ALL_WORKLOADS),
                    help="other BASELINE.json configs measured into `workloads` (comma separated, empty = none): " + ",".join(ALL_WORKLOADS))
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
