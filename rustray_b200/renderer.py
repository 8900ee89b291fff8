"""Host-side mirror of the reference's render interface, on top of the C ABI.

The reference drives its hot path through `RendererManager` (reference src/renderer.rs:
`new(width, height, raytracing)`, `start()`, `is_done()`, `get_message_receiver()`) and picks with
`Raytracing::pick(x, y)` (src/raytracing.rs:237-273); `Run::apply_pixels` scatters the per-pixel
messages into four frame buffers (src/run.rs:506-545).  On the GPU one call renders the frame, so
`RendererManager.start()` here blocks and leaves the same four buffers in `image`, `normals`,
`depth`, `objects`.

The CUDA library is mandatory: there is no CPU fallback and no import of anything under oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import abi

_LIB_NAME = "librtx_b200.so"
_lib: Optional[C.CDLL] = None


class RtxError(RuntimeError):
    pass


def lib_path() -> str:
    """rustray_b200/librtx_b200.so; RTX_LIB overrides it (tuning variants built by tools/build_variants.sh)."""
    return os.environ.get("RTX_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library() -> C.CDLL:
    """Load rustray_b200/librtx_b200.so (built by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise RtxError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA library is the product; there is no CPU fallback)" % p)
        _lib = C.CDLL(p)
        abi.bind(_lib, "rtx_")
        if _lib.rtx_abi_version() != 2:
            raise RtxError("ABI version mismatch")
    return _lib


class Frame:
    """The four buffers of Run (src/run.rs:117-120)."""

    def __init__(self, width: int, height: int, pinned: bool = False):
        """`pinned`: allocate the buffers in page-locked host memory (through torch) so that the library's device-to-host
        copies run at full PCIe/C2C rate; falls back to ordinary memory when CUDA is not available."""
        self.width, self.height = width, height
        self._pinned = None
        if pinned:
            try:
                import torch
                if torch.cuda.is_available():
                    self._pinned = [torch.zeros(shape, dtype=dt, pin_memory=True) for shape, dt in
                                    (((height, width, 4), torch.uint8), ((height, width, 3), torch.float32),
                                     ((height, width), torch.float32), ((height, width), torch.int32))]
            except Exception:
                self._pinned = None
        if self._pinned is not None:
            self.image, self.normals, self.depth = (t.numpy() for t in self._pinned[:3])
            self.objects = self._pinned[3].numpy().view(np.uint32)
        else:
            self.image = np.zeros((height, width, 4), dtype=np.uint8)
            self.normals = np.zeros((height, width, 3), dtype=np.float32)
            self.depth = np.zeros((height, width), dtype=np.float32)
            self.objects = np.zeros((height, width), dtype=np.uint32)
        self.stats = abi.RtxStats()


class AbiRenderer:
    """Thin object wrapper over any library exporting the rtx ABI under `prefix`."""

    def __init__(self, lib: C.CDLL, prefix: str, flat_scene: abi.FlatScene, device: int = 0, devices=None, device_bvh: bool = False):
        """`devices`: list of CUDA ordinals -> one handle that renders every frame on all of them (rtx_scene_create_multi:
        scene replicated device-to-device, interleaved tiles, peer-memory stores into the first device's frame buffers)."""
        self._lib, self._p = lib, prefix
        self.flat = flat_scene
        self.devices = [int(d) for d in devices] if devices else [int(device)]
        self.device = self.devices[0]
        self._h = C.c_void_p()
        desc = flat_scene.desc()
        if len(self.devices) > 1:
            arr = (C.c_int * len(self.devices))(*self.devices)
            rc = self._fn("scene_create_multi")(C.byref(desc), arr, len(self.devices), C.byref(self._h))
        elif device_bvh:
            rc = self._fn("scene_create_ex")(C.byref(desc), self.device, 1, C.byref(self._h))    # RTX_SCENE_DEVICE_BVH
        else:
            rc = self._fn("scene_create")(C.byref(desc), self.device, C.byref(self._h))
        self._check(rc)

    def _fn(self, name):
        return getattr(self._lib, self._p + name)

    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self._fn("last_error")()
            raise RtxError("%s error %d: %s" % (self._p, rc, msg.decode() if msg else ""))

    def close(self) -> None:
        if self._h:
            self._fn("scene_destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- frame ----------------------------------------------------------------------------
    def render(self, cam: abi.RtxCamera, cfg: abi.RtxConfig, frame: Optional[Frame] = None) -> Frame:
        f = frame or Frame(cam.width, cam.height)
        rc = self._fn("render_frame")(self._h, C.byref(cam), C.byref(cfg), f.image.ctypes.data, f.normals.ctypes.data,
                                      f.depth.ctypes.data, f.objects.ctypes.data, C.byref(f.stats))
        self._check(rc)
        return f

    # -- Raytracing::trace / pick -------------------------------------------------------
    def trace(self, origins, dirs, for_shadow=False, stop_on_first_hit=False, depth=1) -> np.ndarray:
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        rays = np.zeros(o.shape[0], dtype=abi.RAY_DTYPE)
        rays["origin"], rays["dir"] = o, d
        hits = np.zeros(o.shape[0], dtype=abi.HIT_DTYPE)
        if o.shape[0]:
            rc = self._fn("trace_probe")(self._h, rays.ctypes.data, o.shape[0], int(for_shadow), int(stop_on_first_hit),
                                         int(depth), hits.ctypes.data)
            self._check(rc)
        return hits

    def shadow_probe(self, origins, dirs, light_distance=None, receiver_item=None, depth=1) -> np.ndarray:
        """Shadow rays through the production shadow kernels (rtx_shadow_probe): -> abi.SHADOW_HIT_DTYPE records."""
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        rays = np.zeros(n, dtype=abi.RAY_DTYPE)
        rays["origin"], rays["dir"] = o, d
        out = np.zeros(n, dtype=abi.SHADOW_HIT_DTYPE)
        ld = None if light_distance is None else np.ascontiguousarray(np.broadcast_to(np.asarray(light_distance, dtype=np.float32), (n,)))
        rv = None if receiver_item is None else np.ascontiguousarray(np.broadcast_to(np.asarray(receiver_item, dtype=np.int32), (n,)))
        if n:
            rc = self._fn("shadow_probe")(self._h, rays.ctypes.data, ld.ctypes.data if ld is not None else None,
                                          rv.ctypes.data if rv is not None else None, n, int(depth), out.ctypes.data)
            self._check(rc)
        return out

    def update_items(self, updates) -> None:
        """updates: iterable of (item_index, trans 4x4 (row, col), tran_inverse 4x4)."""
        arr = [abi.RtxItemXform(int(i), abi.colmajor16(t), abi.colmajor16(ti)) for i, t, ti in updates]
        a = (abi.RtxItemXform * len(arr))(*arr)
        self._check(self._fn("scene_update_items")(self._h, a, len(arr)))

    def sample_table(self, samples: int):
        cell = C.c_uint32()
        xy = np.zeros((samples, 2), dtype=np.uint16)
        self._check(self._fn("sample_table")(int(samples), C.byref(cell), xy.ctypes.data))
        return int(cell.value), xy


def primary_ray(cam: abi.RtxCamera, x: int, y: int):
    """Ray of Raytracing::pick (src/raytracing.rs:241-262), normalised direction."""
    pinv = np.array(cam.projection_inverse, dtype=np.float32).reshape(4, 4).T
    vinv = np.array(cam.view_inverse, dtype=np.float32).reshape(4, 4).T
    sx = np.float32((np.float32(x) + np.float32(0.5)) / np.float32(cam.width)) * np.float32(2) - np.float32(1)
    sy = np.float32(1) - np.float32((np.float32(y) + np.float32(0.5)) / np.float32(cam.height)) * np.float32(2)
    pp = pinv @ np.array([sx, sy, -1, 1], dtype=np.float32)
    pp[3] = 1
    rd = pp.copy(); rd[3] = 0
    o = (vinv @ pp)[:3]
    d = (vinv @ rd)[:3]
    return o.astype(np.float32), (d / np.linalg.norm(d)).astype(np.float32)


class RendererManager(AbiRenderer):
    """B200 stand-in for reference `RendererManager` (src/renderer.rs:19-60)."""

    def __init__(self, width: int, height: int, flat_scene: abi.FlatScene, device: int = 0, devices=None, device_bvh: bool = False):
        super().__init__(load_library(), "rtx_", flat_scene, device, devices, device_bvh)
        self.width, self.height = width, height
        self.frame = Frame(width, height, pinned=True)
        self._done = False

    # RendererManager::start (src/renderer.rs:105-172) — blocking on the GPU
    def start(self, cam: abi.RtxCamera, cfg: abi.RtxConfig) -> Frame:
        self._done, self._async = False, False
        self.render(cam, cfg, self.frame)
        self._done = True
        return self.frame

    # Non-blocking path of the GUI: start_async / is_running / is_done / get_rendered_pixels / stop (src/renderer.rs:105-231)
    def start_async(self, cam: abi.RtxCamera, cfg: abi.RtxConfig) -> None:
        self._done, self._async = False, True
        f = self.frame
        self._keep = (cam, cfg)
        self._check(self._lib.rtx_render_frame_async(self._h, C.byref(cam), C.byref(cfg), f.image.ctypes.data, f.normals.ctypes.data,
                                                     f.depth.ctypes.data, f.objects.ctypes.data))

    def _poll(self):
        px, run, done, res = C.c_uint64(), C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.rtx_render_poll(self._h, C.byref(px), C.byref(run), C.byref(done), C.byref(res), C.byref(self.frame.stats)))
        return int(px.value), bool(run.value), bool(done.value), int(res.value)

    def is_running(self) -> bool:            # src/renderer.rs:221-227
        return self._poll()[1] if getattr(self, "_async", False) else False

    def is_done(self) -> bool:               # src/renderer.rs:228-231
        if getattr(self, "_async", False):
            self._done = self._poll()[2]
        return self._done

    def get_rendered_pixels(self) -> int:    # src/renderer.rs:215-219
        if getattr(self, "_async", False):
            return self._poll()[0]
        return self.width * self.height if self._done else 0

    def snapshot(self) -> int:
        """Progressive display (the per-pixel mpsc stream of src/renderer.rs:305-312 -> src/run.rs:506-545): refresh
        `self.frame` with the current state of the frame in flight; returns get_rendered_pixels()."""
        px = C.c_uint64()
        self._check(self._lib.rtx_render_snapshot(self._h, C.byref(px)))
        return int(px.value)

    def stop(self) -> None:                  # src/renderer.rs:174-199
        self._check(self._lib.rtx_render_stop(self._h))

    def pick(self, cam: abi.RtxCamera, x: int, y: int):
        """Raytracing::pick (src/raytracing.rs:237-273): (item id, name, distance) or None."""
        o, d = primary_ray(cam, x, y)
        h = self.trace([o], [d])[0]
        if h["t"] < 0:
            return None
        return int(h["item_id"]), self.flat.item_names[int(h["item_index"])], float(h["t"])

    def bvh_info(self) -> abi.RtxBvhInfo:
        info = abi.RtxBvhInfo()
        self._check(self._lib.rtx_scene_bvh_info(self._h, C.byref(info)))
        return info

    # -- device-buffer path (multi-GPU bench): torch tensors in, same layouts ---------------
    def render_device(self, cam, cfg, shard, t_rgba, t_normals, t_depth, t_ids, stream_ptr=0) -> abi.RtxStats:
        st = abi.RtxStats()
        sh = C.byref(shard) if shard is not None else None
        rc = self._lib.rtx_render_frame_device(self._h, C.byref(cam), C.byref(cfg), sh, t_rgba.data_ptr(),
                                               t_normals.data_ptr(), t_depth.data_ptr(), t_ids.data_ptr(),
                                               C.c_void_p(stream_ptr), C.byref(st))
        self._check(rc)
        return st
