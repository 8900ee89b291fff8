"""Multi-GPU frames: interleaved image tiles, one process per GPU, one gather of the G-buffer.

The reference has no distributed path (threads only).  Pixels are independent and the
scene is read-only during a frame, so the path shards with no data-path collective: the flattened
scene is replicated on every GPU, the row-major tiles of tile_w x tile_h pixels are dealt out in groups of
`world` (one tile per rank and group, rotated per group so that a rank does not keep the same image columns),
and each rank renders and resolves its own pixels.  The finished pixels reach rank 0 in one of two ways:

  * `PeerFrame` (default on one NVLink box): rank 0 owns the 24 B/pixel frame buffers and exports them through CUDA IPC;
    every rank's resolve kernel STORES its pixels directly into them over NVLink peer memory (resolve and gather are one
    kernel, csrc/rtx_api.cu rtx_gbuffer_*); a barrier ends the frame.
  * `ShardedRenderer.gather`: ONE NCCL gather of the packed G-buffer (rgba8 | normal 3xf32 | depth f32 | id u32) into one
    contiguous buffer on rank 0 and ONE scatter kernel (the same host logic runs over gloo on CPU in the tests with a
    numpy pack/unpack).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import abi

DEFAULT_TILE = (8, 4)     # one warp of primary rays per tile


def shard_pixels(width: int, height: int, rank: int, world: int, tile_w: int = DEFAULT_TILE[0], tile_h: int = DEFAULT_TILE[1]) -> np.ndarray:
    """Frame pixel indices (y*width+x) owned by `rank`, in the order the library renders and packs them
    (must match get_pixel_list in csrc/rtx_api.cu)."""
    tx, ty = (width + tile_w - 1) // tile_w, (height + tile_h - 1) // tile_h
    g = np.arange((tx * ty + world - 1) // world, dtype=np.int64)               # groups of `world` consecutive tiles
    rot = ((g * 0x9E3779B1) & 0xFFFFFFFF) >> 16                                  # per-group rotation (shard_rot in csrc/rtx_api.cu)
    t = g * world + (rank + world - rot % world) % world
    t = t[t < tx * ty]
    x0, y0 = (t % tx) * tile_w, (t // tx) * tile_h
    dy, dx = np.mgrid[0:tile_h, 0:tile_w]
    x = x0[:, None] + dx.reshape(1, -1)
    y = y0[:, None] + dy.reshape(1, -1)
    ok = (x < width) & (y < height)
    return (y * width + x)[ok].astype(np.uint32)


def packed_bytes(n_pixels: int) -> int:
    return 24 * n_pixels


def pack_numpy(pixels: np.ndarray, image, normals, depth, objects) -> np.ndarray:
    """CPU mirror of rtx_shard_pack: rgba[n] | normals[n*3] | depth[n] | ids[n] as one byte buffer."""
    n = pixels.size
    out = np.zeros(24 * n, dtype=np.uint8)
    out[: 4 * n] = image.reshape(-1, 4)[pixels].reshape(-1)
    out[4 * n: 16 * n] = normals.reshape(-1, 3)[pixels].astype(np.float32).view(np.uint8).reshape(-1)
    out[16 * n: 20 * n] = depth.reshape(-1)[pixels].astype(np.float32).view(np.uint8).reshape(-1)
    out[20 * n: 24 * n] = objects.reshape(-1)[pixels].astype(np.uint32).view(np.uint8).reshape(-1)
    return out


def unpack_numpy(pixels: np.ndarray, packed: np.ndarray, image, normals, depth, objects) -> None:
    n = pixels.size
    image.reshape(-1, 4)[pixels] = packed[: 4 * n].reshape(-1, 4)
    normals.reshape(-1, 3)[pixels] = packed[4 * n: 16 * n].view(np.float32).reshape(-1, 3)
    depth.reshape(-1)[pixels] = packed[16 * n: 20 * n].view(np.float32)
    objects.reshape(-1)[pixels] = packed[20 * n: 24 * n].view(np.uint32)


def gather_frame_cpu(rank: int, world: int, width: int, height: int, local_frame, tile=DEFAULT_TILE, group=None):
    """gloo path used by the CPU tests: every rank packs its owned pixels from `local_frame`
    (renderer.Frame with at least those pixels valid); rank 0 returns the assembled Frame."""
    import torch
    import torch.distributed as dist
    from .renderer import Frame
    mine = shard_pixels(width, height, rank, world, *tile)
    buf = torch.from_numpy(pack_numpy(mine, local_frame.image, local_frame.normals, local_frame.depth, local_frame.objects))
    sizes = [packed_bytes(shard_pixels(width, height, r, world, *tile).size) for r in range(world)]
    pad = max(sizes)
    send = torch.zeros(pad, dtype=torch.uint8)
    send[: buf.numel()] = buf
    recv = [torch.zeros(pad, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0, group=group)
    if rank != 0:
        return None
    out = Frame(width, height)
    for r in range(world):
        px = shard_pixels(width, height, r, world, *tile)
        unpack_numpy(px, recv[r][: sizes[r]].numpy(), out.image, out.normals, out.depth, out.objects)
    return out


class PeerFrame:
    """Frame buffers on rank 0 that every rank renders into through NVLink peer memory (CUDA IPC handle broadcast once).
    `pointers()` are the d_rgba / d_normals / d_depth / d_object_ids to pass to RendererManager.render_device together
    with this rank's shard; `finish()` is the barrier after which rank 0 holds the whole frame.  `ok` is False on EVERY rank when
    any rank could not create / open the buffers (no peer access between the GPUs): the caller then gathers with NCCL
    (ShardedRenderer)."""

    def __init__(self, lib, width: int, height: int, rank: int, world: int, device_index: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        self.lib, self.w, self.h, self.rank, self.world, self.group = lib, width, height, rank, world, group
        self._g = C.c_void_p()
        self.ptrs = [0, 0, 0, 0]
        handle = (C.c_uint8 * 64)()
        good = 1
        if rank == 0:
            if lib.rtx_gbuffer_create(device_index, width, height, C.byref(self._g)) != 0:
                good = 0
            elif world > 1 and lib.rtx_gbuffer_export(self._g, handle) != 0:
                good = 0
        if world > 1:
            dev = torch.device("cuda", device_index)
            t = torch.tensor(list(bytes(handle)) + [good], dtype=torch.uint8, device=dev)
            dist.broadcast(t, src=0, group=group)
            raw = bytes(t.cpu().tolist())
            good = raw[64]
            if rank != 0 and good:
                C.memmove(handle, raw[:64], 64)
                if lib.rtx_gbuffer_open(device_index, width, height, handle, C.byref(self._g)) != 0:
                    good = 0
            f = torch.tensor([good], dtype=torch.int32, device=dev)
            dist.all_reduce(f, op=dist.ReduceOp.MIN, group=group)
            good = int(f.item())
        self.ok = bool(good)
        self._token = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", device_index)) if world > 1 else None
        if not self.ok:
            self.close()
            return
        p = [C.c_void_p() for _ in range(4)]
        _check(lib, lib.rtx_gbuffer_pointers(self._g, *[C.byref(x) for x in p]))
        self.ptrs = [x.value for x in p]

    def pointers(self):
        return self.ptrs

    def finish(self) -> None:
        """End of the frame on the DEVICE timeline: a one-element all-reduce enqueued behind this rank's frame.  On rank 0 whatever is
        enqueued next (the D2H copy of the frame, the timing event) runs after every rank's resolve kernel has finished — its peer
        stores are complete at kernel end.  Unlike dist.barrier() the host does not block here."""
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(self._token, group=self.group)

    def download(self, frame, stream_ptr=0) -> None:
        """rank 0: copy the assembled frame into a renderer.Frame (page-locked host buffers)."""
        import ctypes as C
        _check(self.lib, self.lib.rtx_gbuffer_download(self._g, frame.image.ctypes.data, frame.normals.ctypes.data, frame.depth.ctypes.data,
                                                       frame.objects.ctypes.data, C.c_void_p(stream_ptr)))

    def close(self) -> None:
        if self._g:
            self.lib.rtx_gbuffer_destroy(self._g)
            self._g = None
        self.ptrs = [0, 0, 0, 0]


def _check(lib, rc: int) -> None:
    if rc != 0:
        from .renderer import RtxError
        msg = lib.rtx_last_error()
        raise RtxError("rtx error %d: %s" % (rc, msg.decode() if msg else ""))


class ShardedRenderer:
    """One rank of a multi-GPU render.  `rm` is this rank's RendererManager (scene replicated)."""

    def __init__(self, rm, width: int, height: int, rank: int, world: int, tile=DEFAULT_TILE, device=None):
        import torch
        self.rm, self.w, self.h, self.rank, self.world, self.tile = rm, width, height, rank, world, tile
        self.dev = device if device is not None else torch.device("cuda", rm.device)
        self.shard = abi.RtxShard(rank, world, tile[0], tile[1])
        n = width * height
        self.rgba = torch.zeros(n * 4, dtype=torch.uint8, device=self.dev)
        self.normals = torch.zeros(n * 3, dtype=torch.float32, device=self.dev)
        self.depth = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.ids = torch.zeros(n, dtype=torch.int32, device=self.dev)
        lib = rm._lib
        self.sizes = [int(lib.rtx_shard_packed_bytes(width, height, abi.RtxShard(r, world, tile[0], tile[1]))) for r in range(world)]
        self.pad = (max(self.sizes) + 15) // 16 * 16
        self.send = torch.zeros(self.pad, dtype=torch.uint8, device=self.dev)
        # one contiguous receive buffer (world x pad): the gather lands in it and ONE kernel scatters all ranks' pixels
        self.recv_all = torch.zeros(self.pad * world, dtype=torch.uint8, device=self.dev) if rank == 0 else None
        self.recv = list(self.recv_all.split(self.pad)) if rank == 0 else None

    def render_local(self, cam, cfg) -> abi.RtxStats:
        import torch
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        return self.rm.render_device(cam, cfg, self.shard, self.rgba, self.normals, self.depth, self.ids, stream)

    def gather(self, group=None) -> None:
        """Pack owned pixels, gather to rank 0, scatter into rank 0's frame buffers."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        lib = self.rm._lib
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        rc = lib.rtx_shard_pack(self.w, self.h, C.byref(self.shard), self.rgba.data_ptr(), self.normals.data_ptr(), self.depth.data_ptr(),
                                self.ids.data_ptr(), self.send.data_ptr(), stream)
        self.rm._check(rc)
        if self.world > 1:
            dist.gather(self.send, self.recv, dst=0, group=group)
            if self.rank == 0:
                rc = lib.rtx_shard_unpack_all(self.w, self.h, self.world, self.tile[0], self.tile[1], 1, self.recv_all.data_ptr(), self.pad,
                                              self.rgba.data_ptr(), self.normals.data_ptr(), self.depth.data_ptr(), self.ids.data_ptr(), stream)
                self.rm._check(rc)


# ---------------------------------------------------------------------------------------------------------
# frame-parallel animation: the frames of a keyframe animation are independent once
# Scene::apply_frame has produced the item transforms (reference src/scene.rs:1695-1713, src/run.rs:422-485),
# so rank r renders whole frames r, r+N, r+2N, ... on its own GPU and rank 0 receives them in frame order.
# ---------------------------------------------------------------------------------------------------------
def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_frames, world))


def render_animation_frame_parallel(render_frame, n_frames: int, width: int, height: int, rank: int, world: int,
                                    on_frame=None, device=None, group=None) -> List[int]:
    """`render_frame(f) -> renderer.Frame` renders animation frame f on this rank (apply the frame's transforms with
    RendererManager.update_items, then start()).  Frames are produced in rounds of `world`; after each round ONE gather
    brings the round's packed frames (24 B/pixel) to rank 0, which calls `on_frame(f, Frame)` in increasing f — the order
    Run::save_image numbers its output files in.  Returns the frames this rank rendered."""
    import torch
    from .renderer import Frame
    n = width * height
    all_px = np.arange(n, dtype=np.uint32)
    dev = device if device is not None else torch.device("cpu")
    mine = []
    for first in range(0, n_frames, world):
        f = first + rank
        send = torch.zeros(24 * n, dtype=torch.uint8)
        if f < n_frames:
            fr = render_frame(f)
            mine.append(f)
            send = torch.from_numpy(pack_numpy(all_px, fr.image, fr.normals, fr.depth, fr.objects))
        if world == 1:
            if on_frame is not None:
                on_frame(f, fr)
            continue
        import torch.distributed as dist
        send = send.to(dev)
        recv = [torch.zeros(24 * n, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(send, recv, dst=0, group=group)
        if rank == 0 and on_frame is not None:
            for r in range(world):
                if first + r >= n_frames:
                    break
                out = Frame(width, height)
                unpack_numpy(all_px, recv[r].cpu().numpy(), out.image, out.normals, out.depth, out.objects)
                on_frame(first + r, out)
    return mine
