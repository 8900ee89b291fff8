"""Run under torchrun with N >= 2 ranks on one box: the same frame (a) rendered by all ranks into rank 0's frame buffers over
CUDA IPC / NVLink peer memory (PeerFrame) and (b) gathered with one NCCL gather + one scatter kernel (ShardedRenderer), both
compared on rank 0 with the single-GPU frame."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rustray_b200 import abi  # noqa: E402
from rustray_b200.distributed import PeerFrame, ShardedRenderer  # noqa: E402
from rustray_b200.renderer import Frame, RendererManager  # noqa: E402


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fs, cam, cfg = abi.load_fixture("c2_floor_monkey", samples=4, monte_carlo=1)
    cam = abi.resize_camera(cam, 640, 360)
    w, h = 640, 360
    rm = RendererManager(w, h, fs, device=local)
    ref = rm.start(cam, cfg) if rank == 0 else None
    shard = abi.RtxShard(rank, world, 8, 4)
    pf = PeerFrame(rm._lib, w, h, rank, world, local)
    p = pf.pointers()
    st = abi.RtxStats()
    stream = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(2):
        rm._check(rm._lib.rtx_render_frame_device(rm._h, C.byref(cam), C.byref(cfg), C.byref(shard), p[0], p[1], p[2], p[3], C.c_void_p(stream), C.byref(st)))
        pf.finish()
    if rank == 0:
        out = Frame(w, h)
        pf.download(out, stream)
        same = (np.abs(out.image.astype(int) - ref.image.astype(int)).max(axis=-1) <= 1).mean()
        assert np.array_equal(out.objects, ref.objects) and np.array_equal(out.depth, ref.depth) and same >= 0.9999, same
        print("PEER_FRAME_OK", flush=True)
    dist.barrier()
    sr = ShardedRenderer(rm, w, h, rank, world, device=dev)
    sr.render_local(cam, cfg)
    sr.gather()
    torch.cuda.synchronize()
    if rank == 0:
        ids = sr.ids.cpu().numpy().view(np.uint32).reshape(h, w)
        depth = sr.depth.cpu().numpy().reshape(h, w)
        assert np.array_equal(ids, ref.objects) and np.array_equal(depth, ref.depth)
        print("NCCL_GATHER_OK", flush=True)
    dist.barrier()
    pf.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
