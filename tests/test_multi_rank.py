"""CPU suite, part 4: the N > 1 host logic over gloo (world_size 2): interleaved tile ownership, packed
G-buffer gather to rank 0 and scatter — the same code path the NCCL run uses, with numpy pack/unpack
standing in for the device kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rustray_b200.distributed import gather_frame_cpu, pack_numpy, shard_pixels, unpack_numpy
from rustray_b200.renderer import Frame


def _free_port() -> int:
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _synthetic_frame(w, h, seed=0) -> Frame:
    rng = np.random.default_rng(seed)
    f = Frame(w, h)
    f.image[:] = rng.integers(0, 255, size=f.image.shape)
    f.normals[:] = rng.normal(size=f.normals.shape)
    f.normals[0, 0] = np.nan                                        # miss pixels carry NaN normals
    f.depth[:] = rng.uniform(0, 50, size=f.depth.shape)
    f.objects[:] = rng.integers(0, 100, size=f.objects.shape)
    return f


def _worker(rank, world, port, w, h, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _synthetic_frame(w, h)
    # each rank only "renders" its own pixels: blank everything else to prove nothing leaks through
    mine = shard_pixels(w, h, rank, world)
    local = Frame(w, h)
    for src, dst in ((full.image.reshape(-1, 4), local.image.reshape(-1, 4)), (full.normals.reshape(-1, 3), local.normals.reshape(-1, 3)),
                     (full.depth.reshape(-1), local.depth.reshape(-1)), (full.objects.reshape(-1), local.objects.reshape(-1))):
        dst[mine] = src[mine]
    out = gather_frame_cpu(rank, world, w, h, local)
    if rank == 0:
        ok = (np.array_equal(out.image, full.image) and np.array_equal(out.depth, full.depth) and np.array_equal(out.objects, full.objects)
              and np.array_equal(np.nan_to_num(out.normals, nan=-7), np.nan_to_num(full.normals, nan=-7)))
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h", [(64, 36), (37, 23)])
def test_two_rank_gather_reassembles_the_frame(tmp_path, w, h):
    out = str(tmp_path / "res.txt")
    mp.spawn(_worker, args=(2, _free_port(), w, h, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_pack_unpack_roundtrip_and_ragged_tiles():
    for (w, h, world) in [(33, 17, 3), (8, 4, 1), (5, 3, 8)]:
        full = _synthetic_frame(w, h, 3)
        out = Frame(w, h)
        for r in range(world):
            px = shard_pixels(w, h, r, world)
            unpack_numpy(px, pack_numpy(px, full.image, full.normals, full.depth, full.objects), out.image, out.normals, out.depth, out.objects)
        assert np.array_equal(out.image, full.image) and np.array_equal(out.objects, full.objects) and np.array_equal(out.depth, full.depth)
    assert shard_pixels(5, 3, 7, 8).size == 0                      # more ranks than tiles: empty shard is fine
