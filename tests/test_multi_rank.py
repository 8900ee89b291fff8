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


# ---- frame-parallel animation over 2 ranks (the oracle stands in for the GPU renderer on CPU) ----------------
def _anim_setup():
    from oracle.oracle import OracleRenderer
    from rustray_b200 import abi
    from rustray_b200.animation import Animation
    from rustray_b200.scene_loader import Item, SHAPE_MESH, Material, mat_identity
    fs, cam, cfg = abi.load_fixture("monkey_gltf", samples=1, monte_carlo=0)
    w, h = 48, 27
    cam = abi.resize_camera(cam, w, h)
    kf = lambda ry: {"rotation": {"x": 0.0, "y": ry, "z": 0.0}, "scale": {"x": 1.0, "y": 1.0, "z": 1.0}, "translation": {"x": 0.0, "y": 0.0, "z": 0.0}}
    an = Animation({"fps": 25, "enabled": True, "keyframes": [{"time": 0, "objects": [{"name": "Suzanne", "transformation": kf(0.0)}]},
                                                              {"time": 200, "objects": [{"name": "Suzanne", "transformation": kf(90.0)}]}]})
    items = [Item(id=0, name=n, shape=SHAPE_MESH, material=Material(), trans=mat_identity()) for n in fs.item_names]
    r = OracleRenderer(fs)

    def render_frame(f):
        r.update_items(an.updates_for_frame(items, f))
        return r.render(cam, cfg)
    return render_frame, an.frames_to_render(), w, h


def _anim_worker(rank, world, port, out_path):
    from rustray_b200.distributed import render_animation_frame_parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    render_frame, n_frames, w, h = _anim_setup()
    got = {}
    mine = render_animation_frame_parallel(render_frame, n_frames, w, h, rank, world, on_frame=lambda f, fr: got.__setitem__(f, fr.image.copy()))
    assert mine == list(range(rank, n_frames, world))
    if rank == 0:
        ref_render, _, _, _ = _anim_setup()
        ok = list(got.keys()) == list(range(n_frames)) and all(np.array_equal(got[f], ref_render(f).image) for f in range(n_frames))
        moved = any((got[0] != got[f]).any() for f in range(1, n_frames))
        open(out_path, "w").write("ok" if ok and moved else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_frame_parallel_animation_two_ranks(tmp_path):
    """5 frames over 2 ranks (ragged last round): rank 0 receives every frame, in order, identical to a serial render."""
    render_frame, n_frames, w, h = _anim_setup()
    assert n_frames == 5
    out = str(tmp_path / "anim.txt")
    mp.spawn(_anim_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"
