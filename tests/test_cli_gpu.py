"""GPU suite: the stand-in CLI (python -m rustray_b200, reference src/main.rs flags) end to end on a small JSON scene in the
reference's schema with a 5-frame keyframe animation — on one GPU, and frame-parallel under torchrun on two."""
import glob
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from rustray_b200 import abi
from tests.util import lsb_stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE = os.path.join("tests", "golden", "json", "bouncing_spheres.json")
pytestmark = pytest.mark.gpu


def _expected_frames(w, h):
    from oracle.oracle import OracleRenderer
    from rustray_b200.animation import Animation
    from rustray_b200.scene_loader import load_scene
    sc = load_scene([SCENE], w, h, asset_root=ROOT)
    fs, cam, cfg = abi.FlatScene.from_scene(sc), abi.make_camera(sc.cam), abi.make_config(sc.config)
    an, r, out = Animation(sc.animation), OracleRenderer(fs), []
    for f in range(an.frames_to_render()):
        r.update_items(an.updates_for_frame(sc.items, f))
        out.append(r.render(cam, cfg).image.copy())
    return out


def _read_frames(out_dir):
    from PIL import Image
    files = sorted(glob.glob(os.path.join(out_dir, "output_*.png")), key=lambda p: int(p.rsplit("_", 1)[1].split(".")[0]))
    return [np.asarray(Image.open(p).convert("RGBA")) for p in files]


def test_cli_renders_the_animation_on_one_gpu(tmp_path):
    out = str(tmp_path / "frames")
    r = subprocess.run([sys.executable, "-m", "rustray_b200", SCENE, "cmd", "320x180", "out=" + out], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got, want = _read_frames(out), _expected_frames(320, 180)
    assert len(got) == len(want) == 5
    for g, w in zip(got, want):
        assert lsb_stats(g, w)[0] >= 0.999
    assert (got[0] != got[4]).any()                                  # the ball moved


def test_cli_frame_parallel_under_torchrun_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "frames2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), "-m", "rustray_b200", SCENE, "cmd", "320x180", "out=" + out],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    got, want = _read_frames(out), _expected_frames(320, 180)
    assert len(got) == 5                                             # rank 0 wrote every frame, in order
    for g, w in zip(got, want):
        assert lsb_stats(g, w)[0] >= 0.999
