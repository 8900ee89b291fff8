"""Shared helpers of the parity tests."""
import ctypes as C

import numpy as np

from rustray_b200 import abi


def clone_cfg(cfg: abi.RtxConfig, **over) -> abi.RtxConfig:
    out = abi.RtxConfig()
    C.memmove(C.byref(out), C.byref(cfg), C.sizeof(cfg))
    for k, v in over.items():
        setattr(out, k, v)
    return out


def scene_to_abi(sc, **cfg_over):
    fs = abi.FlatScene.from_scene(sc)
    return fs, abi.make_camera(sc.cam), abi.make_config(sc.config, **cfg_over)


def lsb_stats(a: np.ndarray, b: np.ndarray):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32)).max(axis=-1)
    return float((d <= 1).mean()), float((d == 0).mean()), int(d.max())


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def random_rays(n: int, seed: int, center=(0, 0, -10), radius=12.0):
    """Rays from points on a sphere around the scene towards jittered points near the centre + a few
    axis-aligned / degenerate directions (zero components exercise the dir == 0 slab branch)."""
    rng = np.random.default_rng(seed)
    c = np.array(center, dtype=np.float32)
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    o = (c + v * radius).astype(np.float32)
    tgt = (c + rng.normal(size=(n, 3)) * radius * 0.35).astype(np.float32)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    k = min(n // 10, 64)
    axes = np.eye(3, dtype=np.float32)
    for i in range(k):
        d[i] = axes[i % 3] * (1 if (i // 3) % 2 == 0 else -1)
        o[i] = c - d[i] * radius + rng.normal(size=3).astype(np.float32) * 2.0
    return o, d
