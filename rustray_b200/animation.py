"""Keyframe animation -> per-frame item transforms (reference src/animation.rs:140-205, Scene::apply_frame
src/scene.rs:1695-1713).  Host side, cold path: each frame yields the (item_index, trans, tran_inverse) updates that
go to rtx_scene_update_items."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

from .scene_loader import Item, mat_euler, mat_identity, mat_inverse, mat_mul, mat_scaling, mat_translation, to_radians

F = np.float32


class Animation:
    def __init__(self, spec: Optional[dict]):
        spec = spec or {}
        self.enabled = bool(spec.get("enabled", False))
        self.fps = int(spec.get("fps", 25))
        self.keyframes: List[Tuple[int, Dict[str, dict]]] = []
        for kf in spec.get("keyframes") or []:
            if kf.get("time") is None:
                continue
            objs = {}
            for o in kf.get("objects") or []:
                tr = o.get("transformation") or {}

                def vec(key, rad=False):
                    v = tr.get(key)
                    if not isinstance(v, dict) or any(v.get(k) is None for k in "xyz"):
                        return None
                    a = np.array([v["x"], v["y"], v["z"]], dtype=F)
                    return np.array([to_radians(c) for c in a], dtype=F) if rad else a
                if o["name"] not in objs:                              # first match wins (animation.rs:147-163)
                    objs[o["name"]] = {"translation": vec("translation"), "scale": vec("scale"), "rotation": vec("rotation", True)}
            self.keyframes.append((int(kf["time"]), objs))

    def frames_to_render(self) -> int:                                  # animation.rs:90-101
        last = self.keyframes[-1][0] if self.keyframes else 0
        return int(math.floor(self.fps * (last / 1000.0)))

    def has_animation(self) -> bool:                                    # :76-79
        return self.enabled and self.frames_to_render() > 0 and bool(self.keyframes) and self.keyframes[0][0] == 0 and len(self.keyframes) >= 2

    def _keyframes_for(self, frame: int):                               # :103-131
        ts = int(math.floor((1000.0 / self.fps) * frame))
        first = last = self.keyframes[0]
        for i, kf in enumerate(self.keyframes):
            if kf[0] <= ts:
                first = kf
                last = kf if i + 1 >= len(self.keyframes) else self.keyframes[i + 1]
        diff = last[0] - first[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = np.float64(1.0) / np.float64(diff) * np.float64(ts - first[0])
        return first, last, factor

    def trans_for_frame(self, frame: int, name: str) -> Optional[np.ndarray]:   # :133-205
        first, last, factor = self._keyframes_for(frame)
        a, b = first[1].get(name), last[1].get(name)
        if a is None or b is None:
            return None
        f = F(factor)

        def lerp(key, default):
            if a[key] is None or b[key] is None:
                return np.array(default, dtype=F)
            return np.array([F(a[key][i] + F(f * F(b[key][i] - a[key][i]))) for i in range(3)], dtype=F)   # helper::interpolate
        t, s, r = lerp("translation", (0, 0, 0)), lerp("scale", (1, 1, 1)), lerp("rotation", (0, 0, 0))
        m = mat_identity()
        m = mat_mul(m, mat_translation(*t))
        m = mat_mul(m, mat_euler(0.0, 0.0, r[2])); m = mat_mul(m, mat_euler(0.0, r[1], 0.0)); m = mat_mul(m, mat_euler(r[0], 0.0, 0.0))
        return mat_mul(m, mat_scaling(*s))

    def updates_for_frame(self, items: List[Item], frame: int):
        """Scene::apply_frame: (item_index, trans, tran_inverse) for every item whose name is animated."""
        if not self.has_animation() or frame > self.frames_to_render():
            return []
        out = []
        for i, it in enumerate(items):
            m = self.trans_for_frame(frame, it.name)
            if m is not None:
                out.append((i, m, mat_inverse(m)))
        return out
