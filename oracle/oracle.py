"""Python binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under rustray_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rustray_b200 import abi
from rustray_b200.renderer import AbiRenderer, Frame

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rt_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _lib = C.CDLL(so)
        abi.bind(_lib, "oracle_")
        P = C.POINTER
        _lib.oracle_render_frame_ex.restype = C.c_int
        _lib.oracle_render_frame_ex.argtypes = [C.c_void_p, P(abi.RtxCamera), P(abi.RtxConfig), C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, P(abi.RtxStats), C.c_int, C.c_int, C.c_int]
        _lib.oracle_scene_set_options.argtypes = [C.c_void_p, C.c_uint32]
        _lib.oracle_chacha_block.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        _lib.oracle_seed_from_u64.argtypes = [C.c_uint64, C.c_void_p]
        _lib.oracle_fresnel.restype = C.c_float
        _lib.oracle_fresnel.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        _lib.oracle_approx_equal.argtypes = [C.c_float, C.c_float]
        _lib.oracle_gen_ray.argtypes = [P(abi.RtxCamera), P(abi.RtxConfig), C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.oracle_mc_uniform.restype = C.c_float
        _lib.oracle_mc_uniform.argtypes = [C.c_uint32] * 5
        _lib.oracle_tex_fetch.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        _lib.oracle_tri_cast.argtypes = [C.c_void_p] * 5 + [P(C.c_float), C.c_void_p, P(C.c_int)]
    return _lib


class OracleRenderer(AbiRenderer):
    def __init__(self, flat_scene: abi.FlatScene):
        super().__init__(load(), "oracle_", flat_scene, 0)

    def set_options(self, brute_force: bool = False, ball_normal_outward_inside: bool = False) -> None:
        self._lib.oracle_scene_set_options(self._h, int(brute_force) | (int(ball_normal_outward_inside) << 1))

    def render_ex(self, cam, cfg, threads: int = 0, cell_step: int = 1, faithful: bool = False, frame: Frame = None) -> Frame:
        f = frame or Frame(cam.width, cam.height)
        if threads <= 0:
            threads = os.cpu_count() or 1
        rc = self._lib.oracle_render_frame_ex(self._h, C.byref(cam), C.byref(cfg), f.image.ctypes.data,
                                              f.normals.ctypes.data, f.depth.ctypes.data, f.objects.ctypes.data,
                                              C.byref(f.stats), int(threads), int(cell_step), int(faithful))
        self._check(rc)
        return f


def gen_ray(cam, cfg, x, y, x_i=0, y_i=0, cell=1):
    o = np.zeros(3, dtype=np.float32); d = np.zeros(3, dtype=np.float32)
    load().oracle_gen_ray(C.byref(cam), C.byref(cfg), x, y, x_i, y_i, cell, o.ctypes.data, d.ctypes.data)
    return o, d
