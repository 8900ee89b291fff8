"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/rtx.h declares, agrees with the
header on struct layouts, and refuses to work without a GPU (no CPU fallback).  No device compute here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from rustray_b200 import abi, renderer
from rustray_b200.distributed import shard_pixels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtx.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(renderer.lib_path())
    names = declared_functions()
    assert len(names) >= 17 and set(names) == set(abi.ABI_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.rtx_abi_version() == 2


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a tiny C program against include/rtx.h and compare sizeof/offsetof with the ctypes mirror."""
    structs = {"RtxTexture": abi.RtxTexture, "RtxMaterial": abi.RtxMaterial, "RtxMesh": abi.RtxMesh, "RtxItem": abi.RtxItem,
               "RtxLight": abi.RtxLight, "RtxSceneDesc": abi.RtxSceneDesc, "RtxCamera": abi.RtxCamera, "RtxConfig": abi.RtxConfig,
               "RtxShard": abi.RtxShard, "RtxStats": abi.RtxStats, "RtxRay": abi.RtxRay, "RtxHit": abi.RtxHit,
               "RtxBvhInfo": abi.RtxBvhInfo, "RtxItemXform": abi.RtxItemXform, "RtxShadowHit": abi.RtxShadowHit}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rtx.h"', "int main(void){"]
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, fname, name, fname))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (name, fname)]) == getattr(cls, fname).offset, (name, fname)


def test_sample_table_matches_the_oracle():
    """rtx_sample_table is host code (ChaCha12 shuffle of raytracing.rs:300-313); it must agree with the oracle."""
    from oracle import oracle
    lib = renderer.load_library()
    olib = oracle.load()
    for samples in (1, 2, 5, 32, 64, 128, 500):
        a = np.zeros((samples, 2), dtype=np.uint16); b = np.zeros((samples, 2), dtype=np.uint16)
        ca, cb = C.c_uint32(), C.c_uint32()
        assert lib.rtx_sample_table(samples, C.byref(ca), a.ctypes.data) == 0
        assert olib.oracle_sample_table(samples, C.byref(cb), b.ctypes.data) == 0
        assert ca.value == cb.value and (a == b).all()
    assert lib.rtx_sample_table(0, None, None) == -1 and lib.rtx_sample_table(70000, None, None) == -1
    assert b"samples" in lib.rtx_last_error()


def test_shard_pixel_counts_partition_the_frame():
    lib = renderer.load_library()
    for (w, h, world, tw, th) in [(1280, 720, 8, 8, 4), (800, 600, 3, 32, 8), (37, 23, 5, 8, 4), (3840, 2160, 8, 16, 16), (5, 3, 8, 8, 4)]:
        counts = [lib.rtx_shard_pixel_count(w, h, C.byref(abi.RtxShard(r, world, tw, th))) for r in range(world)]
        assert sum(counts) == w * h
        px = [shard_pixels(w, h, r, world, tw, th) for r in range(world)]
        assert [p.size for p in px] == counts
        allpx = np.concatenate(px)
        assert np.array_equal(np.sort(allpx), np.arange(w * h, dtype=np.uint32))       # each pixel exactly once
        assert lib.rtx_shard_packed_bytes(w, h, C.byref(abi.RtxShard(0, world, tw, th))) == 24 * counts[0]
        if w * h >= 64 * world:
            assert max(counts) - min(counts) <= tw * th * 2 + w * th                     # interleaving balances the ranks
    assert lib.rtx_shard_pixel_count(10, 10, C.byref(abi.RtxShard(3, 2, 8, 4))) == 0     # rank >= world
    # the per-group rotation keeps a rank from owning fixed image columns (1280/8 = 160 tiles per row, 160 % 8 == 0: plain
    # t % world would give rank 0 the tile columns 0, 8, 16, ... only — 8 % time imbalance on config 2)
    px0 = shard_pixels(1280, 720, 0, 8, 8, 4)
    cols = np.unique((px0 % 1280) // 8 % 8)
    assert cols.size == 8


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    lib = renderer.load_library()
    if lib.rtx_device_count() > 0:
        pytest.skip("a GPU is present")
    fs, cam, cfg = abi.load_fixture("c1_spheres")
    with pytest.raises(renderer.RtxError, match="no CUDA device"):
        renderer.RendererManager(cam.width, cam.height, fs)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/ (task rule ③)."""
    pkg = os.path.join(ROOT, "rustray_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                code = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith(("#", "//", "*", '"""')))
                assert "liboracle" not in code and "from oracle" not in code and "import oracle" not in code and "rt_oracle" not in code, f
    out = subprocess.check_output(["ldd", renderer.lib_path()]).decode()
    assert "oracle" not in out


def test_traversal_kernel_votes_are_protected_in_the_built_sass():
    """The persistent kernels' warp votes must be compiled with the BRA.DIV / WARPSYNC slow path (a build with bare VOTEs
    lost ray indices on B200 — DESIGN.md 'A codegen trap'); __graft_entry__.build() enforces it, this test re-checks the
    library that is actually shipped."""
    import __graft_entry__ as g
    g.check_vote_convergence(renderer.lib_path())


def test_bvh_depth_limit_falls_back_to_a_balanced_tree():
    """an unbalanced SAH tree must not make rtx_scene_create refuse a scene the reference accepts.  Boxes that
    shrink geometrically towards a corner make the binned SAH peel off one primitive per level; with the limit the tree is
    rebuilt with object-median splits and stays within it, and every primitive is still in exactly one leaf."""
    lib = renderer.load_library()
    n = 400
    k = np.arange(n, dtype=np.float64)
    size = 1000.0 * 0.9 ** k
    boxes = np.zeros((n, 6), dtype=np.float32)
    boxes[:, 3:] = size[:, None]                                    # [0, s]^3, nested
    depth, nodes = C.c_uint32(), C.c_uint32()
    order = np.zeros(n, dtype=np.uint32)
    assert lib.rtx_bvh_build_probe(boxes.ctypes.data, n, 0, C.byref(depth), C.byref(nodes), order.ctypes.data) == 0
    unlimited = depth.value
    assert sorted(order.tolist()) == list(range(n))
    assert lib.rtx_bvh_build_probe(boxes.ctypes.data, n, 6, C.byref(depth), C.byref(nodes), order.ctypes.data) == 0
    assert depth.value <= 6 and sorted(order.tolist()) == list(range(n))
    assert unlimited > 6, "the test scene no longer provokes a deep SAH tree (%d)" % unlimited
    # a well-behaved set is left alone by the limit
    rng = np.random.default_rng(0)
    c = rng.uniform(-10, 10, (5000, 3)).astype(np.float32)
    b2 = np.concatenate([c - 0.1, c + 0.1], axis=1).astype(np.float32)
    d0, d1 = C.c_uint32(), C.c_uint32()
    lib.rtx_bvh_build_probe(b2.ctypes.data, 5000, 0, C.byref(d0), C.byref(nodes), None)
    lib.rtx_bvh_build_probe(b2.ctypes.data, 5000, 16, C.byref(d1), C.byref(nodes), None)
    assert d0.value == d1.value <= 8
