"""Generate tests/golden/textured_pbr.glb: a few-KB binary glTF that exercises every texture branch of Scene::load_gltf /
get_dyn_image_from_gltf_material (reference src/scene.rs:895-960, 980-1124): base colour (RGBA with an alpha cut-out), normal,
metallic-roughness (B -> Reflectivity, G -> Roughness), occlusion (R * strength), emissive (+ emissiveFactor -> ambient
colour), a second untextured material with baseColorFactor alpha < 1, a node hierarchy with TRS and matrix transforms,
a perspective camera node and a KHR_lights_punctual point light (intensity / 10, :747).  The file is synthetic (no asset of the
reference is involved); its pixel values are kept next to it in textured_pbr_images.npz for the loader test.
    python tests/golden/make_gltf_fixture.py"""
import io
import json
import os
import struct

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(0x617F)
N = 16


def png(a: np.ndarray) -> bytes:
    b = io.BytesIO()
    Image.fromarray(a).save(b, format="PNG")
    return b.getvalue()


def main() -> None:
    y, x = np.mgrid[0:N, 0:N]
    base = np.zeros((N, N, 4), dtype=np.uint8)
    base[..., 0] = 40 + 12 * x; base[..., 1] = 255 - 10 * y; base[..., 2] = rng.integers(0, 255, (N, N)); base[..., 3] = np.where((x // 4 + y // 4) % 2, 255, 96)
    nv = rng.normal(size=(N, N, 3)) * 0.3 + np.array([0, 0, 1.0]); nv /= np.linalg.norm(nv, axis=-1, keepdims=True)
    normal = np.clip((nv * 0.5 + 0.5) * 255, 0, 255).astype(np.uint8)                       # RGB
    mr = rng.integers(0, 255, (N, N, 3)).astype(np.uint8)                                   # R unused, G roughness, B metallic
    occ = rng.integers(60, 255, (N, N, 3)).astype(np.uint8)                                 # R used
    emis = np.zeros((N, N, 3), dtype=np.uint8); emis[4:8, 4:12] = (200, 120, 30)
    images = {"base": base, "normal": normal, "mr": mr, "occ": occ, "emis": emis}

    # geometry: a 6x6 grid patch (72 triangles, indexed u16) and a tetrahedron-ish fan (u32 indices, no uv / no normals)
    g = 6
    u, v = np.meshgrid(np.linspace(0, 1, g + 1), np.linspace(0, 1, g + 1), indexing="ij")
    pos = np.stack([u * 4 - 2, 0.3 * np.sin(u * 3) * np.cos(v * 2), v * 4 - 2], -1).reshape(-1, 3).astype(np.float32)
    nrm = np.tile(np.array([0, 1, 0], dtype=np.float32), (pos.shape[0], 1))
    uv = np.stack([u * 2.0, v * 1.5], -1).reshape(-1, 2).astype(np.float32)                 # > 1: wrap is exercised
    idx = []
    for i in range(g):
        for j in range(g):
            a, b, c, d = i * (g + 1) + j, (i + 1) * (g + 1) + j, (i + 1) * (g + 1) + j + 1, i * (g + 1) + j + 1
            idx += [a, d, c, a, c, b]
    idx = np.array(idx, dtype=np.uint16)
    pos2 = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1], [0.3, 1.2, 0.3]], dtype=np.float32)
    idx2 = np.array([0, 1, 3, 1, 2, 3, 2, 0, 3, 0, 2, 1], dtype=np.uint32)

    blobs, views, accessors = [], [], []

    def add_view(data: bytes, target=None) -> int:
        off = sum(len(b) for b in blobs)
        pad = (-len(data)) % 4
        blobs.append(data + b"\0" * pad)
        v_ = {"buffer": 0, "byteOffset": off, "byteLength": len(data)}
        if target:
            v_["target"] = target
        views.append(v_)
        return len(views) - 1

    def add_acc(arr: np.ndarray, ctype: int, typ: str, target: int) -> int:
        v_ = add_view(arr.tobytes(), target)
        a = {"bufferView": v_, "componentType": ctype, "count": int(arr.shape[0]), "type": typ}
        if typ == "VEC3" and ctype == 5126:
            a["min"] = [float(q) for q in arr.min(axis=0)]; a["max"] = [float(q) for q in arr.max(axis=0)]
        accessors.append(a)
        return len(accessors) - 1
    a_pos, a_nrm, a_uv = add_acc(pos, 5126, "VEC3", 34962), add_acc(nrm, 5126, "VEC3", 34962), add_acc(uv, 5126, "VEC2", 34962)
    a_idx = add_acc(idx, 5123, "SCALAR", 34963)
    a_pos2, a_idx2 = add_acc(pos2, 5126, "VEC3", 34962), add_acc(idx2, 5125, "SCALAR", 34963)
    img_views = [add_view(png(images[k])) for k in ("base", "normal", "mr", "occ", "emis")]
    doc = {
        "asset": {"version": "2.0", "generator": "tests/golden/make_gltf_fixture.py"},
        "extensionsUsed": ["KHR_lights_punctual"],
        "extensions": {"KHR_lights_punctual": {"lights": [{"type": "point", "color": [1.0, 0.9, 0.8], "intensity": 900.0, "name": "lamp"}]}},
        "scene": 0,
        "scenes": [{"nodes": [0, 3, 4]}],
        "nodes": [
            {"name": "root", "translation": [0.0, -1.0, -6.0], "rotation": [0.0, 0.3826834, 0.0, 0.9238795], "scale": [1.2, 1.0, 1.2], "children": [1, 2]},
            {"name": "patch", "mesh": 0},
            {"name": "tetra", "mesh": 1, "matrix": [1.5, 0, 0, 0, 0, 1.5, 0, 0, 0, 0, 1.5, 0, -0.5, 0.4, 0.2, 1]},
            {"name": "cam", "camera": 0, "translation": [0.0, 2.0, 1.0], "rotation": [-0.1736482, 0.0, 0.0, 0.9848078]},
            {"name": "lamp", "translation": [2.0, 4.0, -3.0], "extensions": {"KHR_lights_punctual": {"light": 0}}}],
        "cameras": [{"type": "perspective", "perspective": {"yfov": 0.9, "znear": 0.05, "zfar": 200.0, "aspectRatio": 1.5}}],
        "meshes": [{"name": "patch", "primitives": [{"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "TEXCOORD_0": a_uv}, "indices": a_idx, "material": 0}]},
                   {"name": "tetra", "primitives": [{"attributes": {"POSITION": a_pos2}, "indices": a_idx2, "material": 1}]}],
        "materials": [
            {"name": "pbr_all_maps", "pbrMetallicRoughness": {"baseColorFactor": [0.9, 0.8, 0.7, 1.0], "metallicFactor": 0.6, "roughnessFactor": 0.5,
                                                              "baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 2}},
             "normalTexture": {"index": 1}, "occlusionTexture": {"index": 3, "strength": 0.75}, "emissiveTexture": {"index": 4}, "emissiveFactor": [1.0, 0.5, 0.25]},
            {"name": "glass", "pbrMetallicRoughness": {"baseColorFactor": [0.2, 0.6, 0.9, 0.4], "metallicFactor": 0.4, "roughnessFactor": 0.0}}],
        "textures": [{"source": i} for i in range(5)],
        "images": [{"bufferView": v_, "mimeType": "image/png"} for v_ in img_views],
        "accessors": accessors, "bufferViews": views,
        "buffers": [{"byteLength": sum(len(b) for b in blobs)}],
    }
    js = json.dumps(doc, separators=(",", ":")).encode()
    js += b" " * ((-len(js)) % 4)
    binc = b"".join(blobs)
    out = struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(binc)) + struct.pack("<II", len(js), 0x4E4F534A) + js + \
        struct.pack("<II", len(binc), 0x004E4942) + binc
    open(os.path.join(HERE, "textured_pbr.glb"), "wb").write(out)
    np.savez_compressed(os.path.join(HERE, "textured_pbr_images.npz"), **images)
    print("wrote textured_pbr.glb (%d bytes)" % len(out))


if __name__ == "__main__":
    main()
