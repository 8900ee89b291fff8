"""glTF 2.0 (.gltf / .glb) loading with the semantics the reference gets from easy-gltf 1.1 followed by
`Scene::load_gltf` (reference src/scene.rs:722-978) and `get_dyn_image_from_gltf_material` (:980-1124).

  * every scene of the file is walked; node transforms are baked into positions (and normals) of each primitive;
  * one Mesh item per primitive, FULLY DE-INDEXED (3 fresh vertices per triangle, :853-891), uv.y := 1 - v;
  * ids: per primitive `object_id = next_id()` first, then (for a material not seen before) `material_id = next_id()`;
  * material: base = baseColorFactor.rgb, specular = base*0.8, alpha = baseColorFactor.a, reflectivity =
    metallicFactor*0.5, roughness = roughnessFactor/(2*pi); textures: base (RGBA), normal (RGB, a=255),
    metallic -> Reflectivity (B channel as grey), roughness (G channel as grey), emissive -> AmbientEmissive
    (+ ambient_color = emissive factor), occlusion (R channel * strength, truncated to u8);
  * KHR_lights_punctual lights (point intensity / 10, :747) and the first camera.
Stand-in host code (cold path) for a third-party crate; not a copy of reference code.
"""
from __future__ import annotations

import base64
import io
import json
import math
import os
import struct
from typing import Dict, List, Optional

import numpy as np

from .scene_loader import (Item, Light, Material, MeshData, LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_SPOT, SHAPE_MESH,
                           TEX_AMBIENT, TEX_AO, TEX_BASE, TEX_NORMAL, TEX_REFLECTIVITY, TEX_ROUGHNESS, mat_identity)

F = np.float32
_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class _Gltf:
    def __init__(self, path: str):
        self.dir = os.path.dirname(path)
        raw = open(path, "rb").read()
        self.bin: Optional[bytes] = None
        if raw[:4] == b"glTF":
            _, _, length = struct.unpack_from("<III", raw, 0)
            off = 12
            self.doc = None
            while off < length:
                clen, ctype = struct.unpack_from("<II", raw, off)
                chunk = raw[off + 8: off + 8 + clen]
                if ctype == 0x4E4F534A:
                    self.doc = json.loads(chunk.decode("utf-8"))
                elif ctype == 0x004E4942:
                    self.bin = chunk
                off += 8 + clen
        else:
            self.doc = json.loads(raw.decode("utf-8"))
        self._buffers: Dict[int, bytes] = {}

    def buffer(self, i: int) -> bytes:
        if i not in self._buffers:
            b = self.doc["buffers"][i]
            uri = b.get("uri")
            if uri is None:
                self._buffers[i] = self.bin
            elif uri.startswith("data:"):
                self._buffers[i] = base64.b64decode(uri.split(",", 1)[1])
            else:
                self._buffers[i] = open(os.path.join(self.dir, uri), "rb").read()
        return self._buffers[i]

    def view(self, i: int) -> bytes:
        v = self.doc["bufferViews"][i]
        b = self.buffer(v["buffer"])
        off = v.get("byteOffset", 0)
        return b[off: off + v["byteLength"]]

    def accessor(self, i: int) -> np.ndarray:
        a = self.doc["accessors"][i]
        dt, nc, cnt = _COMP[a["componentType"]], _NCOMP[a["type"]], a["count"]
        v = self.doc["bufferViews"][a["bufferView"]]
        b = self.buffer(v["buffer"])
        off = v.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = v.get("byteStride", 0)
        item = np.dtype(dt).itemsize * nc
        if stride and stride != item:
            out = np.zeros((cnt, nc), dtype=dt)
            for k in range(cnt):
                out[k] = np.frombuffer(b, dtype=dt, count=nc, offset=off + k * stride)
        else:
            out = np.frombuffer(b, dtype=dt, count=cnt * nc, offset=off).reshape(cnt, nc).copy()
        if a.get("normalized") and dt != np.float32:
            out = out.astype(np.float32) / float(np.iinfo(dt).max)
        return out

    def image(self, tex_index: int) -> np.ndarray:
        from PIL import Image
        src = self.doc["textures"][tex_index]["source"]
        img = self.doc["images"][src]
        if "bufferView" in img:
            data = self.view(img["bufferView"])
        elif img["uri"].startswith("data:"):
            data = base64.b64decode(img["uri"].split(",", 1)[1])
        else:
            data = open(os.path.join(self.dir, img["uri"]), "rb").read()
        return np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"), dtype=np.uint8)


def _node_matrix(n: dict) -> np.ndarray:
    if "matrix" in n:
        return np.array(n["matrix"], dtype=np.float64).reshape(4, 4).T
    t = np.eye(4); r = np.eye(4); s = np.eye(4)
    if "translation" in n:
        t[:3, 3] = n["translation"]
    if "rotation" in n:
        x, y, z, w = n["rotation"]
        r[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    if "scale" in n:
        s[0, 0], s[1, 1], s[2, 2] = n["scale"]
    return t @ r @ s


def load_gltf(scene, path: str) -> List[int]:
    """Append the file's lights / camera / primitives to `scene` (a scene_loader.Scene); returns the new item ids."""
    g = _Gltf(scene._path(path))
    doc = g.doc
    loaded: List[int] = []
    seen_mats: Dict[int, Material] = {}
    tex_cache: Dict[tuple, str] = {}

    def texture(mat: Material, tex_index: int, kind: str, tt: int, factor: float = 1.0) -> None:
        key = (os.path.normpath(path), tex_index, kind, factor)
        if key not in tex_cache:
            px = g.image(tex_index)
            out = np.zeros_like(px)
            if kind == "rgba":
                out = px.copy()
            elif kind == "rgb":
                out[..., :3] = px[..., :3]; out[..., 3] = 255
            elif kind in ("metallic", "roughness"):                      # easy-gltf: metallic = B, roughness = G
                ch = px[..., 2] if kind == "metallic" else px[..., 1]
                out[...] = ch[..., None]
            elif kind == "occlusion":                                    # R channel * strength, `as u8`
                ch = np.clip(np.trunc(px[..., 0].astype(np.float32) * F(factor)), 0, 255).astype(np.uint8)
                out[...] = ch[..., None]
            name = "%s#%d:%s:%g" % key
            scene.texture_data[name] = np.ascontiguousarray(out)
            tex_cache[key] = name
        mat.textures[tt] = tex_cache[key]

    def make_material(mi: Optional[int]) -> Material:
        gm = doc["materials"][mi] if mi is not None else {}
        pbr = gm.get("pbrMetallicRoughness", {})
        m = Material(id=scene.get_next_id(), name=gm.get("name", "default"))
        bc = [F(x) for x in pbr.get("baseColorFactor", [1, 1, 1, 1])]
        m.base_color = np.array(bc[:3], dtype=F)
        m.specular_color = (m.base_color * F(0.8)).astype(F)
        m.alpha = float(bc[3])
        m.reflectivity = float(F(pbr.get("metallicFactor", 1.0)) * F(0.5))
        m.roughness = float(F(F(F(1.0) / F(math.pi)) / F(2.0)) * F(pbr.get("roughnessFactor", 1.0)))
        if "baseColorTexture" in pbr:
            texture(m, pbr["baseColorTexture"]["index"], "rgba", TEX_BASE)
        if "normalTexture" in gm:
            texture(m, gm["normalTexture"]["index"], "rgb", TEX_NORMAL)
        if "metallicRoughnessTexture" in pbr:
            texture(m, pbr["metallicRoughnessTexture"]["index"], "metallic", TEX_REFLECTIVITY)
        if "emissiveTexture" in gm:
            texture(m, gm["emissiveTexture"]["index"], "rgb", TEX_AMBIENT)
            m.ambient_color = np.array(gm.get("emissiveFactor", [0, 0, 0]), dtype=F)
        if "metallicRoughnessTexture" in pbr:
            texture(m, pbr["metallicRoughnessTexture"]["index"], "roughness", TEX_ROUGHNESS)
        if "occlusionTexture" in gm:
            texture(m, gm["occlusionTexture"]["index"], "occlusion", TEX_AO, float(gm["occlusionTexture"].get("strength", 1.0)))
        scene.materials.append(m)
        return m

    lights_ext = doc.get("extensions", {}).get("KHR_lights_punctual", {}).get("lights", [])

    def visit(ni: int, parent: np.ndarray, lights: list, cams: list, prims: list) -> None:
        n = doc["nodes"][ni]
        m = parent @ _node_matrix(n)
        if "mesh" in n:
            for p in doc["meshes"][n["mesh"]]["primitives"]:
                if p.get("mode", 4) == 4:
                    prims.append((m, p, doc["meshes"][n["mesh"]].get("name", "unknown")))
        if "camera" in n:
            cams.append((m, doc["cameras"][n["camera"]]))
        le = n.get("extensions", {}).get("KHR_lights_punctual")
        if le is not None:
            lights.append((m, lights_ext[le["light"]]))
        for c in n.get("children", []):
            visit(c, m, lights, cams, prims)

    for sc in doc.get("scenes", []):
        lights, cams, prims = [], [], []
        for root in sc.get("nodes", []):
            visit(root, np.eye(4), lights, cams, prims)
        for m, l in lights:                                            # scene.rs:732-787
            pos = (m @ np.array([0, 0, 0, 1.0]))[:3]
            d = (m @ np.array([0, 0, -1.0, 0]))[:3]; d = d / np.linalg.norm(d)
            col = np.array(l.get("color", [1, 1, 1]), dtype=F)
            inten = float(l.get("intensity", 1.0))
            if l["type"] == "point":
                scene.lights.append(Light(scene.get_next_id(), l.get("name", "light"), pos.astype(F), np.array([0, -1, 0], dtype=F), col,
                                          float(F(inten) / F(10.0)), math.pi / 2, LIGHT_POINT))
            elif l["type"] == "directional":
                scene.lights.append(Light(scene.get_next_id(), l.get("name", "light"), np.zeros(3, dtype=F), d.astype(F), col, inten, math.pi / 2,
                                          LIGHT_DIRECTIONAL))
            else:
                scene.lights.append(Light(scene.get_next_id(), l.get("name", "light"), pos.astype(F), d.astype(F), col, inten,
                                          float(l.get("spot", {}).get("outerConeAngle", math.pi / 4)), LIGHT_SPOT))
        if cams:                                                       # :790-821
            m, c = cams[0]
            if c.get("type") == "perspective":
                p = c["perspective"]
                scene.cam.eye_pos = (m @ np.array([0, 0, 0, 1.0]))[:3].astype(F)
                fwd = (m @ np.array([0, 0, -1.0, 0]))[:3]; up = (m @ np.array([0, 1.0, 0, 0]))[:3]
                scene.cam.dir = (fwd / np.linalg.norm(fwd)).astype(F)
                scene.cam.up = (up / np.linalg.norm(up)).astype(F)
                scene.cam.fov = F(p["yfov"])
                scene.cam.clipping_near = float(p.get("znear", 0.001)); scene.cam.clipping_far = float(p.get("zfar", 1000.0))
        for m, p, mesh_name in prims:                                  # :824-975
            object_id = scene.get_next_id()
            pos = g.accessor(p["attributes"]["POSITION"]).astype(np.float64)
            idx = g.accessor(p["indices"]).reshape(-1).astype(np.int64) if "indices" in p else np.arange(len(pos))
            idx = idx[: (len(idx) // 3) * 3]
            wpos = (np.concatenate([pos, np.ones((len(pos), 1))], axis=1) @ m.T)[:, :3]
            verts = wpos[idx].astype(F)
            nrm = np.zeros((0, 3), dtype=F); uvs = np.zeros((0, 2), dtype=F)
            if "NORMAL" in p["attributes"]:
                n0 = g.accessor(p["attributes"]["NORMAL"]).astype(np.float64)
                wn = n0 @ m[:3, :3].T
                wn /= np.maximum(np.linalg.norm(wn, axis=1, keepdims=True), 1e-30)
                nrm = wn[idx].astype(F)
            if "TEXCOORD_0" in p["attributes"]:
                t0 = g.accessor(p["attributes"]["TEXCOORD_0"]).astype(F)
                uv = t0[idx]
                uvs = np.stack([uv[:, 0], F(1.0) - uv[:, 1]], axis=1).astype(F)
            mi = p.get("material")
            key = mi if mi is not None else -1
            if key in seen_mats:
                mat = seen_mats[key]
            else:
                mat = make_material(mi)
                seen_mats[key] = mat
            tri = np.arange(len(verts), dtype=np.uint32).reshape(-1, 3)
            z3 = np.zeros((0, 3), dtype=np.uint32)
            mesh = MeshData(verts, tri, uvs, tri.copy() if len(uvs) else z3, nrm, tri.copy() if len(nrm) else z3)
            scene.items.append(Item(id=object_id, name=mesh_name, shape=SHAPE_MESH, material=mat, trans=mat_identity(), mesh=mesh))
            loaded.append(object_id)
    return loaded
