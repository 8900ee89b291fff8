"""Wavefront OBJ/MTL parse with the semantics the reference gets from tobj 4.0 called with
`LoadOptions { triangulate: true, single_index: true }` (reference src/scene.rs:1130-1137).

  * a model is emitted at every `o` / `g` / `usemtl` change that has pending faces;
  * polygons are fan-triangulated: (0, i, i+1);
  * single_index: every distinct (v, vt, vn) triple becomes one output vertex, numbered in order
    of first use, per model; positions / texcoords / normals are re-emitted per output vertex and
    the separate texcoord / normal index lists stay empty;
  * negative indices are relative to the current end of the respective list.
Host-side, cold path; stand-in for a third-party crate (tobj), not a copy of reference code.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple


def _parse_mtl(path: str) -> Tuple[List[dict], Dict[str, int]]:
    mats: List[dict] = []
    index: Dict[str, int] = {}
    cur: Optional[dict] = None
    if not os.path.exists(path):
        return mats, index
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            parts = line.split()
            key, args = parts[0], parts[1:]
            if key == "newmtl":
                cur = {"name": " ".join(args)}
                index[cur["name"]] = len(mats)
                mats.append(cur)
            elif cur is None:
                continue
            elif key in ("Ka", "Kd", "Ks") and len(args) >= 3:
                cur[key] = [float(a) for a in args[:3]]
            elif key in ("Ns", "Ni", "d") and args:
                cur[key] = float(args[0])
            elif key == "illum" and args:
                cur["illum"] = int(float(args[0]))
            elif key in ("map_Kd", "map_Ka", "map_Ks", "map_d", "map_Ns") and args:
                cur[key] = args[-1]
            elif key in ("map_Bump", "map_bump", "bump") and args:
                cur["map_Bump"] = args[-1]
    return mats, index


def load_obj(path: str):
    positions: List[Tuple[float, float, float]] = []
    texcoords: List[Tuple[float, float]] = []
    normals: List[Tuple[float, float, float]] = []
    faces: List[List[Tuple[int, int, int]]] = []
    models: List[dict] = []
    mtls: List[dict] = []
    mtl_index: Dict[str, int] = {}
    name = "unnamed_object"
    mat_id: Optional[int] = None

    def export():
        nonlocal faces
        if not faces:
            return
        remap: Dict[Tuple[int, int, int], int] = {}
        out_p: List[float] = []
        out_t: List[float] = []
        out_n: List[float] = []
        out_i: List[int] = []

        def add(v):
            i = remap.get(v)
            if i is None:
                i = len(remap)
                remap[v] = i
                out_p.extend(positions[v[0]])
                if v[1] >= 0:
                    out_t.extend(texcoords[v[1]])
                if v[2] >= 0:
                    out_n.extend(normals[v[2]])
            out_i.append(i)

        for f in faces:
            if len(f) < 3:
                continue
            for k in range(1, len(f) - 1):
                add(f[0]); add(f[k]); add(f[k + 1])
        models.append({"name": name, "positions": out_p, "texcoords": out_t, "normals": out_n,
                       "indices": out_i, "material_id": mat_id})
        faces = []

    def fix(i: int, n: int) -> int:
        return i - 1 if i > 0 else n + i

    with open(path, "r", errors="replace") as fh:
        for line in fh:
            if not line or line[0] == "#":
                continue
            parts = line.split()
            if not parts:
                continue
            key = parts[0]
            if key == "v":
                positions.append((float(parts[1]), float(parts[2]), float(parts[3])))
            elif key == "vt":
                texcoords.append((float(parts[1]), float(parts[2]) if len(parts) > 2 else 0.0))
            elif key == "vn":
                normals.append((float(parts[1]), float(parts[2]), float(parts[3])))
            elif key == "f":
                face = []
                for tok in parts[1:]:
                    sp = tok.split("/")
                    vi = fix(int(sp[0]), len(positions))
                    ti = fix(int(sp[1]), len(texcoords)) if len(sp) > 1 and sp[1] else -1
                    ni = fix(int(sp[2]), len(normals)) if len(sp) > 2 and sp[2] else -1
                    face.append((vi, ti, ni))
                faces.append(face)
            elif key in ("o", "g"):
                export()
                name = " ".join(parts[1:]) if len(parts) > 1 else "unnamed_object"
            elif key == "mtllib":
                for lib in parts[1:]:
                    m, _ = _parse_mtl(os.path.join(os.path.dirname(path), lib))
                    for mm in m:
                        mtl_index[mm["name"]] = len(mtls)
                        mtls.append(mm)
            elif key == "usemtl":
                new_id = mtl_index.get(" ".join(parts[1:]))
                if new_id != mat_id:
                    export()
                mat_id = new_id
    export()
    return models, mtls
