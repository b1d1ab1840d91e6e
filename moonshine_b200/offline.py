"""`offline <in.glb> <skybox.exr> <out.exr> [spp]` — the reference's headless entry point on the B200 library.

Mirrors /root/reference/offline/main.zig:27-50,80-203 and the loaders it calls:
  * GLB + PNG import with the reference's exact rules — engine/hrtsystem/World.zig:44-228 (material mapping),
    :233-363 (one instance per node-with-mesh, one mesh per primitive, `sampled` <=> material name starts with
    "Emitter", rows (x, z, y) of the global transform = Y-up -> Z-up), engine/hrtsystem/Camera.zig:26-51 (first camera
    node), deps/zgltf/src/{main.zig:236-276, helpers.zig:17-101} (TRS recompose, parent chain, f32 matrix product);
  * EXR: engine/fileformats/exr.zig:137-231 — skybox loaded as RGBA f32, result saved as three FLOAT channels named
    B, G, R, uncompressed scanlines (tinyexr defaults after InitEXRHeader).
The reference hard-codes a 1280x720 sensor (offline/main.zig:48, "TODO: cli"); `--extent WxH` overrides it.

`write_glb` is the inverse of the importer for scenes expressible in glTF: it lets the procedural BASELINE scenes of
scenes.py be fed to the reference's own `offline` binary on a Vulkan-RT machine for cross-checking.
"""
from __future__ import annotations

import io
import json
import math
import os
import struct
import sys
import zlib
from typing import List, Optional, Tuple

import numpy as np

from . import scenes as S

# ---------------------------------------------------------------------------------------------------------------
# EXR
# ---------------------------------------------------------------------------------------------------------------
_EXR_MAGIC = 20000630


def write_exr(path: str, rgba: np.ndarray) -> None:
    """Three FLOAT channels B, G, R (exr.zig:187-200), NO_COMPRESSION, increasing-Y scanlines."""
    h, w = rgba.shape[:2]
    img = np.ascontiguousarray(rgba[..., :3], np.float32)

    def attr(name, typ, value):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(value)) + value

    chlist = b"".join(n + b"\0" + struct.pack("<iBxxxii", 2, 0, 1, 1) for n in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = b"".join([
        attr("channels", "chlist", chlist), attr("compression", "compression", b"\0"), attr("dataWindow", "box2i", box),
        attr("displayWindow", "box2i", box), attr("lineOrder", "lineOrder", b"\0"), attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)),
        attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)), attr("screenWindowWidth", "float", struct.pack("<f", 1.0))]) + b"\0"
    head = struct.pack("<ii", _EXR_MAGIC, 2) + header
    line_bytes = 3 * w * 4
    offsets = [len(head) + 8 * h + y * (8 + line_bytes) for y in range(h)]
    with open(path, "wb") as f:
        f.write(head)
        f.write(struct.pack(f"<{h}Q", *offsets))
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            f.write(img[y, :, 2].tobytes() + img[y, :, 1].tobytes() + img[y, :, 0].tobytes())  # B, G, R planes


def read_exr(path: str) -> np.ndarray:
    """Scanline EXR -> (H, W, 4) f32 RGBA like tinyexr's LoadEXR (missing A = 1). Supports NONE/ZIPS/ZIP, HALF/FLOAT."""
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", data, 0)
    if magic != _EXR_MAGIC:
        raise ValueError("not an OpenEXR file")
    if version & 0x200:
        raise ValueError("tiled EXR not supported")
    pos, attrs = 8, {}
    while data[pos] != 0:
        e = data.index(b"\0", pos); name = data[pos:e].decode(); pos = e + 1
        e = data.index(b"\0", pos); typ = data[pos:e].decode(); pos = e + 1
        (size,) = struct.unpack_from("<i", data, pos); pos += 4
        attrs[name] = (typ, data[pos:pos + size]); pos += size
    pos += 1
    channels, cp = [], 0
    cl = attrs["channels"][1]
    while cl[cp] != 0:
        e = cl.index(b"\0", cp); name = cl[cp:e].decode(); cp = e + 1
        ptype, _pl, xs, ys = struct.unpack_from("<iBxxxii", cl, cp); cp += 16
        if xs != 1 or ys != 1 or ptype not in (1, 2):
            raise ValueError("unsupported EXR channel layout")
        channels.append((name, ptype))
    comp = attrs["compression"][1][0]
    if comp not in (0, 2, 3):
        raise ValueError(f"EXR compression {comp} not supported (NONE, ZIPS, ZIP only)")
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    lines_per_block = 16 if comp == 3 else 1
    nblocks = (h + lines_per_block - 1) // lines_per_block
    offsets = struct.unpack_from(f"<{nblocks}Q", data, pos)
    planes = {n: np.zeros((h, w), np.float32) for n, _ in channels}
    bpp = [2 if t == 1 else 4 for _, t in channels]
    for off in offsets:
        y, size = struct.unpack_from("<ii", data, off)
        raw = data[off + 8:off + 8 + size]
        nl = min(lines_per_block, y1 - y + 1)
        expect = nl * w * sum(bpp)
        if comp and size < expect:
            d = np.frombuffer(zlib.decompress(raw), np.uint8).astype(np.int32)
            d = np.cumsum(np.concatenate([[d[0]], d[1:] - 128])).astype(np.uint8) if len(d) else d.astype(np.uint8)
            half = (len(d) + 1) // 2
            out = np.empty(len(d), np.uint8)
            out[0::2] = d[:half]; out[1::2] = d[half:]
            raw = out.tobytes()
        p = 0
        for ln in range(nl):
            for (n, t), b in zip(channels, bpp):
                row = np.frombuffer(raw, np.float16 if t == 1 else np.float32, w, p).astype(np.float32)
                planes[n][y - y0 + ln] = row
                p += w * b
    out = np.ones((h, w, 4), np.float32)
    for i, n in enumerate("RGB"):
        if n in planes:
            out[..., i] = planes[n]
        elif "Y" in planes:
            out[..., i] = planes["Y"]
    if "A" in planes:
        out[..., 3] = planes["A"]
    return out


# ---------------------------------------------------------------------------------------------------------------
# glTF (GLB) import
# ---------------------------------------------------------------------------------------------------------------
_F = np.float32
_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def _mat_mul(left, right):  # zgltf helpers.zig:82-101; matrices are [column][row], f32 accumulation in column order
    res = np.zeros((4, 4), _F)
    for c in range(4):
        for r in range(4):
            s = _F(0)
            for k in range(4):
                s = _F(s + _F(left[k][r] * right[c][k]))
            res[c][r] = s
    return res


def _recompose(t, q, s):  # helpers.zig:17-78, T * R * S
    T = np.eye(4, dtype=_F); T[3][0:3] = t
    x, y, z, w = (_F(v) for v in q)
    xx, yy, zz, xy, xz, yz, wx, wy, wz = x * x, y * y, z * z, x * y, x * z, y * z, w * x, w * y, w * z
    R = np.eye(4, dtype=_F)
    R[0][0] = _F(1) - _F(2) * (yy + zz); R[0][1] = _F(2) * (xy + wz); R[0][2] = _F(2) * (xz - wy)
    R[1][0] = _F(2) * (xy - wz); R[1][1] = _F(1) - _F(2) * (xx + zz); R[1][2] = _F(2) * (yz + wx)
    R[2][0] = _F(2) * (xz + wy); R[2][1] = _F(2) * (yz - wx); R[2][2] = _F(1) - _F(2) * (xx + yy)
    Sc = np.eye(4, dtype=_F); Sc[0][0], Sc[1][1], Sc[2][2] = s
    return _mat_mul(T, _mat_mul(R, Sc))


class _Glb:
    def __init__(self, buf: bytes):
        magic, _ver, _len = struct.unpack_from("<III", buf, 0)
        if magic != 0x46546C67:
            raise ValueError("OnlySupportsGlbInput: not a GLB file")
        pos, self.json, self.bin = 12, None, b""
        while pos < len(buf):
            clen, ctype = struct.unpack_from("<II", buf, pos)
            chunk = buf[pos + 8:pos + 8 + clen]
            if ctype == 0x4E4F534A:
                self.json = json.loads(chunk.decode("utf-8"))
            elif ctype == 0x004E4942:
                self.bin = chunk
            pos += 8 + clen
        g = self.json
        self.nodes = g.get("nodes", [])
        self.parent = [None] * len(self.nodes)
        for i, n in enumerate(self.nodes):
            for c in n.get("children", []):
                self.parent[c] = i

    def view(self, idx) -> bytes:
        bv = self.json["bufferViews"][idx]
        o = bv.get("byteOffset", 0)
        return self.bin[o:o + bv["byteLength"]]

    def accessor(self, idx) -> np.ndarray:
        a = self.json["accessors"][idx]
        bv = self.json["bufferViews"][a["bufferView"]]
        dt, nc = np.dtype(_COMP[a["componentType"]]), _NCOMP[a["type"]]
        start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * nc
        if stride == dt.itemsize * nc:
            arr = np.frombuffer(self.bin, dt, a["count"] * nc, start).reshape(a["count"], nc)
        else:
            arr = np.stack([np.frombuffer(self.bin, dt, nc, start + i * stride) for i in range(a["count"])])
        return arr

    def local(self, i):  # main.zig:245-262
        n = self.nodes[i]
        if "matrix" in n:
            return np.array(n["matrix"], _F).reshape(4, 4)
        return _recompose(np.array(n.get("translation", [0, 0, 0]), _F), np.array(n.get("rotation", [0, 0, 0, 1]), _F), np.array(n.get("scale", [1, 1, 1]), _F))

    def global_transform(self, i):  # main.zig:264-276
        m = self.local(i)
        p = self.parent[i]
        while p is not None:
            m = _mat_mul(self.local(p), m)
            p = self.parent[p]
        return m

    def png_rgb(self, texture_index) -> np.ndarray:
        from PIL import Image
        img = self.json["images"][self.json["textures"][texture_index]["source"]]
        if img.get("mimeType") != "image/png":
            raise ValueError("only image/png textures are supported (World.zig:50)")
        im = Image.open(io.BytesIO(self.view(img["bufferView"])))
        if im.mode != "RGB":
            raise ValueError("PNG textures must be RGB8 (World.zig:59: img.pixels.rgb24)")
        return np.asarray(im, np.uint8)


def _zup(m) -> np.ndarray:
    """World.zig:341-345 / Camera.zig:33-41: rows (x, z, y) of the column-major global matrix."""
    return np.array([[m[0][0], m[1][0], m[2][0], m[3][0]], [m[0][2], m[1][2], m[2][2], m[3][2]], [m[0][1], m[1][1], m[2][1], m[3][1]]], _F)


def load_glb(path: str, extent=(1280, 720)) -> S.Scene:
    g = _Glb(open(path, "rb").read())
    J = g.json
    sc = S.Scene(name=path, extent=tuple(extent))
    # camera: first node with a camera (Camera.zig:26-51)
    cam_node = next((i for i, n in enumerate(g.nodes) if "camera" in n), None)
    if cam_node is None:
        raise ValueError("NoCameraInGlb")
    T = _zup(g.global_transform(cam_node))
    mul_vec = lambda v: np.array([(T[r][0] * v[0] + T[r][1] * v[1] + T[r][2] * v[2]) for r in range(3)], _F)
    fwd = mul_vec(np.array([0, 0, -1], _F))
    fwd = (fwd / _F(np.sqrt(_F(fwd[0] * fwd[0] + fwd[1] * fwd[1] + fwd[2] * fwd[2])))).astype(_F)
    yfov = J["cameras"][g.nodes[cam_node]["camera"]]["perspective"]["yfov"]
    sc.lens = S.Lens(tuple(float(T[r][3]) for r in range(3)), tuple(float(x) for x in fwd), tuple(float(x) for x in mul_vec(np.array([0, 1, 0], _F))), float(_F(yfov)), 0.0, 1.0)
    # materials (World.zig:44-228)
    for m in J.get("materials", []):
        pbr = m.get("pbrMetallicRoughness", {})
        ext = m.get("extensions", {})
        if "normalTexture" in m:
            rgb = g.png_rgb(m["normalTexture"]["index"])
            normal = sc.raw_texture(np.ascontiguousarray(rgb[..., :2]), S.TEX_RG8_UNORM)
        else:
            normal = sc.solid2(0.5, 0.5)
        if "emissiveTexture" in m:
            rgb = g.png_rgb(m["emissiveTexture"]["index"])
            emissive = sc.raw_texture(np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], -1), S.TEX_RGBA8_SRGB)
        else:
            ef = np.array(m.get("emissiveFactor", [0, 0, 0]), _F)
            strength = _F(ext.get("KHR_materials_emissive_strength", {}).get("emissiveStrength", 1.0))
            e = (ef * strength).astype(_F)
            emissive = sc.solid3(float(e[0]), float(e[1]), float(e[2]))
        ior = float(_F(ext.get("KHR_materials_ior", {}).get("ior", 1.5)))
        if float(_F(ext.get("KHR_materials_transmission", {}).get("transmissionFactor", 0.0))) == 1.0:
            sc.add_material(type=S.MAT_GLASS, normal=normal, emissive=emissive, ior=ior)
            continue
        if "baseColorTexture" in pbr:
            rgb = g.png_rgb(pbr["baseColorTexture"]["index"])
            color = sc.raw_texture(np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], -1), S.TEX_RGBA8_SRGB)
        else:
            c = np.array(pbr.get("baseColorFactor", [1, 1, 1, 1]), _F)
            color = sc.solid3(float(c[0]), float(c[1]), float(c[2]))
        metallic, roughness = float(_F(pbr.get("metallicFactor", 1.0))), float(_F(pbr.get("roughnessFactor", 1.0)))
        if "metallicRoughnessTexture" in pbr:
            rgb = g.png_rgb(pbr["metallicRoughnessTexture"]["index"])
            # sic: metalness <- R, roughness <- G (World.zig:172-174); glTF itself stores roughness in G, metalness in B
            met = sc.raw_texture(np.ascontiguousarray(rgb[..., 0]), S.TEX_R8_UNORM)
            rou = sc.raw_texture(np.ascontiguousarray(rgb[..., 1]), S.TEX_R8_UNORM)
            sc.add_material(type=S.MAT_STANDARD_PBR, normal=normal, emissive=emissive, color=color, metalness=met, roughness=rou, ior=ior)
        elif metallic == 0.0 and roughness == 1.0:
            sc.add_material(type=S.MAT_LAMBERT, normal=normal, emissive=emissive, color=color)
        elif metallic == 1.0 and roughness == 0.0:
            sc.add_material(type=S.MAT_PERFECT_MIRROR, normal=normal, emissive=emissive)
        else:
            sc.add_material(type=S.MAT_STANDARD_PBR, normal=normal, emissive=emissive, color=color, metalness=sc.solid1(metallic), roughness=sc.solid1(roughness), ior=ior)
    # nodes with meshes (World.zig:262-349)
    for i, n in enumerate(g.nodes):
        if "mesh" not in n:
            continue
        geoms = []
        for prim in J["meshes"][n["mesh"]]["primitives"]:
            idx = g.accessor(prim["indices"]).reshape(-1, 3).astype(np.uint32)
            pos = nrm = uv = None
            for name, acc in prim["attributes"].items():
                if name == "POSITION":
                    pos = g.accessor(acc).astype(_F)
                elif name.startswith("TEXCOORD"):
                    uv = g.accessor(acc).astype(_F)
                elif name == "NORMAL":
                    nrm = g.accessor(acc).astype(_F)
                else:
                    raise ValueError(f"UnhandledAttribute: {name} (World.zig:308-311)")
            mat = prim["material"]
            mesh = sc.add_mesh(pos, idx, normals=nrm, texcoords=uv)
            geoms.append((mesh, mat, J["materials"][mat].get("name", "").startswith("Emitter")))
        sc.add_instance(geoms, transform=_zup(g.global_transform(i)))
    sc.constants = S.Constants()  # offline/main.zig:106-111
    return sc


# ---------------------------------------------------------------------------------------------------------------
# GLB export (inverse of the importer)
# ---------------------------------------------------------------------------------------------------------------
def write_glb(scene: S.Scene, path: str) -> None:
    from PIL import Image
    bin_parts: List[bytes] = []
    views, accessors, images, textures, materials, meshes, nodes = [], [], [], [], [], [], []

    def add_view(b: bytes) -> int:
        off = sum(len(p) for p in bin_parts)
        pad = (-len(b)) % 4
        bin_parts.append(b + b"\0" * pad)
        views.append({"buffer": 0, "byteOffset": off, "byteLength": len(b)})
        return len(views) - 1

    def add_accessor(arr: np.ndarray, ctype: int, typ: str) -> int:
        a = {"bufferView": add_view(np.ascontiguousarray(arr).tobytes()), "componentType": ctype, "count": int(arr.shape[0]), "type": typ}
        if typ == "VEC3" and ctype == 5126:
            a["min"], a["max"] = [float(x) for x in arr.min(0)], [float(x) for x in arr.max(0)]
        accessors.append(a)
        return len(accessors) - 1

    def add_png(rgb: np.ndarray) -> int:
        b = io.BytesIO()
        Image.fromarray(np.ascontiguousarray(rgb, np.uint8), "RGB").save(b, "PNG")
        images.append({"bufferView": add_view(b.getvalue()), "mimeType": "image/png"})
        textures.append({"source": len(images) - 1})
        return len(textures) - 1

    def fl(x):
        return float(np.float32(x))

    def solid(tex: S.Texture):
        return None if (tex.width, tex.height) != (1, 1) else np.asarray(tex.data, np.float32).reshape(-1)

    T = scene.textures
    emitters = {g[1] for inst in scene.instances for g in inst.geometries if g[2]}
    for mi, m in enumerate(scene.materials):
        out = {"name": ("Emitter " if mi in emitters else "") + f"material{mi}", "pbrMetallicRoughness": {}, "extensions": {}}
        n = T[m.normal]
        if solid(n) is None:
            d = np.asarray(n.data, np.uint8).reshape(n.height, n.width, 2)
            out["normalTexture"] = {"index": add_png(np.concatenate([d, np.full((n.height, n.width, 1), 255, np.uint8)], -1))}
        e = T[m.emissive]
        if solid(e) is None:
            out["emissiveTexture"] = {"index": add_png(np.asarray(e.data, np.uint8).reshape(e.height, e.width, 4)[..., :3])}
        else:
            out["emissiveFactor"] = [fl(v) for v in solid(e)[:3]]
        if m.type == S.MAT_GLASS:
            out["extensions"]["KHR_materials_transmission"] = {"transmissionFactor": 1.0}
            out["extensions"]["KHR_materials_ior"] = {"ior": fl(m.ior)}
            materials.append(out)
            continue
        p = out["pbrMetallicRoughness"]
        if m.type != S.MAT_PERFECT_MIRROR:
            c = T[m.color]
            if solid(c) is None:
                if c.fmt != S.TEX_RGBA8_SRGB:
                    raise ValueError("only sRGB8 colour textures can be exported")
                p["baseColorTexture"] = {"index": add_png(np.asarray(c.data, np.uint8).reshape(c.height, c.width, 4)[..., :3])}
            else:
                p["baseColorFactor"] = [fl(v) for v in solid(c)[:3]] + [1.0]
        if m.type == S.MAT_LAMBERT:
            p["metallicFactor"], p["roughnessFactor"] = 0.0, 1.0
        elif m.type == S.MAT_PERFECT_MIRROR:
            p["metallicFactor"], p["roughnessFactor"] = 1.0, 0.0
        else:
            out["extensions"]["KHR_materials_ior"] = {"ior": fl(m.ior)}
            me, ro = T[m.metalness], T[m.roughness]
            if solid(me) is None or solid(ro) is None:
                if (me.width, me.height) != (ro.width, ro.height) and solid(me) is None and solid(ro) is None:
                    raise ValueError("metalness/roughness textures must share a size to be packed into one PNG")
                w, h = (me.width, me.height) if solid(me) is None else (ro.width, ro.height)
                pack = np.zeros((h, w, 3), np.uint8)
                pack[..., 0] = np.asarray(me.data, np.uint8).reshape(h, w) if solid(me) is None else int(round(float(solid(me)[0]) * 255))
                pack[..., 1] = np.asarray(ro.data, np.uint8).reshape(h, w) if solid(ro) is None else int(round(float(solid(ro)[0]) * 255))
                p["metallicRoughnessTexture"] = {"index": add_png(pack)}  # R = metalness, G = roughness (World.zig:172-174)
            else:
                p["metallicFactor"], p["roughnessFactor"] = fl(solid(me)[0]), fl(solid(ro)[0])
                if (p["metallicFactor"], p["roughnessFactor"]) in ((0.0, 1.0), (1.0, 0.0)):
                    raise ValueError("a constant StandardPBR with (metal, rough) = (0,1) or (1,0) would be imported as Lambert/Mirror")
        if not out["extensions"]:
            del out["extensions"]
        materials.append(out)
    for inst in scene.instances:
        prims = []
        for (mesh_i, mat_i, _s) in inst.geometries:
            me = scene.meshes[mesh_i]
            if len(me.positions) > 65535:
                raise ValueError("primitives must have <= 65535 vertices (u16 indices, World.zig:278)")
            attrs = {"POSITION": add_accessor(me.positions.astype(np.float32), 5126, "VEC3")}
            if me.normals is not None:
                attrs["NORMAL"] = add_accessor(me.normals.astype(np.float32), 5126, "VEC3")
            if me.texcoords is not None:
                attrs["TEXCOORD_0"] = add_accessor(me.texcoords.astype(np.float32), 5126, "VEC2")
            prims.append({"attributes": attrs, "indices": add_accessor(me.indices.astype(np.uint16).reshape(-1), 5123, "SCALAR"), "material": mat_i})
        meshes.append({"primitives": prims})
        t = inst.transform
        rows = [t[0], t[2], t[1], np.array([0, 0, 0, 1], np.float32)]  # undo the (x, z, y) row swizzle
        nodes.append({"mesh": len(meshes) - 1, "matrix": [fl(rows[r][c]) for c in range(4) for r in range(4)]})
    # camera node: T' columns = (side, up, -forward, origin) in engine space, then undo the row swizzle
    L = scene.lens
    f, up, o = np.array(L.forward, np.float32), np.array(L.up, np.float32), np.array(L.origin, np.float32)
    side = np.cross(f, up).astype(np.float32)
    Tc = np.stack([side, up, -f, o], 1)  # 3x4, engine rows (x, y, z)
    rows = [Tc[0], Tc[2], Tc[1], np.array([0, 0, 0, 1], np.float32)]
    nodes.append({"camera": 0, "matrix": [fl(rows[r][c]) for c in range(4) for r in range(4)]})
    blob = b"".join(bin_parts)
    J = {"asset": {"version": "2.0", "generator": "moonshine_b200.offline.write_glb"}, "scene": 0, "scenes": [{"nodes": list(range(len(nodes)))}], "nodes": nodes,
         "meshes": meshes, "materials": materials, "accessors": accessors, "bufferViews": views, "buffers": [{"byteLength": len(blob)}],
         "cameras": [{"type": "perspective", "perspective": {"yfov": fl(L.vfov), "znear": 0.01}}],
         "extensionsUsed": ["KHR_materials_ior", "KHR_materials_transmission", "KHR_materials_emissive_strength"]}
    if images:
        J["images"], J["textures"] = images, textures
    js = json.dumps(J).encode()
    js += b" " * ((-len(js)) % 4)
    with open(path, "wb") as fo:
        fo.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(blob)))
        fo.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        fo.write(struct.pack("<II", len(blob), 0x004E4942) + blob)


# ---------------------------------------------------------------------------------------------------------------
# CLI (offline/main.zig:27-50)
# ---------------------------------------------------------------------------------------------------------------
def main(argv: Optional[List[str]] = None) -> int:
    """offline <in.glb> <skybox.exr> <out.exr> [spp=16] (offline/main.zig:27-50) [--extent WxH] [--device N] [--gpus N].

    --gpus N (one node): the process re-launches itself as N ranks (RANK / WORLD_SIZE / LOCAL_RANK, the torchrun convention, so
    `torchrun --nproc-per-node N -m moonshine_b200.offline ...` works as well); every rank loads the scene, renders all samples
    of its own pixel tiles on its own GPU, and one ncclReduce inside the library assembles the frame on rank 0, which writes it.
    """
    import time
    argv = list(sys.argv[1:] if argv is None else argv)
    extent, device, gpus = (1280, 720), None, 1
    while "--extent" in argv:
        i = argv.index("--extent"); w, h = argv[i + 1].lower().split("x"); extent = (int(w), int(h)); del argv[i:i + 2]
    while "--device" in argv:
        i = argv.index("--device"); device = int(argv[i + 1]); del argv[i:i + 2]
    while "--gpus" in argv:
        i = argv.index("--gpus"); gpus = int(argv[i + 1]); del argv[i:i + 2]
    if len(argv) < 3:
        print("usage: offline <in.glb> <skybox.exr> <out.exr> [spp=16] [--extent WxH] [--device N] [--gpus N]", file=sys.stderr)
        return 2  # error.BadArgs
    in_path, sky_path, out_path = argv[:3]
    if not in_path.endswith(".glb"):
        raise SystemExit("OnlySupportsGlbInput")
    if not sky_path.endswith(".exr"):
        raise SystemExit("OnlySupportsExrSkybox")
    if not out_path.endswith(".exr"):
        raise SystemExit("OnlySupportsExrOutput")
    spp = int(argv[3]) if len(argv) > 3 else 16
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if gpus > 1 and world == 1:
        return _launch_ranks(gpus, [in_path, sky_path, out_path, str(spp), "--extent", f"{extent[0]}x{extent[1]}"])
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", 0))
    from . import hydra
    t0 = time.perf_counter()
    scene = load_glb(in_path, extent)
    scene.background = read_exr(sky_path)
    m = hydra.HdMoonshine(device)
    S.load_scene(m, scene)
    if world > 1:
        tag = os.environ.get("MSNE_JOB_TAG", os.environ.get("MASTER_PORT", "0"))
        m.CommInit(hydra.exchange_unique_id(rank, world, tag), rank, world)
        if rank == 0:
            hydra.retire_unique_id(tag)      # CommInit is collective: every rank has read the id by now
    sensor, lens = m.CreateSensor(*extent), m.CreateLens(scene.lens)
    if rank == 0:
        print(f"{time.perf_counter() - t0:.3f} seconds to load world")
    t0 = time.perf_counter()
    m.RenderLaunches(sensor, lens, spp)
    if world > 1:
        m.CommReduceSensor(sensor, 0)     # rank 0's host buffer receives the assembled frame
        img = m.GetSensorData(sensor)
    else:
        img = m.ReadSensor(sensor)
    st = m.GetStats()
    rays = int(st.rays_closest + st.rays_shadow)
    if world > 1:
        rays = int(m.CommAllReduceU64([rays])[0])
    if rank == 0:
        print(f"{time.perf_counter() - t0:.3f} seconds to render ({rays / max(st.device_ms, 1e-9) / 1e3:.1f} Mrays/s on {world} GPU{'s' if world > 1 else ''})")
        write_exr(out_path, img)
    m.Destroy()
    return 0


def _launch_ranks(n: int, args: List[str]) -> int:
    """One process per GPU on this node (the torchrun environment convention, without needing torch)."""
    import subprocess
    import uuid
    tag = uuid.uuid4().hex[:12]
    procs = []
    for r in range(n):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(n), MSNE_JOB_TAG=tag)
        procs.append(subprocess.Popen([sys.executable, "-m", "moonshine_b200.offline"] + args, env=env,
                                      cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    codes = [p.wait() for p in procs]
    return max(codes)


if __name__ == "__main__":
    sys.exit(main())
