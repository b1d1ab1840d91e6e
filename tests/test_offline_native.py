"""The native `offline` front end (moonshine_b200/csrc/offline_main.cpp: the reference's offline/main.zig over the C ABI) imports a GLB
by the reference's rules on its own - JSON, accessors, zgltf's binary32 transform chain, PNG, material mapping. Without a GPU: its
`--dump` digest of the imported scene must equal, field by field and bit by bit, the digest of what the Python front end
(moonshine_b200/offline.py, whose renders the GPU tests compare with the oracle) imports from the same file."""
import os
import subprocess

import numpy as np
import pytest

from moonshine_b200 import offline, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "moonshine_b200", "offline")


def _exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["bash", os.path.join(ROOT, "moonshine_b200", "csrc", "build.sh")])
    return EXE


def _fnv(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _digest(sc: scenes.Scene) -> str:
    def arr(a, dt):
        return b"" if a is None else np.ascontiguousarray(a, dt).tobytes()
    out = [f"textures {len(sc.textures)} materials {len(sc.materials)} meshes {len(sc.meshes)} instances {len(sc.instances)}",
           f"lens {_fnv(sc.lens.as_array().tobytes()):016x}"]
    for t in sc.textures:
        out.append(f"tex {t.width} {t.height} {t.fmt} {_fnv(np.ascontiguousarray(t.data).tobytes()):016x}")
    for m in sc.materials:
        out.append(f"mat {m.type} {m.normal} {m.emissive} {m.color} {m.metalness} {m.roughness} {np.float32(m.ior).view(np.uint32):08x}")
    for m in sc.meshes:
        out.append(f"mesh {len(m.positions)} {len(m.indices)} {_fnv(arr(m.positions, np.float32)):016x} {_fnv(arr(m.indices, np.uint32)):016x} "
                   f"{_fnv(arr(m.normals, np.float32)):016x} {_fnv(arr(m.texcoords, np.float32)):016x}")
    for i in sc.instances:
        out.append(f"inst {_fnv(np.ascontiguousarray(i.transform, np.float32).tobytes()):016x}" + "".join(f" {g[0]}/{g[1]}/{int(bool(g[2]))}" for g in i.geometries))
    return "\n".join(out) + "\n"


def _expressible_zoo():
    sc = scenes.material_zoo((48, 32))
    # expressible in glTF: no half-float colour texture, metal / rough maps of one size (as tests/test_offline_io.py)
    sc.materials[1].color = sc.solid3(0.7, 0.6, 0.2)
    sc.textures[sc.materials[0].metalness] = scenes.Texture(((np.indices((32, 32)).sum(0) % 2) * 255).astype(np.uint8), 32, 32, scenes.TEX_R8_UNORM)
    return sc


def _same_import(path):
    want = _digest(offline.load_glb(path))
    r = subprocess.run([_exe(), path, "--dump"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == want


@pytest.mark.parametrize("maker", [_expressible_zoo, lambda: scenes.cornell_box((32, 32)), lambda: scenes.bathroom((64, 36), detail=0.02)])
def test_native_import_equals_python_import(tmp_path, maker):
    """PNG textures (RG8 normal maps, sRGB colour, packed metal / rough), all four material mappings, matrix nodes."""
    path = str(tmp_path / "scene.glb")
    offline.write_glb(maker(), path)
    _same_import(path)


def test_native_import_of_trs_nodes_and_the_reference_fixture(tmp_path):
    """zgltf's recompose (T * R * S in binary32) under a rotated parent, bit for bit; and a file this repo's writer did not produce:
    the Khronos Box the reference ships with its glTF parser (u16 indices, normals, no texcoords), given a camera node."""
    import json
    import math
    import struct
    from test_offline_io import BOX_GLB, _glb_with_camera
    p = str(tmp_path / "t.glb")
    offline.write_glb(scenes.single_triangle(), p)
    buf = open(p, "rb").read()
    jl = struct.unpack_from("<I", buf, 12)[0]
    J = json.loads(buf[20:20 + jl])
    binchunk = buf[20 + jl:]
    J["nodes"][0].pop("matrix")
    J["nodes"][0].update(translation=[1.0, 2.0, 3.0], scale=[2.0, 0.7, 1.3], rotation=[0.1825742, 0.3651484, 0.5477226, 0.7302967])
    s = math.sin(math.pi / 5)
    J["nodes"].append({"children": [0], "rotation": [0.0, s, 0.0, math.cos(math.pi / 5)], "translation": [0.5, -0.25, 0.125], "scale": [1.5, 1.5, 0.5]})
    J["scenes"][0]["nodes"] = [1, 2]
    js = json.dumps(J).encode()
    js += b" " * ((-len(js)) % 4)
    open(p, "wb").write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + len(binchunk)) + struct.pack("<II", len(js), 0x4E4F534A) + js + binchunk)
    _same_import(p)
    if os.path.exists(BOX_GLB):
        q = str(tmp_path / "box_cam.glb")
        _glb_with_camera(BOX_GLB, q)
        _same_import(q)


def test_native_front_end_reports_the_reference_errors(tmp_path):
    box = "/root/reference/deps/zgltf/test-samples/box_binary/Box.glb"
    if os.path.exists(box):   # the reference's own fixture has no camera (Camera.zig:30)
        r = subprocess.run([_exe(), box, "--dump"], capture_output=True, text=True, timeout=60)
        assert r.returncode == 1 and "NoCameraInGlb" in r.stderr
    r = subprocess.run([_exe(), "scene.gltf", "--dump"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "OnlySupportsGlbInput" in r.stderr
    r = subprocess.run([_exe()], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2   # error.BadArgs


def test_native_exr_reader_and_writer_match_the_python_ones(tmp_path):
    """Scanline EXR in (NONE as written here; ZIP / HALF on the reference-held deps/tinyexr/asakusa.exr), three FLOAT channels B, G, R
    out (exr.zig:187-200): same pixels as offline.read_exr, same bytes as offline.write_exr."""
    rs = np.random.RandomState(3)
    img = np.concatenate([rs.randn(23, 37, 3).astype(np.float32) * 5, np.ones((23, 37, 1), np.float32)], -1)
    src = str(tmp_path / "a.exr")
    offline.write_exr(src, img)
    cases = [src]
    if os.path.exists("/root/reference/deps/tinyexr/asakusa.exr"):
        cases.append("/root/reference/deps/tinyexr/asakusa.exr")
    for path in cases:
        want = offline.read_exr(path)
        out = str(tmp_path / "b.exr")
        r = subprocess.run([_exe(), "--exr-copy", path, out], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        w, h, digest = r.stdout.split()
        assert (int(w), int(h)) == (want.shape[1], want.shape[0])
        assert digest == f"{_fnv(np.ascontiguousarray(want, np.float32).tobytes()):016x}"
        ref = str(tmp_path / "c.exr")
        offline.write_exr(ref, want)
        assert open(out, "rb").read() == open(ref, "rb").read()
