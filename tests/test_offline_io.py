"""§8(f) rows 1-2 on CPU: GLB/PNG import rules (World.zig:44-363, Camera.zig:26-51) and EXR I/O (fileformats/exr.zig)."""
import math
import os

import numpy as np
import pytest

from moonshine_b200 import offline, scenes


def test_exr_write_read_roundtrip_is_exact(tmp_path):
    rs = np.random.RandomState(0)
    img = (rs.rand(9, 17, 4) * 1000).astype(np.float32)
    p = str(tmp_path / "a.exr")
    offline.write_exr(p, img)
    back = offline.read_exr(p)
    np.testing.assert_array_equal(back[..., :3], img[..., :3])
    np.testing.assert_array_equal(back[..., 3], 1.0)          # three channels B,G,R only (exr.zig:187-200)
    raw = open(p, "rb").read()
    assert raw[:4] == b"\x76\x2f\x31\x01" and b"B\0" in raw and b"chlist" in raw


@pytest.mark.skipif(not os.path.exists("/root/reference/deps/tinyexr/asakusa.exr"), reason="reference fixture not present on this machine")
def test_exr_reader_on_the_reference_fixture():
    """deps/tinyexr/asakusa.exr (ZIP-compressed HALF scanlines), checked against OpenCV's decoder when available."""
    a = offline.read_exr("/root/reference/deps/tinyexr/asakusa.exr")
    assert a.shape == (440, 660, 4) and np.isfinite(a).all()
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    c = cv2.imread("/root/reference/deps/tinyexr/asakusa.exr", cv2.IMREAD_UNCHANGED)
    if c is None:
        pytest.skip("OpenCV built without OpenEXR")
    np.testing.assert_array_equal(c[..., 2::-1], a[..., :3])


def _glb_scene():
    sc = scenes.material_zoo((64, 48))
    # make it expressible in glTF: no half-float colour texture, metal/rough maps of one size
    sc.materials[1].color = sc.solid3(0.7, 0.6, 0.2)
    sc.textures[sc.materials[0].metalness] = scenes.Texture(((np.indices((32, 32)).sum(0) % 2) * 255).astype(np.uint8), 32, 32, scenes.TEX_R8_UNORM)
    return sc


def test_glb_export_import_roundtrip(tmp_path):
    sc = _glb_scene()
    p = str(tmp_path / "zoo.glb")
    offline.write_glb(sc, p)
    back = offline.load_glb(p, sc.extent)
    assert len(back.instances) == len(sc.instances)
    for a, b in zip(sc.instances, back.instances):
        np.testing.assert_array_equal(a.transform, b.transform)          # (x,z,y) row swizzle undone exactly
        assert [g[2] for g in a.geometries] == [g[2] for g in b.geometries]   # "Emitter" name prefix <=> sampled
        for ga, gb in zip(a.geometries, b.geometries):
            ma, mb = sc.meshes[ga[0]], back.meshes[gb[0]]
            np.testing.assert_array_equal(ma.positions, mb.positions)
            np.testing.assert_array_equal(ma.indices, mb.indices)
            assert (ma.normals is None) == (mb.normals is None) and (ma.texcoords is None) == (mb.texcoords is None)
            assert sc.materials[ga[1]].type == back.materials[gb[1]].type
    # material payloads: constants and texels survive
    for ma, mb in zip(sc.materials, back.materials):
        for f in ("normal", "emissive") + (("color",) if ma.type in (scenes.MAT_LAMBERT, scenes.MAT_STANDARD_PBR) else ()) + (("metalness", "roughness") if ma.type == scenes.MAT_STANDARD_PBR else ()):
            ta, tb = sc.textures[getattr(ma, f)], back.textures[getattr(mb, f)]
            assert (ta.width, ta.height, ta.fmt) == (tb.width, tb.height, tb.fmt), f
            np.testing.assert_array_equal(np.asarray(ta.data).reshape(-1)[: 3 if ta.fmt == scenes.TEX_RGBA32F else None],
                                          np.asarray(tb.data).reshape(-1)[: 3 if ta.fmt == scenes.TEX_RGBA32F else None])
        if ma.type in (scenes.MAT_GLASS, scenes.MAT_STANDARD_PBR):
            assert np.float32(ma.ior) == np.float32(mb.ior)
    np.testing.assert_allclose(back.lens.as_array()[:10], sc.lens.as_array()[:10], rtol=0, atol=2e-7)
    assert (back.lens.aperture, back.lens.focus_distance) == (0.0, 1.0)   # Camera.zig:47-48: glTF carries no thin-lens data
    assert back.constants == scenes.Constants()   # offline constants (offline/main.zig:106-111)


def test_gltf_material_mapping_rules(tmp_path):
    """World.zig:119-227: transmission 1 -> Glass; no MR texture & (0,1) -> Lambert; (1,0) -> mirror; else StandardPBR;
    glTF defaults (metallic 1, roughness 1) -> rough metal."""
    sc = scenes.single_triangle()
    n, e = sc.materials[0].normal, sc.materials[0].emissive
    sc.materials = [
        scenes.Material(scenes.MAT_GLASS, n, e, ior=1.33),
        scenes.Material(scenes.MAT_LAMBERT, n, e, color=sc.solid3(0.1, 0.2, 0.3)),
        scenes.Material(scenes.MAT_PERFECT_MIRROR, n, e),
        scenes.Material(scenes.MAT_STANDARD_PBR, n, e, color=sc.solid3(1, 1, 1), metalness=sc.solid1(1.0), roughness=sc.solid1(1.0)),
    ]
    p = str(tmp_path / "m.glb")
    offline.write_glb(sc, p)
    back = offline.load_glb(p)
    assert [m.type for m in back.materials] == [scenes.MAT_GLASS, scenes.MAT_LAMBERT, scenes.MAT_PERFECT_MIRROR, scenes.MAT_STANDARD_PBR]
    assert np.float32(back.materials[0].ior) == np.float32(1.33)
    assert back.extent == (1280, 720)  # offline/main.zig:48


def test_trs_nodes_and_parent_chain(tmp_path):
    """zgltf recompose (T*R*S) and getGlobalTransform: a child under a rotated/scaled parent."""
    import json, struct
    sc = scenes.single_triangle()
    p = str(tmp_path / "t.glb")
    offline.write_glb(sc, p)
    buf = open(p, "rb").read()
    jl = struct.unpack_from("<I", buf, 12)[0]
    J = json.loads(buf[20:20 + jl])
    binchunk = buf[20 + jl:]
    J["nodes"][0].pop("matrix")
    J["nodes"][0].update(translation=[1.0, 2.0, 3.0], scale=[2.0, 2.0, 2.0])
    s = math.sin(math.pi / 4)
    J["nodes"].append({"children": [0], "rotation": [0.0, s, 0.0, s], "translation": [0.5, 0.0, 0.0]})   # 90 deg about +Y
    J["scenes"][0]["nodes"] = [1, 2]
    js = json.dumps(J).encode(); js += b" " * ((-len(js)) % 4)
    open(p, "wb").write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + len(binchunk)) + struct.pack("<II", len(js), 0x4E4F534A) + js + binchunk)
    back = offline.load_glb(p)
    t = back.instances[0].transform
    # glTF point (0,0,0) of the child -> parent: R*(1,2,3)+ (0.5,0,0) = (3+0.5, 2, -1); engine rows are (x, z, y)
    np.testing.assert_allclose(t[:, 3], [3.5, -1.0, 2.0], atol=1e-6)
    np.testing.assert_allclose(np.abs(np.linalg.det(t[:, :3].astype(np.float64))), 8.0, rtol=1e-5)
