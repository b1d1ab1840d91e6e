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


BOX_GLB = "/root/reference/deps/zgltf/test-samples/box_binary/Box.glb"


def _glb_with_camera(src_path, dst_path):
    """Repack a GLB with one perspective camera node appended to the scene (JSON chunk edited, BIN chunk untouched)."""
    import json
    import struct
    buf = open(src_path, "rb").read()
    pos, chunks = 12, []
    while pos < len(buf):
        clen, ctype = struct.unpack_from("<II", buf, pos)
        chunks.append((ctype, buf[pos + 8:pos + 8 + clen]))
        pos += 8 + clen
    j = json.loads(chunks[0][1].decode())
    j["cameras"] = [{"type": "perspective", "perspective": {"yfov": 0.7, "znear": 0.01, "aspectRatio": 1.5}}]
    j["nodes"].append({"camera": 0, "translation": [0.0, 0.3, 3.0]})
    j["scenes"][0]["nodes"].append(len(j["nodes"]) - 1)
    js = json.dumps(j).encode()
    js += b" " * ((-len(js)) % 4)
    out = struct.pack("<II", len(js), 0x4E4F534A) + js
    for ctype, data in chunks[1:]:
        out += struct.pack("<II", len(data), ctype) + data
    open(dst_path, "wb").write(struct.pack("<III", 0x46546C67, 2, 12 + len(out)) + out)


@pytest.mark.skipif(not os.path.exists(BOX_GLB), reason="reference fixture not present on this machine")
def test_reference_box_glb_has_no_camera():
    """deps/zgltf/test-samples/box_binary/Box.glb holds no camera node: Scene.fromGlbExr fails with error.NoCameraInGlb
    (engine/hrtsystem/Camera.zig:30) - a file this repo's writer did not produce."""
    with pytest.raises(ValueError, match="NoCameraInGlb"):
        offline.load_glb(BOX_GLB)


@pytest.mark.skipif(not os.path.exists(BOX_GLB), reason="reference fixture not present on this machine")
def test_reference_box_glb_loads_once_it_has_a_camera(tmp_path):
    """The Khronos Box sample as shipped with the reference's glTF parser, plus a camera: 24 vertices / 12 triangles with
    normals and no texcoords, u16 indices, one red dielectric material (metallic 0, default roughness 1 -> Lambert,
    World.zig:119-227), mesh node under a Y-up -> Z-up root matrix that the (x, z, y) row swizzle (World.zig:341-345) undoes."""
    p = str(tmp_path / "box_cam.glb")
    _glb_with_camera(BOX_GLB, p)
    sc = offline.load_glb(p, (96, 64))
    assert len(sc.instances) == 1 and len(sc.instances[0].geometries) == 1
    mesh = sc.meshes[sc.instances[0].geometries[0][0]]
    assert mesh.positions.shape == (24, 3) and mesh.indices.shape == (12, 3) and mesh.indices.max() == 23
    assert mesh.normals is not None and mesh.normals.shape == (24, 3) and mesh.texcoords is None
    np.testing.assert_array_equal(np.abs(mesh.positions), 0.5)                 # a unit cube centred at the origin
    np.testing.assert_array_equal(np.sort(np.abs(mesh.normals), axis=1), np.tile([0.0, 0.0, 1.0], (24, 1)))
    mat = sc.materials[sc.instances[0].geometries[0][1]]
    assert mat.type == scenes.MAT_LAMBERT                                      # metallicFactor 0, roughnessFactor default 1
    np.testing.assert_allclose(sc.textures[mat.color].data[:3], [0.8, 0.0, 0.0], rtol=1e-6)
    assert not sc.instances[0].geometries[0][2]                                # material name does not start with "Emitter"
    # node 0 carries the sample's Y-up -> Z-up matrix; composed with the importer's row swizzle the cube lands axis-aligned
    t = sc.instances[0].transform
    assert t.shape == (3, 4) and np.allclose(np.abs(t[:, :3]).sum(axis=0), 1.0) and np.allclose(t[:, 3], 0.0)
    # camera: translation (0, 0.3, 3) in glTF's Y-up frame -> (0, 3, 0.3) Z-up, looking down -Z_gltf = -Y here
    np.testing.assert_allclose(sc.lens.origin, (0.0, 3.0, 0.3), atol=1e-6)
    np.testing.assert_allclose(sc.lens.forward, (0.0, -1.0, 0.0), atol=1e-6)
    np.testing.assert_allclose(sc.lens.up, (0.0, 0.0, 1.0), atol=1e-6)
    assert abs(sc.lens.vfov - 0.7) < 1e-6
    # and the CPU side of the pipeline accepts it (the oracle is the checker here, as everywhere in tests/)
    from oracle.binding import Oracle
    o = Oracle()
    scenes.load_scene(o, sc)
    img, counts = o.render(sc.lens, 96, 64)
    assert counts[0] > 96 * 64 and (img[..., 0] > img[..., 1] + 0.05).mean() > 0.05   # a red box in a white furnace
