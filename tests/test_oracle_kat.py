"""Known-answer and invariant tests that pin the CPU oracle (no GPU needed).

The reference's own tests hold no numeric vectors for these functions ("parity
unpinned"); the RNG values below are the KATs derived from shaders/utils/random.hlsl:8-45.
"""
import math

import numpy as np
import pytest

from moonshine_b200 import scenes
from oracle import binding
from oracle.binding import Oracle


def test_hash_kats(oracle_lib):
    L = oracle_lib
    assert L.OracleHashPcg(0) == 0x07BB2FE2
    assert L.OracleHashPcg(1) == 0xA8BEEA3C
    assert L.OracleHashLcg(0) == 0xAC564B05


@pytest.mark.parametrize("seed,state,mant", [
    ((0, 0, 0), 0x7FDDB461, (0x8D3248, 0x9E7D27, 0x4EC34E, 0x14B186)),
    ((0, 1, 0), 0xCB8B5B66, (0x77D38C, 0xA748E0, 0xAA536B, 0xB647E9)),
    ((0, 0, 1), 0x2370168F, (0x2118BC, 0xB8B8EF, 0xB778A0, 0xD86AAA)),
    ((1, 0, 0), 0xC9369417, (0x2008E2, 0xA4E863, 0x9EC988, 0xA941E0)),
    ((5, 63, 63), 0x707D3AAB, (0xDF96CE, 0x62544A, 0xF82DF9, 0xA2ACD0)),
    ((1023, 1919, 1079), 0x7DE787B1, (0xC5E63B, 0xEEBA96, 0x04E41C, 0x5DDD3B)),
])
def test_rng_kats(oracle_lib, seed, state, mant):
    assert oracle_lib.OracleRngSeed(*seed) == state
    f = binding.rng_floats(*seed, 4)
    assert tuple(int(x * 2 ** 24) for x in f) == mant
    assert all(0.0 <= x < 1.0 for x in f)


def test_rng_float_kat_values():
    f = binding.rng_floats(0, 0, 0, 4)
    np.testing.assert_array_equal(f, np.array([0.5515484809875488, 0.6190971732139587, 0.30766761302948, 0.0808337926864624], np.float32))


def test_equal_area_sphere_roundtrip(oracle_lib):
    rs = np.random.RandomState(1)
    out3 = np.zeros(3, np.float32)
    out2 = np.zeros(2, np.float32)
    for _ in range(2000):
        u, v = rs.rand(2).astype(np.float32)
        oracle_lib.OracleSquareToEqualAreaSphere(float(u), float(v), out3.ctypes.data)
        assert abs(np.linalg.norm(out3.astype(np.float64)) - 1.0) < 1e-5
        oracle_lib.OracleSquareToEqualAreaSphereInverse(float(out3[0]), float(out3[1]), float(out3[2]), out2.ctypes.data)
        assert abs(out2[0] - u) < 2e-4 and abs(out2[1] - v) < 2e-4


def test_offset_along_normal(oracle_lib):
    # math.hlsl:32-42: near the origin a float offset n/65536, elsewhere an integer-ULP offset of int(256 n)
    out = np.zeros(3, np.float32)
    p = np.array([0.01, 5.0, -5.0], np.float32)
    n = np.array([1.0, 1.0, 1.0], np.float32)
    oracle_lib.OracleOffsetAlongNormal(p.ctypes.data, n.ctypes.data, out.ctypes.data)
    assert out[0] == np.float32(0.01) + np.float32(1.0 / 65536.0)
    assert out[1].view(np.int32) == p[1].view(np.int32) + 256
    assert out[2].view(np.int32) == p[2].view(np.int32) - 256  # negative coordinate: bits move the other way
    assert out[1] > p[1] and out[2] > p[2]


def test_inverse_affine_is_inverse():
    sc = scenes.single_triangle()
    rs = np.random.RandomState(3)
    a = rs.randn(3, 3).astype(np.float32) + 2 * np.eye(3, dtype=np.float32)
    t = np.concatenate([a, rs.randn(3, 1).astype(np.float32)], 1)
    sc.instances[0].transform = t
    o = Oracle()
    scenes.load_scene(o, sc)
    inv = o.world_to_instance(0).astype(np.float64)
    m = np.vstack([t.astype(np.float64), [0, 0, 0, 1]])
    mi = np.vstack([inv, [0, 0, 0, 1]])
    np.testing.assert_allclose(mi @ m, np.eye(4), atol=2e-5)
    # identity stays bit-exact identity
    sc2 = scenes.single_triangle()
    o2 = Oracle()
    scenes.load_scene(o2, sc2)
    np.testing.assert_array_equal(o2.world_to_instance(0), np.eye(3, 4, dtype=np.float32))


def test_alias_table_reconstructs_weights():
    """Vose table (alias_table.zig:28-91): P(i) = (select_i + sum_{j: alias_j = i} (1 - select_j)) / n."""
    sc = scenes.cornell_box()
    # make every geometry a sampled emitter so the table covers all 36 triangles
    for inst in sc.instances:
        inst.geometries = [(m, mat, True) for (m, mat, _) in inst.geometries]
    o = Oracle()
    scenes.load_scene(o, sc)
    tab = o.alias_table()
    n = int(tab[0]["alias"])
    assert n == 36 and len(tab) == 37
    ent = tab[1:]
    areas = []
    for i, inst in enumerate(sc.instances):
        for (m, _, _) in inst.geometries:
            mesh = sc.meshes[m]
            p = mesh.positions.astype(np.float64)
            for tri in mesh.indices:
                areas.append(0.5 * np.linalg.norm(np.cross(p[tri[1]] - p[tri[0]], p[tri[2]] - p[tri[0]])))
    areas = np.array(areas)
    assert abs(float(tab[0]["select"]) - areas.sum()) < 1e-4
    prob = np.zeros(n)
    for i in range(n):
        s = min(float(ent[i]["select"]), 1.0)
        prob[i] += s
        if s < 1.0:
            prob[int(ent[i]["alias"])] += 1.0 - s
    np.testing.assert_allclose(prob / n, areas / areas.sum(), atol=1e-5)
    # entry order is (instance, geometry, primitive) lexicographic (Accel.zig:503-521)
    keys = [(int(e["instance"]), int(e["geometry"]), int(e["primitive"])) for e in ent]
    assert keys == sorted(keys)


def test_env_preprocess_1x1_and_pyramid():
    o = Oracle()
    o.set_background(np.array([[[0.25, 0.5, 1.0, 1.0]]], np.float32))
    rgba, lum = o.env_data()
    assert rgba.shape == (1, 1, 4)
    np.testing.assert_allclose(rgba[0, 0, :3], [0.25, 0.5, 1.0], rtol=1e-6)
    assert abs(lum[0] - (0.2126 * 0.25 + 0.7152 * 0.5 + 0.0722 * 1.0)) < 1e-6
    # 8x16 equirect -> S = min(floorPow2(8), 1024) = 8; pyramid levels 8,4,2,1 with SUM folding (fold.hlsl)
    rs = np.random.RandomState(5)
    img = rs.rand(8, 16, 4).astype(np.float32)
    o.set_background(img)
    rgba, lum = o.env_data()
    assert rgba.shape == (8, 8, 4)
    l0 = lum[:64].reshape(8, 8)
    l1 = lum[64:80].reshape(4, 4)
    l3 = lum[84]
    np.testing.assert_allclose(l0, rgba[..., :3] @ np.array([0.2126, 0.7152, 0.0722], np.float32), rtol=1e-6)
    np.testing.assert_allclose(l1, l0.reshape(4, 2, 4, 2).sum((1, 3)), rtol=1e-6)
    np.testing.assert_allclose(l3, l0.sum(), rtol=1e-5)
    # constant image stays constant under the resample
    o.set_background(np.full((4, 8, 4), 0.75, np.float32))
    rgba, _ = o.env_data()
    np.testing.assert_allclose(rgba[..., :3], 0.75, rtol=1e-6)


def _random_rays(rs, n, center, radius):
    o = (rs.randn(n, 3) * radius + center).astype(np.float32)
    d = rs.randn(n, 3)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = 0.0
    rays[:, 4:7] = d
    rays[:, 7] = 1e12
    return rays


@pytest.mark.parametrize("maker", [lambda: scenes.furnace_white_sphere(0, order=3), scenes.cornell_box])
def test_bvh_matches_brute_force(maker):
    sc = maker()
    o = Oracle()
    scenes.load_scene(o, sc)
    rs = np.random.RandomState(11)
    rays = _random_rays(rs, 4000, np.array([0.3, 0.4, 0.5]), 0.5)
    a = o.trace(rays, brute=False)
    b = o.trace(rays, brute=True)
    np.testing.assert_array_equal(a, b)
    assert (a["instance"] != 0xFFFFFFFF).sum() > 1500
    sa = o.trace(rays, any_hit=True, brute=False)
    sb = o.trace(rays, any_hit=True, brute=True)
    np.testing.assert_array_equal(sa["instance"] != 0xFFFFFFFF, sb["instance"] != 0xFFFFFFFF)
    np.testing.assert_array_equal(sa["instance"] != 0xFFFFFFFF, a["instance"] != 0xFFFFFFFF)


def test_closest_hit_tie_break_is_lowest_id():
    """Two coincident triangles in different geometries: the hit must report the smaller geometry index."""
    sc = scenes.single_triangle()
    mesh2 = sc.add_mesh(sc.meshes[0].positions.copy(), sc.meshes[0].indices.copy())
    sc.instances[0].geometries = [(0, 0, False), (mesh2, 0, False)]
    o = Oracle()
    scenes.load_scene(o, sc)
    rays = np.array([[-3, 0, 0, 0, 1, 0, 0, 1e12]], np.float32)
    for brute in (False, True):
        h = o.trace(rays, brute=brute)
        assert h["instance"][0] == 0 and h["geometry"][0] == 0 and h["primitive"][0] == 0


def test_invisible_instance_is_skipped():
    sc = scenes.single_triangle()
    o = Oracle()
    _, _, _, insts = scenes.load_scene(o, sc)
    rays = np.array([[-3, 0, 0, 0, 1, 0, 0, 1e12]], np.float32)
    assert o.trace(rays)["instance"][0] == 0
    o.set_instance_visibility(insts[0], False)
    assert o.trace(rays)["instance"][0] == 0xFFFFFFFF


def test_bsdf_lambert_and_mirror_closed_forms():
    wi = np.array([0.3, 0.2, 0.9327379], np.float32)
    wo = np.array([-0.1, 0.4, 0.9110434], np.float32)
    p = [0.8, 0.6, 0.4, 0.0, 1.0, 1.5]
    e = binding.bsdf(scenes.MAT_LAMBERT, p, 0, wi, wo)
    np.testing.assert_allclose(e[:3], np.array([0.8, 0.6, 0.4]) / math.pi, rtol=1e-6)
    assert abs(binding.bsdf(scenes.MAT_LAMBERT, p, 1, wi, wo)[0] - wi[2] / math.pi) < 1e-6
    assert binding.bsdf(scenes.MAT_LAMBERT, p, 1, wi, -wo)[0] == 0.0
    s = binding.bsdf(scenes.MAT_PERFECT_MIRROR, p, 2, [0.1, 0.2, 0], wo)
    np.testing.assert_array_equal(s, np.array([-wo[0], -wo[1], wo[2], 1.0], np.float32))
    # glass at normal incidence: Fresnel ((n1-n2)/(n1+n2))^2
    f = binding.bsdf(scenes.MAT_GLASS, p, 0, [0, 0, 1], [0, 0, 1])[0]
    assert abs(f - ((1.5 - 1.000277) / (1.5 + 1.000277)) ** 2) < 1e-6


def test_bsdf_pbr_sample_pdf_consistency():
    """StandardPBR.sample's pdf must equal StandardPBR.pdf of the sampled direction (material.hlsl:217-249)."""
    rs = np.random.RandomState(2)
    p = [0.9, 0.5, 0.3, 0.4, 0.5, 1.5]
    wo = np.array([0.3, -0.2, 0.9327379], np.float32)
    checked = 0
    for _ in range(300):
        sq = rs.rand(2).astype(np.float32)
        s = binding.bsdf(scenes.MAT_STANDARD_PBR, p, 2, [sq[0], sq[1], 0], wo)
        if s[3] == 0:
            continue
        pdf = binding.bsdf(scenes.MAT_STANDARD_PBR, p, 1, s[:3], wo)[0]
        assert abs(pdf - s[3]) <= 2e-4 * max(1.0, abs(pdf))
        checked += 1
    assert checked > 200


def test_object_picker_follows_input_hlsl():
    """input.hlsl:24-57: centre of the sensor through the test lens hits the single triangle; a corner misses;
    y is flipped (coords origin top-left, uv origin bottom-left)."""
    from moonshine_b200 import scenes
    from oracle.binding import Oracle
    sc = scenes.single_triangle()
    o = Oracle()
    scenes.load_scene(o, sc)
    hit = o.pick(sc.lens, 64, 64, 0.5, 0.5)
    assert hit is not None and hit[:3] == (0, 0, 0)
    assert o.pick(sc.lens, 64, 64, 0.01, 0.01) is None
    # the triangle's apex points up (+z): it is hit near the top of the image (small y) but not near the bottom
    top, bottom = o.pick(sc.lens, 64, 64, 0.5, 0.3), o.pick(sc.lens, 64, 64, 0.5, 0.95)
    assert top is not None and bottom is None


def _ulp_err(got, ref64):
    ulp = np.maximum(np.spacing(np.abs(ref64.astype(np.float32))).astype(np.float64), np.finfo(np.float32).tiny)
    return np.abs(got.astype(np.float64) - ref64) / ulp


def test_defined_transcendentals_accuracy():
    """sin/cos/log/atan2/acos are DEFINED as fixed IEEE binary32 operation sequences (oracle/utils.h) because the
    HLSL intrinsics they stand for have implementation-defined precision. They must still BE those functions: error against
    float64 in the class of a GPU's intrinsics, over the ranges the shaders use (mappings.hlsl:11-99)."""
    from oracle import binding as ob
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.uniform(-2 * np.pi, 4 * np.pi, 400_000), np.linspace(0, 2 * np.pi, 50_001)]).astype(np.float32)
    for fn, f in ((0, np.sin), (1, np.cos)):
        g, ref = ob.math_probe(fn, x), f(x.astype(np.float64))
        assert np.abs(g - ref).max() <= 1.2e-7                                  # absolute: 1 ulp of 1.0
        assert _ulp_err(g, ref)[np.abs(ref) > 1e-3].max() <= 2.0
    u = np.concatenate([rs.uniform(2.0 ** -24, 1, 400_000), 1 - np.arange(1, 5000) * 2.0 ** -24, np.arange(1, 5000) * 2.0 ** -24]).astype(np.float32)
    g, ref = ob.math_probe(2, u), np.log(u.astype(np.float64))
    assert _ulp_err(g, ref)[np.abs(ref) > 1e-3].max() <= 2.0 and np.abs(g - ref)[np.abs(ref) <= 1e-3].max() <= 1e-9
    y, xx = rs.uniform(-1, 1, 400_000).astype(np.float32), rs.uniform(-1, 1, 400_000).astype(np.float32)
    assert _ulp_err(ob.math_probe(3, y, xx), np.arctan2(y.astype(np.float64), xx.astype(np.float64))).max() <= 3.5
    z = rs.uniform(-1, 1, 400_000).astype(np.float32)
    assert _ulp_err(ob.math_probe(4, z), np.arccos(z.astype(np.float64))).max() <= 4.5
    # exact anchors
    np.testing.assert_array_equal(ob.math_probe(0, np.array([0.0], np.float32)), [0.0])
    np.testing.assert_array_equal(ob.math_probe(1, np.array([0.0], np.float32)), [1.0])
    np.testing.assert_array_equal(ob.math_probe(2, np.array([1.0], np.float32)), [0.0])
    np.testing.assert_array_equal(ob.math_probe(3, np.array([0.0, 1.0], np.float32), np.array([1.0, 0.0], np.float32)), np.array([0.0, np.pi / 2], np.float32))
