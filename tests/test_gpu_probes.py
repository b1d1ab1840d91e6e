"""Function-level GPU-vs-oracle probes: MaterialVariant::{eval,pdf,sample} (material.hlsl:395-487) for all four variants and
EnvMap::{sample,eval,incomingRadiance} (light.hlsl:47-102) on 10^5 random inputs each, through the C ABI
(HdMoonshineExtBsdfProbe / HdMoonshineExtEnvProbe).

eval and pdf are pure IEEE +,-,*,/,sqrt chains on both sides -> compared BIT FOR BIT. sample goes through sin/cos (libm on the
host, libdevice on the GPU, <= 2 ulp apart) -> directions within 4e-6, discrete decisions (lobe, hemisphere, texel) must agree
except where an input sits within an ulp of a decision boundary.
"""
import numpy as np
import pytest

from moonshine_b200 import hydra, scenes
from oracle import binding as ob
from oracle.binding import Oracle

pytestmark = pytest.mark.gpu
N = 100_000


def _dirs(rs, n, grazing_share=0.1):
    d = rs.randn(n, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = int(n * grazing_share)                 # a share of grazing directions (|cos| tiny): where Lambda / Fresnel branch
    d[:k, 2] *= 1e-3
    d[:k] /= np.linalg.norm(d[:k], axis=1, keepdims=True)
    return d.astype(np.float32)


def _queries(rs, n, mtype, op):
    q = np.zeros((n, 14), np.float32)
    q[:, 0] = mtype
    q[:, 1:4] = rs.rand(n, 3)
    q[:, 4] = rs.choice([0.0, 1.0, 0.3, 0.77], n)                      # metalness
    q[:, 5] = rs.choice([0.0, 0.02, 0.05, 0.2, 0.5, 1.0], n) * (0.5 + 0.5 * rs.rand(n))   # roughness (alpha clamps at 0.001)
    q[:, 6] = rs.choice([1.33, 1.5, 2.4], n)
    q[:, 7] = op
    q[:, 8:11] = rs.rand(n, 3) if op == 2 else _dirs(rs, n)
    q[:, 11:14] = _dirs(rs, n)
    return q


@pytest.mark.parametrize("mtype", [scenes.MAT_GLASS, scenes.MAT_LAMBERT, scenes.MAT_PERFECT_MIRROR, scenes.MAT_STANDARD_PBR])
def test_bsdf_eval_and_pdf_bit_exact(mtype):
    m = hydra.HdMoonshine()
    rs = np.random.RandomState(100 + mtype)
    for op in (0, 1):
        q = _queries(rs, N, mtype, op)
        g, r = m.bsdf_batch(q), ob.bsdf_batch(q)
        same = (g.view(np.uint32) == r.view(np.uint32)) | (np.isnan(g) & np.isnan(r))
        assert same.all(), f"type {mtype} op {op}: {np.count_nonzero(~same.all(axis=1))} of {N} differ, first {q[np.argmin(same.all(axis=1))]}"
        assert np.isfinite(r).mean() > 0.9


@pytest.mark.parametrize("mtype", [scenes.MAT_GLASS, scenes.MAT_LAMBERT, scenes.MAT_PERFECT_MIRROR, scenes.MAT_STANDARD_PBR])
def test_bsdf_sample_matches(mtype):
    m = hydra.HdMoonshine()
    rs = np.random.RandomState(200 + mtype)
    q = _queries(rs, N, mtype, 2)
    g, r = m.bsdf_batch(q), ob.bsdf_batch(q)
    if mtype in (scenes.MAT_GLASS, scenes.MAT_PERFECT_MIRROR):   # no transcendental on these paths
        np.testing.assert_array_equal(g.view(np.uint32), r.view(np.uint32))
        return
    close = (np.abs(g[:, :3] - r[:, :3]) <= 4e-6).all(axis=1)
    assert close.mean() >= 0.9999, f"type {mtype}: {np.count_nonzero(~close)} sampled directions differ"
    pd = np.abs(g[:, 3] - r[:, 3]) <= 1e-4 * np.maximum(1.0, np.abs(r[:, 3])) + 2e-2 * (q[:, 5] < 0.1)   # sharp lobes: D(m) amplifies ulps
    assert (pd | ~close).mean() >= 0.999


def _env_pair():
    sky = scenes.procedural_sky(256, 128, sun_power=5e4)
    m, o = hydra.HdMoonshine(), Oracle()
    m.set_background(sky)
    o.set_background(sky)
    return m, o


def test_env_sample_matches_oracle():
    m, o = _env_pair()
    rs = np.random.RandomState(5)
    a = np.zeros((N, 3), np.float32)
    a[:, :2] = rs.rand(N, 2)
    g, r = m.env_probe(0, a), o.env_probe(0, a)
    # the texel walk is pure IEEE arithmetic on the same pyramid only if the two pyramids agree bit for bit; they are built by
    # different code (GPU kernel vs host loop, sin/cos/atan2 inside the resampling) and agree to 2e-5 relative, so a walk can
    # take the other branch when a random number lands within that margin of a split: rare, and then everything differs.
    same_texel = (np.abs(g[:, 0:3] - r[:, 0:3]) <= 2e-2).all(axis=1)
    assert same_texel.mean() >= 0.995
    s = same_texel
    np.testing.assert_allclose(g[s, 3:6], r[s, 3:6], rtol=3e-5, atol=1e-6)      # radiance of the chosen texel
    np.testing.assert_allclose(g[s, 6], r[s, 6], rtol=1e-4, atol=1e-7)          # pdf
    close_dir = (np.abs(g[s, 0:3] - r[s, 0:3]) <= 1e-4).all(axis=1)             # rand is REMAPPED level after level: ulps of the pyramid grow
    assert close_dir.mean() >= 0.99
    assert (r[:, 6] > 0).mean() > 0.99


def test_env_eval_and_incoming_radiance_match_oracle():
    m, o = _env_pair()
    rs = np.random.RandomState(6)
    d = _dirs(rs, N, grazing_share=0.0)
    g, r = m.env_probe(1, d), o.env_probe(1, d)
    ok = np.abs(g[:, 3] - r[:, 3]) <= 1e-4 * np.maximum(np.abs(r[:, 3]), 1e-3)   # same texel -> same pdf up to the pyramid's 2e-5
    assert ok.mean() >= 0.998                                                   # atan2 ulps at texel borders pick the neighbour
    np.testing.assert_allclose(g[ok, 0:3], r[ok, 0:3], rtol=3e-5, atol=1e-6)
    g2, r2 = m.env_probe(2, d), o.env_probe(2, d)
    ok2 = (np.abs(g2[:, :3] - r2[:, :3]) <= 2e-3 * np.maximum(1.0, np.abs(r2[:, :3]))).all(axis=1)   # bilinear: continuous in the direction
    assert ok2.mean() >= 0.999


def test_constant_env_probes_are_exact():
    """1x1 white background (the reference's default, BackgroundManager.zig:116-126): no resampling error to hide behind."""
    m, o = hydra.HdMoonshine(), Oracle()
    rs = np.random.RandomState(7)
    d = _dirs(rs, 20000, grazing_share=0.0)
    for op in (1, 2):
        np.testing.assert_array_equal(m.env_probe(op, d).view(np.uint32), o.env_probe(op, d).view(np.uint32))
    a = np.zeros((20000, 3), np.float32)
    a[:, :2] = rs.rand(20000, 2)
    g, r = m.env_probe(0, a), o.env_probe(0, a)
    np.testing.assert_array_equal(g[:, 3:].view(np.uint32), r[:, 3:].view(np.uint32))   # radiance + pdf exact
    np.testing.assert_allclose(g[:, :3], r[:, :3], atol=4e-7)                              # direction: sin/cos ulps


def test_defined_transcendentals_bit_exact():
    """device_math.cuh and oracle/utils.h state the same operation sequences independently: 10^6 inputs per function, every bit."""
    m = hydra.HdMoonshine()
    rs = np.random.RandomState(8)
    n = 1_000_000
    x = rs.uniform(-2 * np.pi, 4 * np.pi, n).astype(np.float32)
    for fn in (0, 1):
        np.testing.assert_array_equal(m.math_probe(fn, x).view(np.uint32), ob.math_probe(fn, x).view(np.uint32))
    u = np.concatenate([rs.uniform(2.0 ** -24, 1, n - 20000), 1 - np.arange(1, 10001) * 2.0 ** -24, np.arange(1, 10001) * 2.0 ** -24]).astype(np.float32)
    np.testing.assert_array_equal(m.math_probe(2, u).view(np.uint32), ob.math_probe(2, u).view(np.uint32))
    y, xx = rs.uniform(-1, 1, n).astype(np.float32), rs.uniform(-1, 1, n).astype(np.float32)
    y[:1000] = 0.0
    xx[500:1500] = 0.0
    np.testing.assert_array_equal(m.math_probe(3, y, xx).view(np.uint32), ob.math_probe(3, y, xx).view(np.uint32))
    z = rs.uniform(-1, 1, n).astype(np.float32)
    z[:4] = [1.0, -1.0, 0.0, -0.0]
    np.testing.assert_array_equal(m.math_probe(4, z).view(np.uint32), ob.math_probe(4, z).view(np.uint32))
