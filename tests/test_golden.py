"""Committed golden vectors (tests/golden/*.npz, written by tools/make_golden.py from the CPU oracle).

CPU: the oracle still reproduces them (integers / bit patterns exactly; radiance to 1e-6 relative, because sin/cos/log of
another libm build may differ in the last place). GPU: the CUDA path through the C ABI reproduces them at the parity bar.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from moonshine_b200 import hydra, scenes  # noqa: E402
from oracle.binding import Oracle  # noqa: E402
import make_golden  # noqa: E402

NAMES = sorted(make_golden.CASES)


def _load(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    g, now = _load(name), make_golden.build(name)
    for k in ("hit_instance", "hit_geometry", "hit_primitive", "hit_u_bits", "hit_v_bits", "occluded_within_0p8", "rng_state_sample3", "alias_table"):
        np.testing.assert_array_equal(now[k], g[k], err_msg=k)
    np.testing.assert_allclose(now["image"], g["image"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(now["ray_counts"].astype(np.float64), g["ray_counts"].astype(np.float64), rtol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_golden(name):
    g = _load(name)
    sc = make_golden.CASES[name][0]()
    m = hydra.HdMoonshine()
    scenes.load_scene(m, sc)
    hits = m.trace(g["rays"])
    np.testing.assert_array_equal(hits["instance"], g["hit_instance"])
    hit = g["hit_instance"] != 0xFFFFFFFF
    np.testing.assert_array_equal(hits["geometry"][hit], g["hit_geometry"][hit])
    np.testing.assert_array_equal(hits["primitive"][hit], g["hit_primitive"][hit])
    np.testing.assert_array_equal(hits["u"].view(np.uint32)[hit], g["hit_u_bits"][hit])
    np.testing.assert_array_equal(hits["v"].view(np.uint32)[hit], g["hit_v_bits"][hit])
    short = g["rays"].copy()
    short[:, 7] = 0.8
    np.testing.assert_array_equal(m.trace(short, any_hit=True)["instance"] != 0xFFFFFFFF, g["occluded_within_0p8"])
    np.testing.assert_array_equal(m.alias_table().view(np.uint8), g["alias_table"])
    sensor, lens = m.CreateSensor(*sc.extent), m.CreateLens(sc.lens)
    _, state = m.primary_rays(sensor, lens, 3)
    np.testing.assert_array_equal(state, g["rng_state_sample3"])
    m.RenderLaunches(sensor, lens, int(g["launches"]))
    img = m.ReadSensor(sensor)
    d = np.abs(img[..., :3].astype(np.float64) - g["image"][..., :3])
    ok = (d <= 1e-4 * np.maximum(1.0, np.abs(g["image"][..., :3]))).mean()
    assert ok >= 0.998, ok
