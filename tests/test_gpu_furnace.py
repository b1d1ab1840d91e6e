"""The reference's GPU furnace tests (engine/tests.zig:257-364, 366-455) on the CUDA path, same bounds."""
import numpy as np
import pytest

from moonshine_b200 import scenes
from gpu_common import gpu_render, oracle_render, within_tol_fraction

pytestmark = pytest.mark.gpu


def test_white_sphere_on_white_background_is_white():
    img, stats, _ = gpu_render(scenes.furnace_white_sphere(env_samples=0))
    assert np.abs(img[..., :3] - 1.0).max() <= 1e-5  # tests.zig:340-344
    np.testing.assert_array_equal(img[..., 3], 1.0)
    assert stats.rays_shadow == 0


def test_white_sphere_with_env_nee_is_white():
    img, stats, _ = gpu_render(scenes.furnace_white_sphere(env_samples=1))
    assert np.abs(img[..., :3] - 1.0).max() <= 0.1  # tests.zig:359-363
    assert stats.rays_shadow > 0


def test_inside_illuminating_sphere_is_white():
    img, _, _ = gpu_render(scenes.furnace_emissive_interior(mesh_samples=0))
    assert np.abs(img[..., :3] - 1.0).max() <= 0.02  # tests.zig:450-454


def test_inside_illuminating_sphere_with_mesh_nee_is_white():
    img, stats, _ = gpu_render(scenes.furnace_emissive_interior(mesh_samples=1, spr=512))
    assert np.abs(img[..., :3] - 1.0).max() <= 0.1  # tests.zig:457-487 (commented out in the reference)
    assert stats.rays_shadow > 0


def test_mirror_and_glass_furnace_conserve_energy():
    """Same bounds as tests/test_oracle_furnace.py on the CUDA path (PerfectMirror / Glass, material.hlsl:313-393)."""
    img, _, _ = gpu_render(scenes.furnace_material_sphere("mirror"))
    assert np.abs(img[..., :3] - 1.0).max() <= 1e-5
    img, _, _ = gpu_render(scenes.furnace_material_sphere("glass"))
    assert np.abs(img[..., :3] - 1.0).max() <= 0.25
    assert abs(float(img[..., :3].mean()) - 1.0) <= 5e-3


@pytest.mark.parametrize("kind,env", [("metal", 0), ("metal", 1), ("dielectric", 1), ("glass", 1), ("mirror", 1)])
def test_material_furnaces_match_the_oracle(kind, env):
    """Each non-Lambert BSDF alone on a sphere, 64 spp: the GPU image against the oracle's, pixel by pixel, and the energy bound."""
    sc = scenes.furnace_material_sphere(kind, spr=64, env_samples=env)
    g, _, _ = gpu_render(sc)
    r, _, _ = oracle_render(sc)
    assert within_tol_fraction(g, r, rel=2e-4) >= 0.99
    assert abs(float(g[..., :3].mean()) - float(r[..., :3].mean())) <= 1e-3
    if kind == "metal":
        assert float(g[..., :3].mean()) <= 1.0 + 5e-3
