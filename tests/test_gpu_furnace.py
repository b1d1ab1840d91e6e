"""The reference's GPU furnace tests (engine/tests.zig:257-364, 366-455) on the CUDA path, same bounds."""
import numpy as np
import pytest

from moonshine_b200 import scenes
from gpu_common import gpu_render

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
