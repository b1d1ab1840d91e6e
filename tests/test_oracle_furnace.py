"""The reference's own GPU furnace tests (engine/tests.zig) run against the CPU oracle.

These are the only results the reference's test-suite pins for the hot path,
so they are what ties the oracle to the reference. Same scenes, extents, constants and bounds.
"""
import numpy as np

from moonshine_b200 import scenes
from oracle.binding import Oracle


def _render(sc):
    o = Oracle()
    scenes.load_scene(o, sc)
    img, counts = o.render(sc.lens, *sc.extent)
    return img, counts


def test_white_sphere_on_white_background_is_white():
    """engine/tests.zig:257-344: env=0, mesh=0, 32x32, samples_per_run=512 -> |c-1| <= 1e-5."""
    img, _ = _render(scenes.furnace_white_sphere(env_samples=0))
    assert np.abs(img[..., :3] - 1.0).max() <= 1e-5
    np.testing.assert_array_equal(img[..., 3], 1.0)  # alpha = number of launches (main.hlsl:43-51)


def test_white_sphere_with_env_nee_is_white():
    """engine/tests.zig:347-363: env=1 -> |c-1| <= 0.1."""
    img, counts = _render(scenes.furnace_white_sphere(env_samples=1))
    assert np.abs(img[..., :3] - 1.0).max() <= 0.1
    assert counts[1] > 0


def test_inside_illuminating_sphere_is_white():
    """engine/tests.zig:366-455: albedo .5 + emissive .5, samples_per_run=1024 -> |c-1| <= 0.02."""
    img, _ = _render(scenes.furnace_emissive_interior(mesh_samples=0))
    assert np.abs(img[..., :3] - 1.0).max() <= 0.02


def test_inside_illuminating_sphere_with_mesh_nee_is_white():
    """The commented-out engine/tests.zig:457-487 (mesh_samples=1, 512 spp, bound 0.1), revived."""
    img, counts = _render(scenes.furnace_emissive_interior(mesh_samples=1, spr=512))
    assert np.abs(img[..., :3] - 1.0).max() <= 0.1
    assert counts[1] > 0


def test_running_mean_matches_batch_mean():
    """storeColor (main.hlsl:43-51): N launches of spr=1 give the running mean of the N per-sample colours."""
    sc = scenes.single_triangle(extent=(16, 16))
    o = Oracle()
    scenes.load_scene(o, sc)
    acc, _ = o.render(sc.lens, 16, 16, n_launches=8)
    singles = []
    for s in range(8):
        img = np.zeros((16, 16, 4), np.float32)
        # a launch with sample_count=s on a zeroed image is NOT what the reference does; instead use
        # sample_count=0 semantics per sample by seeding through spr: render one launch at index s with prior 0
        o2 = Oracle()
        sc2 = scenes.single_triangle(extent=(16, 16))
        scenes.load_scene(o2, sc2)
        one, _ = o2.render(sc2.lens, 16, 16, n_launches=1, image=img, sample_count=s)
        # with prior = 0 and sample_count = s>0: out = c/(s+1)
        singles.append(one[..., :3].astype(np.float64) * (s + 1))
    np.testing.assert_allclose(acc[..., :3], np.mean(singles, 0), rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(acc[..., 3], 8.0)
