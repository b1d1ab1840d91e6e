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


def _sphere_mask(img):
    """pixels whose primary ray hits the unit sphere of the test lens (origin (-3,0,0), vfov pi/4), away from the silhouette"""
    h, w = img.shape[:2]
    y, x = np.mgrid[0:h, 0:w]
    u, v = (x + 0.5) / w * 2 - 1, (y + 0.5) / h * 2 - 1
    t = np.tan(np.pi / 8)
    return (u * t) ** 2 + (v * t) ** 2 < (0.8 / 3.0) ** 2


def test_mirror_and_glass_furnace_conserve_energy():
    """Delta BSDFs (material.hlsl:313-393) neither absorb nor gain: a mirror / glass sphere in a white furnace is white.
    Not pinned by any reference test - this ties PerfectMirror/Glass eval*|cos|/pdf == 1 to the analytic expectation."""
    img, _ = _render(scenes.furnace_material_sphere("mirror"))
    assert np.abs(img[..., :3] - 1.0).max() <= 1e-5
    img, _ = _render(scenes.furnace_material_sphere("glass"))
    assert np.abs(img[..., :3] - 1.0).max() <= 0.25         # Russian roulette after bounce 3 (p = 0.95) adds variance, not bias
    assert abs(float(img[..., :3].mean()) - 1.0) <= 5e-3


def test_ggx_metal_furnace_never_gains_energy():
    """StandardPBR (material.hlsl:179-270), white metal, roughness 0.5: single-scattering GGX with the Smith G term can only lose
    energy, so the sphere is at most white - with and without env NEE (MIS must not change the expectation)."""
    means = []
    for env in (0, 1):
        img, _ = _render(scenes.furnace_material_sphere("metal", env_samples=env))
        on = _sphere_mask(img)
        means.append(float(img[on][:, :3].mean()))
        assert 0.7 <= means[-1] <= 1.0 + 5e-3, means
        np.testing.assert_allclose(img[~on & (img[..., 0] == 1.0)][:, :3], 1.0)   # background pixels stay exactly white
    assert abs(means[0] - means[1]) <= 0.02                  # BSDF sampling alone == MIS of BSDF + env sampling


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
