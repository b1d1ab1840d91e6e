"""Physics-level checks of the oracle's BSDFs and environment sampling (CPU only). The GPU kernels are compared with the oracle bit
for bit (tests/test_gpu_probes.py), and both restate the same HLSL - so a shared misreading of material.hlsl / light.hlsl would pass
every GPU-vs-oracle test. These tests hold the oracle to what any correct implementation of those models must satisfy, whatever the
source text says: Helmholtz reciprocity, energy conservation, pdf normalisation, agreement between sample() and pdf(), Snell's law
and the Fresnel split of glass, and env-map sampling proportional to luminance."""
import math

import numpy as np

from moonshine_b200 import scenes
from oracle import binding
from oracle.binding import Oracle

PBR, GLASS, LAMBERT = scenes.MAT_STANDARD_PBR, scenes.MAT_GLASS, scenes.MAT_LAMBERT


def _bsdf(type_, color, metal, rough, ior, op, a, wo):
    a = np.atleast_2d(np.asarray(a, np.float32))
    q = np.zeros((len(a), 14), np.float32)
    q[:, 0], q[:, 1:4], q[:, 4], q[:, 5], q[:, 6], q[:, 7], q[:, 8:11] = type_, color, metal, rough, ior, op, a
    q[:, 11:14] = wo
    return binding.bsdf_batch(q).astype(np.float64)


def _upper_hemisphere(nz=256, nphi=512):
    """midpoint grid, equal solid angle per cell: directions (n, 3), cell solid angle"""
    z = (np.arange(nz) + 0.5) / nz
    phi = (np.arange(nphi) + 0.5) / nphi * 2 * np.pi
    zz, pp = np.meshgrid(z, phi, indexing="ij")
    r = np.sqrt(1 - zz * zz)
    d = np.stack([r * np.cos(pp), r * np.sin(pp), zz], -1).reshape(-1, 3).astype(np.float32)
    return d, 2 * np.pi / len(d)


def _unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


CONFIGS = [(0.5, 0.0), (0.5, 1.0), (0.2, 1.0), (0.25, 0.3), (0.9, 0.5)]   # (roughness, metalness)
VIEWS = [(0, 0, 1), (0.6, 0, 0.8), (0.95, 0.0, 0.31225)]


def test_standard_pbr_is_reciprocal():
    """f(w_i, w_o) = f(w_o, w_i): GGX with Smith G and Schlick F plus a Lambert lobe is symmetric in its two directions."""
    rs = np.random.RandomState(1)

    def hemi(n):
        v = rs.randn(n, 3)
        v[:, 2] = np.abs(v[:, 2]) + 0.05
        return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)

    for rough, metal in CONFIGS:
        wi, wo = hemi(1500), hemi(1500)
        a = _bsdf(PBR, (0.9, 0.5, 0.3), metal, rough, 1.5, 0, wi, wo)[:, :3]
        b = _bsdf(PBR, (0.9, 0.5, 0.3), metal, rough, 1.5, 0, wo, wi)[:, :3]
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-7)


def test_standard_pbr_pdf_is_a_sub_normalised_density_and_white_metal_never_gains_energy():
    d, dw = _upper_hemisphere()
    for rough, metal in CONFIGS:
        for view in VIEWS:
            wo = _unit(view)
            pdf = _bsdf(PBR, (1, 1, 1), metal, rough, 1.5, 1, d, wo)[:, 0]
            assert (pdf >= 0).all()
            total = pdf.sum() * dw
            # half-vector sampling reflects some directions below the horizon, where the pdf is 0: the integral is at most 1
            assert 0.6 < total < 1.002, (rough, metal, view, total)
            if metal == 1.0:
                f = _bsdf(PBR, (1, 1, 1), metal, rough, 1.5, 0, d, wo)[:, 0]
                albedo = (f * d[:, 2]).sum() * dw
                assert 0.5 < albedo < 1.002, (rough, view, albedo)   # single-scattering microfacets lose energy, never gain


def test_standard_pbr_samples_are_distributed_as_its_pdf():
    """Histogram of sample() over a coarse grid of the hemisphere == integral of pdf() over each cell (and the pdf sample() reports
    is pdf() of the direction - tests/test_oracle_kat.py). Directions sampled below the horizon carry pdf 0 and are dropped."""
    n_side = 700
    u = (np.stack(np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij"), -1).reshape(-1, 2) + 0.5) / n_side   # stratified
    sq = np.concatenate([u, np.zeros((len(u), 1))], 1).astype(np.float32)
    d, dw = _upper_hemisphere(256, 512)
    cz, cphi = 4, 8   # coarse cells: 4 bands of cos(theta) x 8 sectors of phi
    cell_of = lambda v: np.minimum((v[:, 2] * cz).astype(int), cz - 1) * cphi + ((np.arctan2(v[:, 1], v[:, 0]) % (2 * np.pi)) / (2 * np.pi) * cphi).astype(int) % cphi
    grid_cells = cell_of(d.astype(np.float64))
    for rough, metal in [(0.5, 0.0), (0.3, 1.0), (0.8, 0.5)]:
        wo = _unit((0.5, 0.2, 0.84))
        s = _bsdf(PBR, (0.8, 0.8, 0.8), metal, rough, 1.5, 2, sq, wo)
        ok = s[:, 3] > 0
        assert ok.mean() > 0.6
        hist = np.bincount(cell_of(s[ok, :3]), minlength=cz * cphi) / len(sq)
        pdf = _bsdf(PBR, (0.8, 0.8, 0.8), metal, rough, 1.5, 1, d, wo)[:, 0]
        expect = np.bincount(grid_cells, weights=pdf * dw, minlength=cz * cphi)
        assert abs(hist.sum() - expect.sum()) < 0.002
        np.testing.assert_allclose(hist, expect, atol=0.002, err_msg=str((rough, metal)))   # 32 cells holding up to 0.64 of the mass


def test_glass_obeys_snell_and_splits_energy_by_fresnel():
    ior, air = 1.5, 1.000277   # material.hlsl:348-351 (AIR_IOR)
    rs = np.random.RandomState(3)
    for cos_o in (1.0, 0.8, 0.5, 0.2):
        sin_o = math.sqrt(1 - cos_o * cos_o)
        wo = np.array([sin_o, 0, cos_o], np.float32)
        sq = np.concatenate([rs.rand(20000, 2), np.zeros((20000, 1))], 1).astype(np.float32)
        s = _bsdf(GLASS, (1, 1, 1), 0, 0, ior, 2, sq, wo)
        refl = s[:, 2] > 0
        # reflected: mirror direction; transmitted: Snell, in the plane of incidence, on the other side
        np.testing.assert_allclose(s[refl, :3], np.broadcast_to([-sin_o, 0, cos_o], (refl.sum(), 3)), atol=2e-6)
        sin_t = sin_o * air / ior
        np.testing.assert_allclose(s[~refl, :3], np.broadcast_to([-sin_t, 0, -math.sqrt(1 - sin_t * sin_t)], ((~refl).sum(), 3)), atol=2e-6)
        # unpolarised Fresnel reflectance of a dielectric interface
        cos_t = math.sqrt(1 - sin_t * sin_t)
        rp = (ior * cos_o - air * cos_t) / (ior * cos_o + air * cos_t)
        rs_ = (air * cos_o - ior * cos_t) / (air * cos_o + ior * cos_t)
        fresnel = 0.5 * (rp * rp + rs_ * rs_)
        assert abs(refl.mean() - fresnel) < 4 * math.sqrt(fresnel * (1 - fresnel) / 20000) + 1e-3, (cos_o, refl.mean(), fresnel)
        # each branch's pdf is its selection probability; eval * |cos| / pdf = 1 (no energy lost or gained at the interface)
        e = _bsdf(GLASS, (1, 1, 1), 0, 0, ior, 0, s[:, :3].astype(np.float32), wo)[:, 0]
        np.testing.assert_allclose(e * np.abs(s[:, 2]) / s[:, 3], 1.0, rtol=1e-5)


def test_lambert_is_a_normalised_cosine_lobe():
    d, dw = _upper_hemisphere()
    wo = _unit((0.3, -0.4, 0.866))
    pdf = _bsdf(LAMBERT, (0.7, 0.7, 0.7), 0, 0, 1.5, 1, d, wo)[:, 0]
    assert abs(pdf.sum() * dw - 1.0) < 1e-3
    f = _bsdf(LAMBERT, (0.7, 0.7, 0.7), 0, 0, 1.5, 0, d, wo)[:, 0]
    assert abs((f * d[:, 2]).sum() * dw - 0.7) < 1e-3   # albedo = colour


def test_env_map_is_sampled_in_proportion_to_luminance():
    """EnvMap::sample picks texel (x, y) with probability lum / sum and reports pdf = that probability * texels / 4 pi
    (equal-area mapping: every texel subtends the same solid angle); eval() of the sampled direction agrees."""
    rs = np.random.RandomState(5)
    n = 8
    rgba = np.ones((n, n, 4), np.float32)
    rgba[..., :3] = rs.rand(n, n, 3).astype(np.float32) ** 3 * 4 + 0.01   # uneven, strictly positive
    o = Oracle()
    o.set_background(rgba)
    rgb_dev, lum_levels = o.env_data()
    size = rgb_dev.shape[0]
    assert size == n
    lum = lum_levels[: n * n].reshape(n, n).astype(np.float64)
    m = 400
    u = (np.stack(np.meshgrid(np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 2) + 0.5) / m
    out = o.env_probe(0, np.concatenate([u, np.zeros((len(u), 1))], 1).astype(np.float32)).astype(np.float64)
    dirs, radiance, pdf = out[:, :3], out[:, 3:6], out[:, 6]
    np.testing.assert_allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-5)
    # which texel a sample landed in: its radiance is that texel's colour (colours are distinct)
    flat = rgb_dev.reshape(-1, 4)[:, :3].astype(np.float64)
    tex = np.argmin(((radiance[:, None, :] - flat[None, :, :]) ** 2).sum(-1), axis=1)
    freq = np.bincount(tex, minlength=n * n) / len(u)
    prob = (lum / lum.sum()).reshape(-1)
    np.testing.assert_allclose(freq, prob, atol=2e-3)
    np.testing.assert_allclose(pdf, prob[tex] * n * n / (4 * math.pi), rtol=1e-4)
    ev = o.env_probe(1, dirs.astype(np.float32)).astype(np.float64)
    np.testing.assert_allclose(ev[:, 3], pdf, rtol=1e-4)
    np.testing.assert_allclose(ev[:, :3], radiance, rtol=1e-6)


# ---- integrator level: direct lighting against closed forms / plain quadrature -------------------------------------------------
def _floor_scene(albedo, extent=(24, 24), spr=256):
    s = scenes.Scene("floor", extent=extent, lens=scenes.Lens((3.0, 0.0, 1.5), (-3.0, 0.0, -1.5), (0, 0, 1), math.radians(1.0), 0.0, 1.0))
    n, black, c = s.solid2(0.5, 0.5), s.solid3(0, 0, 0), s.solid3(albedo, albedo, albedo)
    floor_mat = s.add_material(type=LAMBERT, normal=n, emissive=black, color=c)
    quad = np.array([[-10, -10, 0], [10, -10, 0], [10, 10, 0], [-10, 10, 0]], np.float32)
    floor = s.add_mesh(quad, np.array([[0, 1, 2], [0, 2, 3]], np.uint32))
    s.add_instance([(floor, floor_mat, False)])
    return s, n, black


def _mean_rgb(sc):
    o = Oracle()
    scenes.load_scene(o, sc)
    img, counts = o.render(sc.lens, *sc.extent)
    assert np.isfinite(img).all()
    return img[..., :3].reshape(-1, 3).mean(0).astype(np.float64), counts


def test_square_light_over_a_lambert_floor_matches_the_form_factor():
    """Radiance leaving a Lambert floor under a one-sided square emitter = albedo * L * F, F the closed-form configuration factor of a
    differential element to a parallel rectangle. Holds MeshLights::sample, the area -> solid-angle measure, the emission test and the
    power-heuristic weights (light.hlsl:105-158, integrator.hlsl:10-35, :109-124) to the radiometry they implement: BSDF sampling
    alone, light sampling with MIS, and two light samples per bounce must all land on the same number."""
    albedo, L, a, h = 0.6, 5.0, 1.0, 1.0
    x = (a / 2) / h
    corner = (x / math.sqrt(1 + x * x) * math.atan(x / math.sqrt(1 + x * x))) * 2 / (2 * math.pi)
    expect = albedo * L * 4 * corner
    assert abs(expect - 0.7183) < 1e-3
    for mesh_samples in (0, 1, 2):
        sc, n, black = _floor_scene(albedo)
        light_mat = sc.add_material(type=LAMBERT, normal=n, emissive=sc.solid3(L, L, L), color=black)
        q = a / 2
        light = sc.add_mesh(np.array([[-q, -q, h], [-q, q, h], [q, q, h], [q, -q, h]], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32))   # faces -z
        sc.add_instance([(light, light_mat, mesh_samples > 0)])
        sc.background = np.array([[[0, 0, 0, 1]]], np.float32)
        sc.constants = scenes.Constants(samples_per_run=256, max_bounces=1, env_samples_per_bounce=0, mesh_samples_per_bounce=mesh_samples)
        sc.build_emitter_table = mesh_samples > 0
        rgb, counts = _mean_rgb(sc)
        assert (counts[1] > 0) == (mesh_samples > 0)
        np.testing.assert_allclose(rgb, expect, rtol=0.006, err_msg=f"mesh_samples={mesh_samples}")


def test_uneven_sky_over_a_lambert_floor_matches_quadrature_of_the_same_sky():
    """Radiance leaving a Lambert floor under an environment map = albedo / pi * integral of L cos over the hemisphere. The integral is
    taken by plain quadrature through EnvMap::incomingRadiance (the lookup the integrator itself uses for escaping rays); BSDF sampling
    alone, env importance sampling with MIS and three env samples per bounce must agree with it."""
    rs = np.random.RandomState(11)
    sky = np.ones((16, 16, 4), np.float32)
    sky[..., :3] = (rs.rand(16, 16, 1) ** 4 * 6 + 0.05).astype(np.float32) * np.array([1.0, 0.8, 0.6], np.float32)
    albedo = 0.7
    d, dw = _upper_hemisphere(384, 768)
    o = Oracle()
    o.set_background(sky)
    # escaping rays read the sky bilinearly when there is no env sampling (EnvMap::incomingRadiance, op 2) and by nearest texel when
    # there is (EnvMap::eval, op 1, light.hlsl:83-102): each estimator is held to the quadrature of the sky it integrates
    expect_by_lookup = {op: albedo / math.pi * (o.env_probe(op, d)[:, :3].astype(np.float64) * d[:, 2:3]).sum(0) * dw for op in (1, 2)}
    for env_samples in (0, 1, 3):
        expect = expect_by_lookup[2 if env_samples == 0 else 1]
        sc, _, _ = _floor_scene(albedo, spr=512)
        sc.background = sky
        sc.constants = scenes.Constants(samples_per_run=512, max_bounces=1, env_samples_per_bounce=env_samples, mesh_samples_per_bounce=0)
        sc.build_emitter_table = False
        rgb, _ = _mean_rgb(sc)
        np.testing.assert_allclose(rgb, expect, rtol=0.006, err_msg=f"env_samples={env_samples}")


# ---- texturing: the software sampler against the Vulkan filtering rules, written independently in float64 -----------------------
def _vulkan_linear(texels, uv, mirror):
    """VK_FILTER_LINEAR at LOD 0 with REPEAT / MIRRORED_REPEAT addressing (Vulkan spec, 'Texel Filtering' and 'Wrapping Operation'):
    unnormalised coordinate u = s * width - 0.5, i0 = floor(u), weight = frac(u), neighbours i0 and i0 + 1 wrapped per axis."""
    h, w = texels.shape[:2]
    u, v = uv[:, 0].astype(np.float64) * w - 0.5, uv[:, 1].astype(np.float64) * h - 0.5
    i0, j0 = np.floor(u).astype(np.int64), np.floor(v).astype(np.int64)
    a, b = (u - i0)[:, None], (v - j0)[:, None]

    def wrap(i, n):
        if not mirror:
            return i % n
        m = i % (2 * n)
        return np.where(m < n, m, 2 * n - 1 - m)   # mirror(n - 1 - mirror((i mod 2n) - n)) of the spec, simplified

    t = texels.astype(np.float64)
    i1, j1 = wrap(i0 + 1, w), wrap(j0 + 1, h)
    i0, j0 = wrap(i0, w), wrap(j0, h)
    return (t[j0, i0] * (1 - a) + t[j0, i1] * a) * (1 - b) + (t[j1, i0] * (1 - a) + t[j1, i1] * a) * b


def _srgb_eotf(c8):
    s = c8.astype(np.float64) / 255.0
    return np.where(s <= 0.04045, s / 12.92, ((s + 0.055) / 1.055) ** 2.4)


def test_texture_sampler_follows_the_vulkan_filtering_rules_for_every_format():
    rs = np.random.RandomState(9)
    w, h = 13, 7   # not powers of two, not square
    uv = np.concatenate([rs.rand(4000, 2) * 6 - 3,                                   # several wraps either side of [0, 1)
                         (rs.randint(-3 * w, 3 * w, (500, 2)) + 0.5) / np.array([w, h]),   # texel centres: no blending
                         rs.randint(-2, 3, (100, 2)).astype(np.float64)]).astype(np.float32)   # texture corners
    o = Oracle()
    cases = []
    f32 = rs.randn(h, w, 4).astype(np.float32)
    cases.append((scenes.TEX_R32F, f32[..., 0].copy(), np.concatenate([f32[..., :1], np.zeros((h, w, 2)), np.ones((h, w, 1))], -1)))
    cases.append((scenes.TEX_RG32F, f32[..., :2].copy(), np.concatenate([f32[..., :2], np.zeros((h, w, 1)), np.ones((h, w, 1))], -1)))
    cases.append((scenes.TEX_RGBA32F, f32.copy(), f32))
    u8 = rs.randint(0, 256, (h, w, 4)).astype(np.uint8)
    cases.append((scenes.TEX_R8_UNORM, u8[..., 0].copy(), np.concatenate([u8[..., :1] / 255.0, np.zeros((h, w, 2)), np.ones((h, w, 1))], -1)))
    cases.append((scenes.TEX_RG8_UNORM, u8[..., :2].copy(), np.concatenate([u8[..., :2] / 255.0, np.zeros((h, w, 1)), np.ones((h, w, 1))], -1)))
    cases.append((scenes.TEX_RGBA8_SRGB, u8.copy(), np.concatenate([_srgb_eotf(u8[..., :3]), u8[..., 3:] / 255.0], -1)))   # alpha stays linear
    f16 = (rs.randn(h, w, 4) * 3).astype(np.float16)
    cases.append((scenes.TEX_RGBA16F, f16.copy(), f16.astype(np.float64)))
    for fmt, data, decoded in cases:
        tex = o.create_texture(data, w, h, fmt)
        for mirror in (False, True):
            got = o.texture_probe(tex, uv, mirror).astype(np.float64)
            want = _vulkan_linear(decoded, uv, mirror)
            np.testing.assert_allclose(got, want, rtol=0, atol=2e-5 * max(1.0, np.abs(decoded).max()), err_msg=f"format {fmt} mirror {mirror}")
    # a 1x1 image is its texel whatever the coordinate
    one = o.create_texture(np.array([0.25, 0.5, 0.75, 1.0], np.float32), 1, 1, scenes.TEX_RGBA32F)
    np.testing.assert_array_equal(o.texture_probe(one, uv[:50]), np.broadcast_to(np.array([0.25, 0.5, 0.75, 1.0], np.float32), (50, 4)))


# ---- camera: the thin lens against geometric optics ---------------------------------------------------------------------------
def test_thin_lens_rays_focus_on_the_focal_plane_and_cover_the_field_of_view():
    """camera.hlsl:14-43 + the pixel filter of main.hlsl:85-88, held to optics: (a) with a pinhole the mean direction of a pixel's
    samples is the ray through that pixel's centre of a vfov / aspect frustum (top row at +vfov/2: image y runs downwards);
    (b) with an aperture, origins fill a disc of diameter `aperture` around the lens centre, perpendicular to the view direction, and
    every ray of a sample passes through the point of the focal plane its pinhole twin passes through."""
    w, h, vfov, focus = 48, 24, math.radians(50.0), 2.5
    origin, forward, up = np.array([0.3, -1.0, 0.7]), _unit((1.0, 0.4, -0.2)).astype(np.float64), np.array([0.0, 0.0, 1.0])
    right = np.cross(forward, up); right /= np.linalg.norm(right)
    true_up = np.cross(right, forward)
    o = Oracle()
    pin = scenes.Lens(tuple(origin), tuple(forward), tuple(up), vfov, 0.0, focus)
    acc = np.zeros((h, w, 3))
    n_samples = 48
    for s in range(n_samples):
        rays, _ = o.primary_rays(pin, w, h, s)
        np.testing.assert_allclose(rays[..., :3], np.broadcast_to(origin.astype(np.float32), (h, w, 3)), atol=1e-6)
        d = rays[..., 3:6].astype(np.float64)
        np.testing.assert_allclose(np.linalg.norm(d, axis=-1), 1.0, atol=1e-5)
        # where the ray meets the plane one unit in front of the eye, in (right, up) coordinates
        t = 1.0 / (d @ forward)
        acc += np.stack([(d @ right) * t, (d @ true_up) * t, t], -1)
    mean = acc / n_samples
    th = math.tan(vfov / 2)
    px = ((np.arange(w) + 0.5) / w * 2 - 1) * th * (w / h)
    py = (1 - (np.arange(h) + 0.5) / h * 2) * th
    pixel = 2 * th / h   # one pixel on that plane
    assert np.abs(mean[..., 0] - px[None, :]).max() < 0.35 * pixel   # mean of 48 filter offsets (sigma ~ half a pixel)
    assert np.abs(mean[..., 1] - py[:, None]).max() < 0.35 * pixel
    # (b)
    aperture = 0.2
    lens = scenes.Lens(tuple(origin), tuple(forward), tuple(up), vfov, aperture, focus)
    radii = []
    for s in range(8):
        rp, _ = o.primary_rays(pin, w, h, s)
        ra, _ = o.primary_rays(lens, w, h, s)
        off = ra[..., :3].astype(np.float64) - origin
        assert np.abs(off @ forward).max() < 1e-6                     # the lens disc is perpendicular to the view direction
        radii.append(np.linalg.norm(off, axis=-1).ravel())
        # both rays of a sample meet on the focal plane (the plane `focus` in front of the eye)
        def on_focal_plane(r):
            oo, dd = r[..., :3].astype(np.float64), r[..., 3:6].astype(np.float64)
            t = (focus - (oo - origin) @ forward) / (dd @ forward)
            return oo + dd * t[..., None]
        np.testing.assert_allclose(on_focal_plane(ra), on_focal_plane(rp), atol=5e-5)
    radii = np.concatenate(radii)
    assert radii.max() <= aperture / 2 + 1e-6 and radii.max() > 0.49 * aperture
    assert abs((radii ** 2).mean() / (aperture / 2) ** 2 - 0.5) < 0.02   # uniform over the disc: E[r^2] = R^2 / 2


# ---- normal mapping: tangent-space convention ------------------------------------------------------------------------------------
def test_normal_map_tilts_the_shading_normal_along_increasing_u():
    """A tangent-space normal (x, y, z) means x along dP/du, y along dP/dv, z along the surface normal (world.hlsl:86-100 builds the
    frame from the texcoord derivatives, material.hlsl:489-517 decodes and applies the map). A floor whose u runs along world +x, lit
    by a small far emitter 45 degrees towards +x: tilting the map's normal 30 degrees towards +x must brighten the Lambert response
    by cos 15 / cos 45, tilting it towards -x must darken it by cos 75 / cos 45 - the geometry term does not move."""
    rho, L = 0.8, 4000.0
    ldir = np.array([math.sin(math.radians(45)), 0.0, math.cos(math.radians(45))])

    def render(nx):
        s = scenes.Scene("normal-map", extent=(16, 16), lens=scenes.Lens((0.0, -3.0, 3.0), (0.0, 3.0, -3.0), (0, 0, 1), math.radians(1.0), 0.0, 1.0))
        black, white = s.solid3(0, 0, 0), s.solid3(rho, rho, rho)
        floor_mat = s.add_material(type=LAMBERT, normal=s.solid2(0.5 + 0.5 * nx, 0.5), emissive=black, color=white)   # rg = n.xy * 0.5 + 0.5
        quad = np.array([[-4, -4, 0], [4, -4, 0], [4, 4, 0], [-4, 4, 0]], np.float32)
        uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)                                                    # u along +x, v along +y
        nrm = np.array([[0, 0, 1]] * 4, np.float32)
        floor = s.add_mesh(quad, np.array([[0, 1, 2], [0, 2, 3]], np.uint32), normals=nrm, texcoords=uv)
        s.add_instance([(floor, floor_mat, False)])
        # emitter: a 0.2 x 0.2 square 10 units away along ldir, facing the origin
        c = ldir * 10.0
        e1 = np.array([0.0, 1.0, 0.0])
        e2 = np.cross(ldir, e1)
        corners = np.array([c + 0.1 * (-e1 - e2), c + 0.1 * (e1 - e2), c + 0.1 * (e1 + e2), c + 0.1 * (-e1 + e2)], np.float32)
        light_mat = s.add_material(type=LAMBERT, normal=s.solid2(0.5, 0.5), emissive=s.solid3(L, L, L), color=black)
        tri = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
        n_geo = np.cross(corners[0] - corners[2], corners[1] - corners[2])
        if np.dot(n_geo, -ldir) < 0:          # emission leaves the side the geometric normal points to: towards the origin
            tri = tri[:, ::-1].copy()
        light = s.add_mesh(corners, tri)
        s.add_instance([(light, light_mat, True)])
        s.background = np.array([[[0, 0, 0, 1]]], np.float32)
        s.constants = scenes.Constants(samples_per_run=512, max_bounces=1, env_samples_per_bounce=0, mesh_samples_per_bounce=1)
        rgb, counts = _mean_rgb(s)
        assert counts[1] > 0
        return rgb[0]

    flat, plus, minus = render(0.0), render(0.5), render(-0.5)
    irradiance = L * 0.04 / 100.0 * math.cos(math.radians(45))          # L * A / r^2 * cos at the floor (the emitter faces the floor point)
    assert abs(flat / (rho / math.pi * irradiance) - 1) < 0.005
    assert abs(plus / flat - math.cos(math.radians(15)) / math.cos(math.radians(45))) < 0.005
    assert abs(minus / flat - math.cos(math.radians(75)) / math.cos(math.radians(45))) < 0.005


# ---- instancing: a transformed instance is the transformed mesh -------------------------------------------------------------------
def test_an_instance_transform_renders_the_transformed_mesh():
    """Rays enter an instance through worldToInstance, hit attributes leave it through inWorld: positions by the instance matrix,
    frames by the inverse transpose (world.hlsl:160-176, reflection_frame.hlsl:23-29). (a) Rotation, uniform scale and translation:
    the instanced unit sphere is, pixel for pixel, the sphere baked into world-space vertices (same random streams, same frames up
    to rounding). (b) Non-uniform scale: tangents carried by the inverse transpose differ from tangents of the baked mesh, so samples
    are rotated about the normal - the images agree in expectation only - but the NORMAL must be the inverse-transpose one: baking
    normals like directions shades the ellipsoid visibly differently."""
    pos, idx = scenes.icosphere(3)
    pos = np.asarray(pos, np.float64)
    nrm = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    a = math.radians(35.0)
    rot = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
    t = np.array([0.2, -0.1, 0.15])
    rs = np.random.RandomState(4)
    sky = np.ones((8, 8, 4), np.float32)
    sky[..., :3] = (rs.rand(8, 8, 1) ** 2 * 3 + 0.1).astype(np.float32)   # uneven light: shading depends on the normal

    def render(p, n, xf, spr):
        s = scenes.Scene("instancing", extent=(48, 48))
        mat = s.add_material(type=LAMBERT, normal=s.solid2(0.5, 0.5), emissive=s.solid3(0, 0, 0), color=s.solid3(0.7, 0.7, 0.7))
        mesh = s.add_mesh(p.astype(np.float32), idx, normals=n.astype(np.float32))
        s.add_instance([(mesh, mat, False)], transform=xf)
        s.background = sky
        s.constants = scenes.Constants(samples_per_run=spr, max_bounces=4, env_samples_per_bounce=1, mesh_samples_per_bounce=0)
        s.build_emitter_table = False
        o = Oracle()
        scenes.load_scene(o, s)
        img, _ = o.render(s.lens, *s.extent)
        return img[..., :3].astype(np.float64)

    def baked(m, normals_by):
        n = nrm @ normals_by
        return pos @ m.T + t, n / np.linalg.norm(n, axis=1, keepdims=True)

    # (a) similarity transform
    m = rot * 0.9
    instanced = render(pos, nrm, np.concatenate([m, t[:, None]], 1), 8)
    ref = render(*baked(m, np.linalg.inv(m)), None, 8)
    close = np.abs(instanced - ref).max(-1) <= 2e-3 * np.maximum(1.0, ref.max(-1))
    assert close.mean() > 0.99, close.mean()          # only silhouette / grazing pixels may take other paths
    # (b) non-uniform scale, compared over 6 x 6 pixel blocks at 64 spp
    m = rot @ np.diag([1.6, 0.8, 0.5])
    blocks = lambda img: img.reshape(8, 6, 8, 6, 3).mean((1, 3))
    instanced = blocks(render(pos, nrm, np.concatenate([m, t[:, None]], 1), 64))
    right = blocks(render(*baked(m, np.linalg.inv(m)), None, 64))       # rows: (M^-T n)^T
    wrong = blocks(render(*baked(m, m.T), None, 64))                    # normals carried like directions
    err_right, err_wrong = np.abs(instanced - right).mean(), np.abs(instanced - wrong).mean()
    assert err_right < 0.005 * right.mean(), (err_right, right.mean())
    assert err_wrong > 8 * err_right, (err_wrong, err_right)


def test_unevenly_tessellated_one_sided_emitter():
    """The same square light cut into four triangles of very different areas (a fan around an off-centre point) must light the floor
    exactly as before: MeshLights::sample has to pick triangles with the probability its pdf assumes (area / total area, alias table
    of Accel.zig:491-538). Seen from behind, the emitter is black: emission leaves the side of the geometric normal only."""
    albedo, L, a, h = 0.6, 5.0, 1.0, 1.0
    x = (a / 2) / h
    expect = albedo * L * 4 * (x / math.sqrt(1 + x * x) * math.atan(x / math.sqrt(1 + x * x))) * 2 / (2 * math.pi)
    q = a / 2
    verts = np.array([[-q, -q, h], [-q, q, h], [q, q, h], [q, -q, h], [0.35, -0.4, h]], np.float32)   # 4 = fan centre, near one corner
    fan = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], np.uint32)                               # faces -z like the two-triangle quad
    areas = [0.5 * np.linalg.norm(np.cross(verts[t[1]] - verts[t[0]], verts[t[2]] - verts[t[0]])) for t in fan]
    assert max(areas) / min(areas) > 8 and abs(sum(areas) - a * a) < 1e-6
    sc, n, black = _floor_scene(albedo)
    light_mat = sc.add_material(type=LAMBERT, normal=n, emissive=sc.solid3(L, L, L), color=black)
    sc.add_instance([(sc.add_mesh(verts, fan), light_mat, True)])
    sc.background = np.array([[[0, 0, 0, 1]]], np.float32)
    sc.constants = scenes.Constants(samples_per_run=256, max_bounces=1, env_samples_per_bounce=0, mesh_samples_per_bounce=1)
    rgb, counts = _mean_rgb(sc)
    assert counts[1] > 0
    np.testing.assert_allclose(rgb, expect, rtol=0.006)
    # the back of the emitter, looked at from above
    sc.lens = scenes.Lens((0.0, 0.0, 4.0), (0.0, 0.0, -1.0), (0, 1, 0), math.radians(2.0), 0.0, 1.0)
    o = Oracle()
    scenes.load_scene(o, sc)
    img, _ = o.render(sc.lens, *sc.extent)
    assert np.abs(img[..., :3]).max() == 0.0     # back face: no emission; it is black (albedo 0) and hides the lit floor below it
    sc.lens = scenes.Lens((0.0, 0.0, 0.2), (0.0, 0.0, 1.0), (0, 1, 0), math.radians(2.0), 0.0, 1.0)
    o = Oracle()
    scenes.load_scene(o, sc)
    img, _ = o.render(sc.lens, *sc.extent)
    np.testing.assert_allclose(img[..., :3], L, rtol=1e-6)   # front face, looked at from the floor: its radiance


# ---- the scenes above, for the GPU twin (tests/test_gpu_parity.py::test_physics_scenes_are_bit_identical_on_the_gpu) ----------------
def physics_scenes():
    """Small scenes in the style of the checks above that lean on code paths the procedural BASELINE scenes use lightly: explicit
    texcoords + vertex normals + a tilted constant normal map, an unevenly tessellated one-sided emitter with two light samples per
    bounce, an uneven sky with three env samples per bounce, a non-uniformly scaled instance."""
    out = []
    # normal-mapped floor under a far emitter
    s = scenes.Scene("phys-normal-map", extent=(16, 16), lens=scenes.Lens((0.0, -3.0, 3.0), (0.0, 3.0, -3.0), (0, 0, 1), math.radians(4.0), 0.0, 1.0))
    black, white = s.solid3(0, 0, 0), s.solid3(0.8, 0.8, 0.8)
    floor_mat = s.add_material(type=LAMBERT, normal=s.solid2(0.75, 0.4), emissive=black, color=white)
    quad = np.array([[-4, -4, 0], [4, -4, 0], [4, 4, 0], [-4, 4, 0]], np.float32)
    floor = s.add_mesh(quad, np.array([[0, 1, 2], [0, 2, 3]], np.uint32), normals=np.array([[0, 0, 1]] * 4, np.float32),
                       texcoords=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32))
    s.add_instance([(floor, floor_mat, False)])
    light_mat = s.add_material(type=LAMBERT, normal=s.solid2(0.5, 0.5), emissive=s.solid3(400, 300, 200), color=black)
    verts = np.array([[-0.5, -0.5, 2], [-0.5, 0.5, 2], [0.5, 0.5, 2], [0.5, -0.5, 2], [0.35, -0.4, 2]], np.float32)
    fan = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], np.uint32)
    s.add_instance([(s.add_mesh(verts, fan), light_mat, True)])
    s.background = np.array([[[0, 0, 0, 1]]], np.float32)
    s.constants = scenes.Constants(samples_per_run=64, max_bounces=3, env_samples_per_bounce=0, mesh_samples_per_bounce=2)
    out.append(s)
    # uneven sky, three env samples per bounce, a non-uniformly scaled and rotated instance with vertex normals
    rs = np.random.RandomState(11)
    sky = np.ones((16, 16, 4), np.float32)
    sky[..., :3] = (rs.rand(16, 16, 1) ** 4 * 6 + 0.05).astype(np.float32) * np.array([1.0, 0.8, 0.6], np.float32)
    s = scenes.Scene("phys-sky-instance", extent=(40, 40))
    mat = s.add_material(type=LAMBERT, normal=s.solid2(0.5, 0.5), emissive=s.solid3(0, 0, 0), color=s.solid3(0.7, 0.7, 0.7))
    pos, idx = scenes.icosphere(3)
    pos = np.asarray(pos, np.float32)
    mesh = s.add_mesh(pos, idx, normals=pos / np.linalg.norm(pos, axis=1, keepdims=True))
    a = math.radians(35.0)
    m = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]]) @ np.diag([1.6, 0.8, 0.5])
    s.add_instance([(mesh, mat, False)], transform=np.concatenate([m, [[0.2], [-0.1], [0.15]]], 1))
    s.add_instance([(mesh, mat, False)], transform=np.array([[0.4, 0, 0, -0.2], [0, 0.4, 0, 0.9], [0, 0, 0.4, 0.6]], np.float32))
    s.background = sky
    s.constants = scenes.Constants(samples_per_run=16, max_bounces=6, env_samples_per_bounce=3, mesh_samples_per_bounce=0)
    s.build_emitter_table = False
    out.append(s)
    return out


def test_physics_scenes_render_on_the_oracle():
    for sc in physics_scenes():
        rgb, counts = _mean_rgb(sc)
        assert (rgb > 0).all() and counts[0] > 0 and counts[1] > 0


def test_mirror_floor_shows_the_sky_in_the_reflected_direction():
    """One sample per pixel on a perfect-mirror floor: the pixel must hold the sky's radiance in the direction of its own primary ray
    reflected about the floor normal (d.x, d.y, -d.z) - law of reflection, unit throughput (eval * |cos| / pdf of material.hlsl:313-330),
    and the pixel <-> primary-ray correspondence of raygen (main.hlsl:54-95), all at once and per pixel."""
    rs = np.random.RandomState(21)
    sky = np.ones((32, 32, 4), np.float32)
    sky[..., :3] = (rs.rand(32, 32, 3) ** 2 * 3 + 0.05).astype(np.float32)
    s = scenes.Scene("mirror-floor", extent=(40, 24), lens=scenes.Lens((0.0, -2.0, 2.0), (0.0, 1.0, -1.0), (0, 0, 1), math.radians(40.0), 0.0, 1.0))
    mat = s.add_material(type=scenes.MAT_PERFECT_MIRROR, normal=s.solid2(0.5, 0.5), emissive=s.solid3(0, 0, 0))
    quad = np.array([[-50, -50, 0], [50, -50, 0], [50, 50, 0], [-50, 50, 0]], np.float32)
    s.add_instance([(s.add_mesh(quad, np.array([[0, 1, 2], [0, 2, 3]], np.uint32)), mat, False)])
    s.background = sky
    s.constants = scenes.Constants(samples_per_run=1, max_bounces=8, env_samples_per_bounce=1, mesh_samples_per_bounce=0)
    s.build_emitter_table = False
    o = Oracle()
    scenes.load_scene(o, s)
    img, counts = o.render(s.lens, *s.extent)
    rays, _ = o.primary_rays(s.lens, *s.extent, 0)
    d = rays[..., 3:6].reshape(-1, 3).copy()
    assert (d[:, 2] < 0).all()                       # every primary ray goes down to the floor
    d[:, 2] = -d[:, 2]
    expect = o.env_probe(2, d)[:, :3].reshape(img.shape[0], img.shape[1], 3)
    np.testing.assert_allclose(img[..., :3], expect, rtol=2e-5, atol=1e-6)
    assert counts[0] == 2 * img.shape[0] * img.shape[1] and counts[1] == 0   # primary + reflected ray per pixel, no shadow rays on a delta surface


def test_glass_slab_transmits_undeviated_and_reflects_like_a_stack_of_two_interfaces():
    """A plane-parallel glass slab under a sky, one sample per pixel: whatever sequence of refractions and internal reflections a path
    takes, it leaves either upwards in the mirrored direction or downwards in its ORIGINAL direction (Snell at the way in undone at the
    way out) with unit throughput; and the share of paths that come back up is the incoherent two-interface reflectance 2R / (1 + R),
    R the Fresnel reflectance of one interface at that angle."""
    rs = np.random.RandomState(22)
    sky = np.ones((32, 32, 4), np.float32)
    sky[..., :3] = (rs.rand(32, 32, 3) ** 2 * 3 + 0.05).astype(np.float32)
    ior, air, incidence = 1.5, 1.000277, math.radians(55.0)
    fwd = (0.0, math.sin(incidence), -math.cos(incidence))
    s = scenes.Scene("glass-slab", extent=(96, 96), lens=scenes.Lens((0.0, -3 * math.sin(incidence), 3 * math.cos(incidence)), fwd, (0, 0, 1), math.radians(1.0), 0.0, 1.0))
    mat = s.add_material(type=GLASS, normal=s.solid2(0.5, 0.5), emissive=s.solid3(0, 0, 0), ior=ior)
    e = 200.0
    top = np.array([[-e, -e, 0], [e, -e, 0], [e, e, 0], [-e, e, 0]], np.float32)
    bottom = top[::-1].copy() - np.array([0, 0, 0.5], np.float32)          # faces -z
    tri = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    s.add_instance([(s.add_mesh(top, tri), mat, False), (s.add_mesh(bottom, tri), mat, False)])
    s.background = sky
    s.constants = scenes.Constants(samples_per_run=1, max_bounces=64, env_samples_per_bounce=1, mesh_samples_per_bounce=0)
    s.build_emitter_table = False
    o = Oracle()
    scenes.load_scene(o, s)
    img, counts = o.render(s.lens, *s.extent)
    rays, _ = o.primary_rays(s.lens, *s.extent, 0)
    d = rays[..., 3:6].reshape(-1, 3)
    through = o.env_probe(2, d)[:, :3]
    mirrored = o.env_probe(2, d * np.array([1, 1, -1], np.float32))[:, :3]
    got = img[..., :3].reshape(-1, 3).astype(np.float64)

    def weight_if_match(target):
        """got = w * target with one w for the three channels and w = 0.95^-k, k = 0, 1, 2 ...: unit throughput, times one division by
        the survival probability for every Russian-roulette test passed after the third bounce (integrator.hlsl:128-135) -> w, else nan"""
        ratio = got / target
        same = np.abs(ratio - ratio[:, :1]).max(1) < 3e-4
        k = np.log(np.maximum(ratio[:, 0], 1e-30)) / -math.log(0.95)
        ok = same & (np.abs(k - np.round(k)) < 2e-2) & (k > -0.5)
        return np.where(ok, ratio[:, 0], np.nan)

    w_through, w_mirrored = weight_if_match(through.astype(np.float64)), weight_if_match(mirrored.astype(np.float64))
    killed = (got == 0).all(1)                          # ended by Russian roulette inside the slab
    is_through, is_mirrored = ~np.isnan(w_through), ~np.isnan(w_mirrored)
    assert (is_through | is_mirrored | killed).all()    # nothing else comes out of a slab
    assert not (is_through & is_mirrored).any()
    assert killed.mean() < 0.002 and (np.nan_to_num(w_through) + np.nan_to_num(w_mirrored) > 1.01).mean() < 0.002   # roulette only touches paths of 4+ bounces
    cos_i = math.cos(incidence)
    sin_t = math.sin(incidence) * air / ior
    cos_t = math.sqrt(1 - sin_t * sin_t)
    rp = (ior * cos_i - air * cos_t) / (ior * cos_i + air * cos_t)
    rs_ = (air * cos_i - ior * cos_t) / (air * cos_i + ior * cos_t)
    r1 = 0.5 * (rp * rp + rs_ * rs_)
    expect = 2 * r1 / (1 + r1)
    n = len(got)
    reflected, transmitted = np.nansum(w_mirrored) / n, np.nansum(w_through) / n      # unbiased estimates of the two energy shares
    assert abs(reflected - expect) < 4 * math.sqrt(expect * (1 - expect) / n) + 2e-3, (reflected, expect)
    assert abs(reflected + transmitted - 1.0) < 2e-3                                   # energy is conserved
    assert counts[1] == 0
