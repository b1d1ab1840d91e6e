"""Backend-neutral scene descriptions + the procedural scenes of BASELINE.json's configs.

The reference ships no assets, so every benchmark/test scene is generated here with
fixed seeds and pushed through the C ABI (`HdMoonshine*` / `HdMoonshineExt*`). Engine space is +Z up.
A `Scene` is plain data; `load_scene(backend, scene)` replays it into any object exposing the
backend protocol (moonshine_b200.hydra.HdMoonshine for the CUDA library, oracle/binding.py for the
CPU oracle in tests).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# TextureFormat codes shared with include/moonshine_ext.h (MsneTexFormat)
TEX_R32F, TEX_RG32F, TEX_RGBA32F, TEX_R8_UNORM, TEX_RG8_UNORM, TEX_RGBA8_SRGB, TEX_RGBA16F = range(7)
# MaterialType (shaders/hrtsystem/world.hlsl:31-36)
MAT_GLASS, MAT_LAMBERT, MAT_PERFECT_MIRROR, MAT_STANDARD_PBR = range(4)


@dataclass
class Texture:
    data: np.ndarray  # raw bytes/floats, C-contiguous
    width: int
    height: int
    fmt: int


@dataclass
class Material:
    type: int
    normal: int
    emissive: int
    color: int = 0
    metalness: int = 0
    roughness: int = 0
    ior: float = 1.5


@dataclass
class Mesh:
    positions: np.ndarray  # (n,3) f32
    indices: np.ndarray  # (m,3) u32
    normals: Optional[np.ndarray] = None  # (n,3) when indexed_attributes else (3m,3)
    texcoords: Optional[np.ndarray] = None  # (n,2) / (3m,2)


@dataclass
class Instance:
    transform: np.ndarray  # (3,4) f32 row-major
    geometries: List[Tuple[int, int, bool]]  # (mesh, material, sampled)
    visible: bool = True


@dataclass
class Lens:
    origin: Tuple[float, float, float]
    forward: Tuple[float, float, float]
    up: Tuple[float, float, float]
    vfov: float
    aperture: float = 0.0
    focus_distance: float = 1.0

    def as_array(self) -> np.ndarray:
        return np.array([*self.origin, *self.forward, *self.up, self.vfov, self.aperture, self.focus_distance], dtype=np.float32)


@dataclass
class Constants:
    """Specialization constants (shaders/hrtsystem/main.hlsl:34-40). Defaults = offline/main.zig:106-111."""
    samples_per_run: int = 1
    max_bounces: int = 1024
    env_samples_per_bounce: int = 1
    mesh_samples_per_bounce: int = 1
    flip_image: bool = True
    indexed_attributes: bool = True
    two_component_normal_texture: bool = True


@dataclass
class Scene:
    name: str
    textures: List[Texture] = field(default_factory=list)
    materials: List[Material] = field(default_factory=list)
    meshes: List[Mesh] = field(default_factory=list)
    instances: List[Instance] = field(default_factory=list)
    background: Optional[np.ndarray] = None  # (H,W,4) f32 equirect; None = default 1x1 white
    constants: Constants = field(default_factory=Constants)
    lens: Lens = field(default_factory=lambda: Lens((-3, 0, 0), (1, 0, 0), (0, 0, 1), math.pi / 4))
    extent: Tuple[int, int] = (64, 64)  # (width, height)
    build_emitter_table: bool = True  # Accel.create semantics (Accel.zig:491-538)

    # -- small builders ---------------------------------------------------------------------
    def solid1(self, v: float) -> int:
        self.textures.append(Texture(np.array([v], np.float32), 1, 1, TEX_R32F))
        return len(self.textures) - 1

    def solid2(self, x: float, y: float) -> int:
        self.textures.append(Texture(np.array([x, y], np.float32), 1, 1, TEX_RG32F))
        return len(self.textures) - 1

    def solid3(self, r: float, g: float, b: float) -> int:
        # stored as f32x4 (MaterialManager.zig:364-373); the 4th lane is never sampled
        self.textures.append(Texture(np.array([r, g, b, 1.0], np.float32), 1, 1, TEX_RGBA32F))
        return len(self.textures) - 1

    def raw_texture(self, data: np.ndarray, fmt: int) -> int:
        h, w = data.shape[:2]
        self.textures.append(Texture(np.ascontiguousarray(data), w, h, fmt))
        return len(self.textures) - 1

    def add_material(self, **kw) -> int:
        self.materials.append(Material(**kw))
        return len(self.materials) - 1

    def add_mesh(self, positions, indices, normals=None, texcoords=None) -> int:
        self.meshes.append(Mesh(np.ascontiguousarray(positions, np.float32), np.ascontiguousarray(indices, np.uint32),
                                None if normals is None else np.ascontiguousarray(normals, np.float32),
                                None if texcoords is None else np.ascontiguousarray(texcoords, np.float32)))
        return len(self.meshes) - 1

    def add_instance(self, geometries, transform=None, visible=True) -> int:
        t = np.eye(3, 4, dtype=np.float32) if transform is None else np.ascontiguousarray(transform, np.float32)
        self.instances.append(Instance(t, list(geometries), visible))
        return len(self.instances) - 1

    def triangle_count(self) -> int:
        return sum(len(self.meshes[g[0]].indices) for i in self.instances for g in i.geometries)


def load_scene(backend, scene: Scene):
    """Replay `scene` into a backend; returns (texture, material, mesh, instance) handle lists."""
    backend.set_constants(scene.constants)  # first: indexed_attributes decides how mesh attributes are read
    tex = [backend.create_texture(t.data, t.width, t.height, t.fmt) for t in scene.textures]
    mats = [backend.create_material(m.type, tex[m.normal], tex[m.emissive], tex[m.color], tex[m.metalness], tex[m.roughness], m.ior)
            for m in scene.materials]
    meshes = [backend.create_mesh(m.positions, m.normals, m.texcoords, m.indices) for m in scene.meshes]
    insts = [backend.create_instance(i.transform, [(meshes[g[0]], mats[g[1]], g[2]) for g in i.geometries], i.visible)
             for i in scene.instances]
    if scene.background is not None:
        backend.set_background(scene.background)
    if scene.build_emitter_table:
        backend.build_emitter_table()
    return tex, mats, meshes, insts


# ---------------------------------------------------------------------------------------------
# geometry helpers
# ---------------------------------------------------------------------------------------------
def icosphere(order: int, reverse_winding: bool = False):
    """Unit icosphere, restating engine/tests.zig:115-247 (same vertex/triangle order)."""
    t = np.float32((1.0 + math.sqrt(5.0)) / 2.0)
    pos = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
           (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    pos = [np.array(p, np.float32) for p in pos]
    tris = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
            (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    cache = {}

    def midpoint(a, b):
        key = (min(a, b), max(a, b))
        if key in cache:
            return cache[key]
        pos.append(((pos[a] + pos[b]) / np.float32(2.0)).astype(np.float32))
        cache[key] = len(pos) - 1
        return cache[key]

    for _ in range(order):
        nxt = []
        for (x, y, z) in tris:
            a, b, c = midpoint(x, y), midpoint(y, z), midpoint(z, x)
            nxt += [(x, a, c), (y, b, a), (z, c, b), (a, b, c)]
        tris = nxt
    p = np.stack(pos).astype(np.float32)
    # Vec3.unit(): div_scalar(length), length = sqrt(x*x + y*y + z*z) in f32
    ln = np.sqrt((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2]).astype(np.float32)).astype(np.float32)
    p = (p / ln[:, None]).astype(np.float32)
    idx = np.array(tris, np.uint32)
    if reverse_winding:
        idx = idx[:, ::-1].copy()
    return p, idx


def _quad(p0, p1, p2, p3):
    """Two triangles (p0,p1,p2),(p0,p2,p3)."""
    return np.array([p0, p1, p2, p3], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)


def _merge(parts):
    ps, idx, off = [], [], 0
    for p, i in parts:
        ps.append(p)
        idx.append(i + off)
        off += len(p)
    return np.concatenate(ps).astype(np.float32), np.concatenate(idx).astype(np.uint32)


def _box(lo, hi, rot_z=0.0, center=None):
    lo, hi = np.array(lo, np.float32), np.array(hi, np.float32)
    c = [(lo[0], lo[1], lo[2]), (hi[0], lo[1], lo[2]), (hi[0], hi[1], lo[2]), (lo[0], hi[1], lo[2]),
         (lo[0], lo[1], hi[2]), (hi[0], lo[1], hi[2]), (hi[0], hi[1], hi[2]), (lo[0], hi[1], hi[2])]
    v = np.array(c, np.float32)
    if rot_z:
        ctr = v.mean(0) if center is None else np.array(center, np.float32)
        cs, sn = math.cos(rot_z), math.sin(rot_z)
        d = v - ctr
        v = np.stack([d[:, 0] * cs - d[:, 1] * sn, d[:, 0] * sn + d[:, 1] * cs, d[:, 2]], 1) + ctr
    # outward-facing (counter-clockwise seen from outside)
    f = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]
    idx = []
    for a, b, c_, d in f:
        idx += [(a, b, c_), (a, c_, d)]
    return v.astype(np.float32), np.array(idx, np.uint32)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs + parity scenes
# ---------------------------------------------------------------------------------------------
def test_lens() -> Lens:
    """engine/tests.zig:302-309."""
    return Lens((-3, 0, 0), (1, 0, 0), (0, 0, 1), math.pi / 4.0, 0.0, 1.0)


def single_triangle(extent=(64, 64)) -> Scene:
    """cfg1: one Lambert(0.8) triangle, 1x1 white env through addBackground, test lens, offline constants."""
    s = Scene("cfg1-single-triangle", extent=extent, lens=test_lens())
    n, e, c = s.solid2(0.5, 0.5), s.solid3(0, 0, 0), s.solid3(0.8, 0.8, 0.8)
    m = s.add_material(type=MAT_LAMBERT, normal=n, emissive=e, color=c)
    mesh = s.add_mesh([(0, -1, -1), (0, 1, -1), (0, 0, 1)], [(0, 1, 2)])
    s.add_instance([(mesh, m, False)])
    s.background = np.ones((1, 1, 4), np.float32)
    return s


def furnace_white_sphere(env_samples: int, spr: int = 512, order: int = 5, extent=(32, 32)) -> Scene:
    """engine/tests.zig:257-364 — "white sphere on white background is white"."""
    s = Scene("furnace-white-sphere", extent=extent, lens=test_lens())
    n, c, e = s.solid2(0.5, 0.5), s.solid3(1, 1, 1), s.solid3(0, 0, 0)
    m = s.add_material(type=MAT_LAMBERT, normal=n, emissive=e, color=c)
    p, i = icosphere(order, False)
    mesh = s.add_mesh(p, i)
    s.add_instance([(mesh, m, False)])
    s.background = np.ones((1, 1, 4), np.float32)
    s.constants = Constants(samples_per_run=spr, max_bounces=1024, env_samples_per_bounce=env_samples, mesh_samples_per_bounce=0)
    s.build_emitter_table = False  # uploadInstance path: "alias table currently unimplemented" (Accel.zig:187)
    return s


def furnace_material_sphere(kind: str, spr: int = 256, env_samples: int = 0, order: int = 4, extent=(32, 32)) -> Scene:
    """White-furnace variants of engine/tests.zig:257-364 for the BSDFs the reference's own tests never exercise:
    kind = "mirror" (PerfectMirror), "glass" (Glass, ior 1.5), "metal" (StandardPBR white, metalness 1, roughness 0.5),
    "dielectric" (StandardPBR white, metalness 0, roughness 0.5). Mirror and glass neither absorb nor gain (radiance 1);
    single-scattering GGX metal can only lose energy (radiance <= 1)."""
    s = Scene(f"furnace-{kind}-sphere", extent=extent, lens=test_lens())
    n, e, white = s.solid2(0.5, 0.5), s.solid3(0, 0, 0), s.solid3(1, 1, 1)
    if kind == "mirror":
        m = s.add_material(type=MAT_PERFECT_MIRROR, normal=n, emissive=e)
    elif kind == "glass":
        m = s.add_material(type=MAT_GLASS, normal=n, emissive=e, ior=1.5)
    elif kind == "metal":
        m = s.add_material(type=MAT_STANDARD_PBR, normal=n, emissive=e, color=white, metalness=s.solid1(1.0), roughness=s.solid1(0.5), ior=1.5)
    elif kind == "dielectric":
        m = s.add_material(type=MAT_STANDARD_PBR, normal=n, emissive=e, color=white, metalness=s.solid1(0.0), roughness=s.solid1(0.5), ior=1.5)
    else:
        raise ValueError(kind)
    p, i = icosphere(order, False)
    s.add_instance([(s.add_mesh(p, i), m, False)])
    s.background = np.ones((1, 1, 4), np.float32)
    s.constants = Constants(samples_per_run=spr, max_bounces=1024, env_samples_per_bounce=env_samples, mesh_samples_per_bounce=0)
    s.build_emitter_table = False
    return s


def furnace_emissive_interior(mesh_samples: int = 0, spr: int = 1024, order: int = 5, extent=(32, 32)) -> Scene:
    """engine/tests.zig:366-455 — "inside illuminating sphere is white" (mesh_samples=1 revives :457-487)."""
    s = Scene("furnace-emissive-interior", extent=extent, lens=Lens((0, 0, 0), (1, 0, 0), (0, 0, 1), math.pi / 3.0, 0.0, 1.0))
    n, c, e = s.solid2(0.5, 0.5), s.solid3(0.5, 0.5, 0.5), s.solid3(0.5, 0.5, 0.5)
    m = s.add_material(type=MAT_LAMBERT, normal=n, emissive=e, color=c)
    p, i = icosphere(order, True)
    mesh = s.add_mesh(p, i)
    s.add_instance([(mesh, m, mesh_samples > 0)])
    s.background = np.array([[[0, 0, 0, 1]]], np.float32)
    s.constants = Constants(samples_per_run=spr, max_bounces=1024, env_samples_per_bounce=0, mesh_samples_per_bounce=mesh_samples)
    s.build_emitter_table = mesh_samples > 0
    return s


def cornell_box(extent=(512, 512)) -> Scene:
    """cfg2: Cornell box in the unit cube, 36 triangles: 5 walls (10) + 2 boxes (24) + quad light (2)."""
    s = Scene("cfg2-cornell-box", extent=extent)
    n, black = s.solid2(0.5, 0.5), s.solid3(0, 0, 0)
    white = s.add_material(type=MAT_LAMBERT, normal=n, emissive=black, color=s.solid3(0.73, 0.73, 0.73))
    red = s.add_material(type=MAT_LAMBERT, normal=n, emissive=black, color=s.solid3(0.65, 0.05, 0.05))
    green = s.add_material(type=MAT_LAMBERT, normal=n, emissive=black, color=s.solid3(0.12, 0.45, 0.15))
    light = s.add_material(type=MAT_LAMBERT, normal=n, emissive=s.solid3(17, 12, 4), color=s.solid3(0.78, 0.78, 0.78))
    # room: x in [0,1] (depth, camera looks +x), y in [0,1] (left/right), z in [0,1] (up); normals face inward
    floor = _quad((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0))
    ceil_ = _quad((0, 0, 1), (0, 1, 1), (1, 1, 1), (1, 0, 1))
    back = _quad((1, 0, 0), (1, 0, 1), (1, 1, 1), (1, 1, 0))
    left = _quad((0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1))
    right = _quad((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0))
    m_white = s.add_mesh(*_merge([floor, ceil_, back]))
    m_left = s.add_mesh(*left)
    m_right = s.add_mesh(*right)
    tall = _box((0.55, 0.52, 0.0), (0.85, 0.82, 0.6), rot_z=math.radians(-17))
    short = _box((0.25, 0.15, 0.0), (0.55, 0.45, 0.3), rot_z=math.radians(18))
    m_boxes = s.add_mesh(*_merge([tall, short]))
    # emitter faces down (front face = side the normal points to; lights emit from the front face only)
    lq = _quad((0.37, 0.35, 0.998), (0.63, 0.35, 0.998), (0.63, 0.65, 0.998), (0.37, 0.65, 0.998))
    lq = (lq[0], lq[1][:, ::-1].copy())
    m_light = s.add_mesh(*lq)
    s.add_instance([(m_white, white, False), (m_left, red, False), (m_right, green, False), (m_boxes, white, False)])
    s.add_instance([(m_light, light, True)])
    s.background = np.array([[[0, 0, 0, 1]]], np.float32)
    s.lens = Lens((-1.4, 0.5, 0.5), (1, 0, 0), (0, 0, 1), math.radians(39.3), 0.0, 1.0)
    s.constants = Constants()
    return s


def uv_sphere(nu: int, nv: int, radius=1.0, center=(0, 0, 0)):
    """Lat-long sphere with per-vertex normals and texcoords (indexed attributes)."""
    u = np.linspace(0.0, 1.0, nu + 1, dtype=np.float32)
    v = np.linspace(0.0, 1.0, nv + 1, dtype=np.float32)
    uu, vv = np.meshgrid(u, v)
    phi, theta = uu * np.float32(2 * math.pi), vv * np.float32(math.pi)
    n = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], -1).reshape(-1, 3).astype(np.float32)
    p = (n * np.float32(radius) + np.array(center, np.float32)).astype(np.float32)
    t = np.stack([uu, vv], -1).reshape(-1, 2).astype(np.float32)
    idx = []
    for j in range(nv):
        for i in range(nu):
            a, b = j * (nu + 1) + i, j * (nu + 1) + i + 1
            c, d = (j + 1) * (nu + 1) + i, (j + 1) * (nu + 1) + i + 1
            if j != 0:
                idx.append((a, c, b))
            if j != nv - 1:
                idx.append((b, c, d))
    return p, np.array(idx, np.uint32), n, t


def procedural_sky(width: int, height: int, sun_dir=(0.3, 0.4, 0.8), sun_power=200.0, seed=0x4D4F4F4E, sharpness=400.0) -> np.ndarray:
    """Equirectangular RGBA32F sky: gradient + sun lobe; row = theta in [0,pi] from +Z, column = phi."""
    phi = (np.arange(width, dtype=np.float64) + 0.5) / width * 2 * math.pi
    theta = (np.arange(height, dtype=np.float64) + 0.5) / height * math.pi
    pp, tt = np.meshgrid(phi, theta)
    d = np.stack([np.sin(tt) * np.cos(pp), np.sin(tt) * np.sin(pp), np.cos(tt)], -1)
    s = np.array(sun_dir, np.float64)
    s /= np.linalg.norm(s)
    cosang = np.clip((d * s).sum(-1), -1, 1)
    up = np.clip(d[..., 2], 0, 1)
    sky = np.stack([0.25 + 0.35 * (1 - up), 0.35 + 0.35 * (1 - up), 0.6 + 0.3 * up], -1)
    ground = np.array([0.08, 0.07, 0.06])
    col = np.where(d[..., 2:3] >= 0, sky, ground)
    col = col + sun_power * np.exp((cosang[..., None] - 1.0) * sharpness) * np.array([1.0, 0.9, 0.75])
    out = np.ones((height, width, 4), np.float32)
    out[..., :3] = col.astype(np.float32)
    return out


def _checker_srgb(n: int, a=(200, 60, 50), b=(230, 230, 220), cells=8) -> np.ndarray:
    y, x = np.mgrid[0:n, 0:n]
    m = (((x * cells) // n + (y * cells) // n) % 2).astype(bool)
    img = np.zeros((n, n, 4), np.uint8)
    img[m] = (*a, 255)
    img[~m] = (*b, 255)
    return img


def material_zoo(extent=(96, 64)) -> Scene:
    """Parity scene touching every branch of the hot path: StandardPBR with textures of every format, Lambert,
    Glass, PerfectMirror, a sampled emitter, an env map, vertex normals + texcoords, a transformed instance
    sharing a BLAS with an untransformed one, and a thin-lens aperture."""
    s = Scene("material-zoo", extent=extent)
    rs = np.random.RandomState(0x4D4F4F4E & 0x7FFFFFFF)
    n_flat, black = s.solid2(0.5, 0.5), s.solid3(0, 0, 0)
    # textures
    checker = s.raw_texture(_checker_srgb(64), TEX_RGBA8_SRGB)
    rough = s.raw_texture((np.clip(rs.rand(32, 32) * 0.6 + 0.2, 0, 1) * 255).astype(np.uint8), TEX_R8_UNORM)
    metal = s.raw_texture(((np.indices((16, 16)).sum(0) % 2) * 255).astype(np.uint8), TEX_R8_UNORM)
    nm = np.zeros((32, 32, 2), np.uint8)
    yy, xx = np.mgrid[0:32, 0:32]
    nm[..., 0] = (127.5 + 40 * np.sin(xx * 0.8)).astype(np.uint8)
    nm[..., 1] = (127.5 + 40 * np.cos(yy * 0.6)).astype(np.uint8)
    normal_map = s.raw_texture(nm, TEX_RG8_UNORM)
    half = s.raw_texture(np.stack([rs.rand(8, 8), rs.rand(8, 8), rs.rand(8, 8), np.ones((8, 8))], -1).astype(np.float16), TEX_RGBA16F)
    # materials
    floor_m = s.add_material(type=MAT_STANDARD_PBR, normal=normal_map, emissive=black, color=checker, metalness=metal, roughness=rough, ior=1.5)
    pbr_m = s.add_material(type=MAT_STANDARD_PBR, normal=n_flat, emissive=black, color=half, metalness=s.solid1(0.3), roughness=s.solid1(0.35), ior=1.45)
    lambert_m = s.add_material(type=MAT_LAMBERT, normal=n_flat, emissive=black, color=s.solid3(0.2, 0.5, 0.8))
    glass_m = s.add_material(type=MAT_GLASS, normal=n_flat, emissive=black, ior=1.5)
    mirror_m = s.add_material(type=MAT_PERFECT_MIRROR, normal=n_flat, emissive=black)
    light_m = s.add_material(type=MAT_LAMBERT, normal=n_flat, emissive=s.solid3(12, 10, 8), color=s.solid3(0.5, 0.5, 0.5))
    # meshes
    fp = np.array([(-4, -4, 0), (4, -4, 0), (4, 4, 0), (-4, 4, 0)], np.float32)
    floor = s.add_mesh(fp, [(0, 1, 2), (0, 2, 3)], normals=np.tile([[0, 0, 1]], (4, 1)), texcoords=[(0, 0), (3, 0), (3, 3), (0, 3)])
    sp, si, sn, st = uv_sphere(24, 12, 0.5)
    sphere = s.add_mesh(sp, si, normals=sn, texcoords=st)
    ip, ii = icosphere(2)
    ico = s.add_mesh(ip * np.float32(0.45), ii)
    bx = s.add_mesh(*_box((-0.4, -0.4, 0.0), (0.4, 0.4, 0.9), rot_z=0.5))
    lq = _quad((-0.6, -0.6, 2.6), (0.6, -0.6, 2.6), (0.6, 0.6, 2.6), (-0.6, 0.6, 2.6))
    lightq = s.add_mesh(lq[0], lq[1][:, ::-1].copy())
    # instances
    s.add_instance([(floor, floor_m, False)])
    s.add_instance([(sphere, pbr_m, False)], transform=[[1, 0, 0, -1.2], [0, 1, 0, 0.2], [0, 0, 1, 0.5]])
    c, sn_ = math.cos(0.7), math.sin(0.7)
    s.add_instance([(sphere, lambert_m, False)], transform=[[1.4 * c, -0.8 * sn_, 0, 1.3], [1.4 * sn_, 0.8 * c, 0, -0.9], [0, 0, 1.1, 0.55]])
    s.add_instance([(ico, glass_m, False)], transform=[[1, 0, 0, 0.1], [0, 1, 0, -0.3], [0, 0, 1, 0.46]])
    s.add_instance([(bx, mirror_m, False)], transform=[[1, 0, 0, 0.2], [0, 1, 0, 1.6], [0, 0, 1, 0.0]])
    s.add_instance([(lightq, light_m, True)])
    s.add_instance([(ico, light_m, True)], transform=[[0.3, 0, 0, -0.4], [0, 0.3, 0, -1.5], [0, 0, 0.3, 0.2]])
    s.background = procedural_sky(32, 16, sun_power=60.0)
    s.lens = Lens((-4.2, -1.0, 1.6), tuple(np.array([0.95, 0.2, -0.25]) / np.linalg.norm([0.95, 0.2, -0.25])), (0, 0, 1), math.radians(42), 0.05, 4.0)
    s.constants = Constants()
    return s


# ---------------------------------------------------------------------------------------------
# cfg3/cfg4: procedural "Salle-de-bain-class" bathroom: ~1.2 M triangles in primitives of
# <= 65 535 vertices, ~40 materials (textured GGX, glass, mirror), window to a sun+sky equirect.
# ---------------------------------------------------------------------------------------------
def param_surface(fn, nu: int, nv: int, wrap_u=False, flip=False):
    """Tessellate p = fn(u, v), u,v in [0,1], into (nu x nv) quads with smooth normals and texcoords."""
    assert (nu + 1) * (nv + 1) <= 65535
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    uu, vv = np.meshgrid(u, v)
    p = fn(uu, vv).astype(np.float64)  # (nv+1, nu+1, 3)
    e = 1e-4
    du = (fn(np.clip(uu + e, 0, 1), vv) - fn(np.clip(uu - e, 0, 1), vv)).astype(np.float64)
    dv = (fn(uu, np.clip(vv + e, 0, 1)) - fn(uu, np.clip(vv - e, 0, 1))).astype(np.float64)
    n = np.cross(du, dv)
    ln = np.linalg.norm(n, axis=-1, keepdims=True)
    n = np.where(ln > 1e-20, n / np.maximum(ln, 1e-20), np.array([0.0, 0.0, 1.0]))
    if flip:
        n = -n
    j, i = np.mgrid[0:nv, 0:nu]
    a = (j * (nu + 1) + i).reshape(-1)
    b, c, d = a + 1, a + (nu + 1), a + (nu + 2)
    tri = np.concatenate([np.stack([a, b, d], 1), np.stack([a, d, c], 1)])
    if flip:
        tri = tri[:, ::-1]
    t = np.stack([uu, vv], -1).reshape(-1, 2)
    return p.reshape(-1, 3).astype(np.float32), tri.astype(np.uint32), n.reshape(-1, 3).astype(np.float32), t.astype(np.float32)


def _superellipsoid(center, radii, e1=0.4, e2=0.4, v0=0.0, v1=1.0):
    cx, cy, cz = center
    rx, ry, rz = radii

    def spow(x, p):
        return np.sign(x) * np.abs(x) ** p

    def fn(u, v):
        phi = u * 2 * math.pi
        th = (v0 + (v1 - v0) * v) * math.pi - math.pi / 2
        return np.stack([cx + rx * spow(np.cos(th), e1) * spow(np.cos(phi), e2),
                         cy + ry * spow(np.cos(th), e1) * spow(np.sin(phi), e2),
                         cz + rz * spow(np.sin(th), e1)], -1)
    return fn


def _torus(center, R, r, axis=2):
    def fn(u, v):
        a, b = u * 2 * math.pi, v * 2 * math.pi
        x, y, z = (R + r * np.cos(b)) * np.cos(a), (R + r * np.cos(b)) * np.sin(a), r * np.sin(b)
        q = [x, y, z]
        if axis == 0:
            q = [z, x, y]
        elif axis == 1:
            q = [y, z, x]
        return np.stack([center[0] + q[0], center[1] + q[1], center[2] + q[2]], -1)
    return fn


def _lathe(center, profile_r, height):
    """Surface of revolution around +Z; profile_r(v) -> radius at height v*height."""
    def fn(u, v):
        a = u * 2 * math.pi
        r = profile_r(v)
        return np.stack([center[0] + r * np.cos(a), center[1] + r * np.sin(a), center[2] + v * height], -1)
    return fn


def _cloth(corner, ex, ey, normal, amp, fx, fy, phase=0.0):
    corner, ex, ey, normal = (np.array(x, np.float64) for x in (corner, ex, ey, normal))

    def fn(u, v):
        w = amp * (np.sin(u * fx + phase) * np.cos(v * fy) + 0.35 * np.sin(u * fx * 2.7 + v * 3.1))
        return corner + u[..., None] * ex + v[..., None] * ey + w[..., None] * normal
    return fn


def _tile_textures(n=1024, tiles=8, base=(225, 228, 230), grout=(120, 118, 115), seed=1):
    """sRGB colour + RG8 tangent-space normal for a glazed tile grid with bevelled grout lines."""
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:n, 0:n]
    cell = n // tiles
    fx, fy = (x % cell) / cell, (y % cell) / cell
    edge = np.minimum(np.minimum(fx, 1 - fx), np.minimum(fy, 1 - fy))
    g = np.clip(edge / 0.035, 0, 1)
    tint = rs.rand(tiles, tiles, 3) * 18 - 9
    tile_col = np.array(base)[None, None] + tint[(y // cell) % tiles, (x // cell) % tiles]
    col = np.array(grout)[None, None] * (1 - g[..., None]) + tile_col * g[..., None]
    rgba = np.zeros((n, n, 4), np.uint8)
    rgba[..., :3] = np.clip(col, 0, 255).astype(np.uint8)
    rgba[..., 3] = 255
    h = g + 0.004 * rs.rand(n, n)
    gx, gy = np.gradient(h, axis=1), np.gradient(h, axis=0)
    nrm = np.zeros((n, n, 2), np.uint8)
    # tangent-space xy kept inside the unit disk (|x|, |y| <= 0.56): a texel on its rim decodes to z = 0, a normal lying IN the
    # surface, for which Frame::reorthogonalize (reflection_frame.hlsl:31-35) divides 0 by 0 whenever it lines up with the tangent
    nrm[..., 0] = np.clip(127.5 - gx * 900, 56, 199).astype(np.uint8)
    nrm[..., 1] = np.clip(127.5 - gy * 900, 56, 199).astype(np.uint8)
    rough = np.clip(60 + (1 - g) * 150 + rs.rand(n, n) * 12, 0, 255).astype(np.uint8)
    return rgba, nrm, rough


def bathroom(extent=(1920, 1080), detail: float = 1.0, seed: int = 0x4D4F4F4E) -> Scene:
    """cfg3 (and cfg4). `detail` scales tessellation: 1.0 ~ 1.2 M triangles; tests use ~0.05."""
    s = Scene("cfg3-salle-de-bain-class", extent=extent)
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    flat, black = s.solid2(0.5, 0.5), s.solid3(0, 0, 0)

    def res(nu, nv):
        k = math.sqrt(detail)
        a, b = max(4, int(nu * k)), max(3, int(nv * k))
        while (a + 1) * (b + 1) > 65535:
            a, b = int(a * 0.97), int(b * 0.97)
        return a, b

    def pbr(color, metal, rough, ior=1.5, normal=flat, emissive=black):
        c = color if isinstance(color, int) else s.solid3(*color)
        m = metal if isinstance(metal, int) and not isinstance(metal, bool) and metal > 1 else s.solid1(float(metal))
        r = rough if isinstance(rough, int) and rough > 1 else s.solid1(float(rough))
        return s.add_material(type=MAT_STANDARD_PBR, normal=normal, emissive=emissive, color=c, metalness=m, roughness=r, ior=ior)

    # ---- textures -----------------------------------------------------------------------------
    tex_n = 1024 if detail >= 0.5 else 128
    wall_rgba, wall_nrm, wall_rough = _tile_textures(tex_n, 8, (228, 232, 233), (128, 126, 122), 1)
    floor_rgba, floor_nrm, floor_rough = _tile_textures(tex_n, 4, (70, 82, 92), (35, 35, 38), 2)
    t_wall_c, t_wall_n, t_wall_r = s.raw_texture(wall_rgba, TEX_RGBA8_SRGB), s.raw_texture(wall_nrm, TEX_RG8_UNORM), s.raw_texture(wall_rough, TEX_R8_UNORM)
    t_floor_c, t_floor_n, t_floor_r = s.raw_texture(floor_rgba, TEX_RGBA8_SRGB), s.raw_texture(floor_nrm, TEX_RG8_UNORM), s.raw_texture(floor_rough, TEX_R8_UNORM)
    zero1 = s.solid1(0.0)
    m_wall = s.add_material(type=MAT_STANDARD_PBR, normal=t_wall_n, emissive=black, color=t_wall_c, metalness=zero1, roughness=t_wall_r, ior=1.5)
    m_floor = s.add_material(type=MAT_STANDARD_PBR, normal=t_floor_n, emissive=black, color=t_floor_c, metalness=zero1, roughness=t_floor_r, ior=1.5)
    m_ceiling = s.add_material(type=MAT_LAMBERT, normal=flat, emissive=black, color=s.solid3(0.85, 0.85, 0.82))
    m_ceramic = pbr((0.92, 0.92, 0.9), 0.0, 0.12)
    m_chrome = pbr((0.95, 0.95, 0.97), 1.0, 0.05)
    m_brushed = pbr((0.8, 0.8, 0.82), 1.0, 0.32)
    m_brass = pbr((0.9, 0.72, 0.35), 1.0, 0.2)
    m_glass = s.add_material(type=MAT_GLASS, normal=flat, emissive=black, ior=1.5)
    m_mirror = s.add_material(type=MAT_PERFECT_MIRROR, normal=flat, emissive=black)
    m_wood = pbr((0.42, 0.27, 0.15), 0.0, 0.55)
    m_rubber = pbr((0.05, 0.05, 0.05), 0.0, 0.8)
    m_lamp = s.add_material(type=MAT_LAMBERT, normal=flat, emissive=s.solid3(28, 25, 20), color=s.solid3(0.6, 0.6, 0.6))
    towel_cols = [(0.75, 0.2, 0.2), (0.2, 0.45, 0.7), (0.9, 0.85, 0.7), (0.3, 0.6, 0.4)]
    m_towels = [pbr(c, 0.0, 0.9) for c in towel_cols]
    bottle_mats = [pbr(tuple(rs.rand(3) * 0.8 + 0.1), 0.0, float(rs.rand() * 0.5 + 0.08), ior=1.4 + 0.2 * rs.rand()) for _ in range(14)]
    soap_mats = [pbr(tuple(rs.rand(3) * 0.5 + 0.45), 0.0, 0.6) for _ in range(4)]

    # ---- room: x in [0,4] (camera looks +x), y in [0,3], z in [0,2.6]; window in the y=3 wall -----
    X, Y, Z = 4.0, 3.0, 2.6

    def wall(p0, ex, ey, nu, nv, uvscale):
        p0, ex, ey = np.array(p0, np.float64), np.array(ex, np.float64), np.array(ey, np.float64)
        p, i, n, t = param_surface(lambda u, v: p0 + u[..., None] * ex + v[..., None] * ey, nu, nv)
        return p, i, n, (t * np.array(uvscale, np.float32)).astype(np.float32)

    geoms_room = []
    g = wall((0, 0, 0), (X, 0, 0), (0, Y, 0), *res(64, 48), (X / 1.2, Y / 1.2))
    geoms_room.append((s.add_mesh(g[0], g[1], g[2], g[3]), m_floor, False))
    g = wall((0, 0, Z), (0, Y, 0), (X, 0, 0), *res(16, 16), (1, 1))
    geoms_room.append((s.add_mesh(g[0], g[1], g[2], g[3]), m_ceiling, False))
    walls = [((X, 0, 0), (0, Y, 0), (0, 0, Z), (Y / 1.6, Z / 1.6)),      # far wall x = X
             ((0, Y, 0), (0, -Y, 0), (0, 0, Z), (Y / 1.6, Z / 1.6)),     # behind camera x = 0
             ((0, 0, 0), (X, 0, 0), (0, 0, Z), (X / 1.6, Z / 1.6))]      # y = 0
    for p0, ex, ey, uvs in walls:
        g = wall(p0, ex, ey, *res(64, 48), uvs)
        geoms_room.append((s.add_mesh(g[0], g[1], g[2], g[3]), m_wall, False))
    # y = Y wall with a window hole x in [1.2, 2.8], z in [0.9, 2.1]: four panels
    for (x0, x1, z0, z1) in [(0, 1.2, 0, Z), (2.8, X, 0, Z), (1.2, 2.8, 0, 0.9), (1.2, 2.8, 2.1, Z)]:
        g = wall((x1, Y, z0), (x0 - x1, 0, 0), (0, 0, z1 - z0), *res(24, 24), ((x1 - x0) / 1.6, (z1 - z0) / 1.6))
        geoms_room.append((s.add_mesh(g[0], g[1], g[2], g[3]), m_wall, False))
    s.add_instance(geoms_room)

    def add(fn, nu, nv, mat, sampled=False, flip=False, transform=None):
        p, i, n, t = param_surface(fn, *res(nu, nv), flip=flip)
        mesh = s.add_mesh(p, i, n, t)
        s.add_instance([(mesh, mat, sampled)], transform=transform)
        return mesh

    # bathtub (outer shell + inner basin), sink, toilet
    add(_superellipsoid((2.9, 0.75, 0.32), (0.95, 0.48, 0.32), 0.35, 0.45, 0.0, 0.62), 300, 150, m_ceramic)
    add(_superellipsoid((2.9, 0.75, 0.36), (0.86, 0.40, 0.27), 0.4, 0.5, 0.0, 0.58), 300, 150, m_ceramic, flip=True)
    add(_superellipsoid((1.1, 2.55, 0.88), (0.32, 0.26, 0.16), 0.7, 0.8, 0.0, 0.5), 220, 110, m_ceramic)
    add(_superellipsoid((1.1, 2.55, 0.9), (0.27, 0.21, 0.13), 0.8, 0.9, 0.0, 0.5), 220, 110, m_ceramic, flip=True)
    add(_lathe((1.1, 2.62, 0.0), lambda v: 0.09 + 0.05 * (1 - v) ** 2, 0.74), 160, 80, m_ceramic)
    add(_superellipsoid((0.55, 0.6, 0.28), (0.26, 0.22, 0.28), 0.6, 0.8), 200, 100, m_ceramic)
    add(_superellipsoid((0.32, 0.6, 0.62), (0.1, 0.24, 0.28), 0.5, 0.5), 160, 80, m_ceramic)
    add(_torus((0.6, 0.6, 0.46), 0.2, 0.035), 200, 60, m_wood)
    # chrome: faucets, towel rails, shower head + hose
    add(_torus((1.1, 2.78, 1.08), 0.1, 0.017, axis=0), 220, 48, m_chrome)
    add(_lathe((1.1, 2.8, 0.92), lambda v: 0.022 + 0.0 * v, 0.16), 96, 24, m_chrome)
    add(_torus((3.7, 0.75, 0.78), 0.12, 0.02, axis=1), 220, 48, m_chrome)
    add(_lathe((3.86, 0.75, 0.6), lambda v: 0.025 + 0.0 * v, 0.3), 96, 24, m_chrome)
    for k, zz in enumerate((1.25, 1.0)):
        add(lambda u, v, zz=zz: np.stack([0.03 + 0.012 * np.cos(v * 2 * math.pi), 1.3 + u * 0.9, zz + 0.012 * np.sin(v * 2 * math.pi)], -1), 200, 32, m_brushed)
    add(_superellipsoid((3.55, 0.75, 2.05), (0.13, 0.13, 0.025), 0.6, 1.0), 200, 60, m_chrome)
    add(lambda u, v: np.stack([3.9 - 0.3 * np.sin(u * 1.4) + 0.011 * np.cos(v * 2 * math.pi),
                               0.75 + 0.011 * np.sin(v * 2 * math.pi) * np.ones_like(u),
                               0.95 + u * 1.1 + 0.0 * v], -1), 240, 24, m_brass)
    # shower screen: two glass panes (thin boxes) + mirror on the far wall
    for (x0, x1) in [(1.86, 1.88), (1.92, 1.94)]:
        bp, bi = _box((x0, 0.05 if x0 < 1.9 else 0.75, 0.05), (x1, 0.8 if x0 < 1.9 else 1.45, 2.0))
        s.add_instance([(s.add_mesh(bp, bi), m_glass, False)])
    mq = _quad((X - 0.01, 1.7, 1.0), (X - 0.01, 2.7, 1.0), (X - 0.01, 2.7, 1.9), (X - 0.01, 1.7, 1.9))
    s.add_instance([(s.add_mesh(mq[0], mq[1][:, ::-1].copy()), m_mirror, False)])
    add(lambda u, v: np.stack([X - 0.012 - 0.0 * u, 1.66 + 1.08 * u, 0.96 + 0.98 * v], -1) +
        np.stack([0.02 * (np.minimum(np.minimum(u, 1 - u), np.minimum(v, 1 - v)) < 0.04), 0 * u, 0 * u], -1), 64, 64, m_wood, flip=True)
    # towels (cloth), bath mat
    for k in range(4):
        yy = 1.32 + 0.22 * k
        add(_cloth((0.06, yy, 1.24 if k % 2 == 0 else 0.99), (0.0, 0.2, 0.0), (0.0, 0.0, -0.55), (1, 0, 0), 0.012, 38.0, 9.0, phase=k), 260, 220, m_towels[k])
    add(_cloth((1.9, 1.55, 0.012), (0.9, 0, 0), (0, 0.6, 0), (0, 0, 1), 0.006, 90.0, 60.0), 250, 250, m_rubber)
    # shelf with bottles, soaps
    sp, si = _box((2.2, 2.8, 1.15), (3.6, 2.98, 1.18))
    s.add_instance([(s.add_mesh(sp, si), m_wood, False)])
    for k in range(14):
        bx_, h = 2.28 + 0.095 * k, 0.14 + 0.12 * rs.rand()
        r0, neck = 0.028 + 0.012 * rs.rand(), 0.3 + 0.3 * rs.rand()
        prof = (lambda r0, neck: (lambda v: r0 * np.where(v < 0.7, 1.0 - 0.08 * np.sin(v * 9), np.maximum(neck, 1.0 - (v - 0.7) * 2.6))))(r0, neck)
        add(_lathe((bx_, 2.89, 1.18), prof, h), 128, 64, bottle_mats[k])
    for k in range(4):
        add(_superellipsoid((3.3 + 0.12 * k, 0.32, 0.66), (0.045, 0.03, 0.016), 0.5, 0.6), 120, 60, soap_mats[k])
    # ceiling lamp: sampled emitter (material name prefix "Emitter" in the reference's glTF rule, World.zig:270)
    add(_superellipsoid((2.0, 1.5, Z - 0.09), (0.22, 0.22, 0.07), 0.9, 1.0), 96, 48, m_lamp, sampled=True)

    sky_w = 2048 if detail >= 0.5 else 256
    s.background = procedural_sky(sky_w, sky_w // 2, sun_dir=(0.15, 0.75, 0.64), sun_power=5.0e4)
    fw = np.array([0.93, 0.28, -0.22])
    s.lens = Lens((0.35, 0.55, 1.55), tuple(fw / np.linalg.norm(fw)), (0, 0, 1), math.radians(58), 0.0, 1.0)
    s.constants = Constants()
    return s


# ---------------------------------------------------------------------------------------------
# cfg5: instanced mesh-light stress scene: n_side^2 emissive 20-triangle icosahedra sharing
# one BLAS (all `sampled`, per-instance jittered transform, seed 7) + ground plane + 64 occluders, black env.
# n_side = 1000 is the BASELINE config (10^6 instances, 2*10^7 alias-table entries = 400 MB).
# ---------------------------------------------------------------------------------------------
def mesh_light_stress(n_side: int = 1000, extent=(2048, 2048), seed: int = 7) -> Scene:
    s = Scene(f"cfg5-mesh-light-stress-{n_side}x{n_side}", extent=extent)
    rs = np.random.RandomState(seed)
    flat, black = s.solid2(0.5, 0.5), s.solid3(0, 0, 0)
    m_ground = s.add_material(type=MAT_LAMBERT, normal=flat, emissive=black, color=s.solid3(0.6, 0.6, 0.6))
    m_light = s.add_material(type=MAT_LAMBERT, normal=flat, emissive=s.solid3(6.0, 5.0, 4.0), color=s.solid3(0.3, 0.3, 0.3))
    m_occ = [s.add_material(type=MAT_LAMBERT, normal=flat, emissive=black, color=s.solid3(0.7, 0.35, 0.25)),
             s.add_material(type=MAT_STANDARD_PBR, normal=flat, emissive=black, color=s.solid3(0.85, 0.85, 0.9), metalness=s.solid1(1.0), roughness=s.solid1(0.25), ior=1.5)]
    size = float(n_side)
    gp, gi = _quad((0, 0, 0), (size, 0, 0), (size, size, 0), (0, size, 0))
    s.add_instance([(s.add_mesh(gp, gi), m_ground, False)])
    ip, ii = icosphere(0)
    ico = s.add_mesh(ip, ii)
    # per-instance jittered transform: uniform scale 0.12..0.22, rotation about z, position jitter, height 0.5..1.5
    ang = rs.rand(n_side * n_side) * 2 * math.pi
    sc_ = 0.12 + 0.10 * rs.rand(n_side * n_side)
    jx, jy = rs.rand(n_side * n_side) * 0.5 - 0.25, rs.rand(n_side * n_side) * 0.5 - 0.25
    hz = 0.5 + rs.rand(n_side * n_side)
    gx, gy = np.meshgrid(np.arange(n_side), np.arange(n_side))
    cx, cy = gx.reshape(-1) + 0.5 + jx, gy.reshape(-1) + 0.5 + jy
    c, sn = np.cos(ang) * sc_, np.sin(ang) * sc_
    T = np.zeros((n_side * n_side, 3, 4), np.float32)
    T[:, 0, 0], T[:, 0, 1], T[:, 0, 3] = c, -sn, cx
    T[:, 1, 0], T[:, 1, 1], T[:, 1, 3] = sn, c, cy
    T[:, 2, 2], T[:, 2, 3] = sc_, hz
    geom = [(ico, m_light, True)]
    for k in range(n_side * n_side):
        s.instances.append(Instance(T[k], geom, True))
    # 64 occluders: boxes scattered over the field
    for k in range(64):
        x0, y0 = rs.rand(2) * (size - 2.0) + 0.5
        w, d, h = 0.3 + rs.rand() * min(2.0, size / 8), 0.3 + rs.rand() * min(2.0, size / 8), 0.8 + rs.rand() * 1.5
        bp, bi = _box((x0, y0, 0.0), (x0 + w, y0 + d, h), rot_z=rs.rand() * 1.5)
        s.add_instance([(s.add_mesh(bp, bi), m_occ[k % 2], False)])
    s.background = np.array([[[0, 0, 0, 1]]], np.float32)
    ctr = size / 2
    view = min(size, 40.0)
    s.lens = Lens((ctr - 0.6 * view, ctr - 0.6 * view, 0.45 * view), tuple(np.array([0.6, 0.6, -0.52]) / np.linalg.norm([0.6, 0.6, -0.52])), (0, 0, 1), math.radians(55), 0.0, 1.0)
    s.constants = Constants()
    return s
