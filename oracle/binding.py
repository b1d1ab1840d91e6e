"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs
(`cpu_baseline`, `--impl reference`). The product package (moonshine_b200/) never imports it.
Implements the backend protocol of moonshine_b200.scenes.load_scene.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_DIR, "liboracle.so")
    srcs = [os.path.join(_DIR, f) for f in ("oracle.cpp", "vec.h", "utils.h", "scene.h", "bvh.h", "shading.h", "host.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()  # no-op when liboracle.so is newer than its sources
        L = C.CDLL(so)
        L.OracleCreate.restype = C.c_void_p
        L.OracleDestroy.argtypes = [C.c_void_p]
        L.OracleCreateMesh.restype = C.c_uint32
        L.OracleCreateMesh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
        L.OracleCreateTexture.restype = C.c_uint32
        L.OracleCreateTexture.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.OracleCreateMaterial.restype = C.c_uint32
        L.OracleCreateMaterial.argtypes = [C.c_void_p] + [C.c_uint32] * 6 + [C.c_float]
        L.OracleSetMaterial.argtypes = [C.c_void_p] + [C.c_uint32] * 7 + [C.c_float]
        L.OracleCreateInstance.restype = C.c_uint32
        L.OracleCreateInstance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.OracleSetInstanceTransform.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.OracleSetInstanceVisibility.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.OracleSetBackground.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.OracleBuildEmitterTable.argtypes = [C.c_void_p]
        L.OracleSetConstants.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [C.c_int] * 3
        L.OracleSetBruteForce.argtypes = [C.c_void_p, C.c_int]
        L.OracleSetCrossCheck.argtypes = [C.c_void_p, C.c_int]
        L.OracleCrossCheckMismatches.restype = C.c_uint64
        L.OracleCrossCheckMismatches.argtypes = [C.c_void_p]
        L.OracleRender.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p]
        L.OracleRenderDiag.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
        L.OracleRenderPixels.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.OracleDebugPath.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.OracleEnvProbe.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        L.OracleTextureProbe.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        L.OracleBsdfBatch.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p]
        L.OracleMathProbe.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.OracleHashPcg.restype = C.c_uint32
        L.OracleHashPcg.argtypes = [C.c_uint32]
        L.OracleHashLcg.restype = C.c_uint32
        L.OracleHashLcg.argtypes = [C.c_uint32]
        L.OracleRngSeed.restype = C.c_uint32
        L.OracleRngSeed.argtypes = [C.c_uint32] * 3
        L.OracleRngFloats.argtypes = [C.c_uint32] * 4 + [C.c_void_p]
        L.OraclePrimaryRays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.OracleTrace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.OraclePick.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
        L.OracleEnvSize.restype = C.c_uint32
        L.OracleEnvSize.argtypes = [C.c_void_p]
        L.OracleEnvData.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.OracleAliasCount.restype = C.c_uint32
        L.OracleAliasCount.argtypes = [C.c_void_p]
        L.OracleAliasData.argtypes = [C.c_void_p, C.c_void_p]
        L.OracleWorldToInstance.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.OracleSquareToEqualAreaSphere.argtypes = [C.c_float, C.c_float, C.c_void_p]
        L.OracleSquareToEqualAreaSphereInverse.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.OracleOffsetAlongNormal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.OracleBsdf.argtypes = [C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


ALIAS_DTYPE = np.dtype([("alias", "<u4"), ("select", "<f4"), ("instance", "<u4"), ("geometry", "<u4"), ("primitive", "<u4")])
HIT_DTYPE = np.dtype([("instance", "<u4"), ("geometry", "<u4"), ("primitive", "<u4"), ("u", "<f4"), ("v", "<f4")])


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU oracle scene + renderer (backend protocol of scenes.load_scene)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.OracleCreate()
        self.constants = None

    def close(self):
        if self.h:
            self.L.OracleDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- backend protocol --------------------------------------------------------------------
    def create_texture(self, data, w, h, fmt):
        d = np.ascontiguousarray(data)
        return self.L.OracleCreateTexture(self.h, _p(d), w, h, fmt)

    def create_material(self, type_, normal, emissive, color, metalness, roughness, ior):
        return self.L.OracleCreateMaterial(self.h, type_, normal, emissive, color, metalness, roughness, ior)

    def create_mesh(self, positions, normals, texcoords, indices):
        p = np.ascontiguousarray(positions, np.float32)
        n = None if normals is None else np.ascontiguousarray(normals, np.float32)
        t = None if texcoords is None else np.ascontiguousarray(texcoords, np.float32)
        i = np.ascontiguousarray(indices, np.uint32)
        n_attr = len(n) if n is not None else (len(t) if t is not None else 0)
        return self.L.OracleCreateMesh(self.h, _p(p), _p(n), _p(t), len(p), n_attr, _p(i), len(i))

    def create_instance(self, transform, geometries, visible=True):
        g = np.zeros(len(geometries), dtype=np.dtype([("mesh", "<u4"), ("material", "<u4"), ("sampled", "u1")], align=True))
        for k, (m, mat, s) in enumerate(geometries):
            g[k] = (m, mat, 1 if s else 0)
        t = np.ascontiguousarray(transform, np.float32)
        return self.L.OracleCreateInstance(self.h, _p(t), _p(g), len(g), 1 if visible else 0)

    def set_instance_transform(self, handle, transform):
        t = np.ascontiguousarray(transform, np.float32)
        self.L.OracleSetInstanceTransform(self.h, handle, _p(t))

    def set_instance_visibility(self, handle, visible):
        self.L.OracleSetInstanceVisibility(self.h, handle, 1 if visible else 0)

    def set_background(self, rgba):
        a = np.ascontiguousarray(rgba, np.float32)
        self.L.OracleSetBackground(self.h, _p(a), a.shape[1], a.shape[0])

    def build_emitter_table(self):
        self.L.OracleBuildEmitterTable(self.h)

    def set_constants(self, c):
        self.constants = c
        self.L.OracleSetConstants(self.h, c.samples_per_run, c.max_bounces, c.env_samples_per_bounce, c.mesh_samples_per_bounce,
                                  int(c.flip_image), int(c.indexed_attributes), int(c.two_component_normal_texture))

    # -- rendering / probes ------------------------------------------------------------------
    def render(self, lens, width, height, n_launches=1, image=None, sample_count=0, threads=0, pixel_stride=1, brute=False):
        """Returns (image HxWx4 f32, [closest, shadow] ray counts)."""
        if image is None:
            image = np.zeros((height, width, 4), np.float32)
        counts = np.zeros(2, np.uint64)
        la = lens.as_array()
        self.L.OracleSetBruteForce(self.h, int(brute))
        self.L.OracleRender(self.h, _p(la), width, height, _p(image), sample_count, n_launches, threads, pixel_stride, _p(counts))
        return image, counts

    def render_diag(self, lens, width, height, n_launches=1, sample_count=0, threads=0, pixel_stride=1):
        """render() plus the per-pixel diagnostic: smallest GGX alpha the pixel's paths touched (1e30 = none)."""
        image = np.zeros((height, width, 4), np.float32)
        min_alpha = np.full((height, width), 1e30, np.float32)
        counts = np.zeros(2, np.uint64)
        la = lens.as_array()
        self.L.OracleSetBruteForce(self.h, 0)
        self.L.OracleRenderDiag(self.h, _p(la), width, height, _p(image), sample_count, n_launches, threads, pixel_stride, _p(counts), _p(min_alpha))
        return image, counts, min_alpha

    def render_pixels(self, lens, width, height, xy, sample_count=0, brute=False, threads=0):
        """One launch of the listed pixels (n x {x, y}) -> n x RGBA; brute = no BVH (every triangle tested)."""
        p = np.ascontiguousarray(xy, np.uint32).reshape(-1, 2)
        out = np.zeros((len(p), 4), np.float32)
        la = lens.as_array()
        self.L.OracleRenderPixels(self.h, _p(la), width, height, sample_count, len(p), _p(p), _p(out), int(brute), threads)
        return out

    def set_cross_check(self, on=True):
        """Diagnostics: every ray query is answered by the BVH walk and by testing all triangles; mismatches are counted."""
        self.L.OracleSetCrossCheck(self.h, int(on))

    def cross_check_mismatches(self):
        return int(self.L.OracleCrossCheckMismatches(self.h))

    def debug_path(self, lens, width, height, x, y, sample=0):
        """Diagnostics: prints every bounce of one path on stderr, returns its RGBA."""
        out = np.zeros(4, np.float32)
        la = lens.as_array()
        self.L.OracleDebugPath(self.h, _p(la), width, height, x, y, sample, _p(out))
        return out

    def env_probe(self, op, inputs3):
        """EnvMap::sample (op 0, rand in columns 0-1), ::eval (op 1, direction), ::incomingRadiance (op 2) -> (n, 7)."""
        a = np.ascontiguousarray(inputs3, np.float32).reshape(-1, 3)
        out = np.zeros((len(a), 7), np.float32)
        self.L.OracleEnvProbe(self.h, op, len(a), _p(a), _p(out))
        return out

    def texture_probe(self, tex, uv, mirror=False):
        """SampleLevel(texture `tex`, uv, 0) at n coordinates -> (n, 4)."""
        a = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((len(a), 4), np.float32)
        self.L.OracleTextureProbe(self.h, tex, int(mirror), len(a), _p(a), _p(out))
        return out

    def primary_rays(self, lens, width, height, sample):
        rays = np.zeros((height, width, 6), np.float32)
        st = np.zeros((height, width), np.uint32)
        la = lens.as_array()
        self.L.OraclePrimaryRays(self.h, _p(la), width, height, sample, _p(rays), _p(st))
        return rays, st

    def pick(self, lens, width, height, x, y):
        """input.hlsl:24-57: (instance, geometry, primitive, (u, v)) or None."""
        out = np.zeros(5, np.uint32)
        la = lens.as_array()
        self.L.OraclePick(self.h, _p(la), width, height, float(x), float(y), _p(out))
        if out[0] == 0xFFFFFFFF:
            return None
        uv = out[3:5].view(np.float32)
        return (int(out[0]), int(out[1]), int(out[2]), (float(uv[0]), float(uv[1])))

    def trace(self, rays8, any_hit=False, brute=False):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        hits = np.zeros(len(r), HIT_DTYPE)
        self.L.OracleTrace(self.h, _p(r), len(r), int(any_hit), int(brute), _p(hits))
        return hits

    def env_data(self):
        s = self.L.OracleEnvSize(self.h)
        rgba = np.zeros((s, s, 4), np.float32)
        n = sum((s >> l) ** 2 for l in range(int(np.log2(s)) + 1))
        lum = np.zeros(n, np.float32)
        self.L.OracleEnvData(self.h, _p(rgba), _p(lum))
        return rgba, lum

    def alias_table(self):
        n = self.L.OracleAliasCount(self.h)
        a = np.zeros(n, ALIAS_DTYPE)
        if n:
            self.L.OracleAliasData(self.h, _p(a))
        return a

    def world_to_instance(self, inst):
        m = np.zeros((3, 4), np.float32)
        self.L.OracleWorldToInstance(self.h, inst, _p(m))
        return m


def rng_floats(s, x, y, n):
    out = np.zeros(n, np.float32)
    lib().OracleRngFloats(s, x, y, n, _p(out))
    return out


def math_probe(fn, a, b=None):
    """The defined transcendentals of oracle/utils.h: fn 0 sin, 1 cos, 2 log, 3 atan2(a, b), 4 acos."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.zeros_like(a) if b is None else np.ascontiguousarray(b, np.float32)
    out = np.zeros_like(a)
    lib().OracleMathProbe(fn, a.size, _p(a), _p(b), _p(out))
    return out


def bsdf_batch(queries14):
    """n x {type, color rgb, metalness, roughness, ior, op, a.xyz, w_o.xyz} -> n x 4 (see OracleBsdf)."""
    q = np.ascontiguousarray(queries14, np.float32).reshape(-1, 14)
    out = np.zeros((len(q), 4), np.float32)
    lib().OracleBsdfBatch(len(q), _p(q), _p(out))
    return out


def bsdf(type_, params6, op, w_i, w_o):
    out = np.zeros(4, np.float32)
    p = np.asarray(params6, np.float32)
    a = np.asarray(w_i, np.float32)
    b = np.asarray(w_o, np.float32)
    lib().OracleBsdf(type_, _p(p), op, _p(a), _p(b), _p(out))
    return out
