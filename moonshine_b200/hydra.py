"""Host-side mirror of the reference's Hydra/offline interface over the C ABI of libmoonshine.so.

`HdMoonshine` exposes the entry points of /root/reference/hydra/moonshine.h:71-95 under their reference
names (CreateMesh, CreateInstance, Render, GetSensorData, ...) plus the extensions of
include/moonshine_ext.h, all through ctypes — plain pointers and sizes, no torch types. The library is
CUDA-only: importing works anywhere, but creating an object without the built .so or without a GPU raises
(there is no CPU fallback and the oracle is never used here).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSNE_LIB", os.path.join(_PKG, "libmoonshine.so"))  # MSNE_LIB: tuning builds (tools/)
_LIB = None


class F32x2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class F32x3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class F32x4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Mat3x4(C.Structure):
    _fields_ = [("x", F32x4), ("y", F32x4), ("z", F32x4)]


class Geometry(C.Structure):  # hydra/moonshine.h:37-41
    _fields_ = [("mesh", C.c_uint32), ("material", C.c_uint32), ("sampled", C.c_bool)]


class Extent2D(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32)]


class Lens(C.Structure):  # hydra/moonshine.h:48-55
    _fields_ = [("origin", F32x3), ("forward", F32x3), ("up", F32x3), ("vfov", C.c_float), ("aperture", C.c_float), ("focus_distance", C.c_float)]


class Material(C.Structure):  # hydra/moonshine.h:57-64
    _fields_ = [("normal", C.c_uint32), ("emissive", C.c_uint32), ("color", C.c_uint32), ("metalness", C.c_uint32), ("roughness", C.c_uint32), ("ior", C.c_float)]


class MsneSpecConstants(C.Structure):  # engine/hrtsystem/pipeline.zig:319-327
    _fields_ = [(n, C.c_uint32) for n in ("samples_per_run", "max_bounces", "env_samples_per_bounce", "mesh_samples_per_bounce",
                                          "flip_image", "indexed_attributes", "two_component_normal_texture")]


class MsneMaterialInfo(C.Structure):
    _fields_ = [("type", C.c_uint32), ("normal", C.c_uint32), ("emissive", C.c_uint32), ("color", C.c_uint32), ("metalness", C.c_uint32),
                ("roughness", C.c_uint32), ("ior", C.c_float)]


class MsneStats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("samples", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("device_ms", C.c_double), ("trace_ms", C.c_double), ("trace_launches", C.c_uint64), ("max_bounce_reached", C.c_uint64),
                ("static_blas_builds", C.c_uint64), ("tlas_builds", C.c_uint64), ("tlas_refits", C.c_uint64), ("sort_ms", C.c_double),
                ("waves", C.c_uint64), ("host_syncs", C.c_uint64), ("tail_paths", C.c_uint64), ("accel_sync_ms", C.c_double),
                ("ahead_discarded", C.c_uint64), ("pool_bytes", C.c_uint64)]


class MsneClickData(C.Structure):   # input.hlsl:1-6
    _fields_ = [("instance_index", C.c_int32), ("geometry_index", C.c_uint32), ("primitive_index", C.c_uint32), ("barycentrics", F32x2)]


# every symbol include/moonshine.h and include/moonshine_ext.h declare: (restype, argtypes)
_VP = C.c_void_p
_U32 = C.c_uint32
SIGNATURES = {
    "HdMoonshineCreate": (_VP, []),
    "HdMoonshineDestroy": (None, [_VP]),
    "HdMoonshineRender": (C.c_bool, [_VP, _U32, _U32]),
    "HdMoonshineRebuildPipeline": (C.c_bool, [_VP]),
    "HdMoonshineCreateMesh": (_U32, [_VP, _VP, _VP, _VP, C.c_size_t, _VP, C.c_size_t]),
    "HdMoonshineCreateSolidTexture1": (_U32, [_VP, C.c_float, C.c_char_p]),
    "HdMoonshineCreateSolidTexture2": (_U32, [_VP, F32x2, C.c_char_p]),
    "HdMoonshineCreateSolidTexture3": (_U32, [_VP, F32x3, C.c_char_p]),
    "HdMoonshineCreateRawTexture": (_U32, [_VP, _VP, Extent2D, C.c_int, C.c_char_p]),
    "HdMoonshineCreateMaterial": (_U32, [_VP, Material]),
    "HdMoonshineSetMaterialNormal": (None, [_VP, _U32, _U32]),
    "HdMoonshineSetMaterialEmissive": (None, [_VP, _U32, _U32]),
    "HdMoonshineSetMaterialColor": (None, [_VP, _U32, _U32]),
    "HdMoonshineSetMaterialMetalness": (None, [_VP, _U32, _U32]),
    "HdMoonshineSetMaterialRoughness": (None, [_VP, _U32, _U32]),
    "HdMoonshineSetMaterialIOR": (None, [_VP, _U32, C.c_float]),
    "HdMoonshineCreateInstance": (_U32, [_VP, Mat3x4, _VP, C.c_size_t, C.c_bool]),
    "HdMoonshineDestroyInstance": (None, [_VP, _U32]),
    "HdMoonshineSetInstanceTransform": (None, [_VP, _U32, Mat3x4]),
    "HdMoonshineSetInstanceVisibility": (None, [_VP, _U32, C.c_bool]),
    "HdMoonshineCreateSensor": (_U32, [_VP, Extent2D]),
    "HdMoonshineGetSensorData": (C.POINTER(C.c_float), [_VP, _U32]),
    "HdMoonshineCreateLens": (_U32, [_VP, Lens]),
    "HdMoonshineSetLens": (None, [_VP, _U32, Lens]),
    # moonshine_ext.h
    "HdMoonshineExtSetDevice": (C.c_bool, [C.c_int]),
    "HdMoonshineExtSetSpecConstants": (C.c_bool, [_VP, C.POINTER(MsneSpecConstants)]),
    "HdMoonshineExtGetSpecConstants": (None, [_VP, C.POINTER(MsneSpecConstants)]),
    "HdMoonshineExtSetBackground": (C.c_bool, [_VP, _VP, Extent2D]),
    "HdMoonshineExtCreateTexture": (_U32, [_VP, _VP, Extent2D, C.c_int, C.c_char_p]),
    "HdMoonshineExtCreateMaterial": (_U32, [_VP, C.POINTER(MsneMaterialInfo)]),
    "HdMoonshineExtCreateMeshIndexed": (_U32, [_VP, _VP, _VP, _VP, C.c_size_t, _VP, C.c_size_t]),
    "HdMoonshineExtBuildEmitterTable": (C.c_bool, [_VP]),
    "HdMoonshineExtRenderLaunches": (C.c_bool, [_VP, _U32, _U32, _U32]),
    "HdMoonshineExtReadSensor": (C.c_bool, [_VP, _U32]),
    "HdMoonshineExtClearSensor": (None, [_VP, _U32]),
    "HdMoonshineExtGetSensorSampleCount": (_U32, [_VP, _U32]),
    "HdMoonshineExtGetSensorDevicePtr": (_VP, [_VP, _U32]),
    "HdMoonshineExtSetTileShard": (C.c_bool, [_VP, _U32, _U32]),
    "HdMoonshineExtCommGetUniqueId": (C.c_bool, [_VP]),
    "HdMoonshineExtCommInit": (C.c_bool, [_VP, _VP, _U32, _U32]),
    "HdMoonshineExtCommReduceSensor": (C.c_bool, [_VP, _U32, _U32]),
    "HdMoonshineExtCommAllReduceU64": (C.c_bool, [_VP, _VP, C.c_size_t]),
    "HdMoonshineExtCommDestroy": (None, [_VP]),
    "HdMoonshineExtSetPathsInFlight": (None, [_VP, C.c_uint64]),
    "HdMoonshineExtSetRaySorting": (None, [_VP, C.c_uint64]),
    "HdMoonshineExtSetProfiling": (None, [_VP, C.c_bool]),
    "HdMoonshineExtSetSchedule": (None, [_VP, C.c_uint64, C.c_uint64, _U32]),
    "HdMoonshineExtReleasePools": (C.c_bool, [_VP]),
    "HdMoonshineExtSetTraceSuspend": (None, [_VP, _U32]),
    "HdMoonshineExtSetRenderAhead": (None, [_VP, _U32]),
    "HdMoonshineExtGetStats": (None, [_VP, C.POINTER(MsneStats)]),
    "HdMoonshineExtResetStats": (None, [_VP]),
    "HdMoonshineExtSetGeometryMaterial": (None, [_VP, _U32, _U32, _U32]),
    "HdMoonshineExtSetEmitterTableFollowsEdits": (None, [_VP, C.c_bool]),
    "HdMoonshineExtPick": (C.c_bool, [_VP, _U32, _U32, F32x2, _VP]),
    "HdMoonshineExtTraceRays": (C.c_bool, [_VP, _VP, C.c_size_t, C.c_bool, _VP]),
    "HdMoonshineExtPrimaryRays": (C.c_bool, [_VP, _U32, _U32, _U32, _VP, _VP]),
    "HdMoonshineExtMathProbe": (C.c_bool, [_VP, C.c_int, _VP, _VP, C.c_size_t, _VP]),
    "HdMoonshineExtBsdfProbe": (C.c_bool, [_VP, _VP, C.c_size_t, _VP]),
    "HdMoonshineExtEnvProbe": (C.c_bool, [_VP, C.c_int, _VP, C.c_size_t, _VP]),
    "HdMoonshineExtGetEnvSize": (_U32, [_VP]),
    "HdMoonshineExtGetEnvData": (C.c_bool, [_VP, _VP, _VP]),
    "HdMoonshineExtGetEmitterTableSize": (_U32, [_VP]),
    "HdMoonshineExtGetEmitterTable": (C.c_bool, [_VP, _VP]),
    "HdMoonshineExtLastError": (C.c_char_p, []),
}


def load_library():
    """dlopen libmoonshine.so and bind every declared symbol. Raises if the library was not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(moonshine_b200/csrc/build.sh). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _mat(t) -> Mat3x4:
    a = np.ascontiguousarray(t, np.float32).reshape(3, 4)
    return Mat3x4(F32x4(*a[0]), F32x4(*a[1]), F32x4(*a[2]))


def _lens(l) -> Lens:
    a = l.as_array() if hasattr(l, "as_array") else np.asarray(l, np.float32)
    return Lens(F32x3(*a[0:3]), F32x3(*a[3:6]), F32x3(*a[6:9]), float(a[9]), float(a[10]), float(a[11]))


ALIAS_DTYPE = np.dtype([("alias", "<u4"), ("select", "<f4"), ("instance", "<u4"), ("geometry", "<u4"), ("primitive", "<u4")])
HIT_DTYPE = np.dtype([("instance", "<u4"), ("geometry", "<u4"), ("primitive", "<u4"), ("u", "<f4"), ("v", "<f4")])


class HdMoonshine:
    """One renderer instance on one CUDA device (hydra/hydra.zig:62-558)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        if not self.L.HdMoonshineExtSetDevice(device):
            raise RuntimeError(f"CUDA device {device} not available: {self.last_error()} (libmoonshine has no CPU path)")
        self.h = self.L.HdMoonshineCreate()
        if not self.h:
            raise RuntimeError("HdMoonshineCreate returned NULL: " + self.last_error())
        self.constants = None

    def last_error(self) -> str:
        e = self.L.HdMoonshineExtLastError()
        return e.decode() if e else ""

    def Destroy(self):
        if getattr(self, "h", None):
            self.L.HdMoonshineDestroy(self.h)
            self.h = None

    close = Destroy

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass

    # -- reference ABI, reference names --------------------------------------------------------
    def Render(self, sensor: int, lens: int) -> bool:
        return bool(self.L.HdMoonshineRender(self.h, sensor, lens))

    def RebuildPipeline(self) -> bool:
        return bool(self.L.HdMoonshineRebuildPipeline(self.h))

    def CreateMesh(self, positions, normals, texcoords, indices) -> int:
        """Hydra convention: normals/texcoords are per-corner arrays of length 3*len(indices)."""
        p = np.ascontiguousarray(positions, np.float32)
        n = None if normals is None else np.ascontiguousarray(normals, np.float32)
        t = None if texcoords is None else np.ascontiguousarray(texcoords, np.float32)
        i = np.ascontiguousarray(indices, np.uint32)
        assert n is None or len(n) == 3 * len(i)
        assert t is None or len(t) == 3 * len(i)
        return self.L.HdMoonshineCreateMesh(self.h, _p(p), _p(n), _p(t), len(p), _p(i), len(i))

    def CreateSolidTexture1(self, v: float, name: str = "") -> int:
        return self.L.HdMoonshineCreateSolidTexture1(self.h, v, name.encode())

    def CreateSolidTexture2(self, x: float, y: float, name: str = "") -> int:
        return self.L.HdMoonshineCreateSolidTexture2(self.h, F32x2(x, y), name.encode())

    def CreateSolidTexture3(self, r: float, g: float, b: float, name: str = "") -> int:
        return self.L.HdMoonshineCreateSolidTexture3(self.h, F32x3(r, g, b), name.encode())

    def CreateRawTexture(self, data: np.ndarray, width: int, height: int, fmt: int, name: str = "") -> int:
        d = np.ascontiguousarray(data)
        return self.L.HdMoonshineCreateRawTexture(self.h, _p(d), Extent2D(width, height), fmt, name.encode())

    def CreateMaterial(self, normal, emissive, color, metalness, roughness, ior=1.5) -> int:
        return self.L.HdMoonshineCreateMaterial(self.h, Material(normal, emissive, color, metalness, roughness, ior))

    def SetMaterialNormal(self, m, i): self.L.HdMoonshineSetMaterialNormal(self.h, m, i)
    def SetMaterialEmissive(self, m, i): self.L.HdMoonshineSetMaterialEmissive(self.h, m, i)
    def SetMaterialColor(self, m, i): self.L.HdMoonshineSetMaterialColor(self.h, m, i)
    def SetMaterialMetalness(self, m, i): self.L.HdMoonshineSetMaterialMetalness(self.h, m, i)
    def SetMaterialRoughness(self, m, i): self.L.HdMoonshineSetMaterialRoughness(self.h, m, i)
    def SetMaterialIOR(self, m, ior): self.L.HdMoonshineSetMaterialIOR(self.h, m, ior)

    def CreateInstance(self, transform, geometries: Sequence[Tuple[int, int, bool]], visible: bool = True) -> int:
        g = (Geometry * len(geometries))(*[Geometry(m, mat, bool(s)) for (m, mat, s) in geometries])
        return self.L.HdMoonshineCreateInstance(self.h, _mat(transform), C.cast(g, C.c_void_p), len(geometries), visible)

    def DestroyInstance(self, i): self.L.HdMoonshineDestroyInstance(self.h, i)
    def SetInstanceTransform(self, i, transform): self.L.HdMoonshineSetInstanceTransform(self.h, i, _mat(transform))
    def SetInstanceVisibility(self, i, visible): self.L.HdMoonshineSetInstanceVisibility(self.h, i, bool(visible))

    def CreateSensor(self, width: int, height: int) -> int:
        s = self.L.HdMoonshineCreateSensor(self.h, Extent2D(width, height))
        self._extents = getattr(self, "_extents", {})
        self._extents[s] = (width, height)
        return s

    def GetSensorData(self, sensor: int) -> np.ndarray:
        """Zero-copy view (H, W, 4) of the library-owned host buffer rewritten by Render/ReadSensor."""
        w, h = self._extents[sensor]
        ptr = self.L.HdMoonshineGetSensorData(self.h, sensor)
        return np.ctypeslib.as_array(ptr, shape=(h, w, 4))

    def CreateLens(self, lens) -> int:
        return self.L.HdMoonshineCreateLens(self.h, _lens(lens))

    def SetLens(self, handle, lens):
        self.L.HdMoonshineSetLens(self.h, handle, _lens(lens))

    # -- extensions ----------------------------------------------------------------------------------
    def SetSpecConstants(self, c) -> None:
        sc = MsneSpecConstants(c.samples_per_run, c.max_bounces, c.env_samples_per_bounce, c.mesh_samples_per_bounce,
                               int(c.flip_image), int(c.indexed_attributes), int(c.two_component_normal_texture))
        if not self.L.HdMoonshineExtSetSpecConstants(self.h, C.byref(sc)):
            raise ValueError("HdMoonshineExtSetSpecConstants: " + self.last_error())
        self.constants = c

    def RenderLaunches(self, sensor: int, lens: int, n_launches: int) -> None:
        if not self.L.HdMoonshineExtRenderLaunches(self.h, sensor, lens, n_launches):
            raise RuntimeError("HdMoonshineExtRenderLaunches: " + self.last_error())

    def ReadSensor(self, sensor: int) -> np.ndarray:
        if not self.L.HdMoonshineExtReadSensor(self.h, sensor):
            raise RuntimeError("HdMoonshineExtReadSensor: " + self.last_error())
        return self.GetSensorData(sensor)

    def ClearSensor(self, sensor: int): self.L.HdMoonshineExtClearSensor(self.h, sensor)
    def SensorSampleCount(self, sensor: int) -> int: return self.L.HdMoonshineExtGetSensorSampleCount(self.h, sensor)
    def SensorDevicePtr(self, sensor: int) -> int: return self.L.HdMoonshineExtGetSensorDevicePtr(self.h, sensor)

    def SetTileShard(self, rank: int, world: int):
        if not self.L.HdMoonshineExtSetTileShard(self.h, rank, world):
            raise ValueError("HdMoonshineExtSetTileShard: " + self.last_error())

    # one rank per GPU over NCCL (include/moonshine_ext.h "multi-GPU")
    def CommInit(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        if not self.L.HdMoonshineExtCommInit(self.h, buf, rank, world):
            raise RuntimeError("HdMoonshineExtCommInit: " + self.last_error())

    def CommReduceSensor(self, sensor: int, root: int = 0):
        if not self.L.HdMoonshineExtCommReduceSensor(self.h, sensor, root):
            raise RuntimeError("HdMoonshineExtCommReduceSensor: " + self.last_error())

    def CommAllReduceU64(self, values):
        a = np.ascontiguousarray(values, np.uint64)
        if not self.L.HdMoonshineExtCommAllReduceU64(self.h, _p(a), a.size):
            raise RuntimeError("HdMoonshineExtCommAllReduceU64: " + self.last_error())
        return a

    def CommDestroy(self):
        self.L.HdMoonshineExtCommDestroy(self.h)

    def SetPathsInFlight(self, n: int): self.L.HdMoonshineExtSetPathsInFlight(self.h, n)
    def SetRaySorting(self, min_rays: int): self.L.HdMoonshineExtSetRaySorting(self.h, min_rays)
    def SetProfiling(self, on: bool): self.L.HdMoonshineExtSetProfiling(self.h, bool(on))

    def SetSchedule(self, tail_paths: Optional[int] = None, binned_paths: Optional[int] = None, wave_contexts: int = 0):
        """Scheduling hints (never change a result); None = library default."""
        big = 0xFFFFFFFFFFFFFFFF
        self.L.HdMoonshineExtSetSchedule(self.h, big if tail_paths is None else tail_paths, big if binned_paths is None else binned_paths, wave_contexts)

    def ReleasePools(self) -> bool: return bool(self.L.HdMoonshineExtReleasePools(self.h))
    def SetTraceSuspend(self, lanes: int): self.L.HdMoonshineExtSetTraceSuspend(self.h, int(lanes))
    def SetRenderAhead(self, waves_ahead: int): self.L.HdMoonshineExtSetRenderAhead(self.h, int(waves_ahead))

    def GetStats(self) -> MsneStats:
        st = MsneStats()
        self.L.HdMoonshineExtGetStats(self.h, C.byref(st))
        return st

    def ResetStats(self): self.L.HdMoonshineExtResetStats(self.h)

    # -- backend protocol of scenes.load_scene (same calls, scene-description types) -----------------
    def create_texture(self, data, w, h, fmt):
        d = np.ascontiguousarray(data)
        return self.L.HdMoonshineExtCreateTexture(self.h, _p(d), Extent2D(w, h), fmt, b"")

    def create_material(self, type_, normal, emissive, color, metalness, roughness, ior):
        info = MsneMaterialInfo(type_, normal, emissive, color, metalness, roughness, ior)
        return self.L.HdMoonshineExtCreateMaterial(self.h, C.byref(info))

    def create_mesh(self, positions, normals, texcoords, indices):
        p = np.ascontiguousarray(positions, np.float32)
        n = None if normals is None else np.ascontiguousarray(normals, np.float32)
        t = None if texcoords is None else np.ascontiguousarray(texcoords, np.float32)
        i = np.ascontiguousarray(indices, np.uint32)
        indexed = self.constants is None or self.constants.indexed_attributes
        f = self.L.HdMoonshineExtCreateMeshIndexed if indexed else self.L.HdMoonshineCreateMesh
        return f(self.h, _p(p), _p(n), _p(t), len(p), _p(i), len(i))

    def create_instance(self, transform, geometries, visible=True):
        return self.CreateInstance(transform, geometries, visible)

    def set_instance_transform(self, handle, transform): self.SetInstanceTransform(handle, transform)
    def set_instance_visibility(self, handle, visible): self.SetInstanceVisibility(handle, visible)

    def set_background(self, rgba):
        a = np.ascontiguousarray(rgba, np.float32)
        if not self.L.HdMoonshineExtSetBackground(self.h, _p(a), Extent2D(a.shape[1], a.shape[0])):
            raise RuntimeError("HdMoonshineExtSetBackground: " + self.last_error())

    def build_emitter_table(self):
        if not self.L.HdMoonshineExtBuildEmitterTable(self.h):
            raise RuntimeError("HdMoonshineExtBuildEmitterTable: " + self.last_error())

    def set_constants(self, c): self.SetSpecConstants(c)

    def SetGeometryMaterial(self, instance: int, geometry: int, material: int):
        self.L.HdMoonshineExtSetGeometryMaterial(self.h, instance, geometry, material)

    def SetEmitterTableFollowsEdits(self, on: bool):
        self.L.HdMoonshineExtSetEmitterTableFollowsEdits(self.h, bool(on))

    def Pick(self, sensor: int, lens: int, x: float, y: float):
        """ObjectPicker.getClickedObject (ObjectPicker.zig:89-128): (instance, geometry, primitive, (u, v)) or None."""
        c = MsneClickData()
        if not self.L.HdMoonshineExtPick(self.h, sensor, lens, F32x2(x, y), C.byref(c)):
            raise RuntimeError("HdMoonshineExtPick: " + self.last_error())
        if c.instance_index < 0:
            return None
        return (c.instance_index, c.geometry_index, c.primitive_index, (c.barycentrics.x, c.barycentrics.y))

    # -- debug probes ------------------------------------------------------------------------------
    def trace(self, rays8, any_hit=False):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        hits = np.zeros(len(r), HIT_DTYPE)
        if not self.L.HdMoonshineExtTraceRays(self.h, _p(r), len(r), bool(any_hit), _p(hits)):
            raise RuntimeError("HdMoonshineExtTraceRays: " + self.last_error())
        return hits

    def primary_rays(self, sensor, lens, sample):
        w, h = self._extents[sensor]
        rays = np.zeros((h, w, 6), np.float32)
        st = np.zeros((h, w), np.uint32)
        if not self.L.HdMoonshineExtPrimaryRays(self.h, sensor, lens, sample, _p(rays), _p(st)):
            raise RuntimeError("HdMoonshineExtPrimaryRays: " + self.last_error())
        return rays, st

    def math_probe(self, fn, a, b=None):
        a = np.ascontiguousarray(a, np.float32)
        b = np.zeros_like(a) if b is None else np.ascontiguousarray(b, np.float32)
        out = np.zeros_like(a)
        if not self.L.HdMoonshineExtMathProbe(self.h, fn, _p(a), _p(b), a.size, _p(out)):
            raise RuntimeError("HdMoonshineExtMathProbe: " + self.last_error())
        return out

    def bsdf_batch(self, queries14):
        q = np.ascontiguousarray(queries14, np.float32).reshape(-1, 14)
        out = np.zeros((len(q), 4), np.float32)
        if not self.L.HdMoonshineExtBsdfProbe(self.h, _p(q), len(q), _p(out)):
            raise RuntimeError("HdMoonshineExtBsdfProbe: " + self.last_error())
        return out

    def env_probe(self, op, inputs3):
        a = np.ascontiguousarray(inputs3, np.float32).reshape(-1, 3)
        out = np.zeros((len(a), 7), np.float32)
        if not self.L.HdMoonshineExtEnvProbe(self.h, op, _p(a), len(a), _p(out)):
            raise RuntimeError("HdMoonshineExtEnvProbe: " + self.last_error())
        return out

    def env_data(self):
        s = self.L.HdMoonshineExtGetEnvSize(self.h)
        rgba = np.zeros((s, s, 4), np.float32)
        n = sum((s >> l) ** 2 for l in range(int(np.log2(s)) + 1))
        lum = np.zeros(n, np.float32)
        if not self.L.HdMoonshineExtGetEnvData(self.h, _p(rgba), _p(lum)):
            raise RuntimeError("HdMoonshineExtGetEnvData: " + self.last_error())
        return rgba, lum

    def alias_table(self):
        n = self.L.HdMoonshineExtGetEmitterTableSize(self.h)
        a = np.zeros(n, ALIAS_DTYPE)
        if not self.L.HdMoonshineExtGetEmitterTable(self.h, _p(a)):
            raise RuntimeError("HdMoonshineExtGetEmitterTable: " + self.last_error())
        return a


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and hands the 128 bytes to the other ranks)."""
    buf = C.create_string_buffer(128)
    if not load_library().HdMoonshineExtCommGetUniqueId(buf):
        raise RuntimeError("HdMoonshineExtCommGetUniqueId: " + (load_library().HdMoonshineExtLastError() or b"").decode())
    return buf.raw


def exchange_unique_id(rank: int, world: int, tag: str, timeout_s: float = 120.0) -> bytes:
    """File rendezvous for the NCCL id between the ranks of ONE node (no torch, no MPI): rank 0 writes, the others poll.
    `tag` must be the same on every rank and unique per job; the parent process id is appended (the ranks of one launcher -
    torchrun or `offline --gpus N` - share it), so a file left behind by an earlier job on the same port is never picked up.
    Rank 0 should call `retire_unique_id` once every rank has joined (CommInit is collective)."""
    import tempfile
    import time
    path = os.path.join(tempfile.gettempdir(), f"msne_nccl_id_{tag}_{os.getppid()}")
    if rank == 0:
        uid = comm_unique_id()
        with open(path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(path + ".tmp", path)
        return uid
    t0 = time.time()
    while time.time() - t0 < timeout_s:
        if os.path.exists(path) and os.path.getsize(path) == 128:
            return open(path, "rb").read()
        time.sleep(0.02)
    raise RuntimeError(f"rank {rank}: no NCCL id from rank 0 at {path}")


def retire_unique_id(tag: str) -> None:
    import tempfile
    try:
        os.remove(os.path.join(tempfile.gettempdir(), f"msne_nccl_id_{tag}_{os.getppid()}"))
    except OSError:
        pass


def render_scene(scene, n_launches: int = 1, device: int = 0, paths_in_flight: Optional[int] = None, shard=None, ray_sorting: Optional[int] = None,
                 tail_paths: Optional[int] = None, binned_paths: Optional[int] = None, wave_contexts: int = 0, trace_suspend: Optional[int] = None):
    """Load a scenes.Scene into a fresh HdMoonshine, render `n_launches` launches, return (image HxWx4, stats, obj)."""
    from . import scenes as _sc
    m = HdMoonshine(device)
    if paths_in_flight:
        m.SetPathsInFlight(paths_in_flight)
    if ray_sorting is not None:
        m.SetRaySorting(ray_sorting)
    if tail_paths is not None or binned_paths is not None or wave_contexts:
        m.SetSchedule(tail_paths, binned_paths, wave_contexts)
    if trace_suspend is not None:
        m.SetTraceSuspend(trace_suspend)
    _sc.load_scene(m, scene)
    if shard:
        m.SetTileShard(*shard)
    sensor = m.CreateSensor(*scene.extent)
    lens = m.CreateLens(scene.lens)
    m.RenderLaunches(sensor, lens, n_launches)
    img = m.ReadSensor(sensor).copy()
    return img, m.GetStats(), m
