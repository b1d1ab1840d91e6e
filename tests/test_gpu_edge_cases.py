"""Edge cases of the ABI on the GPU: empty / degenerate inputs, object lifecycle, accumulation resets."""
import math

import numpy as np
import pytest

from moonshine_b200 import hydra, scenes
from oracle.binding import Oracle
from gpu_common import gpu_render, oracle_render, within_tol_fraction

pytestmark = pytest.mark.gpu


def test_empty_scene_renders_the_environment():
    """No instances at all: every primary ray misses -> background (default 1x1 white, BackgroundManager.zig:116-126)."""
    m = hydra.HdMoonshine()
    m.SetSpecConstants(scenes.Constants())
    sensor, lens = m.CreateSensor(33, 17), m.CreateLens(scenes.test_lens())
    m.RenderLaunches(sensor, lens, 2)
    img = m.ReadSensor(sensor)
    np.testing.assert_array_equal(img[..., :3], 1.0)
    np.testing.assert_array_equal(img[..., 3], 2.0)
    st = m.GetStats()
    assert st.rays_closest == 2 * 33 * 17 and st.rays_shadow == 0


def test_all_instances_invisible_equals_empty_scene():
    sc = scenes.cornell_box((24, 24))
    m = hydra.HdMoonshine()
    _, _, _, insts = scenes.load_scene(m, sc)
    for i in insts:
        m.SetInstanceVisibility(i, False)
    sensor, lens = m.CreateSensor(24, 24), m.CreateLens(sc.lens)
    m.RenderLaunches(sensor, lens, 1)
    np.testing.assert_array_equal(m.ReadSensor(sensor)[..., :3], 0.0)  # black env, nothing to hit
    m.DestroyInstance(insts[0])  # = set invisible (hydra.zig:495-497); must not crash


def test_zero_triangle_mesh_and_transformed_empty_instance():
    sc = scenes.single_triangle((32, 32))
    empty = sc.add_mesh(np.zeros((3, 3), np.float32), np.zeros((0, 3), np.uint32))
    sc.add_instance([(empty, 0, False)])
    sc.add_instance([(empty, 0, False)], transform=[[2, 0, 0, 1], [0, 2, 0, 0], [0, 0, 2, 0]])
    g, _, _ = gpu_render(sc, 1)
    r, _, _ = oracle_render(scenes.single_triangle((32, 32)), 1)
    assert within_tol_fraction(g, r) >= 0.999


def test_multiple_sensors_and_lenses_are_independent():
    sc = scenes.cornell_box((32, 32))
    m = hydra.HdMoonshine()
    scenes.load_scene(m, sc)
    s0, s1 = m.CreateSensor(32, 32), m.CreateSensor(16, 24)
    l0 = m.CreateLens(sc.lens)
    l1 = m.CreateLens(scenes.Lens((-1.0, 0.5, 0.5), (1, 0, 0), (0, 0, 1), math.radians(60)))
    m.RenderLaunches(s0, l0, 2)
    m.RenderLaunches(s1, l1, 1)
    assert (m.SensorSampleCount(s0), m.SensorSampleCount(s1)) == (2, 1)
    a = m.ReadSensor(s0).copy()
    ref, _, _ = gpu_render(sc, 2)
    np.testing.assert_array_equal(a.view(np.uint32), ref.view(np.uint32))
    assert m.ReadSensor(s1).shape == (24, 16, 4)
    # SetLens / RebuildPipeline clear every sensor (hydra.zig:539, 370)
    m.SetLens(l0, sc.lens)
    assert (m.SensorSampleCount(s0), m.SensorSampleCount(s1)) == (0, 0)
    m.RenderLaunches(s0, l0, 1)
    assert m.RebuildPipeline() and m.SensorSampleCount(s0) == 0
    # material edits do NOT clear accumulation (hydra.zig:435-481)
    m.RenderLaunches(s0, l0, 1)
    m.SetMaterialIOR(0, 1.4)
    assert m.SensorSampleCount(s0) == 1


def test_hydra_material_setters_take_effect():
    """HdMoonshineSetMaterial* are applied by the next Render (hydra.zig:150-223)."""
    m = hydra.HdMoonshine()
    normal, black = m.CreateSolidTexture2(0.5, 0.5), m.CreateSolidTexture3(0, 0, 0)
    red, white = m.CreateSolidTexture3(0.8, 0.1, 0.1), m.CreateSolidTexture3(0.9, 0.9, 0.9)
    mat = m.CreateMaterial(normal, black, red, m.CreateSolidTexture1(0.0), m.CreateSolidTexture1(0.9), 1.5)
    p = np.array([(0, -1, -1), (0, 1, -1), (0, 0, 1)], np.float32)
    mesh = m.CreateMesh(p, None, None, np.array([(0, 1, 2)], np.uint32))
    m.CreateInstance(np.eye(3, 4, dtype=np.float32), [(mesh, mat, False)])
    sensor, lens = m.CreateSensor(16, 16), m.CreateLens(scenes.test_lens())
    assert m.Render(sensor, lens)
    a = m.GetSensorData(sensor)[8, 8, :3].copy()
    m.SetMaterialColor(mat, white)
    m.ClearSensor(sensor)
    assert m.Render(sensor, lens)
    b = m.GetSensorData(sensor)[8, 8, :3].copy()
    assert a[0] > 2 * a[1] and abs(b[0] - b[1]) < 0.05 * b[0]  # red -> grey


def test_deep_degenerate_bvh_does_not_overflow_the_stack():
    """10 000 coincident-centroid triangles (a fan around one point) force index splits; traversal stays correct."""
    n = 10000
    ang = np.linspace(0, 2 * math.pi, n, endpoint=False)
    pos = np.zeros((2 * n + 1, 3), np.float32)
    pos[1:n + 1] = np.stack([np.zeros(n), np.cos(ang), np.sin(ang)], 1)
    pos[n + 1:] = np.stack([np.zeros(n), -np.cos(ang), -np.sin(ang)], 1)   # each triangle spans the centre symmetrically
    idx = np.stack([np.arange(1, n + 1), np.roll(np.arange(1, n + 1), -1), np.arange(n + 1, 2 * n + 1)], 1).astype(np.uint32)
    sc = scenes.single_triangle((24, 24))
    sc.meshes[0] = scenes.Mesh(pos, idx)
    m = hydra.HdMoonshine()
    scenes.load_scene(m, sc)
    o = Oracle()
    scenes.load_scene(o, sc)
    rs = np.random.RandomState(4)
    rays = np.zeros((2000, 8), np.float32)
    rays[:, 0] = -3
    rays[:, 1:3] = rs.rand(2000, 2) * 1.6 - 0.8
    rays[:, 4] = 1
    rays[:, 7] = 1e12
    g, r = m.trace(rays), o.trace(rays)
    np.testing.assert_array_equal(g["instance"], r["instance"])
    hit = r["instance"] != 0xFFFFFFFF
    assert hit.mean() > 0.5
    np.testing.assert_array_equal(g["primitive"][hit], r["primitive"][hit])


def test_transform_edits_of_instanced_objects_rebuild_only_the_tlas():
    """§8(f) row 3: SetInstanceTransform / SetInstanceVisibility / CreateInstance on TRANSFORMED instances must not
    rebuild the BLAS level (the reference does a TLAS update, hydra.zig:225-311); results still match the oracle."""
    sc = scenes.mesh_light_stress(8, (64, 64))
    m = hydra.HdMoonshine()
    _, mats, meshes, insts = scenes.load_scene(m, sc)
    o = Oracle()
    _, _, _, oinsts = scenes.load_scene(o, sc)
    sensor, lens = m.CreateSensor(64, 64), m.CreateLens(sc.lens)
    m.RenderLaunches(sensor, lens, 1)
    s0 = m.GetStats()
    assert s0.static_blas_builds == 1 and s0.tlas_builds == 1
    moved = sc.instances[5].transform.copy()
    moved[:, 3] += np.array([0.3, -0.2, 0.4], np.float32)
    m.SetInstanceTransform(insts[5], moved)
    o.set_instance_transform(oinsts[5], moved)
    m.SetInstanceVisibility(insts[9], False)
    o.set_instance_visibility(oinsts[9], False)
    extra = np.array([[0.2, 0, 0, 4.0], [0, 0.2, 0, 4.0], [0, 0, 0.2, 2.5]], np.float32)
    ico_mesh, light_mat = sc.instances[5].geometries[0][0], sc.instances[5].geometries[0][1]
    m.CreateInstance(extra, [(meshes[ico_mesh], mats[light_mat], True)])
    o.create_instance(extra, [(ico_mesh, light_mat, True)])
    m.RenderLaunches(sensor, lens, 1)
    s1 = m.GetStats()
    assert s1.static_blas_builds == 1, "BLAS level was rebuilt for a TLAS-only edit"
    assert s1.tlas_builds == 2
    rays = np.zeros((4000, 8), np.float32)
    rs = np.random.RandomState(3)
    rays[:, 0:3] = rs.rand(4000, 3) * np.array([8, 8, 3]) + np.array([0, 0, 0.2])
    d = rs.randn(4000, 3); rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 7] = 1e12
    g, r = m.trace(rays), o.trace(rays)
    np.testing.assert_array_equal(g["instance"], r["instance"])
    hit = r["instance"] != 0xFFFFFFFF
    np.testing.assert_array_equal(g["primitive"][hit], r["primitive"][hit])
    assert not (g["instance"] == insts[9]).any()
    # hiding a member of the flattened static BLAS (the ground plane, 2 of its 770 triangles) rebuilds nothing either: the
    # member stays in the tree and traversal skips it
    m.SetInstanceVisibility(insts[0], False)
    o.set_instance_visibility(oinsts[0], False)
    g, r = m.trace(rays), o.trace(rays)
    np.testing.assert_array_equal(g["instance"], r["instance"])
    assert not (g["instance"] == insts[0]).any()
    s2 = m.GetStats()
    assert (s2.static_blas_builds, s2.tlas_builds) == (1, 2)
    # a TRANSFORM edit of a flattened instance changes the BLAS level and does rebuild it
    lifted = np.eye(3, 4, dtype=np.float32)
    lifted[2, 3] = 0.5
    m.SetInstanceVisibility(insts[0], True)
    m.SetInstanceTransform(insts[0], lifted)
    m.RenderLaunches(sensor, lens, 1)
    assert m.GetStats().static_blas_builds == 2


def _probe_rays(n, seed, lo, size):
    rs = np.random.RandomState(seed)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rs.rand(n, 3) * np.array(size) + np.array(lo)
    d = rs.randn(n, 3)
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 7] = 1e12
    return rays


def test_transform_only_edits_refit_the_tlas_and_hits_stay_exact():
    """§8(f) row 3 (Accel.zig:567-602 + :629-679, TLAS build in update mode): moving transformed instances keeps the TLAS
    topology and refits its boxes; hit records stay bit-identical to the oracle (which rebuilds from scratch) however far
    the instances move; after MSNE_TLAS_MAX_REFITS refits the tree is rebuilt."""
    sc = scenes.mesh_light_stress(8, (64, 64))
    m = hydra.HdMoonshine()
    _, _, _, insts = scenes.load_scene(m, sc)
    o = Oracle()
    _, _, _, oinsts = scenes.load_scene(o, sc)
    rays = _probe_rays(6000, 4, (0, 0, 0.2), (8, 8, 3))
    m.trace(rays)
    s0 = m.GetStats()
    assert (s0.tlas_builds, s0.tlas_refits) == (1, 0)
    rs = np.random.RandomState(9)
    movable = [i for i in range(len(sc.instances)) if not np.array_equal(sc.instances[i].transform, np.eye(3, 4, dtype=np.float32))]
    for step in range(3):
        for i in rs.choice(movable, 12, replace=False):
            t = sc.instances[i].transform.copy()
            t[:, 3] += (rs.rand(3).astype(np.float32) - 0.5) * np.array([6.0, 6.0, 1.5], np.float32)   # far from where the tree was built
            m.SetInstanceTransform(insts[i], t)
            o.set_instance_transform(oinsts[i], t)
        g, r = m.trace(rays), o.trace(rays)
        np.testing.assert_array_equal(g["instance"], r["instance"])
        hit = r["instance"] != 0xFFFFFFFF
        assert hit.mean() > 0.2
        for f in ("geometry", "primitive"):
            np.testing.assert_array_equal(g[f][hit], r[f][hit])
        np.testing.assert_array_equal(g["u"][hit].view(np.uint32), r["u"][hit].view(np.uint32))
        st = m.GetStats()
        assert (st.static_blas_builds, st.tlas_builds, st.tlas_refits) == (1, 1, step + 1)
    # a visibility change rebuilds nothing: the instance keeps its TLAS leaf and traversal skips it (mask 0, Accel.zig:401)
    m.SetInstanceVisibility(insts[movable[0]], False)
    o.set_instance_visibility(oinsts[movable[0]], False)
    g, r = m.trace(rays), o.trace(rays)
    np.testing.assert_array_equal(g["instance"], r["instance"])
    assert not (g["instance"] == insts[movable[0]]).any()
    st = m.GetStats()
    assert (st.tlas_builds, st.tlas_refits) == (1, 3)
    # moving a hidden instance and showing it again: still no rebuild, and it is found at its new place
    t = sc.instances[movable[0]].transform.copy()
    t[:, 3] += np.array([1.0, -1.0, 0.3], np.float32)
    m.SetInstanceTransform(insts[movable[0]], t)
    o.set_instance_transform(oinsts[movable[0]], t)
    m.SetInstanceVisibility(insts[movable[0]], True)
    o.set_instance_visibility(oinsts[movable[0]], True)
    g, r = m.trace(rays), o.trace(rays)
    np.testing.assert_array_equal(g["instance"], r["instance"])
    hit = r["instance"] != 0xFFFFFFFF
    np.testing.assert_array_equal(g["u"][hit].view(np.uint32), r["u"][hit].view(np.uint32))
    st = m.GetStats()
    assert (st.static_blas_builds, st.tlas_builds, st.tlas_refits) == (1, 1, 4)
    # once a quarter of the TLAS leaves are hidden the next sync rebuilds without them
    for i in movable[: len(movable) // 4 + 2]:
        m.SetInstanceVisibility(insts[i], False)
        o.set_instance_visibility(oinsts[i], False)
    g, r = m.trace(rays), o.trace(rays)
    np.testing.assert_array_equal(g["instance"], r["instance"])
    assert m.GetStats().tlas_builds == 2


def test_emitter_table_follows_edits_when_asked():
    """The reference samples only the emitters present at Accel.create (Accel.zig:187, :491-538). With
    HdMoonshineExtSetEmitterTableFollowsEdits the table is rebuilt after an emitter moves (areas change under a
    non-uniform scale) or a new one is added, and equals the oracle's table built from the edited scene, byte for byte."""
    sc = scenes.material_zoo()
    m = hydra.HdMoonshine()
    _, mats, meshes, insts = scenes.load_scene(m, sc)
    o = Oracle()
    _, _, _, oinsts = scenes.load_scene(o, sc)
    m.SetEmitterTableFollowsEdits(True)
    emitter = next(i for i, inst in enumerate(sc.instances) if any(g[2] for g in inst.geometries))
    t = sc.instances[emitter].transform.copy()
    t[0, 0] *= 2.5
    t[:, 3] += np.array([0.2, 0.1, 0.3], np.float32)
    m.SetInstanceTransform(insts[emitter], t)
    o.set_instance_transform(oinsts[emitter], t)
    g = sc.instances[emitter].geometries[0]
    extra = np.array([[0.2, 0, 0, 1.0], [0, 0.4, 0, 1.0], [0, 0, 0.2, 1.5]], np.float32)
    m.CreateInstance(extra, [(meshes[g[0]], mats[g[1]], True)])
    o.create_instance(extra, [(g[0], g[1], True)])
    o.build_emitter_table()
    sensor, lens = m.CreateSensor(*sc.extent), m.CreateLens(sc.lens)
    m.RenderLaunches(sensor, lens, 1)   # the rebuild happens here
    np.testing.assert_array_equal(m.alias_table().view(np.uint8), o.alias_table().view(np.uint8))
    ref, _ = o.render(sc.lens, *sc.extent)
    assert within_tol_fraction(m.ReadSensor(sensor), ref) >= 0.995


def test_set_geometry_material_equals_building_with_that_material():
    """Accel.recordUpdateSingleMaterial (Accel.zig:609-627): switching one geometry's material afterwards gives the same
    image, bit for bit, as creating the instance with it."""
    sc = scenes.cornell_box((48, 48))
    a = hydra.HdMoonshine()
    _, mats, _, insts = scenes.load_scene(a, sc)
    sensor, lens = a.CreateSensor(*sc.extent), a.CreateLens(sc.lens)
    a.RenderLaunches(sensor, lens, 1)
    target, gi = 0, 1                                  # the left (red) wall of the box instance
    old = sc.instances[target].geometries[gi]
    new_mat = sc.instances[target].geometries[2][1]    # becomes green like the right wall
    assert new_mat != old[1]
    a.SetGeometryMaterial(insts[target], gi, mats[new_mat])
    a.ClearSensor(sensor)
    a.RenderLaunches(sensor, lens, 1)
    img_a = a.ReadSensor(sensor).copy()
    sc.instances[target].geometries[gi] = (old[0], new_mat, old[2])
    img_b, _, _ = gpu_render(sc, 1)
    assert not np.array_equal(img_a, a.GetSensorData(sensor) * 0)
    np.testing.assert_array_equal(img_a.view(np.uint32), img_b.view(np.uint32))


def _one_sample_calls(m, sensor, lens, n):
    """n HdMoonshineRender calls (one sample each, sensor read back every call) -> the n frames"""
    frames = []
    for _ in range(n):
        assert m.Render(sensor, lens), m.last_error()
        frames.append(m.GetSensorData(sensor).copy())
    return frames


@pytest.mark.parametrize("waves_ahead", [1, 3])
def test_render_ahead_frames_are_the_bits_of_plain_rendering(waves_ahead):
    """HdMoonshineExtSetRenderAhead: every call's frame and sample count are those of render-ahead off (the speculative wave is the
    sample the next call would have rendered itself); edits, another lens and sensor clears in between drop it unseen."""
    sc = scenes.material_zoo((96, 64))
    sc.constants = scenes.Constants(samples_per_run=1, max_bounces=8, env_samples_per_bounce=1, mesh_samples_per_bounce=1)
    other_lens = scenes.Lens((0.2, -3.0, 1.2), (0, 1, -0.2), (0, 0, 1), math.radians(50))

    def session(ahead):
        m = hydra.HdMoonshine()
        _, _, _, insts = scenes.load_scene(m, sc)
        m.SetRenderAhead(ahead)
        sensor, lens = m.CreateSensor(*sc.extent), m.CreateLens(sc.lens)
        lens2 = m.CreateLens(other_lens)
        out = _one_sample_calls(m, sensor, lens, 4)
        m.SetInstanceTransform(insts[0], [[1, 0, 0, 0.25], [0, 1, 0, 0], [0, 0, 1, 0.1]])   # an edit between two calls
        out += _one_sample_calls(m, sensor, lens, 2)
        out += _one_sample_calls(m, sensor, lens2, 2)                                        # another lens
        m.SetLens(lens, other_lens)                                                          # the same handle, new values
        out += _one_sample_calls(m, sensor, lens, 2)
        m.ClearSensor(sensor)
        out += _one_sample_calls(m, sensor, lens, 3)
        assert m.SensorSampleCount(sensor) == 3
        st = m.GetStats()
        m.Destroy()
        return out, st

    plain, st0 = session(0)
    ahead, st1 = session(waves_ahead)
    assert len(plain) == len(ahead) == 13
    for i, (a, b) in enumerate(zip(plain, ahead)):
        assert (a.view(np.uint32) == b.view(np.uint32)).all(), f"call {i}: render-ahead frame differs"
    assert st0.ahead_discarded == 0 and st1.ahead_discarded >= 4   # the edit, the other lens, SetLens, the clear
    assert st1.samples == st0.samples


def test_released_pools_are_rebuilt_on_demand_with_the_same_bits():
    """HdMoonshineExtReleasePools: the wavefront pools (sized by the largest render so far) go back to the driver; the next render
    allocates what it needs and continues the accumulation bit for bit."""
    sc = scenes.material_zoo((96, 64))

    def session(release):
        m = hydra.HdMoonshine()
        scenes.load_scene(m, sc)
        m.SetRaySorting(1)
        sensor, lens = m.CreateSensor(*sc.extent), m.CreateLens(sc.lens)
        m.RenderLaunches(sensor, lens, 3)
        held = m.GetStats().pool_bytes
        assert held > 96 * 64 * 300
        if release:
            assert m.ReleasePools()
            assert m.GetStats().pool_bytes == 0
        m.RenderLaunches(sensor, lens, 2)
        assert m.GetStats().pool_bytes > 0
        img = m.ReadSensor(sensor).copy()
        m.Destroy()
        return img

    a, b = session(False), session(True)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
