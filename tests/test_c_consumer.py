"""A compiled C11 translation unit is the consumer of the boundary the reference actually has (its Hydra delegate links
libmoonshine.so and includes hydra/moonshine.h). tests/c/abi_consumer.c includes include/*.h as C, links the library, passes
bool and 48-byte structs by value, renders cfg1 and stresses the mutex from 8 threads."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_consumer.c")
LIBDIR = os.path.join(ROOT, "moonshine_b200")


def _build(tmp_path):
    exe = str(tmp_path / "abi_consumer")
    if not os.path.exists(os.path.join(LIBDIR, "libmoonshine.so")):
        import __graft_entry__ as g
        g.build()
    subprocess.check_call(["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC,
                           "-o", exe, "-L", LIBDIR, "-lmoonshine", f"-Wl,-rpath,{LIBDIR}", "-lpthread"])
    return exe


def test_headers_are_valid_c11_and_the_library_links(tmp_path):
    """No GPU needed: the headers compile as strict C11 (not just as C++), and every symbol the program uses resolves at link time."""
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 64 and "usage" in r.stderr


@pytest.mark.gpu
def test_c_program_renders_cfg1_bit_identically_to_the_python_binding(tmp_path):
    from moonshine_b200 import hydra, scenes
    exe = _build(tmp_path)
    out = str(tmp_path / "cfg1.f32")
    r = subprocess.run([exe, "render", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    c_img = np.fromfile(out, np.float32).reshape(64, 64, 4)
    sc = scenes.single_triangle()
    m = hydra.HdMoonshine()
    scenes.load_scene(m, sc)
    sensor, lens = m.CreateSensor(64, 64), m.CreateLens(sc.lens)
    assert m.Render(sensor, lens)
    np.testing.assert_array_equal(c_img.view(np.uint32), m.GetSensorData(sensor).view(np.uint32))
    assert "rays_closest=" in r.stdout


@pytest.mark.gpu
def test_c_program_multithreaded_stress(tmp_path):
    """hydra.zig:78,146: Hydra calls Sync from worker threads while the render thread runs."""
    exe = _build(tmp_path)
    r = subprocess.run([exe, "stress"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "stress ok" in r.stdout


@pytest.mark.gpu
def test_c_program_multithreaded_stress_with_render_ahead(tmp_path):
    """The same with HdMoonshineExtSetRenderAhead(2): the workers' edits keep dropping the waves the render thread started early."""
    exe = _build(tmp_path)
    r = subprocess.run([exe, "stress", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "stress ok" in r.stdout
