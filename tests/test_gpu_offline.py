"""`offline in.glb sky.exr out.exr spp` end to end on the GPU (offline/main.zig), checked against the oracle."""
import numpy as np
import pytest

from moonshine_b200 import offline, scenes
from oracle.binding import Oracle
from gpu_common import gpu_render, within_tol_fraction

pytestmark = pytest.mark.gpu


def test_offline_cli_glb_exr_roundtrip(tmp_path):
    src = scenes.material_zoo((80, 56))
    src.materials[1].color = src.solid3(0.7, 0.6, 0.2)
    src.textures[src.materials[0].metalness] = scenes.Texture(((np.indices((32, 32)).sum(0) % 2) * 255).astype(np.uint8), 32, 32, scenes.TEX_R8_UNORM)
    glb, sky, out = (str(tmp_path / n) for n in ("in.glb", "sky.exr", "out.exr"))
    offline.write_glb(src, glb)
    offline.write_exr(sky, src.background)
    assert offline.main([glb, sky, out, "3", "--extent", "80x56"]) == 0
    img = offline.read_exr(out)
    assert img.shape == (56, 80, 4)
    # same scene through the importer, rendered directly: the CLI adds nothing but file I/O
    sc = offline.load_glb(glb, (80, 56))
    sc.background = offline.read_exr(sky)
    direct, _, _ = gpu_render(sc, 3)
    np.testing.assert_array_equal(img[..., :3].view(np.uint32), direct[..., :3].view(np.uint32))
    # and the imported scene matches the oracle at the usual bar
    o = Oracle()
    scenes.load_scene(o, sc)
    ref, _ = o.render(sc.lens, 80, 56, n_launches=3)
    assert within_tol_fraction(img, ref, rel=2e-4) >= 0.99


def test_offline_cli_argument_errors(tmp_path):
    assert offline.main([]) == 2
    with pytest.raises(SystemExit, match="OnlySupportsGlbInput"):
        offline.main(["scene.gltf", "sky.exr", "out.exr"])
    with pytest.raises(SystemExit, match="OnlySupportsExrOutput"):
        offline.main(["scene.glb", "sky.exr", "out.png"])


def test_native_offline_executable_writes_the_same_exr(tmp_path):
    """moonshine_b200/offline (csrc/offline_main.cpp: offline/main.zig:27-203 as a native program over the C ABI, own GLB / PNG / EXR
    code): the file it writes is, byte for byte, the file the Python front end writes for the same arguments."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "moonshine_b200", "offline")
    if not os.path.exists(exe):
        subprocess.check_call(["bash", os.path.join(root, "moonshine_b200", "csrc", "build.sh")])
    src = scenes.material_zoo((80, 56))
    src.materials[1].color = src.solid3(0.7, 0.6, 0.2)
    src.textures[src.materials[0].metalness] = scenes.Texture(((np.indices((32, 32)).sum(0) % 2) * 255).astype(np.uint8), 32, 32, scenes.TEX_R8_UNORM)
    glb, sky, out_py, out_native = (str(tmp_path / n) for n in ("in.glb", "sky.exr", "py.exr", "native.exr"))
    offline.write_glb(src, glb)
    offline.write_exr(sky, src.background)
    assert offline.main([glb, sky, out_py, "3", "--extent", "80x56"]) == 0
    r = subprocess.run([exe, glb, sky, out_native, "3", "--extent", "80x56"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "seconds to render" in r.stdout
    assert open(out_native, "rb").read() == open(out_py, "rb").read()
