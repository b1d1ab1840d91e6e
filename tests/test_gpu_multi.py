"""Two GPUs of one node through the product's own multi-GPU path: `offline --gpus 2` (one process per GPU, NCCL communicator
inside libmoonshine.so, ncclReduce of the frame to rank 0) must write the same bits as the single-GPU run. Skipped on a
one-GPU box (two NCCL ranks cannot share a device)."""
import numpy as np
import pytest

from moonshine_b200 import offline, scenes

pytestmark = pytest.mark.gpu


def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_offline_two_gpus_bit_identical_to_one(tmp_path):
    src = scenes.material_zoo((96, 64))
    src.materials[1].color = src.solid3(0.7, 0.6, 0.2)
    src.textures[src.materials[0].metalness] = scenes.Texture(((np.indices((32, 32)).sum(0) % 2) * 255).astype(np.uint8), 32, 32, scenes.TEX_R8_UNORM)
    glb, sky, one, two = (str(tmp_path / n) for n in ("in.glb", "sky.exr", "one.exr", "two.exr"))
    offline.write_glb(src, glb)
    offline.write_exr(sky, src.background)
    assert offline.main([glb, sky, one, "5", "--extent", "96x64"]) == 0
    assert offline.main([glb, sky, two, "5", "--extent", "96x64", "--gpus", "2"]) == 0
    a, b = offline.read_exr(one), offline.read_exr(two)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
