"""Host-side BVH builder (moonshine_b200/csrc/bvh_build.cpp, the replacement of the driver's acceleration-structure build,
Accel.zig:94-184 / :440-489) checked without a GPU: tools/bvh_sim.cpp links the product's builder and its --selftest mode
verifies the invariants the kernels rely on (every item in exactly one leaf, boxes contain their items, depth <= 72 so the
single-level kernel needs no stack-overflow test, traversal == brute force, reorderTopBfs is a pure relabelling)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bvh_builder_invariants_and_traversal(tmp_path):
    exe = str(tmp_path / "bvh_sim")
    cuda_inc = "/usr/local/cuda/include"   # dtypes.h includes cuda_runtime.h for the vector types only
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "moonshine_b200", "csrc"),
                           os.path.join(ROOT, "tools", "bvh_sim.cpp"), os.path.join(ROOT, "moonshine_b200", "csrc", "bvh_build.cpp"), "-o", exe])
    out = subprocess.run([exe, "--selftest"], capture_output=True, text=True)
    assert out.returncode == 0 and "selftest ok" in out.stdout, out.stdout + out.stderr
