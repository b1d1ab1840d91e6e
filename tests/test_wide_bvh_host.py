"""Host logic of the compressed wide BVH (bvh_build.cpp collapseToBvh8): structural invariants and a scalar
restatement of the device traversal (tests/native/wide_bvh_check.cpp) against brute force. CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_collapse_and_traversal_match_brute_force(tmp_path):
    exe = str(tmp_path / "wide_bvh_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "native", "wide_bvh_check.cpp"),
                           os.path.join(ROOT, "moonshine_b200", "csrc", "bvh_build.cpp"), "-o", exe])
    out = subprocess.run([exe, "6000", "1500"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
