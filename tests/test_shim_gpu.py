"""GPU: the C++ class-surface shims (ORBextractor / ORBmatcher / PointCloudMapping) run end to end on the device."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shims_run(tmp_path):
    shim = os.path.join(ROOT, "orb_slam2_ssd_semantic_b200", "csrc", "shim")
    libdir = os.path.join(ROOT, "orb_slam2_ssd_semantic_b200")
    exe = str(tmp_path / "shim_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(shim, "shim_check.cpp"),
                           "-I" + os.path.join(ROOT, "oracle", "standin"), "-I" + os.path.join(ROOT, "include"), "-L" + libdir,
                           "-lb200orb", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "keypoints" in out.stdout and "UpdateOctomap lag" in out.stdout and "leaves" in out.stdout
