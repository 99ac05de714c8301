"""The C++ host mirror of hd's API (hd_b200/csrc/hd_api.hpp) over the C ABI, compiled with g++ and run on
the GPU box against the reference's fixtures (tests/cpp/test_hd_api.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    exe = os.path.join(tmp, "test_hd_api")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_hd_api.cpp"), "-o", exe,
           "-L" + os.path.join(ROOT, "hd_b200"), "-lhdb200", "-Wl,-rpath," + os.path.join(ROOT, "hd_b200")]
    subprocess.run(cmd, check=True)
    return exe


def test_cpp_host_mirror_compiles(tmp_path):
    """CPU: the header-only mirror compiles and links against the C ABI."""
    _build(str(tmp_path))


@pytest.mark.gpu
def test_cpp_host_mirror_runs(tmp_path):
    exe = _build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
