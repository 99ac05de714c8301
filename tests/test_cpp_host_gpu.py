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
    rep = os.path.join(str(tmp_path), "beam.txt")
    rep2 = os.path.join(str(tmp_path), "truss.txt")
    r = subprocess.run([exe, rep, rep2], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    # hd::Run's text report of ConsoleBeam against the reference's golden file "beam": same skeleton (one printed
    # number is round-off noise, 5.7e-16 where the reference has an exact 0), and the same text as the Python mirror
    import io
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fixtures
    import report_util
    from hd_b200 import api
    with open(rep, newline="") as f:
        text = f.read()
    assert report_util.digest(report_util.skeleton(text)) == report_util.GOLD_TEXT["beam"]["skeleton_sha256"]
    m, lcs, mcs = fixtures.GOLDEN["beam"]()
    out = io.StringIO()
    assert api.Run(out, m, lcs, mcs) is None
    assert out.getvalue() == text
    # a model with pins: the C++ and Python mirrors print the same report, with the golden file's skeleton up to the
    # load cases that example runs and this one does not
    from hd_b200.model import LoadCase
    m, lcs, _ = fixtures.GOLDEN["truss"]()
    one = [LoadCase(LoadNodes=lcs[0].LoadNodes)]
    out = io.StringIO()
    assert api.Run(out, m, one, []) is None
    with open(rep2, newline="") as f:
        assert out.getvalue() == f.read()
