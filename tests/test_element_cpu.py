"""Host build of the element routines the kernels share with the symbolic phase (hd_b200/csrc/element.cuh):
the written-out matrix of a beam without pins is bit-identical to the literal gonum-ordered product."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rigid_beam_closed_form_is_bit_identical(tmp_path):
    exe = os.path.join(str(tmp_path), "test_element_rigid")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off",
                    os.path.join(ROOT, "tests", "cpp", "test_element_rigid.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
