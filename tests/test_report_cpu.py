"""The text report (hd_b200/report.py: Model.String, LoadCase.String, ModalCase.String and hd.Run's framing,
model.go:132-255, 853-969) against the reference's golden files, the way example_test.go:12-77 and
model_test.go:551 use them.  Results come from the CPU oracle here (no GPU in this suite); the GPU twin is
tests/test_report_gpu.py."""
import numpy as np
import pytest

import fixtures
import report_util
from hd_b200.model import BucklingResult, ModalResult

# files whose every printed digit is fixed by the algorithm; the others contain eigenvectors (sign, or 1e-16 noise
# next to 0.5) or a negative zero that depend on the eigen solver, and are pinned through the skeleton + numbers
EXACT = ["beam", "beam-dc-part1", "beam-dc-part2", "beam-modal", "model.String", "truss", "truss-modal"]


def oracle_report(oracle, name):
    m, lcs, mcs = fixtures.GOLDEN[name]()
    err, static, modal = oracle.run(m, lcs, mcs)
    assert err is None, err
    for ci, lc in enumerate(lcs):
        r = static[ci]
        lc.PointDisplacementGlobal = np.asarray(r["disp"]).tolist()
        lc.BeamForces = np.asarray(r["forces"]).tolist()
        lc.Reactions = np.asarray(r["react"]).tolist()
        if lc.AmountLinearBuckling > 0:
            fac, shp = r["buckling"]
            lc.LinearBucklingResult = [BucklingResult(float(f), np.asarray(s).tolist()) for f, s in zip(fac, shp)]
    for mi, mc in enumerate(mcs):
        mc.Result = [ModalResult(float(h), np.asarray(s).tolist()) for h, s in zip(modal[mi]["hz"], modal[mi]["shapes"])]
    return report_util.run_report(m, lcs, mcs)


@pytest.mark.parametrize("name", sorted(fixtures.GOLDEN))
def test_report_skeleton(oracle, name):
    text = oracle_report(oracle, name)
    g = report_util.GOLD_TEXT[name]
    assert text.count("\n") == g["lines"]
    assert report_util.digest(report_util.skeleton(text)) == g["skeleton_sha256"]


@pytest.mark.parametrize("name", EXACT)
def test_report_byte_exact(oracle, name):
    assert report_util.digest(oracle_report(oracle, name)) == report_util.GOLD_TEXT[name]["sha256"]


def test_skeleton_masks_only_numbers():
    s = report_util.skeleton("    1    -1.00000e+01     2.30000       12 true\n")
    assert s == "    1 #e #f       12 true\n"


def test_string_methods_match_report_functions():
    m, lcs, mcs = fixtures.GOLDEN["model.String"]()
    from hd_b200 import report
    assert m.String() == report.model_string(m)
    assert lcs[0].String() == report.load_case_string(lcs[0])  # no results yet: loads + "is not calculated"
    assert "Point displacement" not in lcs[0].String()
    assert mcs[0].String() == report.modal_case_string(mcs[0])
