"""hd.Run's text report produced from GPU results (api.Run(out, ...)) against the reference's golden files:
the skeleton (all text, indices, layout; numbers masked) of every file, and byte equality where every printed
digit is determined by the algorithm (see tests/test_report_cpu.py)."""
import io

import pytest

import fixtures
import report_util
from hd_b200 import api

pytestmark = pytest.mark.gpu

# byte equality holds where no printed digit is round-off noise (1e-16 next to 1, which follows the elimination
# order: LDL^T here, CSparse LU there) and no eigenvector sign is left to the eigen solver; the remaining files
# are pinned by the skeleton here and number by number in test_static_gpu / test_modal_gpu / test_buckling_gpu
EXACT = ["beam-dc-part1", "beam-modal", "beam-modal-rotate", "model.String", "truss-modal", "truss-modal-rotate"]


def gpu_report(name):
    m, lcs, mcs = fixtures.GOLDEN[name]()
    out = io.StringIO()
    err = api.Run(out, m, lcs, mcs)
    assert err is None, err
    return out.getvalue()


@pytest.mark.parametrize("name", sorted(fixtures.GOLDEN))
def test_gpu_report_skeleton(name):
    text = gpu_report(name)
    g = report_util.GOLD_TEXT[name]
    assert text.count("\n") == g["lines"]
    assert report_util.digest(report_util.skeleton(text)) == g["skeleton_sha256"]


@pytest.mark.parametrize("name", EXACT)
def test_gpu_report_byte_exact(name):
    text = gpu_report(name)
    assert report_util.digest(text) == report_util.GOLD_TEXT[name]["sha256"], text
