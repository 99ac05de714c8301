"""Per-model coordinates in one batch (hdb_static_batch_geom: the geometry half of check.go, check.go:223-245 and
294-300, as a device pre-pass) against the oracle run model by model on the moved
points, at 1e-9 (mod.Compare rule)."""
import numpy as np
import pytest

import fixtures
from hd_b200 import _lib, api, frames
from test_static_gpu import compare_rule, force_rule

pytestmark = pytest.mark.gpu


def skew_frame(oracle):
    """No member is axis-aligned, so moving points never changes the sparsity pattern."""
    from hd_b200.model import BeamProp, Model
    m = Model()
    m.Points = [[0.0, 0.0], [1.0, 3.0], [4.5, 3.5], [5.0, 0.2], [8.0, 3.1], [9.0, 0.4]]
    m.Supports = [[True] * 3, [False] * 3, [False] * 3, [True] * 3, [False] * 3, [True, True, False]]
    m.Beams = [BeamProp(N=n, A=12e-4, J=120e-6, E=2e11) for n in ([0, 1], [1, 2], [3, 2], [2, 4], [5, 4])]
    m.Pins = [[False] * 6, [False] * 6, [False, False, False, False, False, True], [False] * 6, [False, False, True, False, False, False]]
    return m


def oracle_per_model(oracle, m, pts, sec, loads):
    out = []
    for i in range(len(pts)):
        mi = m.clone()
        mi.Points = pts[i].tolist()
        fmi = oracle.flatten_model(mi)
        fmi.sections[:] = sec[i]
        rc, disp, forces, react, fc = oracle.linear_static(fmi, loads[i])
        assert rc == oracle.OK
        out.append((disp, forces, react))
    return out


@pytest.mark.parametrize("nM", [1, 7, 66])
def test_moved_points_vs_oracle(oracle, nM):
    m = skew_frame(oracle)
    fm = oracle.flatten_model(m)
    rng = np.random.default_rng(5 + nM)
    nLC = 3
    pts = np.tile(np.asarray(m.Points), (nM, 1, 1)) + rng.uniform(-0.3, 0.3, size=(nM, fm.nP, 2))
    sec = np.tile(fm.sections, (nM, 1, 1))
    sec[:, :, :2] *= rng.uniform(0.5, 2.0, size=(nM, fm.nB, 2))
    loads = rng.normal(size=(nM, nLC, fm.dof)) * 1e3
    t = api._topology(m.clone())
    for g in t.query(_lib.Q_IGNORE):  # no load on DOFs without stiffness (model.go:424-437)
        loads[:, :, int(g)] = 0.0
    t.close()
    status, fail, disp, forces, react = api.RunBatch(m, sec, loads, points=pts)
    assert not status.any(), status
    want = oracle_per_model(oracle, m, pts, sec, loads)
    for i in range(nM):
        for c in range(nLC):
            compare_rule(want[i][0][c], disp[i, c])
            compare_rule(want[i][2][c], react[i, c])
            force_rule(want[i][1][c], forces[i, c])


@pytest.mark.parametrize("nLC", [4, 45])
def test_uniform_frame_heights_vs_oracle(oracle, nLC):
    """A portal frame whose storey height and bay widths vary per model: columns stay vertical and girders
    horizontal, so the pattern (exact zeros of axis-aligned members) is the topology's for every model.
    45 load cases: two passes of the solve kernel per model (the geometry row is re-staged for each)."""
    N, nM = 8, 9
    m = frames.uniform_frame(N)
    fm = oracle.flatten_model(m)
    rng = np.random.default_rng(11)
    base = np.asarray(m.Points)
    pts = np.tile(base, (nM, 1, 1))
    pts[:, :, 1] *= rng.uniform(0.7, 1.6, size=(nM, 1))          # storey height
    pts[:, :, 0] *= rng.uniform(0.8, 1.3, size=(nM, 1))          # bay width
    sec = frames.uniform_frame_sections(N, 0, nM)
    loads = frames.uniform_frame_loads(N, 0, nM, nLC)
    status, fail, disp, forces, react = api.RunBatch(m, sec, loads, points=pts)
    assert not status.any(), status
    want = oracle_per_model(oracle, m, pts, sec, loads)
    for i in range(nM):
        for c in range(nLC):
            compare_rule(want[i][0][c], disp[i, c])
            compare_rule(want[i][2][c], react[i, c])
            force_rule(want[i][1][c], forces[i, c])
    # the topology's own coordinates for every model: bit-identical to the shared-geometry path
    same = np.tile(base, (nM, 1, 1))
    a = api.RunBatch(m, sec, loads, points=same)
    b = api.RunBatch(m, sec, loads)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


def test_bad_geometry_is_per_model(oracle):
    """checkPoints / checkBeams for a batch: NaN or Inf coordinates and zero-length beams flag their own model;
    a column pushed out of plumb turns exact zeros of the pattern into numbers: HDB_STATUS_PATTERN."""
    N, nM = 4, 8
    m = frames.uniform_frame(N)
    fm = oracle.flatten_model(m)
    pts = np.tile(np.asarray(m.Points), (nM, 1, 1))
    pts[1, 3, 0] = np.nan
    pts[2, 5, 1] = np.inf
    pts[4, 3] = pts[4, 2]            # column of zero length (check.go:294-300)
    pts[6, 1, 0] += 0.25             # skewed column: entries the representative pattern dropped
    sec = frames.uniform_frame_sections(N, 0, nM)
    loads = frames.uniform_frame_loads(N, 0, nM, 2)
    with np.errstate(all="ignore"):
        status, fail, disp, forces, react = api.RunBatch(m, sec, loads, points=pts)
    assert list(status) == [0, _lib.STATUS_NONFINITE, _lib.STATUS_NONFINITE, 0, _lib.STATUS_NONFINITE, 0, _lib.STATUS_PATTERN, 0]
    good = [0, 3, 5, 7]
    want = oracle_per_model(oracle, m, pts[good], sec[good], loads[good])
    for k, i in enumerate(good):
        for c in range(2):
            compare_rule(want[k][0][c], disp[i, c])
            compare_rule(want[k][2][c], react[i, c])
            force_rule(want[k][1][c], forces[i, c])
