"""Linear buckling on the GPU (hdb_buckling_batch, model.go:491-601) against the reference's goldens and the
oracle.  Factors at 1e-9 relative; shapes up to sign (eigenvectors are sign free, as in the reference's own
comparison, hd_test.go)."""
import json
import os

import numpy as np
import pytest

import fixtures
from hd_b200 import _lib, api, frames

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
TOL = 1e-9


def _shape_close(a, b, tol, what):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    scale = max(np.abs(b).max(), 1e-300)
    d = min(np.abs(a - b).max(), np.abs(a + b).max()) / scale
    assert d <= tol, (what, d)


@pytest.mark.parametrize("name", sorted(k for k, v in GOLD.items() if v["static"] and any(c["buckling"] is not None for c in v["static"])))
def test_golden_buckling_through_run(oracle, name):
    m, lcs, _ = fixtures.GOLDEN[name]()
    err = api.LinearStatic(None, m, *lcs)
    assert err is None, err
    m2, lcs2, _ = fixtures.GOLDEN[name]()
    oerr, ostat, _ = oracle.run(m2, lcs2, [])
    assert oerr is None
    for ci, g in enumerate(GOLD[name]["static"]):
        res = lcs[ci].LinearBucklingResult
        if g["buckling"] is None:
            assert res is None
            continue
        assert res is not None and len(res) == len(g["buckling"]), (name, ci, res, g["buckling"])
        ofac, oshp = ostat[ci]["buckling"]
        for bi, gb in enumerate(g["buckling"]):
            assert abs(res[bi].Factor - gb["factor"]) <= 1e-5 * abs(gb["factor"]), (name, ci, bi, res[bi].Factor, gb["factor"])
            assert abs(res[bi].Factor - ofac[bi]) <= TOL * abs(ofac[bi]), (name, ci, bi, res[bi].Factor, ofac[bi])
            _shape_close(res[bi].PointDisplacementGlobal, oshp[bi], 1e-7, f"{name} case {ci} shape {bi}")


def _column(n_el, pinned_top):
    """Euler column along Y, axial load at the top; n_el elements."""
    from hd_b200.model import BeamProp, Model
    m = Model()
    L = 3.0
    for i in range(n_el + 1):
        m.Points.append([0.0, L * i / n_el])
        m.Supports.append([False, False, False])
    m.Supports[0] = [True, True, not pinned_top]
    if pinned_top:
        m.Supports[n_el] = [True, False, False]
    for i in range(n_el):
        m.Beams.append(BeamProp(N=[i, i + 1], A=12e-4, J=120e-8, E=2.0e11))
    return m, L


def test_euler_column_batch(oracle):
    """A batch of columns with random sections: GPU factors vs the oracle at 1e-9 and vs Euler's formula."""
    n_el, nM = 8, 16
    m, L = _column(n_el, pinned_top=False)
    fm = oracle.flatten_model(m)
    rng = np.random.default_rng(5)
    sec = np.tile(fm.sections, (nM, 1, 1))
    sec[:, :, 1] *= rng.uniform(0.5, 2.0, size=(nM, 1))  # one J per model
    loads = np.zeros((nM, 2, fm.dof))
    loads[:, 0, 3 * n_el + 1] = -1000.0  # compression
    loads[:, 1, 3 * n_el + 1] = -500.0
    t = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)
    status, fail, disp, forces, react = t.static_batch(sec, loads)
    assert not status.any()
    bst, nOut, fac, shp = t.buckling_batch(sec, forces, 3)
    assert not bst.any()
    for i in range(nM):
        fmi = oracle.flatten_model(m)
        fmi.sections[:] = sec[i]
        for c in range(2):
            rc, ofac, oshp = oracle.buckling(fmi, forces[i, c], 3)
            assert rc == oracle.OK and nOut[i, c] == len(ofac) == 3
            np.testing.assert_allclose(fac[i, c, :3], ofac, rtol=TOL)
            for r in range(3):
                _shape_close(shp[i, c, r], oshp[r], 1e-7, f"model {i} case {c} mode {r}")
        # cantilever Euler load pi^2 E J / (4 L^2); 8 elements: 1e-4 discretisation error
        euler = np.pi ** 2 * 2.0e11 * sec[i, 0, 1] / (4 * L * L)
        assert abs(fac[i, 0, 0] * 1000.0 - euler) / euler < 1e-3
        assert abs(fac[i, 1, 0] * 500.0 - euler) / euler < 1e-3
    t.close()


def test_tension_gives_no_factor(oracle):
    """All members in tension with rotational releases everywhere: G_ff = 0, H = 0, no result and no error
    (truss-with-buckling's first case); a tensioned rigid column keeps G (the reference uses max(0, N1, -N2))."""
    m, L = _column(4, pinned_top=False)
    fm = oracle.flatten_model(m)
    loads = np.zeros((1, 1, fm.dof))
    loads[0, 0, 3 * 4 + 1] = +1000.0
    t = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)
    status, fail, disp, forces, react = t.static_batch(fm.sections[None], loads)
    bst, nOut, fac, shp = t.buckling_batch(fm.sections[None], forces, 2)
    rc, ofac, oshp = oracle.buckling(fm, forces[0, 0], 2)
    assert rc == oracle.OK and not bst.any() and nOut[0, 0] == len(ofac)
    np.testing.assert_allclose(fac[0, 0, :len(ofac)], ofac, rtol=TOL)
    t.close()


@pytest.mark.parametrize("N", [32, 128])
def test_buckling_beyond_shared_memory(oracle, N):
    """uniform_frame(32) has 99 free DOFs, uniform_frame(128) 387 (BASELINE.json's size): the three s x s work matrices of
    the eigen kernel no longer fit shared memory and live in a global scratch record per (model, case).  Factors against the
    oracle's non-symmetric solve of K^-1 G (model.go:491-601) at 1e-9 (cond(K_ff) = 25)."""
    fm = oracle.flatten_model(frames.uniform_frame(N))
    nM = 2
    secs = frames.uniform_frame_sections(N, 9, nM)
    loads = frames.uniform_frame_loads(N, 9, nM, 1)
    t = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)
    assert t.info.nFree > t.buckling_max_free()
    status, fail, disp, forces, react = t.static_batch(secs, loads)
    assert not status.any()
    bst, nOut, fac, shp = t.buckling_batch(secs, forces, 3)
    assert not bst.any()
    for i in range(nM if N <= 32 else 1):  # the oracle's dense 774 x 774 non-symmetric eigen solve takes a while
        fmi = oracle.flatten_model(frames.uniform_frame(N))
        fmi.sections[:] = secs[i]
        rc, ofac, oshp = oracle.buckling(fmi, forces[i, 0], 3)
        assert rc == oracle.OK and nOut[i, 0] == len(ofac) == 3
        np.testing.assert_allclose(fac[i, 0, :3], ofac, rtol=TOL)
        for r in range(3):
            _shape_close(shp[i, 0, r], oshp[r], 1e-6, f"frame {N} model {i} mode {r}")
    t.close()
