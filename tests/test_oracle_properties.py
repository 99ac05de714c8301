"""Oracle vs the reference's own property / known-answer tests, restated:
TestGeometricBeam2d (feBeam2d_test.go:11-164), TestSplit (mod/split_test.go:31-91),
TestRotateBeamLinear / TestRotateBeamModal (model_test.go:407-506), TestDirectionLoadNode
(model_test.go:352-405), TestModelFail (model_test.go:50-324), TestWrongLoad (model_test.go:18-25),
plus LAPACK cross-checks of the third-party stand-ins and the structural anchors of SURVEY.md App. B."""
import math

import numpy as np
import pytest

import fixtures
from hd_b200 import frames
from hd_b200.model import BeamProp, LoadCase, LoadNode, ModalCase, ModalMass, Model

INF, NAN = float("inf"), float("nan")
T, F = True, False


# ------------------------------------------------------------------ element known answers
def _sym(lower):
    a = np.zeros((6, 6))
    for r, row in enumerate(lower):
        for c, v in enumerate(row):
            a[r, c] = v
            a[c, r] = v
    return a


GEOM = [
    (1.0, (F,) * 6, [0, 0, 0, -1, 0, 0],
     [[0.0], [0.0, 1.2], [0.0, 0.1, 2. / 15.], [0.0], [0.0, -1.2, -0.1, 0.0, 1.2], [0.0, 0.1, -1. / 30., 0.0, -0.1, 2. / 15.]]),
    (2.0, (F,) * 6, [1, 0, 0, 0, 0, 0],
     [[0.0], [0.0, 1.2 / 2.], [0.0, 0.1, 2. / 15. * 2.], [0.0], [0.0, -1.2 / 2., -0.1, 0.0, 1.2 / 2.],
      [0.0, 0.1, -1. / 30. * 2., 0.0, -0.1, 2. / 15. * 2.]]),
    (2.0, (F, F, T, F, F, F), [1, 0, 0, -1, 0, 0],
     [[0.0], [0.0, 1.2 / 2.], [0.0], [0.0], [0.0, -1.2 / 2., 0.0, 0.0, 1.2 / 2.], [0.0, 0.2, 0.0, 0.0, -0.2, 0.2 * 2]]),
    (2.0, (F, F, F, F, F, T), [1, 0, 0, -1, 0, 0],
     [[0.0], [0.0, 1.2 / 2.], [0.0, 0.2, 1. / 5. * 2.], [0.0], [0.0, -1.2 / 2., -0.2, 0, 1.2 / 2.], [0.0]]),
]


@pytest.mark.parametrize("case", range(len(GEOM)))
def test_geometric_beam2d_tables(oracle, case):
    L, pins, bf, lower = GEOM[case]
    g = oracle.geometric_beam2d([0, 0, L, 0], pins, bf)
    exp = _sym(lower)
    eps = 1e-16
    for r in range(6):
        for c in range(6):
            if exp[r, c] == 0.0:
                assert abs(g[r, c]) <= eps
            else:
                assert abs(g[r, c] - exp[r, c]) / abs(exp[r, c]) <= eps


def test_truss_pin_condensation_anchor(oracle):
    """literal pin rule on Truss beam 0 gives a pure axial matrix."""
    k = oracle.stiff_beam2d([0, 0, 0, 12], 40e-4, 1.0, 2e11, (F, T, T, F, T, T))
    ax = 2e11 * 40e-4 / 12
    exp = np.zeros((6, 6))
    exp[0, 0] = exp[3, 3] = ax
    exp[0, 3] = exp[3, 0] = -ax
    assert np.array_equal(k != 0, exp != 0)
    assert np.allclose(k, exp, rtol=1e-15, atol=0)
    # single rotational pin at the end: propped-beam 3EJ/L^3 family
    E, J, L = 2e11, 120e-6, 2.0
    k5 = oracle.stiff_beam2d([0, 0, L, 0], 12e-4, J, E, (F, F, F, F, F, T))
    assert np.isclose(k5[1, 1], 3 * E * J / L ** 3, rtol=1e-14)
    assert np.isclose(k5[1, 2], 3 * E * J / L ** 2, rtol=1e-14)
    assert np.isclose(k5[2, 2], 3 * E * J / L, rtol=1e-14)
    assert np.all(k5[5, :] == 0) and np.all(k5[:, 5] == 0)


def test_unpinned_stiffness_is_textbook(oracle):
    E, A, J = 2e11, 12e-4, 120e-6
    xy = [1.0, 2.0, 4.0, 6.0]  # L = 5
    k = oracle.stiff_beam2d(xy, A, J, E, (F,) * 6)
    L = 5.0
    assert k[0, 0] == E * A / L and k[0, 3] == -(E * A / L)
    assert k[1, 1] == 12.0 * E * J / (L * L * L) and k[1, 2] == 6.0 * E * J / (L * L)
    assert k[2, 5] == 2.0 * E * J / L and k[2, 2] == 2 * (2.0 * E * J / L)
    assert np.array_equal(k, k.T)
    t = oracle.trans_beam2d(xy)
    assert t[0, 0] == 3.0 / 5.0 and t[0, 1] == 4.0 / 5.0 and t[1, 0] == -(4.0 / 5.0) and t[2, 2] == 1.0
    kg = oracle.rotate_to_global(k, t)
    assert np.allclose(kg, t.T @ k @ t, rtol=1e-14, atol=1e-6)


# ------------------------------------------------------------------ third-party stand-ins vs LAPACK
def test_eig_general_matches_lapack(oracle):
    import scipy.linalg as sla

    rng = np.random.default_rng(7)
    for n in (1, 2, 5, 12, 40):
        a = rng.standard_normal((n, n))
        rc, wr, wi, v = oracle.eig_general(a)
        assert rc == 0
        w = wr + 1j * wi
        wl = sla.eigvals(a)
        assert np.allclose(np.sort_complex(w), np.sort_complex(wl), rtol=1e-10, atol=1e-10)
        for j in range(n):
            if wi[j] == 0:
                r = a @ v[:, j] - wr[j] * v[:, j]
                assert np.linalg.norm(r) <= 1e-10 * max(1.0, np.linalg.norm(a))
                assert abs(np.linalg.norm(v[:, j]) - 1) < 1e-12


def test_eig_isolated_zeros_are_exact(oracle):
    """The dof-|S| structurally zero columns of H = K^-1 M must come out as exact zeros
    (model.go:818 drops eigenvalues == 0); balancing's permutation guarantees it."""
    rng = np.random.default_rng(3)
    n, s = 15, 4
    h = np.zeros((n, n))
    cols = [2, 5, 6, 11]
    b = rng.standard_normal((n, n))
    spd = b @ b.T + n * np.eye(n)
    h[:, cols] = np.linalg.inv(spd)[:, cols] * rng.uniform(1, 2, s)
    h[[0, 9], :] = 0.0  # eliminated rows
    rc, wr, wi, v = oracle.eig_general(h)
    assert rc == 0
    assert np.sum(wr == 0.0) == n - s and np.all(wi == 0)


def test_static_matches_dense_lapack_solve(oracle):
    """sparse LU restatement vs numpy (LAPACK dgesv) on a well-conditioned frame."""
    N = 8
    m = frames.uniform_frame(N)
    fm = oracle.flatten_model(m)
    fm = fm.with_sections(frames.uniform_frame_sections(N, 5, 1)[0])
    loads = frames.uniform_frame_loads(N, 5, 1)[0]
    rc, disp, forces, react, _ = oracle.linear_static(fm, loads)
    assert rc == 0
    a = oracle.assemble(fm)
    rp, ci, val = a["csr"]
    K = np.zeros((fm.dof, fm.dof))
    for i in range(fm.dof):
        K[i, ci[rp[i]:rp[i + 1]]] = val[rp[i]:rp[i + 1]]
    assert np.array_equal(K, K.T)
    _, elim, ign = oracle.eliminated(fm)
    free = np.setdiff1d(np.arange(fm.dof), elim)
    d = np.zeros((20, fm.dof))
    d[:, free] = np.linalg.solve(K[np.ix_(free, free)], loads[:, free].T).T
    assert np.allclose(disp.reshape(20, -1), d, rtol=1e-11, atol=1e-18)
    R = d @ K.T - loads
    sup = np.array(fm.supports, dtype=bool).reshape(-1)
    assert np.allclose(react.reshape(20, -1)[:, sup], R[:, sup], rtol=1e-9, atol=1e-6)
    assert np.all(react.reshape(20, -1)[:, ~sup] == 0)
    # global equilibrium: reactions balance the applied loads
    assert np.allclose(react.reshape(20, -1, 3)[:, :, :2].sum(axis=1) + loads.reshape(20, -1, 3)[:, :, :2].sum(axis=1),
                       0, atol=1e-6)


def test_literal_and_row_list_reactions_are_bit_identical(oracle):
    m = frames.uniform_frame(4)
    fm = oracle.flatten_model(m)
    loads = frames.uniform_frame_loads(4, 0, 1)[0]
    try:
        oracle.set_literal_scans(True)
        a = oracle.linear_static(fm, loads)
        oracle.set_literal_scans(False)
        b = oracle.linear_static(fm, loads)
    finally:
        oracle.set_literal_scans(True)
    assert np.array_equal(a[3], b[3])


# ------------------------------------------------------------------ structural anchors
@pytest.mark.parametrize("ic,nnz", [(8, 140), (64, 980), (128, 1940)])
def test_bench_cantilever_pattern_anchor(oracle, ic, nnz):
    m, lcs, mcs = frames.bench_cantilever(ic)
    assert len(m.Points) == ic + 2 and len(m.Beams) == ic + 1
    fm = oracle.flatten_model(m)
    a = oracle.assemble(fm)
    assert a["dof"] == 3 * (ic + 2) and a["nnz"] == nnz and len(a["ignore"]) == 0
    # aliasing of BenchmarkRun's load cases (mod/split_test.go:105-109)
    fy = [sum(l.Forces[1] for l in lc.LoadNodes) for lc in lcs]
    assert len(lcs) == 20 and fy[1] == 2.3 and all(abs(v - 155.3) < 1e-12 for i, v in enumerate(fy) if i != 1)


@pytest.mark.parametrize("N,nnz,nff", [(8, 278, 143), (64, 2070, 1095), (128, 4118, 2183)])
def test_uniform_frame_pattern_anchor(oracle, N, nnz, nff):
    m = frames.uniform_frame(N)
    fm = oracle.flatten_model(m)
    a = oracle.assemble(fm)
    assert (fm.nP, fm.nB, a["dof"]) == (2 * N + 2, 2 * N + 1, 6 * N + 6)
    assert a["nnz"] == nnz and len(a["ignore"]) == 0
    _, elim, _ = oracle.eliminated(fm)
    assert len(elim) == 3 * N + 3
    rp, ci, _ = a["csr"]
    free = np.ones(a["dof"], dtype=bool)
    free[elim] = False
    rows = np.repeat(np.arange(a["dof"]), np.diff(rp))
    assert int(np.sum(free[rows] & free[ci])) == nff


def test_ignore_lists_of_examples(oracle):
    exp = {"truss": [2, 5, 8, 11, 14], "truss-modal": [2, 6, 8], "model.String": [5]}
    for name, ign in exp.items():
        m, _, _ = fixtures.GOLDEN[name]()
        a = oracle.assemble(oracle.flatten_model(m))
        assert list(a["ignore"]) == ign, name


def test_cantilever_tip_anchor(oracle):
    m, lcs, mcs = frames.bench_cantilever(8)
    err, static, modal = oracle.run(m, lcs, mcs)
    assert err is None
    tip = static[0]["disp"][1]
    assert abs(tip[1] - 155.3 / (3 * 2e11 * 120e-6)) < 1e-11 * abs(tip[1])
    assert abs(tip[0] - 10 / (2e11 * 12e-4)) < 1e-12 * abs(tip[0])
    assert len(modal[0]["hz"]) == 10
    assert abs(modal[0]["hz"][0] - 42.16108917705785) < 1e-8


# ------------------------------------------------------------------ TestSplit
def _as_cases(static):
    return [LoadCase(PointDisplacementGlobal=s["disp"], Reactions=s["react"]) for s in static]


@pytest.mark.parametrize("mi", range(len(fixtures.SPLIT_MODELS)))
def test_split(oracle, mi):
    m, lcs, mcs = fixtures.SPLIT_MODELS[mi]()
    err, static, modal = oracle.run(m, lcs, mcs)
    assert err is None
    hz = modal[0]["hz"][0]
    nP0 = len(m.Points)
    for i in range(1, 10):
        ml, lcl, mcl = fixtures.SPLIT_MODELS[mi]()
        for j in range(len(ml.Beams)):
            frames.split_beam(ml, j, i)
        err, st, mo = oracle.run(ml, lcl, mcl)
        assert err is None
        # mod.Compare iterates the expected (unsplit) rows: new points are appended after them
        exp = _as_cases(static)
        act = [LoadCase(PointDisplacementGlobal=s["disp"][:nP0], Reactions=s["react"][:nP0]) for s in st]
        assert frames.compare(exp, act, 1e-9) == []
        h = mo[0]["hz"][0]
        assert abs((hz - h) / h) <= 1e-9


# ------------------------------------------------------------------ TestRotateBeam*
def _rot_model(angle):
    L = 2.200
    x = L * math.cos(math.pi / 180.0 * angle)
    y = L * math.sin(math.pi / 180.0 * angle)
    s = (12e-4, 120e-6, 2.0e11)
    return Model(Points=[[0.0, 0.0], [x / 2, y / 2], [x, y]],
                 Beams=[BeamProp([0, 1], *s), BeamProp([1, 2], *s)],
                 Supports=[[T, T, T], [F, F, F], [F, F, F]])


def test_rotate_beam_linear(oracle):
    Fm = 100.0
    for angle in np.arange(0.0, 361.0, 5.0):
        m = _rot_model(angle)
        lc = LoadCase(LoadNodes=[LoadNode(2, [Fm * math.cos(math.pi / 180 * angle), Fm * math.sin(math.pi / 180 * angle), 0.0])])
        err, st, _ = oracle.run(m, [lc], [])
        assert err is None
        bf, re = st[0]["forces"], st[0]["react"]
        for b, o in ((0, 0), (0, 3), (1, 0), (1, 3)):
            assert abs(math.hypot(bf[b][o], bf[b][o + 1]) - Fm) / Fm <= 1e-6
        assert abs(math.hypot(re[0][0], re[0][1]) - Fm) / Fm <= 1e-6


def test_rotate_beam_modal(oracle):
    hz = None
    for angle in np.arange(0.0, 361.0, 5.0):
        m = _rot_model(angle)
        mc = ModalCase(ModalMasses=[ModalMass(1, 50.0), ModalMass(2, 100.0)])
        err, _, mo = oracle.run(m, [], [mc])
        assert err is None
        if hz is None:
            hz = mo[0]["hz"][:2]
            continue
        for i in range(2):
            assert abs((hz[i] - mo[0]["hz"][i]) / hz[i]) <= 1e-6


def test_direction_load_node(oracle):
    m, _, _ = fixtures.console_beam()
    lcs = [LoadCase(LoadNodes=[LoadNode(1, f)]) for f in ([100.0, 0, 0], [0, 100.0, 0], [0, 0, 100.0])]
    err, st, _ = oracle.run(m, lcs, [])
    assert err is None
    assert st[0]["disp"][1][0] > 0 and st[1]["disp"][1][1] > 0 and st[2]["disp"][1][1] > 0


# ------------------------------------------------------------------ negative tests
def fail_models():
    S = (12e-4, 120e-6, 2.0e11)
    ms = []
    ms.append((Model(Points=[[-1.7976931348623157e308 - 1, NAN], [0.0, 0.0], [0.0, INF]],
                     Beams=[BeamProp([-1, 1], -12e-4, -120e-6, -2.0e11), BeamProp([1, 1], *S), BeamProp([1, 20], 12e-4, NAN, -INF)],
                     Supports=[[T, T, T], [F, F, F], [F, T, F]] + [[F, F, F]] * 6,
                     Pins=[[T, T, T, T, T, F], [T, F, F, T, F, F]]),
               [LoadCase(LoadNodes=[LoadNode(-1, [0, 2.3, 0]), LoadNode(5, [INF, 0, NAN])], AmountLinearBuckling=1)],
               [ModalCase([ModalMass(7, -100)]), ModalCase([ModalMass(-1, NAN)]), ModalCase([ModalMass(0, NAN)]), ModalCase()]))
    ms.append((Model(Points=[[0, 0], [0, 1], [1, 1]], Beams=[BeamProp([0, 1], *S), BeamProp([1, 2], *S)],
                     Supports=[[F, F, F]] * 3),
               [LoadCase(LoadNodes=[LoadNode(1, [0, 2.3, 0])])], [ModalCase([ModalMass(1, 100)])]))
    ms.append((Model(Points=[[0, 0], [-1e300, 1e300], [1e300, 1e300]],
                     Beams=[BeamProp([0, 1], 12e-300, 120e-300, 2.0e300), BeamProp([1, 2], 12e+300, 120e+300, 2.0e300)],
                     Supports=[[T, T, T], [F, F, F], [F, F, F]]),
               [LoadCase(LoadNodes=[LoadNode(1, [0, 2.3, 0])])], [ModalCase([ModalMass(1, 100)])]))
    ms.append((Model(Points=[[0, 0], [0, 1], [1, 1]],
                     Beams=[BeamProp([0, 1], 12e-30, 120e-30, 2.0e30), BeamProp([1, 2], 12e-30, 120e-30, 2.0e30)],
                     Supports=[[T, T, T], [F, F, F], [F, F, F]]),
               [LoadCase(LoadNodes=[LoadNode(2, [1e300, 1e308, 1e300]), LoadNode(2, [1e300, 1e308, 1e300])])],
               [ModalCase([ModalMass(1, 1e300)])]))
    ms.append((Model(Points=[[0, 0], [0, 1], [1, 1]],
                     Beams=[BeamProp([0, 1], 12e-30, 5e-324, 5e-324), BeamProp([1, 2], 12e-30, 120e-30, 2.0e30)],
                     Pins=[[F] * 6, [F, F, T, F, F, T], [F] * 6],
                     Supports=[[T, T, T], [F, F, F], [F, F, F]]),
               [LoadCase(LoadNodes=[LoadNode(20, [0, 1, 0]), LoadNode(-20, [0, 1, 0])])],
               [ModalCase([ModalMass(1, 1), ModalMass(-1, -1), ModalMass(10000, 0)])]))
    ms.append((Model(Points=[[0, 0], [0, 5e-324], [1, 5e-324]],
                     Beams=[BeamProp([0, 1], 1e200, 1e200, 1e200), BeamProp([1, 2], 1e200, 1e200, 1e200)],
                     Supports=[[T, T, T], [F, F, F], [F, F, F]]),
               [LoadCase(LoadNodes=[LoadNode(2, [1e20, 1e20, 0])])],
               [ModalCase([ModalMass(1, 1e20), ModalMass(-10, 1e20), ModalMass(10, 1e20)])]))
    ms.append((Model(Points=[[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 0]],
                     Beams=[BeamProp([0, 1], 1e200, 1e200, 1e200), BeamProp([1, 2], 1e200, 1e200, 1e200),
                            BeamProp([3, 4], 1e200, 1e200, 1e200), BeamProp([4, 5], 1e200, 1e200, 1e200)],
                     Supports=[[T, T, T], [F, F, F], [F, F, F], [T, T, T], [F, F, F], [F, F, F]]),
               [LoadCase(LoadNodes=[LoadNode(2, [1e20, 1e20, 0]), LoadNode(5, [1e20, 1e20, 0])])],
               [ModalCase([ModalMass(1, 1)])]))
    return ms


@pytest.mark.parametrize("i", range(7))
def test_model_fail(oracle, i):
    m, lcs, mcs = fail_models()[i]
    err, _, _ = oracle.run(m, lcs, mcs)
    assert err is not None


def test_wrong_load(oracle):
    m, lcs, _ = fixtures.truss()
    lcs[0].LoadNodes[0].Forces[2] = 42.0  # moment on a fully pinned node: an `ignore` DOF
    err, st, _ = oracle.run(m, [lcs[0]], [])
    assert err == [oracle.ERR_IGNORE_LOAD] and st[0] is None


# ------------------------------------------------------------------ load generator
def test_uniform_frame_loads_vectorised_equals_load_uniform(oracle):
    N = 5
    m = frames.uniform_frame(N)
    dense = frames.uniform_frame_loads(N, 11, 2)
    for mi in range(2):
        lcs = frames.uniform_frame_load_cases(m, N, 11 + mi)
        for k, lc in enumerate(lcs):
            assert np.array_equal(oracle.node_loads(len(m.Points), lc.LoadNodes), dense[mi, k])
    # shards reproduce the batch
    assert np.array_equal(frames.uniform_frame_loads(N, 12, 1)[0], dense[1])
    assert np.array_equal(frames.uniform_frame_sections(N, 12, 1)[0], frames.uniform_frame_sections(N, 11, 2)[1])


def test_load_uniform_quirk_and_projection():
    m = Model(Points=[[0, 0], [0, 3], [4, 3]], Beams=[BeamProp([0, 1], 1, 1, 1), BeamProp([1, 2], 1, 1, 1)],
              Supports=[[T, T, T], [F] * 3, [F] * 3])
    v = frames.load_uniform(m, 0, True, [5.0, -7.0])  # vertical member: dx = 0 => zero fixed-end moments
    assert [l.Forces for l in v[:2]] == [[0.0, 0.0, -0.0], [0.0, 0.0, 0.0]]
    assert v[2].Forces == [5.0 * 3 / 2.0, 0.0, 0.0] and v[3].Forces == [0.0, -0.0, 0.0]
    h = frames.load_uniform(m, 1, True, [0.0, -7.0])
    assert h[0].Forces[2] == (-7.0 * 16) / 12.0 and h[1].Forces[2] == -(-7.0 * 16) / 12.0
    assert h[3].Forces[1] == (-7.0 * 4) / 2.0
    with pytest.raises(ValueError):
        frames.load_uniform(m, 5, True, [0, 0])
    with pytest.raises(ValueError):
        frames.split_beam(m, 0, 0)
