"""CPU-side tests of the C-ABI boundary: the library loads, exports every symbol include/hdb200.h
declares, the host symbolic phase reproduces the oracle's integer structures bit-exactly (DOF
numbering, eliminated set, ignore list, CSR pattern, scatter map, permutation), validation accepts /
rejects what check.go does, and compute entry points refuse to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fixtures
from hd_b200 import _lib, frames
from test_oracle_properties import fail_models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def topo_of(oracle, m, **kw):
    fm = oracle.flatten_model(m)
    return fm, _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, **kw)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    with open(os.path.join(ROOT, "include", "hdb200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(hdb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.hdb_version() == 100
    assert L.hdb_error_string(_lib.ERR_NO_DEVICE).decode().startswith("no CUDA device")


MODELS = {**{k: (lambda f=f: f()[0]) for k, f in fixtures.GOLDEN.items()},
          "uniform_frame(8)": lambda: frames.uniform_frame(8),
          "uniform_frame(128)": lambda: frames.uniform_frame(128),
          "bench_cantilever(8)": lambda: frames.bench_cantilever(8)[0],
          "bench_cantilever(128)": lambda: frames.bench_cantilever(128)[0]}


@pytest.mark.parametrize("name", sorted(MODELS))
def test_symbolic_structures_bit_exact(oracle, name):
    m = MODELS[name]()
    fm, t = topo_of(oracle, m)
    a = oracle.assemble(fm)
    rc, elim, ign = oracle.eliminated(fm)
    assert rc == 0
    assert t.info.dof == a["dof"] == 3 * len(m.Points)
    assert np.array_equal(t.query(_lib.Q_IGNORE), ign)
    assert np.array_equal(t.query(_lib.Q_ELIMINATED), elim)
    rp, ci, val = a["csr"]
    assert t.info.nnzK == a["nnz"]
    assert np.array_equal(t.query(_lib.Q_CSR_ROWPTR), rp)
    assert np.array_equal(t.query(_lib.Q_CSR_COLIND), ci)
    # free numbering = rank among non-eliminated DOFs (model.go:326-328 / sparse.LU ignore list)
    fi = t.query(_lib.Q_FREE_INDEX)
    exp = np.full(a["dof"], -1, dtype=np.int32)
    free = np.setdiff1d(np.arange(a["dof"]), elim)
    exp[free] = np.arange(len(free))
    assert np.array_equal(fi, exp)
    # natural ordering: the permutation is the identity, as in the reference's LU
    assert np.array_equal(t.query(_lib.Q_PERM), np.arange(len(free)))
    # scatter map: replay the triplets of model.go:622-635 through it and recover the oracle's K
    sc = t.query(_lib.Q_SCATTER).reshape(fm.nB, 36)
    K = np.zeros(a["nnz"])
    for b in range(fm.nB):
        xy = np.concatenate([fm.points[fm.beamN[b, 0]], fm.points[fm.beamN[b, 1]]])
        k = oracle.stiff_beam2d(xy, *fm.sections[b], fm.pins[b])
        kg = oracle.rotate_to_global(k, oracle.trans_beam2d(xy))
        q = 0
        for p1 in range(2):
            for p2 in range(2):
                for r1 in range(3):
                    for r2 in range(3):
                        v = kg[p1 * 3 + r1, p2 * 3 + r2]
                        if sc[b, q] < 0:
                            assert v == 0.0
                        else:
                            assert v != 0.0
                            row, col = fm.beamN[b, p1] * 3 + r1, fm.beamN[b, p2] * 3 + r2
                            assert rp[row] <= sc[b, q] < rp[row + 1] and ci[sc[b, q]] == col
                            K[sc[b, q]] += v
                        q += 1
    assert np.array_equal(K, val)  # same values, same (beam-order) summation
    t.close()


@pytest.mark.parametrize("name,nnzL", [("uniform_frame(8)", 108), ("uniform_frame(128)", 1668),
                                        ("bench_cantilever(8)", 76), ("bench_cantilever(128)", 1156)])
def test_factor_pattern_anchor_and_fill(oracle, name, nnzL):
    """nnz(L) in natural order and the fill pattern against a dense symbolic run."""
    fm, t = topo_of(oracle, MODELS[name]())
    assert t.info.nnzL + t.info.nFree == nnzL  # the anchor counts the diagonal
    n = t.info.nFree
    lp, li = t.query(_lib.Q_L_COLPTR), t.query(_lib.Q_L_ROWIND)
    rp, ci = t.query(_lib.Q_CSR_ROWPTR), t.query(_lib.Q_CSR_COLIND)
    fi = t.query(_lib.Q_FREE_INDEX)
    if n <= 60:
        A = np.zeros((n, n), dtype=bool)
        for r in range(t.info.dof):
            for q in range(rp[r], rp[r + 1]):
                if fi[r] >= 0 and fi[ci[q]] >= 0:
                    A[fi[r], fi[ci[q]]] = True
        A |= A.T
        for j in range(n):  # boolean right-looking elimination
            rows = np.nonzero(A[j + 1:, j])[0] + j + 1
            A[np.ix_(rows, rows)] = True
        for j in range(n):
            assert np.array_equal(np.nonzero(A[j + 1:, j])[0] + j + 1, li[lp[j]:lp[j + 1]])
    t.close()


def test_rcm_ordering_is_a_permutation_and_bounds_fill(oracle):
    m = frames.bench_cantilever(64)[0]
    fm, t0 = topo_of(oracle, m)
    _, t1 = topo_of(oracle, m, ordering=1)
    p = t1.query(_lib.Q_PERM)
    assert np.array_equal(np.sort(p), np.arange(t1.info.nFree))
    assert t1.info.nnzL <= 2 * t0.info.nnzL
    t0.close(); t1.close()


def test_automatic_ordering_picks_the_smaller_factor(oracle):
    """ordering=2: natural order on a chain (identity permutation kept), RCM on a multi-storey frame whose
    natural order fills a band as wide as a storey."""
    from hd_b200.model import BeamProp, Model
    m = frames.uniform_frame(16)
    _, t0 = topo_of(oracle, m)
    _, t2 = topo_of(oracle, m, ordering=2)
    assert t2.info.nnzL == t0.info.nnzL and np.array_equal(t2.query(_lib.Q_PERM), t0.query(_lib.Q_PERM))
    t0.close(); t2.close()
    ms, idx = Model(), {}
    N, S = 12, 3
    for s_ in range(S + 1):
        for i in range(N + 1):
            idx[(i, s_)] = len(ms.Points)
            ms.Points.append([6.0 * i, 3.0 * s_])
            ms.Supports.append([s_ == 0] * 3)
    for s_ in range(1, S + 1):
        for i in range(N + 1):
            ms.Beams.append(BeamProp(N=[idx[(i, s_ - 1)], idx[(i, s_)]], A=12e-4, J=120e-6, E=2e11))
        for i in range(1, N + 1):
            ms.Beams.append(BeamProp(N=[idx[(i - 1, s_)], idx[(i, s_)]], A=12e-4, J=120e-6, E=2e11))
    _, a0 = topo_of(oracle, ms)
    _, a1 = topo_of(oracle, ms, ordering=1)
    _, a2 = topo_of(oracle, ms, ordering=2)
    assert a1.info.nnzL < a0.info.nnzL and a2.info.nnzL == a1.info.nnzL
    assert np.array_equal(np.sort(a2.query(_lib.Q_PERM)), np.arange(a2.info.nFree))
    a0.close(); a1.close(); a2.close()


def _check_model(m):
    n_pins, n_sup = len(m.Pins), len(m.Supports)
    pts = np.array([[p[0], p[1]] for p in m.Points], dtype=np.float64).reshape(-1, 2)
    beams = _lib.beams_array([[b.N[0], b.N[1]] for b in m.Beams], [[b.A, b.J, b.E] for b in m.Beams])
    pins = np.array([list(p) for p in m.Pins], dtype=np.uint8).reshape(-1, 6)
    sup = np.array([list(s) for s in m.Supports], dtype=np.uint8).reshape(-1, 3)
    msg = C.create_string_buffer(4096)
    rc = _lib.lib().hdb_check_model(len(pts), _lib._np_ptr(pts), len(beams), _lib._np_ptr(beams), n_pins,
                                    _lib._np_ptr(pins), n_sup, _lib._np_ptr(sup), msg, 4096)
    return rc, msg.value.decode()


def test_check_model_matches_oracle_accept_reject(oracle):
    for name, f in fixtures.GOLDEN.items():
        m = f()[0]
        if not m.Pins:
            m.Pins = [[False] * 6 for _ in m.Beams]
        assert _check_model(m)[0] == 0, name
    expected_reject = {0: True, 1: False, 2: False, 3: False, 4: True, 5: False, 6: True}  # model_test.go:50-324
    for i, (m, lcs, mcs) in enumerate(fail_models()):
        if not m.Pins:
            m.Pins = [[False] * 6 for _ in m.Beams]
        rc, msg = _check_model(m)
        assert (rc != 0) == expected_reject[i], (i, msg)
        if rc != 0:
            assert msg
    # pin rules of check.go:357-391
    m = fixtures.console_beam()[0]
    for pins, ok in (([1, 0, 0, 1, 0, 0], False), ([0, 0, 1, 0, 0, 1], False), ([0, 1, 1, 0, 1, 1], True),
                     ([1, 1, 1, 0, 1, 1], False)):
        m.Pins = [[bool(x) for x in pins]]
        assert (_check_model(m)[0] == 0) == ok, pins


def test_check_loads_modal_and_node_loads(oracle):
    L = _lib.lib()
    loads = np.zeros(3, dtype=_lib.LOAD_DTYPE)
    loads["n"] = [1, 1, 0]
    loads["forces"] = [[0, 2.3, 0], [10, 0, 0], [1, 1, 1]]
    assert L.hdb_check_loads(2, 3, _lib._np_ptr(loads), None, 0) == 0
    p = np.zeros(6)
    assert L.hdb_assemble_node_loads(2, 3, _lib._np_ptr(loads), _lib._np_ptr(p)) == 0
    assert np.array_equal(p, [1, 1, 1, 10, 2.3, 0])
    loads["n"][2] = 2
    assert L.hdb_check_loads(2, 3, _lib._np_ptr(loads), None, 0) == _lib.ERR_INPUT
    loads["n"][2] = 0
    loads["forces"][0][1] = np.nan
    assert L.hdb_check_loads(2, 3, _lib._np_ptr(loads), None, 0) == _lib.ERR_INPUT
    sup = np.array([[1, 1, 1], [0, 0, 0]], dtype=np.uint8)
    mass = np.zeros(1, dtype=_lib.MASS_DTYPE)
    for n, w, ok in ((1, 100.0, True), (0, 100.0, False), (1, 0.0, False), (1, -1.0, False), (5, 1.0, False),
                     (1, np.inf, False)):
        mass["n"], mass["mass"] = n, w
        assert (L.hdb_check_modal(2, 2, _lib._np_ptr(sup), 1, _lib._np_ptr(mass), None, 0) == 0) == ok, (n, w)


def test_topology_create_rejects_bad_models(oracle):
    m = fail_models()[6][0]  # two disconnected structures
    fm = oracle.flatten_model(m)
    with pytest.raises(_lib.HdbError) as e:
        _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)
    assert e.value.code == _lib.ERR_INPUT and "single calculation model" in str(e.value)


def test_modal_mode_count(oracle):
    m, _, mcs = fixtures.truss()
    fm, t = topo_of(oracle, m)
    assert t.modal_mode_count([q.N for q in mcs[0].ModalMasses]) == 2
    m2 = frames.uniform_frame(64)
    _, t2 = topo_of(oracle, m2)
    massN, _ = frames.uniform_frame_masses(64, 0, 1)
    assert t2.modal_mode_count(massN) == 130
    t.close(); t2.close()


def test_compute_refuses_without_gpu(oracle):
    """No CPU fallback: on a box without a CUDA device every compute entry errors out."""
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    fm, t = topo_of(oracle, frames.uniform_frame(2))
    with pytest.raises(_lib.HdbError) as e:
        t.static_batch(fm.sections[None], np.zeros((1, 1, fm.dof)))
    assert e.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(_lib.HdbError):
        t.assemble_batch(fm.sections[None])
    t.close()


def test_go_sources_bind_only_exported_symbols():
    """The Go front cannot be compiled here (no toolchain): at least every C function it names must be one the header
    declares and the library exports, and every exported entry point of the reference must be present in go/hd."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "hdb200.h")).read()
    for rel in ("go/hdb200/hdb200.go", "go/hd/hd.go"):
        src = open(os.path.join(root, rel)).read()
        for name in set(re.findall(r"C\.(hdb_[a-z_0-9]+)\(", src)):
            assert name in _lib.SYMBOLS and re.search(r"\b%s\(" % name, header), (rel, name)
        assert src.count("{") == src.count("}") and src.count("(") == src.count(")"), rel
    front = open(os.path.join(root, "go/hd/hd.go")).read()
    for sig in ("func Run(out io.Writer, m *Model, lcs []LoadCase, mcs []ModalCase) error",
                "func LinearStatic(out io.Writer, m *Model, lcs ...*LoadCase) (err error)",
                "func Modal(out io.Writer, m *Model, mc *ModalCase) (err error)", "func RunBatch(", "func RunBatchMulti("):
        assert sig in front, sig
