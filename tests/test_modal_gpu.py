"""GPU parity tests of the batched Modal path (model.go:719-851) through the C ABI: the reference's
goldens, the oracle (non-symmetric Dgeev-style solve of K^-1 M) at 1e-9 on frequencies, shapes up to the
implementation-defined sign, property tests (TestRotateBeamModal, TestSplit's frequency check)."""
import math

import numpy as np
import pytest

import fixtures
from hd_b200 import _lib, api, frames
from hd_b200.model import ModalCase, ModalMass
from test_oracle_golden import GOLD, close_up_to_sign

pytestmark = pytest.mark.gpu


def topo_of(oracle, m, **kw):
    fm = oracle.flatten_model(m)
    return fm, _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, **kw)


@pytest.mark.parametrize("name", sorted(k for k, v in GOLD.items() if v["modal"]))
def test_golden_modal_through_run(oracle, name):
    m, lcs, mcs = fixtures.GOLDEN[name]()
    for mi, mc in enumerate(mcs):
        err = api.Modal(None, m, mc)
        assert err is None, err
        gm = GOLD[name]["modal"][mi]
        assert len(mc.Result) == len(gm), (name, [r.Hz for r in mc.Result])
        for k, gk in enumerate(gm):
            assert abs(mc.Result[k].Hz - gk["hz"]) <= 0.51e-5 + 1e-9 * gk["hz"], (name, k, mc.Result[k].Hz, gk["hz"])
        # shapes: the goldens pin Dgeev's output; inside a cluster of (nearly) equal frequencies only the
        # subspace is defined, elsewhere the vector up to sign
        hz = np.array([g["hz"] for g in gm])
        for k, gk in enumerate(gm):
            others = np.delete(hz, k)
            if others.size and np.min(np.abs(others - hz[k])) < 1e-3 * hz[k]:
                continue
            close_up_to_sign(mc.Result[k].ModalDisplacement, gk["shape"], f"{name} modal {mi} shape {k}")


@pytest.mark.parametrize("N,nM", [(4, 5), (8, 9), (64, 3)])
def test_modal_uniform_frame_vs_oracle(oracle, N, nM):
    fm, t = topo_of(oracle, frames.uniform_frame(N))
    secs = frames.uniform_frame_sections(N, 40, nM)
    massN, mass = frames.uniform_frame_masses(N, 40, nM)
    status, nModes, hz, shapes = t.modal_batch(secs, massN, mass)
    assert not status.any()
    assert np.all(nModes == 2 * (N + 1))
    ost, on, ohz, oshapes = oracle.modal_batch(fm, secs, massN, mass)
    assert not ost.any() and np.array_equal(on, nModes)
    s = 2 * (N + 1)
    rel = np.abs(hz[:, :s] - ohz[:, :s]) / ohz[:, :s]
    assert rel.max() <= 1e-9, rel.max()
    assert np.all(np.diff(hz[:, :s], axis=1) >= 0)
    for mi in range(nM):
        for k in range(s):
            a, b = shapes[mi, k].reshape(-1), oshapes[mi, k].reshape(-1)
            assert abs(np.linalg.norm(a) - 1) < 1e-12
            gap = min(abs(ohz[mi, k] - ohz[mi, j]) for j in range(s) if j != k) / ohz[mi, k]
            if gap < 1e-6:
                continue  # (nearly) repeated frequency: only the invariant subspace is defined
            assert 1 - abs(a @ b) <= 1e-9 / gap + 1e-12, (mi, k, 1 - abs(a @ b), gap)
    # supported DOFs carry exact zeros
    sup = np.array(fm.supports, dtype=bool)
    assert np.all(shapes[:, :, sup] == 0)
    t.close()


def test_modal_bench_cantilever_vs_oracle(oracle):
    """the reference's benchmark model (mod/split_test.go:110-113): masses on the tip and every even point."""
    m, _, mcs = frames.bench_cantilever(8)
    fm, t = topo_of(oracle, m)
    massN = [q.N for q in mcs[0].ModalMasses]
    mass = [q.Mass for q in mcs[0].ModalMasses]
    status, nModes, hz, shapes = t.modal_batch(fm.sections[None], massN, np.array(mass)[None])
    rc, ohz, oshapes = oracle.modal(fm, massN, mass)
    assert rc == 0 and status[0] == 0 and nModes[0] == len(ohz) == 10
    assert (np.abs(hz[0, :10] - ohz) / ohz).max() <= 1e-9
    t.close()


def test_rotate_beam_modal_on_gpu(oracle):
    from test_oracle_properties import _rot_model
    hz = None
    for angle in np.arange(0.0, 361.0, 45.0):
        m = _rot_model(angle)
        mc = ModalCase(ModalMasses=[ModalMass(1, 50.0), ModalMass(2, 100.0)])
        assert api.Modal(None, m, mc) is None
        assert len(mc.Result) == 4
        if hz is None:
            hz = [mc.Result[0].Hz, mc.Result[1].Hz]
            continue
        for i in range(2):
            assert abs((hz[i] - mc.Result[i].Hz) / hz[i]) <= 1e-6


def test_split_keeps_first_frequency_on_gpu(oracle):
    for ctor in (fixtures.console_beam, fixtures.modal_beam, fixtures.modal_truss):
        m, _, mcs = ctor()
        assert api.Modal(None, m, mcs[0]) is None
        hz = mcs[0].Result[0].Hz
        for i in (1, 5):
            ml, _, mcl = ctor()
            for j in range(len(ml.Beams)):
                frames.split_beam(ml, j, i)
            assert api.Modal(None, ml, mcl[0]) is None
            h = mcl[0].Result[0].Hz
            assert abs((hz - h) / h) <= 1e-9, (ctor.__name__, i)


def test_modal_edge_cases(oracle):
    m, _, _ = fixtures.console_beam()
    mc = ModalCase()  # empty mass list: no results, no error (model.go:772-805)
    assert api.Modal(None, m, mc) is None and mc.Result == []
    mc = ModalCase(ModalMasses=[ModalMass(0, 100.0)])  # mass on a support: check.go:437-446
    assert api.Modal(None, m, mc) is not None
    mc = ModalCase(ModalMasses=[ModalMass(1, 100.0), ModalMass(1, 50.0)])  # masses on one point accumulate
    assert api.Modal(None, m, mc) is None
    mc2 = ModalCase(ModalMasses=[ModalMass(1, 150.0)])
    assert api.Modal(None, m, mc2) is None
    assert abs(mc.Result[0].Hz - mc2.Result[0].Hz) <= 1e-12 * mc2.Result[0].Hz
    # Run aggregates static + modal like the reference
    m, lcs, mcs = fixtures.g_beam()
    assert api.Run(None, m, lcs, mcs) is None
    assert lcs[0].Reactions is not None and len(mcs[0].Result) == 2


def test_modal_zero_mass_is_not_a_mode(oracle):
    """A listed mass of exactly 0 N gives exactly zero eigenvalues of K^-1 M; the reference skips them
    (model.go:818-834) instead of reporting an infinite frequency."""
    N, nM = 4, 3
    fm, t = topo_of(oracle, frames.uniform_frame(N))
    secs = frames.uniform_frame_sections(N, 7, nM)
    massN, mass = frames.uniform_frame_masses(N, 7, nM)
    mass = mass.copy()
    mass[1, 2] = 0.0
    status, nModes, hz, shapes = t.modal_batch(secs, massN, mass)
    assert not status.any()
    ost, on, ohz, oshapes = oracle.modal_batch(fm, secs, massN, mass)
    assert not ost.any() and np.array_equal(on, nModes), (on, nModes)
    assert nModes[1] < nModes[0] == nModes[2]
    for mi in range(nM):
        k = int(nModes[mi])
        assert np.all(np.isfinite(hz[mi, :k]))
        rel = np.abs(hz[mi, :k] - ohz[mi, :k]) / ohz[mi, :k]
        assert rel.max() <= 1e-9, (mi, rel.max())
    t.close()


def test_modal_single_kernel_path_is_reproducible(oracle, monkeypatch):
    """The one-kernel fallback (HDB_MODAL_SINGLE=1): same frequencies as the split pipeline at 1e-12, and its shapes are
    bit-identical run to run (norms summed in a fixed order, no floating-point atomics)."""
    N, nM = 8, 4
    fm, t = topo_of(oracle, frames.uniform_frame(N))
    secs = frames.uniform_frame_sections(N, 3, nM)
    massN, mass = frames.uniform_frame_masses(N, 3, nM)
    _, n0, hz0, sh0 = t.modal_batch(secs, massN, mass)
    monkeypatch.setenv("HDB_MODAL_SINGLE", "1")
    _, n1, hz1, sh1 = t.modal_batch(secs, massN, mass)
    _, n2, hz2, sh2 = t.modal_batch(secs, massN, mass)
    assert np.array_equal(n0, n1) and np.array_equal(n1, n2)
    assert np.array_equal(hz1, hz2) and np.array_equal(sh1, sh2)
    assert np.max(np.abs(hz1 - hz0) / hz0) <= 1e-12
    t.close()
