"""mod.LoadUniform + assemblyNodeLoad on the device (hdb_uniform_loads_dev) against the host mirror of
mod/uniform.go (hd_b200.frames.load_uniform) flattened by the oracle's node_loads: bit-identical."""
import numpy as np
import pytest
import torch

from hd_b200 import _lib, frames
from hd_b200.model import BeamProp, LoadNode, Model

pytestmark = pytest.mark.gpu


def _topo(oracle, m):
    fm = oracle.flatten_model(m)
    return fm, _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)


def test_uniform_frame_loads_bit_identical(oracle):
    N, nM, nLC = 16, 5, 20
    m = frames.uniform_frame(N)
    fm, t = _topo(oracle, m)
    q, fx = frames.uniform_frame_load_params(N, 3, nM, nLC)
    qq = np.zeros((nM, nLC, N, 2))
    qq[..., 1] = q
    ff = np.zeros((nM, nLC, 1, 3))
    ff[..., 0, 0] = fx
    dev = torch.device("cuda:0")
    d_q, d_f = torch.from_numpy(qq).to(dev), torch.from_numpy(ff).to(dev)
    d_loads = torch.empty((nM, nLC, fm.dof), dtype=torch.float64, device=dev)
    ubeam = [frames.girder_index(i) for i in range(1, N + 1)]
    t.uniform_loads_dev(nM, nLC, ubeam, [1] * N, d_q, [1], d_f, d_loads)
    torch.cuda.synchronize()
    want = frames.uniform_frame_loads(N, 3, nM, nLC)
    assert np.array_equal(d_loads.cpu().numpy(), want)
    # and the loads feed the solver unchanged
    sec = frames.uniform_frame_sections(N, 3, nM)
    st, fail, disp, forces, react = t.static_batch(sec, d_loads.cpu().numpy())
    st2, fail2, disp2, forces2, react2 = t.static_batch(sec, want)
    assert not st.any() and np.array_equal(disp, disp2)
    t.close()


def test_general_members_match_host_mirror(oracle):
    """Inclined, reversed and vertical members, both projection modes, both intensity components, a node that
    collects several loads, extra node loads: every addition in the reference's order."""
    m = Model()
    m.Points = [[0.0, 0.0], [4.0, 3.0], [4.0, 0.0], [0.0, 5.0], [-2.0, 1.0]]
    m.Supports = [[True, True, True]] + [[False] * 3] * 4
    ends = [(0, 1), (1, 2), (2, 0), (3, 0), (1, 3), (4, 1), (2, 4)]
    m.Beams = [BeamProp(N=list(e), A=1e-3, J=1e-6, E=2e11) for e in ends]
    fm, t = _topo(oracle, m)
    rng = np.random.default_rng(11)
    ubeam = [0, 1, 2, 3, 4, 5, 6, 0, 4]
    uproj = [1, 0, 1, 0, 1, 0, 1, 0, 0]
    nodes = [1, 3, 1]
    nM, nLC = 3, 4
    q = rng.normal(size=(nM, nLC, len(ubeam), 2)) * 1e3
    f = rng.normal(size=(nM, nLC, len(nodes), 3)) * 1e2
    dev = torch.device("cuda:0")
    d_loads = torch.empty((nM, nLC, fm.dof), dtype=torch.float64, device=dev)
    t.uniform_loads_dev(nM, nLC, ubeam, uproj, torch.from_numpy(q).to(dev), nodes, torch.from_numpy(f).to(dev), d_loads)
    torch.cuda.synchronize()
    got = d_loads.cpu().numpy()
    for i in range(nM):
        for c in range(nLC):
            ln = []
            for u, (b, pr) in enumerate(zip(ubeam, uproj)):
                ln.extend(frames.load_uniform(m, b, bool(pr), [float(q[i, c, u, 0]), float(q[i, c, u, 1])]))
            for j, n in enumerate(nodes):
                ln.append(LoadNode(n, [float(x) for x in f[i, c, j]]))
            want = oracle.node_loads(fm.nP, ln)
            assert np.array_equal(got[i, c], want), (i, c, np.abs(got[i, c] - want).max())
    t.close()
