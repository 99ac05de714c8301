"""Throughput on a topology whose factor is not a chain: S-storey, N-bay frame (generic op runs, fill blocks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hd_b200 import _lib
from hd_b200.model import BeamProp, Model
from oracle import hd_oracle as oracle

def frame(N, S):
    m = Model()
    idx = {}
    for s in range(S + 1):
        for i in range(N + 1):
            idx[(i, s)] = len(m.Points)
            m.Points.append([6.0 * i, 3.0 * s])
            m.Supports.append([s == 0] * 3)
    for s in range(1, S + 1):
        for i in range(N + 1):
            m.Beams.append(BeamProp(N=[idx[(i, s - 1)], idx[(i, s)]], A=12e-4, J=120e-6, E=2e11))
        for i in range(1, N + 1):
            m.Beams.append(BeamProp(N=[idx[(i - 1, s)], idx[(i, s)]], A=12e-4, J=120e-6, E=2e11))
    return m

N, S, M, nLC = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 20
m = frame(N, S)
fm = oracle.flatten_model(m)
for ordering in (2,):
    t = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, ordering=ordering)
    info = t.info
    rng = np.random.default_rng(1)
    sec = np.tile(fm.sections, (M, 1, 1)); sec[:, :, :2] *= rng.uniform(0.5, 2.0, size=(M, fm.nB, 2))
    loads = rng.normal(size=(M, nLC, fm.dof)) * 1e3
    sup = np.repeat(np.asarray(fm.supports, bool).reshape(-1), 1)
    dev = torch.device("cuda:0")
    ds, dl = torch.from_numpy(sec).to(dev), torch.from_numpy(loads).to(dev)
    dd = torch.empty((M, nLC, fm.nP, 3), dtype=torch.float64, device=dev)
    df = torch.empty((M, nLC, fm.nB, 6), dtype=torch.float64, device=dev)
    dr = torch.empty((M, nLC, fm.nP, 3), dtype=torch.float64, device=dev)
    st = torch.zeros((M,), dtype=torch.int32, device=dev); fc = torch.zeros((M,), dtype=torch.int32, device=dev)
    for _ in range(2):
        t.set_profiling(1)
        t.static_batch_dev(M, ds, nLC, dl, dd, df, dr, st, fc, 0)
        torch.cuda.synchronize()
    ms = t.last_kernel_ms()
    byt = M * (fm.nB * 3 + nLC * fm.dof + nLC * (2 * fm.dof + fm.nB * 6)) * 8
    print(f"{N} bays x {S} storeys, ordering {ordering}: dof {fm.dof} nB {fm.nB} nnzL {info.nnzL} status {int(st.abs().sum())} "
          f"kernel ms {[round(x, 3) for x in ms]} -> {M / sum(ms) * 1e3:.0f} solves/s, solve kernel {byt / ms[2] / 1e6:.0f} GB/s")
    t.close()
