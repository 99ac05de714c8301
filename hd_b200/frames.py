"""Model helpers either side of the hot path: the reference's `mod` package and the
two benchmark model families of SURVEY.md §8(d).

* split_beam      — mod/split.go:15-72 (SplitBeam)
* load_uniform    — mod/uniform.go:46-211 (LoadUniform), quirks kept (dx used for the
                    dy-branch moments, :146-173)
* compare         — mod/compare.go:13-73 (Compare), the project's 1e-9 rule
* bench_cantilever(ic) — the model BenchmarkRun times (mod/split_test.go:93-124) with its
                    aliased load cases (19 of 20 identical) and modal masses
* uniform_frame(N)     — our definition of the "N-bay uniform frame" (single-storey portal)
* uniform_frame_batch  — per-model randomised sections / loads / masses, counter-based so
                    that any shard of the batch can be generated independently
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from .model import BeamProp, LoadCase, LoadNode, ModalCase, ModalMass, Model

SEED = 20261019
NOMINAL_A, NOMINAL_J, NOMINAL_E = 12e-4, 120e-6, 2.0e11  # example.go:21 (ConsoleBeam section)


# ----------------------------------------------------------------------------- mod.SplitBeam
def split_beam(m: Model, beam_index: int, amount: int) -> None:
    """mod/split.go:15-72. New points and beams are appended; pins [3:6] move to the last piece."""
    if amount < 1:
        raise ValueError(f"Not valid value of amount intermediate points : {amount}")
    if beam_index < 0 or beam_index >= len(m.Beams):
        raise ValueError(f"Not valid value of beam index : {beam_index}")
    last_point = len(m.Points)
    start = list(m.Points[m.Beams[beam_index].N[0]])
    end = list(m.Points[m.Beams[beam_index].N[1]])
    m.Points.extend([[0.0, 0.0] for _ in range(amount)])
    for j in range(2):
        delta_ = (end[j] - start[j]) / float(amount + 1)
        for i in range(amount):
            if start[j] == end[j]:
                m.Points[last_point + i][j] = start[j]
                continue
            m.Points[last_point + i][j] = start[j] + float(i + 1) * delta_
    last_beam = len(m.Beams)
    src = m.Beams[beam_index]
    for i in range(amount):
        m.Beams.append(BeamProp(N=[last_point + i, last_point + i + 1], A=src.A, J=src.J, E=src.E))
    m.Beams[-1].N[1] = src.N[1]
    src.N[1] = last_point
    m.Supports.extend([[False, False, False] for _ in range(amount)])
    if len(m.Pins) != 0:
        m.Pins.extend([[False] * 6 for _ in range(amount)])
        for i in range(3, 6):
            m.Pins[len(m.Beams) - 1][i] = m.Pins[beam_index][i]
            m.Pins[beam_index][i] = False
    assert last_beam + amount == len(m.Beams)


# ----------------------------------------------------------------------------- mod.LoadUniform
def load_uniform(m: Model, beam_index: int, projection: bool, uf: Sequence[float]) -> List[LoadNode]:
    """mod/uniform.go:46-211."""
    if m is None:
        raise ValueError("Model is nil")
    if beam_index >= len(m.Beams) or beam_index < 0:
        raise ValueError("index of beam is outside of model slice")
    for node in m.Beams[beam_index].N:
        if node >= len(m.Points) or node < 0:
            raise ValueError("index of node of beam is outside slice")
    for v in uf:
        if math.isnan(v) or math.isinf(v):
            raise ValueError("load not valid")
    nodes = m.Beams[beam_index].N
    dx = m.Points[nodes[0]][0] - m.Points[nodes[1]][0]
    dy = m.Points[nodes[0]][1] - m.Points[nodes[1]][1]
    ln: List[LoadNode] = []
    if dx > 0:
        ln.append(LoadNode(nodes[0], [0.0, 0.0, -(uf[1] * dx * dx) / 12.0]))
        ln.append(LoadNode(nodes[1], [0.0, 0.0, (uf[1] * dx * dx) / 12.0]))
    elif dx < 0:
        ln.append(LoadNode(nodes[0], [0.0, 0.0, (uf[1] * dx * dx) / 12.0]))
        ln.append(LoadNode(nodes[1], [0.0, 0.0, -(uf[1] * dx * dx) / 12.0]))
    if dy > 0:  # the reference uses dx here (uniform.go:146-153)
        ln.append(LoadNode(nodes[0], [0.0, 0.0, (uf[0] * dx * dx) / 12.0]))
        ln.append(LoadNode(nodes[1], [0.0, 0.0, -(uf[0] * dx * dx) / 12.0]))
    elif dy < 0:
        ln.append(LoadNode(nodes[0], [0.0, 0.0, -(uf[0] * dx * dx) / 12.0]))
        ln.append(LoadNode(nodes[1], [0.0, 0.0, (uf[0] * dx * dx) / 12.0]))
    if projection:
        dx, dy = abs(dx), abs(dy)
    else:
        dx = math.sqrt(dx * dx + dy * dy)
        dy = dx
    for n in nodes:
        ln.append(LoadNode(n, [(uf[0] * dy) / 2.0, 0.0, 0.0]))
        ln.append(LoadNode(n, [0.0, (uf[1] * dx) / 2.0, 0.0]))
    return ln


# ----------------------------------------------------------------------------- mod.Compare
def compare(exp: Sequence[LoadCase], act: Sequence[LoadCase], eps: float) -> List[str]:
    """mod/compare.go:13-73. Returns the list of differences (empty == nil error)."""
    errs: List[str] = []
    if len(exp) != len(act):
        errs.append(f"length of load cases is not same : {len(exp)} != {len(act)}")
    if eps <= 0.0:
        errs.append(f"epsilon is not valid: {eps}")
    if errs:
        return errs
    for name, get in (("reactions", lambda lc: lc.Reactions), ("displacement", lambda lc: lc.PointDisplacementGlobal)):
        for e in range(len(exp)):
            ee, aa = get(exp[e]), get(act[e])
            for r in range(len(ee)):
                for j in range(3):
                    x, y = float(ee[r][j]), float(aa[r][j])
                    if x == 0.0:
                        if abs(y) > eps:
                            errs.append(f"{name}: expected zero, actual {y:e} (case {e}, row {r}, {j})")
                        continue
                    if abs((x - y) / x) > eps:
                        errs.append(f"{name} `{j}` is not same: {x:e} != {y:e} (case {e}, row {r})")
    return errs


# ----------------------------------------------------------------------------- model families
def console_beam() -> Tuple[Model, List[LoadCase], List[ModalCase]]:
    """example.ConsoleBeam (example/example.go:14-49) re-expressed."""
    m = Model(
        Points=[[0.0, 0.0], [1.0, 0.0]],
        Beams=[BeamProp(N=[0, 1], A=NOMINAL_A, J=NOMINAL_J, E=NOMINAL_E)],
        Supports=[[True, True, True], [False, False, False]],
    )
    lcs = [
        LoadCase(LoadNodes=[LoadNode(1, [0, 2.3, 0]), LoadNode(1, [10, 0, 0])]),
        LoadCase(LoadNodes=[LoadNode(1, [10, 0, 0]), LoadNode(1, [0, 2.3, 0])]),
    ]
    mcs = [ModalCase(ModalMasses=[ModalMass(1, 10000)])]
    return m, lcs, mcs


def bench_cantilever(ic: int, n_cases: int = 20) -> Tuple[Model, List[LoadCase], List[ModalCase]]:
    """The model and cases BenchmarkRun builds (mod/split_test.go:93-124).

    `lc := lcs[0]` copies the struct but shares the LoadNodes backing array, so every
    `lc.LoadNodes[0].Forces[1] += float64(i)` mutates case 0 and all its copies: cases
    0,2..19 end with Fy = 2.3 + sum(0..17) = 155.3; case 1 keeps Fx=10, Fy=2.3.
    """
    m, lcs, mcs = console_beam()
    split_beam(m, 0, ic)
    i = 0
    shared = lcs[0].LoadNodes  # aliased on purpose
    while len(lcs) < n_cases:
        shared[0].Forces[1] += float(i)
        lcs.append(LoadCase(LoadNodes=shared))
        i += 1
    for i in range(2, len(m.Points), 2):
        mcs[0].ModalMasses.append(ModalMass(i, 100))
    return m, lcs, mcs


BAY, HEIGHT = 6.0, 3.0


def uniform_frame(n_bays: int) -> Model:
    """Single-storey n-bay portal frame (model family 2).

    point 2i = (i*6, 0) fixed base, point 2i+1 = (i*6, 3); beams: col_0 = (0->1), then for
    i = 1..N: col_i = (2i -> 2i+1), girder_i = (2i-1 -> 2i+1).  nP = 2N+2, nB = 2N+1.
    """
    N = n_bays
    m = Model()
    for i in range(N + 1):
        m.Points.append([i * BAY, 0.0])
        m.Points.append([i * BAY, HEIGHT])
        m.Supports.append([True, True, True])
        m.Supports.append([False, False, False])
    m.Beams.append(BeamProp(N=[0, 1], A=NOMINAL_A, J=NOMINAL_J, E=NOMINAL_E))
    for i in range(1, N + 1):
        m.Beams.append(BeamProp(N=[2 * i, 2 * i + 1], A=NOMINAL_A, J=NOMINAL_J, E=NOMINAL_E))
        m.Beams.append(BeamProp(N=[2 * i - 1, 2 * i + 1], A=NOMINAL_A, J=NOMINAL_J, E=NOMINAL_E))
    return m


def girder_index(i: int) -> int:
    """beam index of girder_i (i = 1..N) in uniform_frame."""
    return 2 * i


# ----------------------------------------------------------------------------- counter-based RNG
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(model_idx: np.ndarray, stream: int, item_idx: np.ndarray, seed: int = SEED) -> np.ndarray:
    """U[0,1) as a pure function of (seed, model, stream, item): shards reproduce the batch."""
    with np.errstate(over="ignore"):
        key = _splitmix(np.uint64(seed) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))
        z = _splitmix(key ^ _splitmix(model_idx.astype(np.uint64)))
        z = _splitmix(z + item_idx.astype(np.uint64))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_frame_sections(n_bays: int, model0: int, n_models: int) -> np.ndarray:
    """sections[n_models][nB][3]: A, J <- nominal * U[0.5, 2.0) per beam, E fixed."""
    nB = 2 * n_bays + 1
    mi = np.arange(model0, model0 + n_models, dtype=np.uint64)[:, None]
    bi = np.arange(nB, dtype=np.uint64)[None, :]
    sec = np.empty((n_models, nB, 3), dtype=np.float64)
    sec[:, :, 0] = NOMINAL_A * (0.5 + 1.5 * uniform01(mi, 1, bi))
    sec[:, :, 1] = NOMINAL_J * (0.5 + 1.5 * uniform01(mi, 2, bi))
    sec[:, :, 2] = NOMINAL_E
    return sec


def uniform_frame_load_params(n_bays: int, model0: int, n_models: int, n_cases: int = 20):
    """q[n_models][n_cases][N] (girder line load, N/m, negative = down) and fx[n_models][n_cases]."""
    N = n_bays
    mi = np.arange(model0, model0 + n_models, dtype=np.uint64)[:, None, None]
    k = np.arange(n_cases, dtype=np.uint64)[None, :, None]
    g = np.arange(N, dtype=np.uint64)[None, None, :]
    kk = np.arange(n_cases, dtype=np.float64)[None, :, None]
    u = uniform01(mi, 3, k * np.uint64(N) + g)
    q = -(1000.0 + 100.0 * kk) * (0.5 + u)
    u2 = uniform01(mi[:, :, 0], 4, k[:, :, 0])
    fx = (500.0 + 50.0 * kk[:, :, 0]) * (0.5 + u2)
    return q, fx


def uniform_frame_loads(n_bays: int, model0: int, n_models: int, n_cases: int = 20, out=None) -> np.ndarray:
    """Dense loads[n_models][n_cases][dof]: LoadUniform(girder_i, projection, {0, q}) on every
    girder plus a lateral Fx on point 1, flattened by assemblyNodeLoad (model.go:683-698).
    Vectorised; bit-identical to uniform_frame_load_cases + node accumulation (tested)."""
    N = n_bays
    dof = 6 * N + 6
    q, fx = uniform_frame_load_params(N, model0, n_models, n_cases)
    p = np.zeros((n_models, n_cases, dof), dtype=np.float64) if out is None else out
    if out is not None:
        p[...] = 0.0
    dx = -BAY  # x[N0] - x[N1] for a girder running left to right
    mom = (q * dx * dx) / 12.0  # start node gets +mom, end node -mom (uniform.go:114-132)
    fy = (q * abs(dx)) / 2.0
    top = 3 * (2 * np.arange(N + 1) + 1)  # first DOF of top node t
    # girder i (1..N) = top node i-1 -> top node i; accumulate in LoadNodes order
    p[:, :, top[:-1] + 2] += mom
    p[:, :, top[1:] + 2] += -mom
    p[:, :, top[:-1] + 1] += fy
    p[:, :, top[1:] + 1] += fy
    p[:, :, 3 * 1 + 0] += fx
    return p


def uniform_frame_load_cases(m: Model, n_bays: int, model_idx: int, n_cases: int = 20) -> List[LoadCase]:
    """The same loads as LoadCase objects built through load_uniform (slow path, used by tests
    and by the single-model API)."""
    q, fx = uniform_frame_load_params(n_bays, model_idx, 1, n_cases)
    lcs = []
    for k in range(n_cases):
        ln: List[LoadNode] = []
        for i in range(1, n_bays + 1):
            ln.extend(load_uniform(m, girder_index(i), True, [0.0, float(q[0, k, i - 1])]))
        ln.append(LoadNode(1, [float(fx[0, k]), 0.0, 0.0]))
        lcs.append(LoadCase(LoadNodes=ln))
    return lcs


def uniform_frame_masses(n_bays: int, model0: int, n_models: int):
    """Modal masses (config 4): 1000*U[0.5,1.5) N on every top node. -> massN[N+1], mass[n_models][N+1]"""
    N = n_bays
    massN = (2 * np.arange(N + 1) + 1).astype(np.int64)
    mi = np.arange(model0, model0 + n_models, dtype=np.uint64)[:, None]
    t = np.arange(N + 1, dtype=np.uint64)[None, :]
    mass = 1000.0 * (0.5 + uniform01(mi, 5, t))
    return massN, mass


def apply_sections(m: Model, sections: np.ndarray) -> Model:
    mm = m.clone()
    for b, s in zip(mm.Beams, sections):
        b.A, b.J, b.E = float(s[0]), float(s[1]), float(s[2])
    return mm
