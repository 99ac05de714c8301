"""Host mirror of hd's three entry points over the C ABI (reference model.go:926-969, 358-605, 719-851).

    Run(out, m, lcs, mcs)          hd.Run
    LinearStatic(out, m, *lcs)     hd.LinearStatic
    Modal(out, m, mc)              hd.Modal
    RunBatch(m, sections, loads)   new: many models sharing one topology

Same argument meaning, same mutation of the arguments (Pins filled when empty, model.go:344-346;
results reset at entry, :365-367, :726), same error behaviour: nothing raises for bad input —
an error value is returned and Run keeps going after a failed case (:937-939, 951-954).  `out` (any
object with write(str), or None) receives the reference's text report, byte for byte (report.py).
Validation is hdb_check_* (check.go restated in the library); all arithmetic runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from . import report
from .model import BucklingResult, LoadCase, ModalCase, ModalResult, Model


class HdError(Exception):
    """Mirror of the error tree hd returns (errors.Tree): a name and a list of causes."""

    def __init__(self, name: str, causes: Sequence[object]):
        self.name = name
        self.causes = list(causes)
        super().__init__(f"{name}: " + "; ".join(str(c) for c in self.causes))


STATUS_TEXT = {
    _lib.STATUS_NONFINITE: "not valid value (NaN/Inf/non-positive section, stiffness entry or node load)",
    _lib.STATUS_SINGULAR: "LU error factorization: singular matrix",
    _lib.STATUS_IGNORE_LOAD: "Warning: List loads on not valid directions",
    _lib.STATUS_EIGEN: "Eigen factorization is not ok",
    _lib.STATUS_PATTERN: "sparsity pattern differs from the topology's representative model",
}


def _flat(m: Model):
    pts = np.array([[p[0], p[1]] for p in m.Points], dtype=np.float64).reshape(-1, 2)
    bn = np.array([[b.N[0], b.N[1]] for b in m.Beams], dtype=np.int64).reshape(-1, 2)
    sec = np.array([[b.A, b.J, b.E] for b in m.Beams], dtype=np.float64).reshape(-1, 3)
    pins = np.array([list(p) for p in m.Pins], dtype=np.uint8).reshape(-1, 6)
    sup = np.array([list(s) for s in m.Supports], dtype=np.uint8).reshape(-1, 3)
    return pts, bn, sec, pins, sup


def _prepare(m: Model) -> Optional[str]:
    """prepare() of model.go:336-355: fill Pins, run checkInputData. Returns an error text or None."""
    if len(m.Pins) == 0:
        m.Pins = [[False] * 6 for _ in m.Beams]
    pts, bn, sec, pins, sup = _flat(m)
    beams = _lib.beams_array(bn, sec)
    msg = C.create_string_buffer(4096)
    rc = _lib.lib().hdb_check_model(len(pts), _lib._np_ptr(pts), len(beams), _lib._np_ptr(beams), len(m.Pins),
                                    _lib._np_ptr(pins), len(m.Supports), _lib._np_ptr(sup), msg, 4096)
    if rc == _lib.OK:
        return None
    return msg.value.decode() or _lib.lib().hdb_error_string(rc).decode()


def _loads_struct(lc: LoadCase):
    a = np.zeros(len(lc.LoadNodes), dtype=_lib.LOAD_DTYPE)
    for i, ln in enumerate(lc.LoadNodes):
        a["n"][i] = ln.N
        a["forces"][i] = ln.Forces
    return a


def _topology(m: Model) -> _lib.Topology:
    pts, bn, sec, pins, sup = _flat(m)
    return _lib.Topology(pts, bn, sec, pins, sup, ordering=2)  # automatic ordering: results do not depend on it beyond round-off


def LinearStatic(out, m: Model, *lcs: LoadCase) -> Optional[HdError]:
    name = "Linear Elastic Analysis"
    for lc in lcs:
        lc.reset()
    if out is not None:
        out.write(name + "\n")  # model.go:370-371
    errs: List[str] = []
    e = _prepare(m)
    if e:
        errs.append(e)
    L = _lib.lib()
    nP = len(m.Points)
    structs = []
    for lc in lcs:  # check.go:127-163
        a = _loads_struct(lc)
        structs.append(a)
        msg = C.create_string_buffer(2048)
        if L.hdb_check_loads(nP, len(a), _lib._np_ptr(a), msg, 2048) != _lib.OK:
            errs.append(msg.value.decode())
    if errs:
        return HdError(name, [HdError("Prepared input data", errs)])
    try:
        topo = _topology(m)
    except _lib.HdbError as ex:
        return HdError(name, [f"Assembly node load: {ex}"])
    buckling, buckling_err, ok_upto_b = {}, None, None
    try:
        nLC = len(lcs)
        loads = np.zeros((1, max(nLC, 1), topo.dof))
        for i, a in enumerate(structs):  # assemblyNodeLoad, model.go:683-698
            _lib.check(L.hdb_assemble_node_loads(nP, len(a), _lib._np_ptr(a), _lib._np_ptr(loads[0, i])), "node loads")
        sec = np.array([[b.A, b.J, b.E] for b in m.Beams], dtype=np.float64)
        if nLC == 0:
            status, _ = topo.assemble_batch(sec[None])  # getK still runs and may fail (model.go:410)
            st, fc = int(status[0]), -1
            disp = forces = react = None
        else:
            with np.errstate(all="ignore"):
                status, fail, disp, forces, react = topo.static_batch(sec[None], loads[:, :nLC])
            st, fc = int(status[0]), int(fail[0])
            # buckling block (model.go:491-601) of every case that asks for it and has static results
            ok_upto = nLC if st == _lib.STATUS_OK else (fc if fc >= 0 else 0)
            for i in range(ok_upto):
                amount = lcs[i].AmountLinearBuckling
                if amount <= 0:
                    continue
                bst, nOut, fac, shp = topo.buckling_batch(sec[None], forces[:, i:i + 1], amount)
                if int(bst[0]) != _lib.STATUS_OK:
                    buckling_err = HdError(name, [f"Linear buckling: {STATUS_TEXT.get(int(bst[0]), int(bst[0]))}"])
                    ok_upto_b = i
                    break
                k = min(int(nOut[0, 0]), amount)
                buckling[i] = [BucklingResult(Factor=float(fac[0, 0, r]), PointDisplacementGlobal=shp[0, 0, r].tolist())
                               for r in range(k)]
    finally:
        topo.close()
    upto = nLC
    err = None
    if st == _lib.STATUS_IGNORE_LOAD:
        upto = fc  # cases before the offending one keep their results (model.go:415-437)
        err = HdError(name, [STATUS_TEXT[st]])
    elif st == _lib.STATUS_NONFINITE and fc >= 0:
        upto = fc  # assemblyNodeLoad of case fc failed (model.go:692-696): the cases before it keep their results
        err = HdError(name, ["Assembly node load: not valid node load"])
    elif st != _lib.STATUS_OK:
        upto = 0
        err = HdError(name, [f"Assembly node load: {STATUS_TEXT.get(st, st)}"])
    if buckling_err is not None and err is None:  # the reference stops at the first failing buckling block
        upto, err = min(upto, ok_upto_b + 1), buckling_err
    for i in range(upto):
        lcs[i].PointDisplacementGlobal = disp[0, i].tolist()
        lcs[i].BeamForces = forces[0, i].tolist()
        lcs[i].Reactions = react[0, i].tolist()
        if i in buckling:
            lcs[i].LinearBucklingResult = buckling[i]
    return err


def Modal(out, m: Model, mc: ModalCase) -> Optional[HdError]:
    name = "Modal Analysis"
    mc.reset()
    if out is not None:
        out.write(name + "\n")  # model.go:729-730
    errs: List[str] = []
    e = _prepare(m)
    if e:
        errs.append(e)
    L = _lib.lib()
    masses = np.zeros(len(mc.ModalMasses), dtype=_lib.MASS_DTYPE)
    for i, q in enumerate(mc.ModalMasses):
        masses["n"][i], masses["mass"][i] = q.N, q.Mass
    sup = np.array([list(s) for s in m.Supports], dtype=np.uint8).reshape(-1, 3)
    msg = C.create_string_buffer(2048)
    if L.hdb_check_modal(len(m.Points), len(m.Supports), _lib._np_ptr(sup), len(masses), _lib._np_ptr(masses), msg,
                         2048) != _lib.OK:
        errs.append(msg.value.decode())
    if errs:
        return HdError(name, errs)
    try:
        topo = _topology(m)
    except _lib.HdbError as ex:
        return HdError(name, [f"Assembly node load: {ex}"])
    try:
        sec = np.array([[b.A, b.J, b.E] for b in m.Beams], dtype=np.float64)
        massN = np.ascontiguousarray(masses["n"])
        with np.errstate(all="ignore"):
            status, nModes, hz, shapes = topo.modal_batch(sec[None], massN, np.ascontiguousarray(masses["mass"])[None])
    finally:
        topo.close()
    st = int(status[0])
    if st != _lib.STATUS_OK:
        return HdError(name, [STATUS_TEXT.get(st, st)])
    mc.Result = [ModalResult(Hz=float(hz[0, k]), ModalDisplacement=shapes[0, k].tolist()) for k in range(int(nModes[0]))]
    return None


def Run(out, m: Model, lcs: List[LoadCase], mcs: List[ModalCase]) -> Optional[HdError]:
    causes = []
    w = out.write if out is not None else (lambda s: None)
    w(report.model_string(m) + "\n")
    if len(lcs) > 0:
        err = LinearStatic(out, m, *lcs)
        if err is not None:
            causes.append(err)
        w("\n\n")
        for lc in lcs:
            w(report.load_case_string(lc) + "\n")
    w("\n\n")
    for mc in mcs:
        err = Modal(out, m, mc)
        if err is not None:
            causes.append(err)
            continue
        w("\n\n")
        w(report.modal_case_string(mc) + "\n")
    w("\n\n")
    return HdError("Run function", causes) if causes else None


def RunBatch(m: Model, sections, loads, want=("disp", "forces", "react"), points=None):
    """Many models sharing m's topology: sections[nModels][nB][3], loads[nModels][nLC][dof] and, for shape sweeps,
    points[nModels][nP][2] (default: m's own coordinates for every model).
    Returns (status, failCase, disp, forces, react) as numpy arrays (hdb_static_batch / hdb_static_batch_geom)."""
    e = _prepare(m)
    if e:
        raise HdError("RunBatch", [e])
    topo = _topology(m)
    try:
        if points is not None:
            return topo.static_batch_geom(points, sections, loads, want)
        return topo.static_batch(sections, loads, want)
    finally:
        topo.close()
