"""ctypes binding of the CPU oracle (oracle/hd_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module.  The product package (hd_b200) never does.

The functions take flat numpy arrays (the layout of oracle/hd_oracle.h) plus
`run()`, a duck-typed restatement of hd.Run's orchestration (model.go:926-969,
358-605, 719-851) that works on any objects carrying the Go field names
(Points, Beams[].N/A/J/E, Pins, Supports, LoadNodes[].N/Forces, ModalMasses[].N/Mass).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhd_oracle.so")

OK, ERR_INPUT, ERR_SINGULAR, ERR_IGNORE_LOAD, ERR_EIGEN, ERR_ALLOC = range(6)


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "hd_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "hd_oracle.h"))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


class _Model(C.Structure):
    _fields_ = [
        ("nP", C.c_int32),
        ("nB", C.c_int32),
        ("points", C.c_void_p),
        ("beamN", C.c_void_p),
        ("sections", C.c_void_p),
        ("pins", C.c_void_p),
        ("supports", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.hdo_distance.restype = C.c_double
        L.hdo_free.argtypes = [C.c_void_p]
        L.hdo_free.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class FlatModel:
    """Flattened hd.Model.  Arrays are kept alive on the object."""

    def __init__(self, points, beamN, sections, pins, supports):
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        self.beamN = np.ascontiguousarray(beamN, dtype=np.int64).reshape(-1, 2)
        self.sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, 3)
        nB = self.beamN.shape[0]
        nP = self.points.shape[0]
        if pins is None or len(pins) == 0:
            pins = np.zeros((nB, 6), dtype=np.uint8)  # model.go:344-346
        self.pins = np.ascontiguousarray(pins, dtype=np.uint8).reshape(-1, 6)
        self.supports = np.ascontiguousarray(supports, dtype=np.uint8).reshape(-1, 3)
        self.nP, self.nB = nP, nB
        self.dof = 3 * nP

    def with_sections(self, sections):
        return FlatModel(self.points, self.beamN, sections, self.pins, self.supports)

    def c(self):
        m = _Model(self.nP, self.nB, _p(self.points), _p(self.beamN), _p(self.sections),
                   _p(self.pins), _p(self.supports))
        return m


def flatten_model(m) -> FlatModel:
    pts = np.array([[p[0], p[1]] for p in m.Points], dtype=np.float64).reshape(-1, 2)
    bn = np.array([[b.N[0], b.N[1]] for b in m.Beams], dtype=np.int64).reshape(-1, 2)
    sec = np.array([[b.A, b.J, b.E] for b in m.Beams], dtype=np.float64).reshape(-1, 3)
    pins = np.array(m.Pins, dtype=np.uint8).reshape(-1, 6) if len(m.Pins) else None
    sup = np.array(m.Supports, dtype=np.uint8).reshape(-1, 3)
    return FlatModel(pts, bn, sec, pins, sup)


# ----------------------------------------------------------------------------- element level
def stiff_beam2d(xy, A, J, E, pins):
    k = np.zeros(36)
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    pn = np.ascontiguousarray(pins, dtype=np.uint8)
    lib().hdo_stiff_beam2d(_p(xy), C.c_double(A), C.c_double(J), C.c_double(E), _p(pn), _p(k))
    return k.reshape(6, 6)


def trans_beam2d(xy):
    t = np.zeros(36)
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    lib().hdo_trans_beam2d(_p(xy), _p(t))
    return t.reshape(6, 6)


def geometric_beam2d(xy, pins, beam_forces):
    g = np.zeros(36)
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    pn = np.ascontiguousarray(pins, dtype=np.uint8)
    bf = np.ascontiguousarray(beam_forces, dtype=np.float64)
    lib().hdo_geometric_beam2d(_p(xy), _p(pn), _p(bf), _p(g))
    return g.reshape(6, 6)


def rotate_to_global(k, t):
    kg = np.zeros(36)
    k = np.ascontiguousarray(k, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    lib().hdo_rotate_to_global(_p(k), _p(t), _p(kg))
    return kg.reshape(6, 6)


# ----------------------------------------------------------------------------- assembly
def _take(ptr, n, ctype, dtype):
    if n == 0 or not ptr:
        out = np.zeros(0, dtype=dtype)
    else:
        out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).astype(dtype, copy=True)
    if ptr:
        lib().hdo_free(ptr)
    return out


def assemble(fm: FlatModel, which: int = 0, beam_forces=None):
    """-> dict(rc, dof, csc=(colptr,rowind,val), csr=(rowptr,colind,val), ignore)"""
    dof, nnz, nig = C.c_int32(), C.c_int32(), C.c_int32()
    ptrs = [C.c_void_p() for _ in range(7)]
    bf = np.ascontiguousarray(beam_forces, dtype=np.float64) if beam_forces is not None else None
    m = fm.c()
    rc = lib().hdo_assemble(C.byref(m), which, _p(bf), C.byref(dof), C.byref(nnz),
                            C.byref(ptrs[0]), C.byref(ptrs[1]), C.byref(ptrs[2]),
                            C.byref(ptrs[3]), C.byref(ptrs[4]), C.byref(ptrs[5]),
                            C.byref(nig), C.byref(ptrs[6]))
    if rc != OK:
        return {"rc": rc}
    n, z = dof.value, nnz.value
    return {
        "rc": rc, "dof": n, "nnz": z,
        "csc": (_take(ptrs[0].value, n + 1, C.c_int32, np.int32), _take(ptrs[1].value, z, C.c_int32, np.int32),
                _take(ptrs[2].value, z, C.c_double, np.float64)),
        "csr": (_take(ptrs[3].value, n + 1, C.c_int32, np.int32), _take(ptrs[4].value, z, C.c_int32, np.int32),
                _take(ptrs[5].value, z, C.c_double, np.float64)),
        "ignore": _take(ptrs[6].value, nig.value, C.c_int32, np.int32),
    }


def eliminated(fm: FlatModel):
    ne, ni = C.c_int32(), C.c_int32()
    pe, pi = C.c_void_p(), C.c_void_p()
    m = fm.c()
    rc = lib().hdo_eliminated(C.byref(m), C.byref(ne), C.byref(pe), C.byref(ni), C.byref(pi))
    if rc != OK:
        return rc, None, None
    return rc, _take(pe.value, ne.value, C.c_int32, np.int32), _take(pi.value, ni.value, C.c_int32, np.int32)


# ----------------------------------------------------------------------------- analyses
def linear_static(fm: FlatModel, loads):
    """loads [nLC][dof] -> rc, disp[nLC][nP][3], forces[nLC][nB][6], react[nLC][nP][3], failCase"""
    loads = np.ascontiguousarray(loads, dtype=np.float64).reshape(-1, fm.dof)
    nLC = loads.shape[0]
    disp = np.zeros((nLC, fm.nP, 3))
    forces = np.zeros((nLC, fm.nB, 6))
    react = np.zeros((nLC, fm.nP, 3))
    fc = C.c_int32(-1)
    m = fm.c()
    rc = lib().hdo_linear_static(C.byref(m), nLC, _p(loads), _p(disp), _p(forces), _p(react), C.byref(fc))
    return rc, disp, forces, react, fc.value


def buckling(fm: FlatModel, beam_forces, amount: int):
    dof = fm.dof
    bf = np.ascontiguousarray(beam_forces, dtype=np.float64)
    n = C.c_int32()
    factors = np.zeros(dof)
    shapes = np.zeros((dof, fm.nP, 3))
    m = fm.c()
    rc = lib().hdo_buckling(C.byref(m), _p(bf), amount, dof, C.byref(n), _p(factors), _p(shapes))
    k = min(n.value, dof)
    return rc, factors[:k].copy(), shapes[:k].copy()


def modal(fm: FlatModel, massN, mass):
    dof = fm.dof
    massN = np.ascontiguousarray(massN, dtype=np.int64)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    n = C.c_int32()
    hz = np.zeros(dof)
    shapes = np.zeros((dof, fm.nP, 3))
    m = fm.c()
    rc = lib().hdo_modal(C.byref(m), len(massN), _p(massN), _p(mass), dof, C.byref(n), _p(hz), _p(shapes))
    k = min(n.value, dof)
    return rc, hz[:k].copy(), shapes[:k].copy()


def eig_general(a):
    a = np.array(a, dtype=np.float64, order="C")
    n = a.shape[0]
    wr, wi, v = np.zeros(n), np.zeros(n), np.zeros((n, n))
    rc = lib().hdo_eig_general(n, _p(a), _p(wr), _p(wi), _p(v))
    return rc, wr, wi, v


def check_model(fm: FlatModel, n_pins=None, n_supports=None):
    m = fm.c()
    return lib().hdo_check_model(C.byref(m), fm.pins.shape[0] if n_pins is None else n_pins,
                                 fm.supports.shape[0] if n_supports is None else n_supports)


def check_loads(nP, loadN, forces):
    loadN = np.ascontiguousarray(loadN, dtype=np.int64)
    forces = np.ascontiguousarray(forces, dtype=np.float64).reshape(-1, 3)
    return lib().hdo_check_loads(nP, len(loadN), _p(loadN), _p(forces))


def check_modal(fm: FlatModel, massN, mass):
    massN = np.ascontiguousarray(massN, dtype=np.int64)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    m = fm.c()
    return lib().hdo_check_modal(C.byref(m), len(massN), _p(massN), _p(mass))


def set_literal_scans(on: bool):
    lib().hdo_set_literal_scans(1 if on else 0)


def max_threads() -> int:
    return lib().hdo_max_threads()


def static_batch(fm: FlatModel, sections, loads, n_threads=0, want=("disp", "forces", "react")):
    """sections [nModels][nB][3], loads [nModels][nLC][dof]"""
    sections = np.ascontiguousarray(sections, dtype=np.float64)
    nM = sections.shape[0]
    loads = np.ascontiguousarray(loads, dtype=np.float64).reshape(nM, -1, fm.dof)
    nLC = loads.shape[1]
    disp = np.zeros((nM, nLC, fm.nP, 3)) if "disp" in want else None
    forces = np.zeros((nM, nLC, fm.nB, 6)) if "forces" in want else None
    react = np.zeros((nM, nLC, fm.nP, 3)) if "react" in want else None
    status = np.zeros(nM, dtype=np.int32)
    m = fm.c()
    lib().hdo_static_batch(C.byref(m), nM, _p(sections), nLC, _p(loads), _p(disp), _p(forces), _p(react),
                           _p(status), n_threads)
    return status, disp, forces, react


def modal_batch(fm: FlatModel, sections, massN, mass, n_threads=0):
    sections = np.ascontiguousarray(sections, dtype=np.float64)
    nM = sections.shape[0]
    massN = np.ascontiguousarray(massN, dtype=np.int64)
    mass = np.ascontiguousarray(mass, dtype=np.float64).reshape(nM, len(massN))
    mo = fm.dof
    nOut = np.zeros(nM, dtype=np.int32)
    hz = np.zeros((nM, mo))
    shapes = np.zeros((nM, mo, fm.nP, 3))
    status = np.zeros(nM, dtype=np.int32)
    m = fm.c()
    lib().hdo_modal_batch(C.byref(m), nM, _p(sections), len(massN), _p(massN), _p(mass), mo, _p(nOut),
                          _p(hz), _p(shapes), _p(status), n_threads)
    return status, nOut, hz, shapes


# ----------------------------------------------------------------------------- hd.Run restated
def node_loads(nP, load_nodes):
    """assemblyNodeLoad, model.go:683-698 (accumulate in LoadNodes order)."""
    p = np.zeros(3 * nP)
    for ln in load_nodes:
        for i in range(3):
            p[ln.N * 3 + i] += ln.Forces[i]
    return p


def run(m, lcs, mcs):
    """hd.Run (model.go:926-969) on duck-typed Go-shaped objects.

    Returns (err, static, modal): err is None or a list of error codes collected the way
    Run collects them; static = list per load case of dict(disp, forces, react, buckling) or
    None; modal = list per modal case of dict(hz, shapes) or None (error) .
    """
    errs = []
    if len(m.Pins) == 0:  # model.go:344-346 mutates the model
        m.Pins = [[False] * 6 for _ in m.Beams]
    static = [None] * len(lcs)
    n_pins, n_sup = len(m.Pins), len(m.Supports)

    def safe_flat():
        # check.go runs on the Go structs; index errors there never reach array code.
        nP = len(m.Points)
        pins = [list(p) for p in m.Pins][: len(m.Beams)] + [[False] * 6] * max(0, len(m.Beams) - n_pins)
        sup = [list(s) for s in m.Supports][:nP] + [[False] * 3] * max(0, nP - n_sup)
        return FlatModel(
            np.array([[p[0], p[1]] for p in m.Points], dtype=np.float64).reshape(-1, 2),
            np.array([[b.N[0], b.N[1]] for b in m.Beams], dtype=np.int64).reshape(-1, 2),
            np.array([[b.A, b.J, b.E] for b in m.Beams], dtype=np.float64).reshape(-1, 3),
            np.array(pins, dtype=np.uint8).reshape(-1, 6),
            np.array(sup, dtype=np.uint8).reshape(-1, 3),
        )

    fm = safe_flat()
    model_ok = check_model(fm, n_pins, n_sup) == OK
    if len(lcs) > 0:
        ok = model_ok
        for lc in lcs:
            ln = [l.N for l in lc.LoadNodes]
            fr = [list(l.Forces) for l in lc.LoadNodes]
            if check_loads(fm.nP, ln, fr) != OK:
                ok = False
        if not ok:
            errs.append(ERR_INPUT)
        else:
            loads = np.stack([node_loads(fm.nP, lc.LoadNodes) for lc in lcs])
            with np.errstate(all="ignore"):
                rc, disp, forces, react, fc = linear_static(fm, loads)
            upto = len(lcs) if rc == OK else max(fc, 0)
            if rc != OK:
                errs.append(rc)
            for i in range(upto):
                res = {"disp": disp[i], "forces": forces[i], "react": react[i], "buckling": None}
                amount = getattr(lcs[i], "AmountLinearBuckling", 0)
                if amount > 0:
                    brc, fac, shp = buckling(fm, forces[i], amount)
                    if brc != OK:
                        errs.append(brc)
                        static[i] = res
                        break
                    res["buckling"] = (fac, shp)
                static[i] = res
    modal_out = [None] * len(mcs)
    for i, mc in enumerate(mcs):
        mn = [q.N for q in mc.ModalMasses]
        ms = [q.Mass for q in mc.ModalMasses]
        if not model_ok or check_modal(fm, mn, ms) != OK:
            errs.append(ERR_INPUT)
            continue
        rc, hz, shapes = modal(fm, mn, ms)
        if rc != OK:
            errs.append(rc)
            continue
        modal_out[i] = {"hz": hz, "shapes": shapes}
    return (errs or None), static, modal_out
