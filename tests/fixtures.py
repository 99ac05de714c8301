"""Fixture models: the reference's example constructors (example/example.go) re-expressed as
compact tables, plus the model of model_test.go:508-552 (`Test`).  Each entry returns fresh
(Model, [LoadCase], [ModalCase]) objects.  Golden numbers for them live in
tests/golden/golden.json (made by tests/golden/make_golden.py from the reference's
example/testdata reports)."""
from __future__ import annotations

import math

from hd_b200.model import BeamProp, Gravity, LoadCase, LoadNode, ModalCase, ModalMass, Model

T, F = True, False


def _model(points, beams, supports, pins=None):
    return Model(
        Points=[list(map(float, p)) for p in points],
        Beams=[BeamProp(N=[a, b], A=A, J=J, E=E) for (a, b, A, J, E) in beams],
        Pins=[list(p) for p in pins] if pins else [],
        Supports=[list(s) for s in supports],
    )


def _lc(loads, buckling=0):
    return LoadCase(LoadNodes=[LoadNode(n, [float(fx), float(fy), float(mz)]) for (n, fx, fy, mz) in loads],
                    AmountLinearBuckling=buckling)


def _mc(masses):
    return ModalCase(ModalMasses=[ModalMass(n, float(w)) for (n, w) in masses])


_S = (12e-4, 120e-6, 2.0e11)
_ROD_A = math.pi * math.pow(0.050, 2) / 4.0
_ROD_J = math.pi * math.pow(0.050, 4) / 64.0
_ROD = (_ROD_A, _ROD_J, 2e11)
FIX, FREE = (T, T, T), (F, F, F)


def console_beam():  # example.go:14-49, golden "beam"
    m = _model([(0, 0), (1, 0)], [(0, 1, *_S)], [FIX, FREE])
    return m, [_lc([(1, 0, 2.3, 0), (1, 10, 0, 0)]), _lc([(1, 10, 0, 0), (1, 0, 2.3, 0)])], [_mc([(1, 10000)])]


def buckling_beam():  # example.go:56-89, golden "beam-buckling"
    m = _model([(0, 0), (0.5, 0), (1.0, 0), (1.5, 0), (2.0, 0)],
               [(0, 1, *_S), (1, 2, *_S), (2, 3, *_S), (3, 4, *_S)],
               [(T, T, F), FREE, FREE, FREE, (F, T, F)])
    return m, [_lc([(4, -1.0, 0, 0)], buckling=2)], []


def g_beam():  # example.go:95-126, golden "g-beam"
    s = (2.5e-3, 0.5208e-6, 2.0e11)
    m = _model([(0, 0), (1, 0), (1, 1)], [(0, 1, *s), (1, 2, *s)], [FIX, FREE, FIX])
    return m, [_lc([(1, 0, 13e3, 0), (1, 13e3, 0, 0)])], [_mc([(1, 10000)])]


def truss():  # example.go:138-205, golden "truss"
    E = 2.0e11
    m = _model([(0, 0), (0, 12), (4, 0), (4, 6), (8, 0)],
               [(0, 1, 40e-4, 1, E), (0, 2, 64e-4, 1, E), (0, 3, 60e-4, 1, E), (1, 3, 60e-4, 1, E),
                (2, 3, 40e-4, 1, E), (2, 4, 64e-4, 1, E), (3, 4, 60e-4, 1, E)],
               [(T, T, F), FREE, (F, T, F), FREE, (F, T, F)],
               pins=[(F, T, T, F, T, T)] * 7)
    return m, [_lc([(1, -70000, 0, 0), (3, 42000, 0, 0)])], [_mc([(1, 10000)])]


def truss_with_buckling():  # example.go:207-249, golden "truss-with-buckling"
    m = _model([(0.0, 0), (0.3, 0), (0.7, 0), (1.0, 0)], [(0, 1, *_ROD), (1, 2, *_ROD), (2, 3, *_ROD)],
               [FIX, FREE, FREE, (F, T, F)], pins=[(F,) * 6] * 3)
    return m, [_lc([(3, +70000, 0, 0)], buckling=2), _lc([(3, -70000, 0, 0)], buckling=2)], []


def _two_rods(points, supports, pins):
    m = _model(points, [(0, 1, *_ROD), (1, 2, *_ROD)], supports, pins=pins)
    return m, [], [_mc([(1, 100 * Gravity)])]


def modal_truss():  # example.go:251-288, golden "truss-modal"
    return _two_rods([(0, 0), (0.4, 0), (1.0, 0)], [FIX, FREE, FIX], [(F, F, T, F, F, F), (F, F, F, T, F, T)])


def modal_truss_rotate():  # example.go:290-327, golden "truss-modal-rotate"
    return _two_rods([(0, 0), (0, 0.4), (0, 1.0)], [FIX, FREE, FIX], [(F, F, T, F, F, F), (F, F, F, T, F, T)])


def modal_beam():  # example.go:329-366, golden "beam-modal"
    return _two_rods([(0, 0), (0.4, 0), (1.0, 0)], [(T, T, F), FREE, (F, T, F)], [(F,) * 6] * 2)


def modal_beam_rotate():  # example.go:368-405, golden "beam-modal-rotate"
    return _two_rods([(0, 0), (0, 0.4), (0, 1.0)], [(T, T, F), FREE, (T, F, F)], [(F,) * 6] * 2)


def modal_beam_3mass():  # example.go:407-450, golden "beam-modal-3mass"
    E, J, mass, l, A = 2e11, 15e-4, 250.0, 4.0, 12e-2
    s = (A, J, E)
    m = _model([(0, 0), (0, l / 6.0), (0, l / 2.0), (0, l - l / 6.0), (0, l)],
               [(0, 1, *s), (1, 2, *s), (2, 3, *s), (3, 4, *s)],
               [(T, T, F), FREE, FREE, FREE, (T, F, F)])
    return m, [], [_mc([(1, mass * Gravity), (2, mass * Gravity), (3, mass * Gravity)])]


def beam_dc_part1():  # example.go:463-490, golden "beam-dc-part1"
    m = _model([(0, 0), (1, 0), (2, 0), (3, 0)], [(0, 1, *_S), (1, 2, *_S), (2, 3, *_S)], [FIX, FREE, FREE, FIX])
    return m, [_lc([(1, 0, 10, 0), (2, 0, 10, 0)])], []


def beam_dc_part2():  # example.go:491-521, golden "beam-dc-part2"
    m = _model([(0, 0), (1, 0), (1.5, 0), (2, 0), (3, 0)],
               [(0, 1, *_S), (1, 2, *_S), (2, 3, *_S), (3, 4, *_S)], [FIX, FREE, FREE, FREE, FIX])
    return m, [_lc([(1, 0, 10, 0), (3, 0, 10, 0)])], []


def frame_modal():  # example.go:525-576, golden "beam-frame-modal"
    s = (24.0e-4, 72.0e-8, 2.00e10)
    mass = 100.0 * 10
    m = _model([(0.0, 0.4), (0.2, 0.4), (0.4, 0.0), (0.4, 0.2), (0.4, 0.4), (0.7, 0.4), (1.0, 0.4)],
               [(0, 1, *s), (1, 4, *s), (2, 3, *s), (3, 4, *s), (4, 5, *s), (5, 6, *s)],
               [FIX, FREE, FIX, FREE, FREE, FREE, FIX])
    return m, [], [_mc([(1, mass), (5, mass)])]


def model_string():  # model_test.go:508-552, golden "model.String"
    m = _model([(0, 0), (1, 0)], [(0, 1, *_S)], [FIX, FREE], pins=[(F, F, F, F, F, T)])
    lcs = [_lc([(1, 0, 2.3, 0), (1, -10, 0, 0)], buckling=1), _lc([(1, -10, 0, 0), (1, 0, 2.3, 0)], buckling=2)]
    return m, lcs, [_mc([(1, 10000)])]


# golden file name -> constructor (example_test.go:12-77, model_test.go:551)
GOLDEN = {
    "beam": console_beam,
    "truss": truss,
    "g-beam": g_beam,
    "truss-modal": modal_truss,
    "truss-modal-rotate": modal_truss_rotate,
    "beam-modal": modal_beam,
    "beam-modal-rotate": modal_beam_rotate,
    "truss-with-buckling": truss_with_buckling,
    "beam-modal-3mass": modal_beam_3mass,
    "beam-frame-modal": frame_modal,
    "beam-buckling": buckling_beam,
    "beam-dc-part1": beam_dc_part1,
    "beam-dc-part2": beam_dc_part2,
    "model.String": model_string,
}

# the seven models TestSplit refines (mod/split_test.go:31-39)
SPLIT_MODELS = [console_beam, g_beam, truss, modal_beam, modal_beam_rotate, modal_truss, modal_truss_rotate]
