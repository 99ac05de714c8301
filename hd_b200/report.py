"""Text report of a run, byte-compatible with the reference's (host-side string work, no arithmetic).

    model_string(m)        Model.String      model.go:853-924
    load_case_string(lc)   LoadCase.String   model.go:132-211
    modal_case_string(mc)  ModalCase.String  model.go:235-255
    run_header / run_*     the framing hd.Run, LinearStatic and Modal put around them
                           (model.go:926-969, 370-371, 729-730)

The reference's own example tests compare exactly this text with example/testdata/* (example_test.go:12-77,
model_test.go:551), so a host that swaps the solver keeps passing them only if the report is reproduced to the
byte: %15.5e / %15.5f numbers, the trailing blank after the support flags, Go's `true`/`false` for pins.
"""
from __future__ import annotations

from typing import List

from .model import LoadCase, ModalCase, Model


def _e(v: float) -> str:  # Go %15.5e: two exponent digits at least, like Python's
    return "%15.5e" % v


def _f(v: float) -> str:  # Go %15.5f
    return "%15.5f" % v


def _row3(i: int, v) -> str:
    return "%5d %s %s %s\n" % (i, _e(v[0]), _e(v[1]), _e(v[2]))


def model_string(m: Model) -> str:
    out: List[str] = ["\n", "Point coordinates:\n"]
    out.append("%37s %17s\n" % ("", "   SUPPORT DIRECTION"))
    out.append("%5s %15s %15s " % ("Index", "X, m", "Y, m"))
    out.append("%5s %5s %5s (0 - free, 1 - fixed)\n" % ("SX", "SY", "SM"))
    for i, p in enumerate(m.Points):
        out.append("%5d %s %s " % (i, _f(p[0]), _f(p[1])))
        for j in range(3):
            out.append("%5d " % (1 if m.Supports[i][j] else 0))  # the blank follows all three flags (model.go:873)
        out.append("\n")
    out.append("Beam property:\n")
    out.append("%5s %15s %15s " % ("Index", "Start point", "End point"))
    out.append("%15s %15s %15s\n" % ("Area,sq.m", "Moment inertia,m4", "Elasticity,Pa"))
    for i, b in enumerate(m.Beams):
        out.append("%5d %15d %15d " % (i, b.N[0], b.N[1]))
        out.append("%s %s %s\n" % (_e(b.A), _e(b.J), _e(b.E)))
    header = False
    for k, pin in enumerate(m.Pins):
        if not any(pin):
            continue
        if not header:
            out.append("Pins of beam in local system coordinate:\n")
            out.append("%5s %23s %23s\n" % ("", "START OF BEAM     ", "END OF BEAM       "))
            out.append("%5s %23s %23s\n" % ("", "------------------", "------------------"))
            out.append("%5s %7s %7s %7s %7s %7s %7s\n" % ("Index", "X", "Y", "M", "X", "Y", "M"))
            header = True
        out.append("%5d " % k)
        out.append(" ".join("%7s" % ("true" if f else "false") for f in pin))
        out.append("\n")
    if not header:
        out.append("All beams haven`t pins\n")
    return "".join(out)


def load_case_string(lc: LoadCase) -> str:
    out: List[str] = ["\nLoad case:\n"]
    out.append("%5s %15s %15s %15s\n" % ("Point", "Fx, N", "Fy, N", "M, N*m"))
    for ln in lc.LoadNodes:
        out.append("%5d %s %s %s\n" % (ln.N, _f(ln.Forces[0]), _f(ln.Forces[1]), _f(ln.Forces[2])))
    if lc.PointDisplacementGlobal is not None and len(lc.PointDisplacementGlobal) > 0:
        out.append("Point displacement in global system coordinate:\n")
        out.append("%5s %15s %15s %15s\n" % ("Point", "DX, m", "DY, m", "Angle, rad."))
        for i, d in enumerate(lc.PointDisplacementGlobal):
            out.append(_row3(i, d))
    if lc.BeamForces is not None and len(lc.BeamForces) > 0:
        out.append("Local force in beam:\n")
        out.append("%5s %45s %45s\n" % ("", "START OF BEAM     ", "END OF BEAM       "))
        out.append("%5s %45s %45s\n" % ("", "-" * 36, "-" * 36))
        out.append("%5s %15s %15s %15s %15s %15s %15s\n" % ("Index", "Fx, N", "Fy, N", "M, N*m", "Fx, N", "Fy, N", "M, N*m"))
        for i, bf in enumerate(lc.BeamForces):
            out.append("%5d " % i)
            out.append(" ".join(_e(bf[j]) for j in range(6)))
            out.append("\n")
    if lc.Reactions is not None and len(lc.Reactions) > 0:
        out.append("Reaction on support:\n")
        out.append("%5s %15s %15s %15s\n" % ("Index", "Fx, N", "Fy, N", "M, N*m"))
        for i, r in enumerate(lc.Reactions):
            if r[0] == 0 and r[1] == 0 and r[2] == 0:  # points without reaction are not listed (model.go:184-188)
                continue
            out.append(_row3(i, r))
    if lc.LinearBucklingResult is not None and len(lc.LinearBucklingResult) > 0:
        out.append("\nLinear buckling result:\n")
        for lbr in lc.LinearBucklingResult:
            out.append("Linear buckling factor: %s\n" % _f(lbr.Factor))
            out.append("%5s %15s %15s %15s\n" % ("Point", "X", "Y", "M"))
            for i, d in enumerate(lbr.PointDisplacementGlobal):
                out.append(_row3(i, d))
    elif lc.AmountLinearBuckling == 0:
        out.append("\nLinear buckling result: is not calculated\n")
    else:
        out.append("\nLinear buckling result: haven`t valid data. Probably all beams are tension\n")
    return "".join(out)


def modal_case_string(mc: ModalCase) -> str:
    out: List[str] = ["\nModal case:\n"]
    out.append("%5s %15s\n" % ("Point", "Mass, N"))
    for mn in mc.ModalMasses:
        out.append("%5d %s\n" % (mn.N, _f(mn.Mass)))
    for mr in mc.Result or []:
        out.append("Natural frequency : %s Hz\n" % _f(mr.Hz))
        out.append("%5s %15s %15s %15s\n" % ("Point", "X", "Y", "M"))
        for i, d in enumerate(mr.ModalDisplacement):
            out.append(_row3(i, d))
    return "".join(out)
