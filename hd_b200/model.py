"""Host-side mirror of hd's exported data model (reference model.go:18-130, 222-315).

Field names, meaning and mutability follow the Go structs one to one so that a test
written against hd reads the same against this package:

    Model{Points, Beams[]BeamProp{N,A,J,E}, Pins, Supports}          model.go:18-77
    LoadCase{LoadNodes, PointDisplacementGlobal, BeamForces, Reactions,
             AmountLinearBuckling, LinearBucklingResult}              model.go:80-130
    LoadNode{N, Forces}                                               model.go:304-315
    ModalCase{ModalMasses, Result}, ModalMass{N, Mass}                model.go:222-228, 294-301
    ModalResult{Hz, ModalDisplacement}                                model.go:258-272
    BucklingResult{Factor, PointDisplacementGlobal}                   model.go:275-291
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

Gravity = 9.80665  # model.go:713


@dataclass
class BeamProp:
    N: List[int]
    A: float
    J: float
    E: float


@dataclass
class Model:
    Points: List[List[float]] = field(default_factory=list)
    Beams: List[BeamProp] = field(default_factory=list)
    Pins: List[List[bool]] = field(default_factory=list)
    Supports: List[List[bool]] = field(default_factory=list)

    def clone(self) -> "Model":
        return copy.deepcopy(self)

    def String(self) -> str:  # model.go:853-924
        from .report import model_string
        return model_string(self)


@dataclass
class LoadNode:
    N: int
    Forces: List[float]


@dataclass
class BucklingResult:
    Factor: float = 0.0
    PointDisplacementGlobal: Optional[list] = None


@dataclass
class LoadCase:
    LoadNodes: List[LoadNode] = field(default_factory=list)
    PointDisplacementGlobal: Optional[list] = None
    BeamForces: Optional[list] = None
    Reactions: Optional[list] = None
    AmountLinearBuckling: int = 0
    LinearBucklingResult: Optional[list] = None

    def reset(self) -> None:  # model.go:214-219
        self.PointDisplacementGlobal = None
        self.BeamForces = None
        self.Reactions = None
        self.LinearBucklingResult = None

    def String(self) -> str:  # model.go:132-211
        from .report import load_case_string
        return load_case_string(self)


@dataclass
class ModalMass:
    N: int
    Mass: float


@dataclass
class ModalResult:
    Hz: float = 0.0
    ModalDisplacement: Optional[list] = None


@dataclass
class ModalCase:
    ModalMasses: List[ModalMass] = field(default_factory=list)
    Result: Optional[list] = None

    def reset(self) -> None:  # model.go:231-233
        self.Result = None

    def String(self) -> str:  # model.go:235-255
        from .report import modal_case_string
        return modal_case_string(self)


class FlatModel:
    """hd.Model flattened into the dense arrays the C ABI takes (include/hdb200.h: points[nP][2], beam nodes and
    A, J, E per beam, pins[nB][6], supports[nP][3]).  An empty Pins slice means all rigid (model.go:344-346)."""

    def __init__(self, points, beamN, sections, pins, supports):
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        self.beamN = np.ascontiguousarray(beamN, dtype=np.int64).reshape(-1, 2)
        self.sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, 3)
        self.nP, self.nB = self.points.shape[0], self.beamN.shape[0]
        if pins is None or len(pins) == 0:
            pins = np.zeros((self.nB, 6), dtype=np.uint8)
        self.pins = np.ascontiguousarray(pins, dtype=np.uint8).reshape(-1, 6)
        self.supports = np.ascontiguousarray(supports, dtype=np.uint8).reshape(-1, 3)
        self.dof = 3 * self.nP


def flatten_model(m: Model) -> FlatModel:
    pts = np.array([[p[0], p[1]] for p in m.Points], dtype=np.float64).reshape(-1, 2)
    bn = np.array([[b.N[0], b.N[1]] for b in m.Beams], dtype=np.int64).reshape(-1, 2)
    sec = np.array([[b.A, b.J, b.E] for b in m.Beams], dtype=np.float64).reshape(-1, 3)
    pins = np.array([list(p) for p in m.Pins], dtype=np.uint8).reshape(-1, 6) if len(m.Pins) else None
    sup = np.array([list(s) for s in m.Supports], dtype=np.uint8).reshape(-1, 3)
    return FlatModel(pts, bn, sec, pins, sup)
