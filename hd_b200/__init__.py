"""hd_b200 — B200-native batched implementation of hd's linear-analysis hot path.

Importing the package only pulls in the host-side data model and helpers.  Every analysis
entry point goes through libhdb200.so (hd_b200/csrc, C-ABI in include/hdb200.h) and raises
loudly when that library or a CUDA device is missing: there is no CPU fallback.
"""
from .model import (BeamProp, BucklingResult, Gravity, LoadCase, LoadNode, ModalCase, ModalMass, ModalResult,
                    Model)

__all__ = ["BeamProp", "BucklingResult", "Gravity", "LoadCase", "LoadNode", "ModalCase", "ModalMass",
           "ModalResult", "Model"]
