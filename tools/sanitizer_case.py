"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): LinearStatic through both solve
kernels, Modal through the split pipeline and the single-kernel path, linear buckling, device-built loads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hd_b200 import _lib, frames
from hd_b200.model import flatten_model

N, nM = 8, 20
fm = flatten_model(frames.uniform_frame(N))
sec = frames.uniform_frame_sections(N, 0, nM)
loads = frames.uniform_frame_loads(N, 0, nM, 7)
t = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)
a = t.static_batch(sec, loads)
os.environ["HDB_SOLVE_LEGACY"] = "1"
b = t.static_batch(sec, loads)
del os.environ["HDB_SOLVE_LEGACY"]
assert all(np.array_equal(x, y) for x, y in zip(a, b)) and not a[0].any()
massN, mass = frames.uniform_frame_masses(N, 0, 4)
st, nModes, hz, shapes = t.modal_batch(sec[:4], massN, mass)
os.environ["HDB_MODAL_SINGLE"] = "1"
st2, nModes2, hz2, shapes2 = t.modal_batch(sec[:4], massN, mass)
del os.environ["HDB_MODAL_SINGLE"]
assert not st.any() and np.allclose(hz, hz2, rtol=1e-12)
bst, nOut, fac, shp = t.buckling_batch(sec[:2], a[3][:2, :1], 2)
assert not bst.any()
t.close()
print("sanitizer case ok")
