"""Phase-time probe of k_solve_recover (HDB_DEBUG_CLOCKS=1): which outputs cost what."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if not os.environ.get("PROBE_NOCLK"):
    os.environ["HDB_DEBUG_CLOCKS"] = "1"
import numpy as np, torch
from hd_b200 import _lib, frames
from hd_b200.model import flatten_model
N, M, nLC = 128, int(os.environ.get("PROBE_M", "4096")), 20
fm = flatten_model(frames.uniform_frame(N))
sec = frames.uniform_frame_sections(N, 0, M); loads = frames.uniform_frame_loads(N, 0, M)
t = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports)
dev = torch.device("cuda:0")
ds, dl = torch.from_numpy(sec).to(dev), torch.from_numpy(loads).to(dev)
dd = torch.empty((M, nLC, fm.nP, 3), dtype=torch.float64, device=dev)
df = torch.empty((M, nLC, fm.nB, 6), dtype=torch.float64, device=dev)
dr = torch.empty((M, nLC, fm.nP, 3), dtype=torch.float64, device=dev)
st = torch.zeros((M,), dtype=torch.int32, device=dev); fc = torch.zeros((M,), dtype=torch.int32, device=dev)
cases = {"all": (dd, df, dr), "disp only": (dd, None, None), "forces only": (None, df, None),
         "react only": (None, None, dr), "none": (None, None, None)}
if os.environ.get("PROBE_ONLY"):
    cases = {k: v for k, v in cases.items() if k in os.environ["PROBE_ONLY"].split(",")}
for name, (a, b, c) in cases.items():
    for _ in range(2):
        t.set_profiling(1)
        sys.stderr.write(f"--- {name}\n"); sys.stderr.flush()
        t.static_batch_dev(M, ds, nLC, dl, a, b, c, st, fc, 0)
        torch.cuda.synchronize()
        print(name, "kernel ms", [round(x, 4) for x in t.last_kernel_ms()], flush=True)
