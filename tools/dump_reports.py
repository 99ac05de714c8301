"""Write hd.Run's text report (GPU results) of every golden fixture to gpurun_out/reports/<name>, for diffing
against the reference's example/testdata in the build container."""
import io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
from hd_b200 import api
os.makedirs(os.path.join(ROOT, "gpurun_out", "reports"), exist_ok=True)
for name, ctor in fixtures.GOLDEN.items():
    m, lcs, mcs = ctor()
    out = io.StringIO()
    err = api.Run(out, m, lcs, mcs)
    with open(os.path.join(ROOT, "gpurun_out", "reports", name), "w") as f:
        f.write(out.getvalue())
    print(name, err)
