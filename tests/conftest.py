import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand with oracle/Makefile."""
    from oracle import hd_oracle

    hd_oracle.build()
    return hd_oracle


def pytest_terminal_summary(terminalreporter):
    """How far the GPU results are from the oracle under the PURE mod.Compare rule, and how many components needed the
    floor (tests/test_static_gpu.py::compare_rule): printed once per session and kept as gpurun_out/parity_log.json."""
    mod = sys.modules.get("test_static_gpu")
    log = getattr(mod, "PARITY_LOG", None) if mod else None
    if not log:
        return
    import json
    n = sum(r["n"] for r in log)
    exc = sum(r["exceptions"] for r in log)
    worst = max(r["worst_pure_rel_without_exceptions"] for r in log)
    terminalreporter.write_line(f"parity vs oracle (pure mod.Compare rule): {n} components, {exc} exceptions below the floor "
                                f"({100.0 * exc / max(n, 1):.4f} %), worst pure relative error elsewhere {worst:.3e}")
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_log.json"), "w") as f:
            json.dump({"components": n, "exceptions": exc, "worst_pure_rel": worst, "records": log}, f, indent=1)
    except OSError:
        pass
