"""Helpers for the text-report tests: digests of a report and of its skeleton (every %e / %f number masked),
and the report an hd.Run would print for results that are already filled in.

tests/golden/golden_text.json holds, per golden file of the reference (example/testdata/<name>), the sha256 of the
whole text and of its skeleton (made by tests/golden/make_golden.py; digests only, the text itself is not copied)."""
import hashlib
import io
import json
import os
import re

from hd_b200 import report

_NUM_E = re.compile(r" *-?\d\.\d{5}e[+-]\d{2,3}")
_NUM_F = re.compile(r" *-?\d+\.\d{5}(?![\de])")

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_text.json")) as _f:
    GOLD_TEXT = json.load(_f)


def skeleton(text: str) -> str:
    return _NUM_F.sub(" #f", _NUM_E.sub(" #e", text))


def digest(text: str) -> str:
    return hashlib.sha256(text.encode()).hexdigest()


def run_report(m, lcs, mcs) -> str:
    """What hd.Run writes (model.go:926-969) around results that are already in lcs / mcs."""
    out = io.StringIO()
    out.write(report.model_string(m) + "\n")
    if lcs:
        out.write("Linear Elastic Analysis\n\n\n")
        for lc in lcs:
            out.write(report.load_case_string(lc) + "\n")
    out.write("\n\n")
    for mc in mcs:
        out.write("Modal Analysis\n\n\n")
        out.write(report.modal_case_string(mc) + "\n")
    out.write("\n\n")
    return out.getvalue()
