"""Parse the reference's golden text reports into numeric known-answer vectors.

Run in the build container only (it reads /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

Source: /root/reference/example/testdata/<name> (written by hd.Run through LoadCase.String,
ModalCase.String — model.go:132-211, 235-255 — and compared byte-for-byte by
example_test.go:12-77 and model_test.go:551).  Numbers are printed with %15.5e / %15.5f, i.e.
6 significant digits.  Output: tests/golden/golden.json with, per file,
  static: [ {disp[nP][3], forces[nB][6], react{point:[3]}, buckling:[{factor, shape[nP][3]}] | null} ]
  modal:  [ [ {hz, shape[nP][3]} ] ]
Only numbers are extracted; no reference text or source is copied.  For the report writer
(hd_b200/report.py) golden_text.json adds two sha256 digests per file: of the whole text and of its skeleton
(tests/report_util.skeleton: every %e / %f number masked).
"""
from __future__ import annotations

import json
import os
import sys

REF = "/root/reference/example/testdata"
NAMES = ["beam", "truss", "g-beam", "truss-modal", "truss-modal-rotate", "beam-modal", "beam-modal-rotate",
         "truss-with-buckling", "beam-modal-3mass", "beam-frame-modal", "beam-buckling", "beam-dc-part1",
         "beam-dc-part2", "model.String"]


def _floats(parts):
    return [float(x) for x in parts]


def parse(text: str):
    lines = text.split("\n")
    static, modal = [], []
    i = 0
    cur = None      # current load case dict
    curm = None     # current modal case list
    state = None
    while i < len(lines):
        ln = lines[i]
        s = ln.strip()
        if s == "Load case:":
            cur = {"disp": [], "forces": [], "react": {}, "buckling": None}
            static.append(cur)
            curm = None
            state = "loads"
        elif s == "Modal case:":
            curm = []
            modal.append(curm)
            cur = None
            state = "masses"
        elif s.startswith("Point displacement in global system coordinate"):
            state = "disp"
            i += 1  # header
        elif s.startswith("Local force in beam"):
            state = "forces"
            i += 3
        elif s.startswith("Reaction on support"):
            state = "react"
            i += 1
        elif s.startswith("Linear buckling result:"):
            if s.endswith("is not calculated"):
                state = None
            elif "haven`t valid data" in s:
                cur["buckling"] = []
                state = None
            else:
                cur["buckling"] = []
                state = "buckling"
        elif s.startswith("Linear buckling factor:"):
            cur["buckling"].append({"factor": float(s.split(":")[1]), "shape": []})
            state = "bshape"
            i += 1
        elif s.startswith("Natural frequency"):
            hz = float(s.split(":")[1].replace("Hz", ""))
            curm.append({"hz": hz, "shape": []})
            state = "mshape"
            i += 1
        elif s == "" or s.startswith("Point ") or s.startswith("Index "):
            pass
        else:
            parts = s.split()
            try:
                if state == "disp" and len(parts) == 4:
                    cur["disp"].append(_floats(parts[1:]))
                elif state == "forces" and len(parts) == 7:
                    cur["forces"].append(_floats(parts[1:]))
                elif state == "react" and len(parts) == 4:
                    cur["react"][parts[0]] = _floats(parts[1:])
                elif state == "bshape" and len(parts) == 4:
                    cur["buckling"][-1]["shape"].append(_floats(parts[1:]))
                elif state == "mshape" and len(parts) == 4:
                    curm[-1]["shape"].append(_floats(parts[1:]))
            except ValueError:
                pass
        i += 1
    return {"static": static, "modal": modal}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference testdata not found: run this in the build container")
    out = {}
    for n in NAMES:
        with open(os.path.join(REF, n)) as f:
            out[n] = parse(f.read())
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=None, separators=(",", ":"))
        f.write("\n")
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    text = {}
    dst_text = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_text.json")
    if not os.path.exists(dst_text):
        with open(dst_text, "w") as f:
            f.write("{}\n")  # report_util loads it at import
    import report_util

    for n in NAMES:
        with open(os.path.join(REF, n)) as f:
            t = f.read()
        text[n] = {"sha256": report_util.digest(t), "skeleton_sha256": report_util.digest(report_util.skeleton(t)),
                   "lines": t.count("\n")}
    with open(dst_text, "w") as f:
        json.dump(text, f, indent=1, sort_keys=True)
        f.write("\n")
    for n in NAMES:
        print(n, "static cases:", len(out[n]["static"]), "modal cases:", [len(x) for x in out[n]["modal"]])


if __name__ == "__main__":
    main()
