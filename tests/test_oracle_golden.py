"""Pin the CPU oracle against every golden report the reference holds for the path
(example/testdata/*, compared by example_test.go:12-77 and model_test.go:551).
Numbers there carry 6 significant digits (%15.5e) or 5 decimals (%15.5f)."""
import json
import os

import numpy as np
import pytest

import fixtures

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden.json")) as f:
    GOLD = json.load(f)

RTOL = 1.01e-5  # half a unit of the 6th significant digit, worst case leading digit 1


def close(x, g, what, rtol=RTOL, fixed_decimals=None):
    x, g = np.asarray(x, dtype=float), np.asarray(g, dtype=float)
    assert x.shape == g.shape, (what, x.shape, g.shape)
    scale = max(np.max(np.abs(g)), 1e-300) if g.size else 1.0
    atol = 1e-9 * scale  # entries printed as roundoff noise (e.g. -2.5e-16 next to 0.5)
    if fixed_decimals is not None:
        atol = max(atol, 0.51 * 10.0 ** (-fixed_decimals))
    assert np.allclose(x, g, rtol=rtol, atol=atol), f"{what}\n got {x}\n exp {g}"


def close_up_to_sign(x, g, what):
    x, g = np.asarray(x, dtype=float), np.asarray(g, dtype=float)
    k = np.argmax(np.abs(g))
    if np.sign(x.flat[k]) != np.sign(g.flat[k]):
        x = -x
    close(x, g, what)


@pytest.mark.parametrize("name", sorted(fixtures.GOLDEN))
def test_golden_report(oracle, name):
    m, lcs, mcs = fixtures.GOLDEN[name]()
    err, static, modal = oracle.run(m, lcs, mcs)
    assert err is None, err
    gold = GOLD[name]
    assert len(gold["static"]) == len(lcs)
    for ci, g in enumerate(gold["static"]):
        res = static[ci]
        close(res["disp"], g["disp"], f"{name} case {ci} displacement")
        close(res["forces"], g["forces"], f"{name} case {ci} beam forces")
        react = np.zeros_like(res["react"])
        for pt, v in g["react"].items():
            react[int(pt)] = v
        close(res["react"], react, f"{name} case {ci} reactions")
        if g["buckling"] is None:
            assert lcs[ci].AmountLinearBuckling == 0
        else:
            fac, shp = res["buckling"]
            assert len(fac) == len(g["buckling"]), (name, fac, g["buckling"])
            for bi, gb in enumerate(g["buckling"]):
                # factors are printed %15.5f: up to 13 digits; the eigen solver is not pinned that far
                assert abs(fac[bi] - gb["factor"]) <= 1e-7 * abs(gb["factor"]) + 0.51e-5, (name, fac[bi], gb["factor"])
                close_up_to_sign(shp[bi], gb["shape"], f"{name} case {ci} buckling shape {bi}")
    assert len(gold["modal"]) == len(mcs)
    for mi, gm in enumerate(gold["modal"]):
        hz, shapes = modal[mi]["hz"], modal[mi]["shapes"]
        assert len(hz) == len(gm), (name, hz)
        for k, gk in enumerate(gm):
            assert abs(hz[k] - gk["hz"]) <= 0.51e-5 + 1e-9 * gk["hz"], (name, k, hz[k], gk["hz"])
            close_up_to_sign(shapes[k], gk["shape"], f"{name} modal {mi} shape {k}")
