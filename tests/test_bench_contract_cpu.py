"""The committed bench lines keep the contract the driver reads (no GPU needed): every key of the JSON line, the
roofline arithmetic, and the traffic figure bench.py attaches to it."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _last_line(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith("{")]
    return json.loads(lines[-1])


def test_final_bench_line_has_every_contract_key():
    d = _last_line(os.path.join(ROOT, "profiles", "r2f_bench.json"))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["config"]["workload"] and "model" not in d["config"]
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["kernel"] == "k_solve_bulk" and r["traffic"] is not None
    # achieved = algorithmic bytes per launch / the kernel's launch time
    alg = d["config"]["bytes_per_solve"] * r["models_per_launch"]
    assert abs(r["achieved"] - alg / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert 1.0 <= r["traffic"] / alg < 1.05, "DRAM traffic of the capture against the algorithmic bytes"
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] >= 3 * d["steps"] and not d["clocks"]["reasons"]
    assert d["parity_max_rel"] <= 1e-9


def test_traffic_file_matches_what_bench_reads():
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    with open(os.path.join(ROOT, "bench.py")) as f:
        assert 'get("k_solve_bulk_dram_bytes_per_launch")' in f.read() # the key bench.py attaches as roofline.traffic
    assert abs(t["ratio"] - t["k_solve_bulk_dram_bytes_per_launch"] / t["algorithmic_bytes_per_launch"]) < 1e-3
    assert t["algorithmic_bytes_per_launch"] == 624408 * 8192 and 1.0 <= t["ratio"] < 1.05


def test_reference_arm_line():
    d = _last_line(os.path.join(ROOT, "profiles", "r2f_bench_ref.json"))
    assert d["impl"] == "reference" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["metric"] == _last_line(os.path.join(ROOT, "profiles", "r2f_bench.json"))["metric"]


def test_every_committed_bench_file_parses():
    for p in glob.glob(os.path.join(ROOT, "profiles", "r2*_bench*.json")):
        d = _last_line(p)
        assert d["unit"] == "solves/s" and d["value"] > 0, p
