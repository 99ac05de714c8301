#!/usr/bin/env python
"""bench.py — batched frame solves/s (20 load cases) on B200, next to the CPU restatement.

One "step" = one pass of the LinearStatic hot path (assemble + factorize + 20 solves + displacement /
beam-force / reaction recovery) over one batch of synthetic uniform_frame(128) models with
randomised sections and loads: `--models-per-gpu` models per rank (default 8192 =
BASELINE.json config 5's shard at 8 GPUs, twice config 3), weak scaling over ranks, no collective on
the data path (models are independent); NCCL only reduces the status summary.

    python bench.py --gpus N --steps K --warmup W              our arm (CUDA, libhdb200.so)
    python bench.py --impl reference --gpus N --steps K ...     CPU restatement of the reference's
                                                               algorithm on all host cores (Go is not
                                                               installable here)
Prints ONE JSON line (rank 0).  Besides the contract's keys the line carries: `parity_max_rel` (eight strided models of
the timed batch re-solved by the CPU oracle after the timed region), `e2e_uniform` (the same end-to-end step with the load
cases built on the device from their line-load parameters, hdb_uniform_loads_dev: a third of the H2D bytes), `modal`
(BASELINE.json config 4: 1024 x uniform_frame(64) modal cases, 130 modes each, device resident), `single` (configs 1 and 2:
ONE model through Run, the launch-latency regime, next to the oracle and the reference's README figures) and, at N > 1,
`single_process_multi_gpu` (hdb_multi_static_batch_dev driving all N GPUs from rank 0 alone).
`--total-models T` fixes the whole job at T models (strong scaling, config 5: T = 65536) instead of 8192 per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched frame solves/s (20 load cases)"
UNIT = "solves/s"


def algorithmic_bytes(nB, dof, nP, nLC):
    """in = nB*3*8 + nLC*dof*8, out = nLC*(nP*24 + nB*48 + nP*24)."""
    return nB * 3 * 8 + nLC * dof * 8, nLC * (nP * 24 + nB * 48 + nP * 24)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def build_inputs(n_bays, model0, n_models, n_cases):
    from hd_b200 import frames
    sections = frames.uniform_frame_sections(n_bays, model0, n_models)
    loads = np.empty((n_models, n_cases, 6 * n_bays + 6))
    step = 1024
    for a in range(0, n_models, step):
        b = min(n_models, a + step)
        frames.uniform_frame_loads(n_bays, model0 + a, b - a, n_cases, out=loads[a:b])
    return sections, loads


def host_threads():
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 for every rank, which would make the
    CPU arm single-threaded at N > 1; the oracle takes its thread count explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa(gpu_index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that its pinned host buffers (first touch) and
    its copy-issuing thread sit on the GPU's NUMA node.  Matters for the e2e leg at N > 1; harmless otherwise."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n = (os.cpu_count() + 63) // 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = {64 * i + b for i, m in enumerate(masks) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def cpu_port(model, sections, loads, threads, literal=True):
    """The oracle (CPU restatement of the reference's algorithm) on `threads` host threads.  The only place (with the
    post-run parity check) where bench.py touches oracle/."""
    from oracle import hd_oracle as oracle
    fm = oracle.flatten_model(model)
    oracle.set_literal_scans(literal)
    try:
        t0 = time.perf_counter()
        status, _, _, _ = oracle.static_batch(fm, sections, loads, n_threads=threads)
        dt = time.perf_counter() - t0
    finally:
        oracle.set_literal_scans(True)
    assert not status.any(), status
    return dt


def parity_check(model, sections, loads, d_disp, d_forces, d_react, n_check=8):
    """Eight strided models of the timed batch against the CPU oracle: max |ours - oracle| / max |oracle| per array."""
    from oracle import hd_oracle as oracle
    M = sections.shape[0]
    idx = np.unique(np.linspace(0, M - 1, n_check).astype(int))
    fm = oracle.flatten_model(model)
    ost, od, of, orr = oracle.static_batch(fm, sections[idx], loads[idx])
    assert not ost.any()
    worst = 0.0
    for got, exp in ((d_disp, od), (d_forces, of), (d_react, orr)):
        g = got[idx].cpu().numpy()
        for k in range(len(idx)):
            worst = max(worst, float(np.abs(g[k] - exp[k]).max() / np.abs(exp[k]).max()))
    return worst, [int(i) for i in idx]


def modal_leg(local_rank, steps=3, N=64, M=1024):
    """BASELINE.json config 4 inside the same run (and therefore the same clock record): device-resident modal batch."""
    import torch
    from hd_b200 import _lib, frames
    from hd_b200.model import flatten_model
    fm = flatten_model(frames.uniform_frame(N))
    sections = frames.uniform_frame_sections(N, 0, M)
    massN, mass = frames.uniform_frame_masses(N, 0, M)
    dev = torch.device("cuda", local_rank)
    topo = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, device=local_rank)
    s = topo.modal_mode_count(massN)
    d_sec, d_mass = torch.from_numpy(sections).to(dev), torch.from_numpy(mass).to(dev)
    d_n = torch.zeros((M,), dtype=torch.int32, device=dev)
    d_hz = torch.zeros((M, s), dtype=torch.float64, device=dev)
    d_sh = torch.zeros((M, s, fm.nP, 3), dtype=torch.float64, device=dev)
    d_st = torch.zeros((M,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        topo.modal_batch_dev(M, d_sec, massN, d_mass, s, d_n, d_hz, d_sh, d_st, stream)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum().item()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        topo.modal_batch_dev(M, d_sec, massN, d_mass, s, d_n, d_hz, d_sh, d_st, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # parity of a few models against the oracle's non-symmetric solve of K^-1 M (frequencies, 1e-9)
    from oracle import hd_oracle as oracle
    ost, on, ohz, _ = oracle.modal_batch(oracle.flatten_model(frames.uniform_frame(N)), sections[:3], massN, mass[:3])
    hz = d_hz[:3].cpu().numpy()
    rel = float((np.abs(hz[:, :s] - ohz[:, :s]) / ohz[:, :s]).max())
    nFree = topo.info.nFree
    flops = 4 * nFree * 5 * s + 28 * s ** 3 / 3 + 2 * fm.dof * s * s + 3 * fm.dof * s
    value = M / (ms * 1e-3)
    topo.close()
    del d_sh
    return {"metric": "modal solves/s (130 modes, uniform_frame(64))", "value": value, "ms_per_step": ms, "models": M, "modes": s,
            "hz_max_rel_vs_oracle": rel,
            "fp64": {"algorithmic_mflop_per_solve": flops / 1e6, "achieved_tflops": value * flops / 1e12, "peak_tflops": 36.86,
                     "peak_source": "tools/fp64_microbench.cu on this pool's B200 (DFMA)", "frac": value * flops / 1e12 / 36.86}}


def single_leg():
    """BASELINE.json configs 1 and 2: ONE model through Run (static, 20 cases, + its modal case): the launch-latency regime.
    ms per Run for the GPU path (topology create + symbolic phase + kernels + copies every call, as hd.Run rebuilds everything
    every call), the CPU oracle, and the reference's published figure where it has one."""
    from hd_b200 import api, frames
    from oracle import hd_oracle as oracle
    out = {}

    def time_it(fn, reps):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return 1e3 * (time.perf_counter() - t0) / reps

    for name, ic, published in (("bench_cantilever(8)", 8, 0.807288), ("bench_cantilever(128)", 128, 105.636607)):
        m, lcs, mcs = frames.bench_cantilever(ic)
        gpu = time_it(lambda: api.Run(None, m.clone(), [type(l)(LoadNodes=l.LoadNodes) for l in lcs], [type(c)(ModalMasses=c.ModalMasses) for c in mcs]), 5)
        cpu = time_it(lambda: oracle.run(m.clone(), lcs, mcs), 3 if ic > 32 else 10)
        out[name] = {"gpu_ms_per_run": gpu, "oracle_ms_per_run": cpu, "go_readme_ms_per_run": published,
                     "what": "hd.Run: LinearStatic over 20 load cases + one modal case (mod/split_test.go:93-124)"}
    m = frames.uniform_frame(128)
    lcs = frames.uniform_frame_load_cases(m, 128, 0, 20)
    gpu = time_it(lambda: api.Run(None, m.clone(), [type(l)(LoadNodes=l.LoadNodes) for l in lcs], []), 5)
    cpu = time_it(lambda: oracle.run(m.clone(), lcs, []), 3)
    out["uniform_frame(128)"] = {"gpu_ms_per_run": gpu, "oracle_ms_per_run": cpu, "go_readme_ms_per_run": None,
                                 "what": "hd.Run: LinearStatic over 20 load cases, one model (BASELINE.json config 2)"}
    return out


def modal_main(args, rank, local_rank, world):
    """Auxiliary line for BASELINE.json config 4: ModalCases on a batch of 1024 uniform_frame(64) models,
    one modal case each, 130 modes per model (frequencies + shapes). Not the driver's headline metric."""
    import torch
    from hd_b200 import _lib, frames
    from hd_b200.model import flatten_model
    N = 64 if args.bays == 128 else args.bays
    M = 1024 if args.models_per_gpu == 8192 else args.models_per_gpu
    fm = flatten_model(frames.uniform_frame(N))
    sections = frames.uniform_frame_sections(N, rank * M, M)
    massN, mass = frames.uniform_frame_masses(N, rank * M, M)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    topo = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, device=local_rank)
    s = topo.modal_mode_count(massN)
    d_sec, d_mass = torch.from_numpy(sections).to(dev), torch.from_numpy(mass).to(dev)
    d_n = torch.zeros((M,), dtype=torch.int32, device=dev)
    d_hz = torch.zeros((M, s), dtype=torch.float64, device=dev)
    d_sh = torch.zeros((M, s, fm.nP, 3), dtype=torch.float64, device=dev)
    d_st = torch.zeros((M,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        topo.modal_batch_dev(M, d_sec, massN, d_mass, s, d_n, d_hz, d_sh, d_st, stream)
    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    assert int(d_st.abs().sum().item()) == 0
    topo.set_profiling(2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    kms = topo.last_kernel_ms()
    nFree = topo.info.nFree
    flops = 4 * nFree * 5 * s + 28 * s ** 3 / 3 + 2 * fm.dof * s * s + 3 * fm.dof * s
    value = M * args.steps / (ms * 1e-3)
    cpu = None
    if not args.no_cpu:
        from oracle import hd_oracle as oracle  # CPU baseline leg only
        threads = host_threads()
        sample = args.cpu_sample or 2 * threads
        t0 = time.perf_counter()
        ost, _, _, _ = oracle.modal_batch(oracle.flatten_model(frames.uniform_frame(N)), sections[:sample], massN, mass[:sample], n_threads=threads)
        dt = time.perf_counter() - t0
        cpu = {"value": sample / dt, "unit": "modal solves/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} models, dense non-symmetric eigen solve of K^-1 M as the reference does"}
    if rank == 0:
        print(json.dumps({"metric": "modal solves/s (130 modes, uniform_frame(64))", "value": value, "unit": "modal solves/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                          "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"uniform_frame({N}) modal, {M} models per GPU, {s} modes each"},
                          "kernel_ms_total": dict(zip(["k_assemble", "k_factor", "k_solve_pipe", "k_modal_loads", "k_modal_eigen"], kms)),
                          "fp64": {"algorithmic_mflop_per_solve": flops / 1e6, "achieved_tflops": value * flops / 1e12,
                                   "peak_tflops": 36.86, "peak_source": "tools/fp64_microbench.cu on this pool's B200",
                                   "frac": value * flops / 1e12 / 36.86},
                          "cpu_baseline": cpu}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--models-per-gpu", type=int, default=8192)
    ap.add_argument("--bays", type=int, default=128)
    ap.add_argument("--cases", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=0, help="models in the CPU baseline sample (0 = auto)")
    ap.add_argument("--workload", default="static", choices=["static", "modal"],
                    help="static = the headline LinearStatic batch (default); modal = BASELINE.json config 4 (auxiliary line)")
    ap.add_argument("--total-models", type=int, default=0, help="strong scaling: the whole job has this many models (config 5: 65536)")
    ap.add_argument("--no-extras", action="store_true", help="skip the modal / single-model / multi-GPU-in-one-process legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "modal":
        return modal_main(args, rank, local_rank, world)
    N, nLC, M = args.bays, args.cases, args.models_per_gpu
    strong = args.total_models > 0
    if strong:
        from hd_b200 import _lib
        if args.total_models % world:
            raise SystemExit("--total-models must be a multiple of the number of ranks")
        first, M = _lib.shard_range(args.total_models, world, rank)  # the library's contiguous-range rule; == (rank * M, M)
        assert first == rank * M

    from hd_b200 import frames
    from hd_b200.model import flatten_model
    model = frames.uniform_frame(N)
    fm = flatten_model(model)
    nP, nB, dof = fm.nP, fm.nB, fm.dof
    b_in, b_out = algorithmic_bytes(nB, dof, nP, nLC)
    workload = f"uniform_frame({N}) x {nLC} load cases, randomised sections+loads, {M} models per GPU"
    config = {"workload": workload, "models_per_gpu": M, "n_bays": N, "load_cases": nLC, "dof": dof,
              "bytes_per_solve": b_in + b_out,
              "l2_policy": f"working set per step {(b_in + b_out) * M / 1e9:.2f} GB per GPU >> 126 MB L2, no flush needed",
              "parallelism": f"models sharded over {world} rank(s), no data-path collective",
              "scaling": "strong" if strong else "weak"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = host_threads()
        sample = args.cpu_sample or 8 * threads
        sections, loads = build_inputs(N, 0, sample, nLC)
        for _ in range(args.warmup):
            cpu_port(model, sections[:threads], loads[:threads], threads)
        t = 0.0
        for _ in range(args.steps):
            t += cpu_port(model, sections, loads, threads)
        value = sample * args.steps / t
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{sample} models per step of the same workload, literal restatement incl. the "
                                           f"O(nSup*nnz) reaction scan (model.go:481), OpenMP over models"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "CPU restatement of the reference algorithm in C (oracle/), not Go: no Go toolchain in this image"}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch
    import torch.distributed as dist
    from hd_b200 import _lib

    if not torch.cuda.is_available() or _lib.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device — hd_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else 0  # N = 1 keeps every core for the CPU baseline
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sections, loads = build_inputs(N, rank * M, M, nLC)
    topo = _lib.Topology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, device=local_rank)
    d_sec = torch.from_numpy(sections).to(dev)
    d_loads = torch.from_numpy(loads).to(dev)
    d_disp = torch.empty((M, nLC, nP, 3), dtype=torch.float64, device=dev)
    d_forces = torch.empty((M, nLC, nB, 6), dtype=torch.float64, device=dev)
    d_react = torch.empty((M, nLC, nP, 3), dtype=torch.float64, device=dev)
    d_status = torch.zeros((M,), dtype=torch.int32, device=dev)
    d_fail = torch.zeros((M,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        topo.static_batch_dev(M, d_sec, nLC, d_loads, d_disp, d_forces, d_react, d_status, d_fail, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the first warm-up step to the end of the timed region (one step is a few ms, far
    # below nvidia-smi's sampling period), and over untimed repeats of the same step if that is still too short
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_warm = time.time()
    for _ in range(args.warmup):
        step()
    barrier()
    assert int(d_status.abs().sum().item()) == 0, "a model failed"
    topo.set_profiling(2)  # accumulate per-kernel CUDA-event times over the timed region
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kernel_ms = topo.last_kernel_ms()  # sums over every launch of the timed region
    topo.set_profiling(0)
    launches = _lib.launch_count() - launches0
    extra = 0
    while time.time() - t_warm < 0.6:  # same step, untimed: keeps the GPU under the same load for the sampler
        step()
        extra += 1
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"warm-up + timed region + {extra} untimed repeats of the same step (nvidia-smi -lms 20)"
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    bad = d_status.abs().sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(bad, op=dist.ReduceOp.SUM)  # the only collective: a reduced status summary
    ms = float(t_ms.item())
    value = world * M * args.steps / (ms * 1e-3)

    # ---- the timed batch is correct: eight strided models against the CPU oracle (after the timed region) ----
    parity, parity_idx = parity_check(model, sections, loads, d_disp, d_forces, d_react)
    assert parity <= 1e-9, f"parity lost: {parity}"

    # ---- e2e: the public C-ABI call with pinned HOST buffers; H2D and D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        keep = []

        def pinned(shape, dtype=np.float64):
            a, o = _lib.pinned_empty(shape, dtype)
            keep.append(o)
            return a
        h_sec, h_loads = pinned(sections.shape), pinned(loads.shape)
        h_sec[...] = sections
        h_loads[...] = loads
        h_disp, h_forces, h_react = pinned((M, nLC, nP, 3)), pinned((M, nLC, nB, 6)), pinned((M, nLC, nP, 3))
        h_status, h_fail = pinned((M,), np.int32), pinned((M,), np.int32)
        e_steps = max(2, min(args.steps, 4))
        topo.static_batch_raw(M, h_sec, nLC, h_loads, h_disp, h_forces, h_react, h_status, h_fail)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            topo.static_batch_raw(M, h_sec, nLC, h_loads, h_disp, h_forces, h_react, h_status, h_fail)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert not h_status.any()
        assert np.array_equal(h_react, d_react.cpu().numpy()), "host path and device path disagree"
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * M * e_steps / float(t_e.item()), "unit": UNIT, "h2d_bytes_per_step": int(M * b_in),
               "d2h_bytes_per_step": int(M * (b_out + 8)), "steps": e_steps,
               "path": "hdb_static_batch on cudaHostAlloc buffers: chunked H2D -> 3 kernels -> D2H, three streams"}
        # the same step with the load cases built on the device from their parameters (mod.LoadUniform on the device,
        # hdb_uniform_loads_dev): H2D carries sections + line loads + the lateral force instead of dense load vectors
        e2e_u = None
        try:
            q, fx = frames.uniform_frame_load_params(N, rank * M, M, nLC)
            qq = np.zeros((M, nLC, N, 2))
            qq[..., 1] = q
            ff = np.zeros((M, nLC, 1, 3))
            ff[:, :, 0, 0] = fx
            h_q, h_f = pinned(qq.shape), pinned(ff.shape)
            h_q[...] = qq
            h_f[...] = ff
            ubeam = np.array([frames.girder_index(i) for i in range(1, N + 1)], dtype=np.int32)
            uproj = np.ones(N, dtype=np.uint8)
            d_q, d_f = torch.empty(qq.shape, dtype=torch.float64, device=dev), torch.empty(ff.shape, dtype=torch.float64, device=dev)
            t_q, t_f, t_sec = torch.from_numpy(h_q), torch.from_numpy(h_f), torch.from_numpy(h_sec)
            t_disp, t_forces, t_react = torch.from_numpy(h_disp), torch.from_numpy(h_forces), torch.from_numpy(h_react)

            def step_u():
                d_sec.copy_(t_sec, non_blocking=True)
                d_q.copy_(t_q, non_blocking=True)
                d_f.copy_(t_f, non_blocking=True)
                topo.uniform_loads_dev(M, nLC, ubeam, uproj, d_q, [1], d_f, d_loads, stream)
                topo.static_batch_dev(M, d_sec, nLC, d_loads, d_disp, d_forces, d_react, d_status, d_fail, stream)
                t_disp.copy_(d_disp, non_blocking=True)
                t_forces.copy_(d_forces, non_blocking=True)
                t_react.copy_(d_react, non_blocking=True)
                torch.cuda.synchronize()
            react_ref = h_react.copy()
            step_u()
            assert np.array_equal(h_react, react_ref), "device-built loads differ from the host-built ones"
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                step_u()
            dtu = time.perf_counter() - t0
            t_u = torch.tensor([dtu], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_u, op=dist.ReduceOp.MAX)
            e2e_u = {"value": world * M * e_steps / float(t_u.item()), "unit": UNIT,
                     "h2d_bytes_per_step": int(sections.nbytes + qq.nbytes + ff.nbytes), "d2h_bytes_per_step": int(M * b_out),
                     "path": "pinned sections + line-load parameters -> hdb_uniform_loads_dev -> hdb_static_batch_dev -> pinned results"}
        except Exception as ex:  # the headline does not depend on this leg
            e2e_u = {"error": repr(ex)}
        del keep

    # ---- one process driving every GPU of the box (hdb_multi_*), rank 0 alone; the other ranks wait at the barrier ----
    spmg = None
    if world > 1 and not args.no_extras:
        barrier()
        # the other ranks must not wait inside an NCCL barrier here: its kernel spins on their GPUs, which this leg uses too
        # (two processes time-slice a GPU); they block on the host store instead
        store = dist.distributed_c10d._get_default_store()
        if rank != 0:
            store.wait(["single_process_leg_done"])
        if rank == 0:
            try:
                mt = _lib.MultiTopology(fm.points, fm.beamN, fm.sections, fm.pins, fm.supports, list(range(world)))
                mm = M
                bufs = []
                for d in range(world):
                    dv = torch.device("cuda", d)
                    bufs.append(dict(sec=torch.from_numpy(sections[:mm]).to(dv), loads=torch.from_numpy(loads[:mm]).to(dv),
                                     disp=torch.empty((mm, nLC, nP, 3), dtype=torch.float64, device=dv),
                                     forces=torch.empty((mm, nLC, nB, 6), dtype=torch.float64, device=dv),
                                     react=torch.empty((mm, nLC, nP, 3), dtype=torch.float64, device=dv),
                                     st=torch.zeros((mm,), dtype=torch.int32, device=dv), fc=torch.zeros((mm,), dtype=torch.int32, device=dv)))
                col = lambda k: [b[k] for b in bufs]
                run = lambda: mt.static_batch_dev([mm] * world, col("sec"), nLC, col("loads"), col("disp"), col("forces"), col("react"), col("st"), col("fc"))
                for _ in range(3):
                    run()
                mt.synchronize()
                t0 = time.perf_counter()
                reps = 10
                for _ in range(reps):
                    run()
                mt.synchronize()
                dts = time.perf_counter() - t0
                summ = mt.gather_summary([mm] * world, col("st"), col("disp"), nLC)
                same = all(torch.equal(bufs[0]["react"].cpu(), b["react"].cpu()) for b in bufs[1:])
                spmg = {"value": world * mm * reps / dts, "unit": UNIT, "models_per_gpu": mm, "devices": world, "timing": "host clock around 10 enqueue rounds + device sync",
                        "nccl_summary": {"failed": float(summ[:, 0].sum()), "max_abs_disp": float(summ[:, 1].max()), "models": float(summ[:, 2].sum())},
                        "shards_identical": bool(same)}
                mt.close()
                del bufs
            except Exception as ex:
                spmg = {"error": repr(ex)}
            store.set("single_process_leg_done", "1")
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_solve_bulk) ----
    peaks, peak_src = measured_peaks()
    sub = int(os.environ.get("HDB_SUB_BATCH", "8192"))
    n_launch = args.steps * ((M + sub - 1) // sub)
    names = ["k_assemble_quad", "k_factor", "k_solve_bulk"]
    roof = None
    if len(kernel_ms) >= 3 and kernel_ms[2] > 0:
        per_launch_models = M * args.steps / n_launch
        dur = kernel_ms[2] * 1e-3 / n_launch
        achieved = per_launch_models * (b_in + b_out) / dur / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("k_solve_bulk_dram_bytes_per_launch")
        roof = {"bound": "hbm", "kernel": "k_solve_bulk", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                "launch_ms": dur * 1e3, "models_per_launch": per_launch_models,
                "kernel_ms_total": dict(zip(names, kernel_ms)), "kernel_share": {n: k / sum(kernel_ms) for n, k in zip(names, kernel_ms)}}

    # ---- CPU baseline: the oracle on the box's host cores, bounded sample of the same workload ----
    cpu = None
    if not args.no_cpu and world == 1:
        threads = host_threads()
        sample = args.cpu_sample or 8 * threads
        dt = cpu_port(model, sections[:sample], loads[:sample], threads, literal=True)
        dt_fast = cpu_port(model, sections[:sample * 8], loads[:sample * 8], threads, literal=False)
        cpu = {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {sample} models of the same batch, literal C restatement of the reference algorithm "
                         f"(incl. its O(nSup*nnz) reaction scan), OpenMP over models",
               "tuned_value": sample * 8 / dt_fast,
               "tuned_note": "same oracle with the reaction scan replaced by per-row lists (bit-identical results)",
               "go_published": "README.md:63 BenchmarkRun/__128: 9.47 runs/s (static+modal, cantilever, unknown amd64, 1 core)"}

    extras = {}
    if not args.no_extras and world == 1:
        try:
            extras["modal"] = modal_leg(local_rank)
        except Exception as ex:
            extras["modal"] = {"error": repr(ex)}
        try:
            extras["single"] = single_leg()
        except Exception as ex:
            extras["single"] = {"error": repr(ex)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches), "failed_models": int(bad.item()), "numa_bound_cpus": int(numa_cpus),
            "hbm_frac_whole_step": value / world * (b_in + b_out) / 1e9 / peaks["hbm_gbs"],
            "parity_max_rel": parity, "parity_models": parity_idx, "e2e_uniform": e2e_u if not args.no_e2e else None,
            "single_process_multi_gpu": spmg}
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
