#!/usr/bin/env python3
"""bench.py -- CZ-fidelity Lindblad gate evaluations per second on N B200s.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle port of
                                                           the QuTiP mesolve path) on the host cores

One "step" = one pass of the hot path over one batch of synthetic pulse candidates.  Headline workload
(BASELINE.json configs[1], SURVEY.md section 8d "config 2"): Jandura-Pupillo bang-bang CZ, 5 piecewise-constant
segments, two-atom 3-level Lindblad model with the medium apparatus's noise rates; per GPU 262,144 candidates
(--candidates) drawn inside the reference optimiser's bounds (RG/optimize_cz_gate.py:576-587), seed 1234 + rank.
Every rank evaluates its own shard (no data-path collective; the only exchange is the all-gather of per-candidate
scores, timed inside the step) -> "scaling": "weak".

The same JSON line (rank 0) also carries
  * "parity":  >= 64 randomly chosen candidates of the timed batch against the expm oracle (outside the timed region);
  * "configs": every other BASELINE.json configuration through the product's public API on one GPU (configs 1, 3,
               4, 5), and at --gpus N the sharded product paths (sharding.simulate_sharded, the code behind
               optimize_cz_gate(population_shards=...)) for configs 3 and 4 -- strong scaling, NCCL;
  * "roofline": the dominant kernel against the measured DFMA peak, and the whole step against the same peak.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cz_lindblad_gate_evals_per_sec"
UNIT = "evals/s"
B_PER_GPU = 262144         # default candidates per GPU per step (--candidates); 4.6 waves of resident threads
N_SEG = 5
# Runtime lines saved in the reference's notebooks (author's laptop, one process)
NOTEBOOK_EVALS_PER_S = {"lp_noiseless": 123, "smooth_jp_noiseless": 5.4, "lp_noisy_older_notebook": 59,
                        "jp_noisy_older_notebook": 25}


# ---------------------------------------------------------------------------- workload
MEDIUM = dict(  # the reference's 'medium' apparatus with full noise (RG/optimize_cz_gate.py:196-213, 885-891)
    laser_1_power=50e-6, laser_1_waist=50e-6, laser_2_power=0.3, laser_2_waist=50e-6, Delta_e=2 * np.pi * 1e9,
    laser_1_linewidth_hz=100.0, laser_2_linewidth_hz=100.0, polarization_purity_1=0.99, polarization_purity_2=0.99,
    temperature=2e-6, spacing_factor=2.8, n_rydberg=70, species="Rb87", tweezer_power=0.020, tweezer_waist=0.8e-6,
    B_field=1e-4, NA=0.5, include_noise=True)


def make_candidates(n, seed):
    """Bang-bang candidates inside the reference optimiser's bounds (RG/optimize_cz_gate.py:576-587):
    omega_tau~U[5,40], 4 sorted switching fractions~U[0.01,0.99] (times = fraction*omega_tau, as
    _build_jp_bangbang_inputs does, :501-516), 5 phases~U[-pi,pi]."""
    rng = np.random.default_rng(seed)
    ot = rng.uniform(5.0, 40.0, n)
    fr = np.sort(rng.uniform(0.01, 0.99, (n, N_SEG - 1)), axis=1)
    ph = rng.uniform(-np.pi, np.pi, (n, N_SEG))
    return dict(omega_tau=ot, switching_times=fr * ot[:, None], phases=ph)


def host_plan(cand, idx):
    """Evaluator inputs of a few candidates via the NumPy prologue (for the CPU path and the oracle check)."""
    from noisyquantumsimulator_b200.rydberg_gates import prologue
    sub = {k: v[idx] for k, v in cand.items()}
    plan = prologue.build_plan("jp", B=len(idx), warn=False, **MEDIUM, **sub)
    return plan.points, plan.segs


def population(proto, n, seed=42):
    """Candidates drawn uniformly inside the reference optimiser's search box (RG/optimize_cz_gate.py:544-643):
    protocol keywords of simulate_CZ_gate_batch for n candidates."""
    from noisyquantumsimulator_b200.rydberg_gates import batched_objective as bo
    bounds, _, _ = bo.bounds_and_x0(proto, 5)
    lo, hi = np.array(bounds).T
    X = lo[:, None] + (hi - lo)[:, None] * np.random.default_rng(seed).uniform(size=(len(bounds), n))
    return bo.protocol_kwargs(proto, X, 5)


def sweep_grid():
    """Config 3: 16 x 16 x 16 grid over second-leg power (Omega), Delta/Omega and the pulse area."""
    p2g, dg, og = np.meshgrid(np.geomspace(0.05, 2.0, 16), np.linspace(0.20, 0.50, 16), np.linspace(3.5, 5.5, 16),
                              indexing="ij")
    return dict(laser_2_power=p2g.ravel().copy(), delta_over_omega=dg.ravel().copy(), omega_tau=og.ravel().copy(),
                temperature=10e-6)


# ---------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU every ~10 ms DURING the timed region
    (NVML in-process; falls back to polling nvidia-smi)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self.stop_flag = False
        self.source = "nvml"

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for k, bit in names.items():
                if mask & bit:
                    self.reasons.add(k)
            time.sleep(0.01)

    def _smi_loop(self):
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                    str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                r = [x.strip() for x in o.strip().splitlines()[0].split(",")]
                self.sm.append(float(r[0])); self.sm_max = float(r[1]); self.power.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "power_w_max": float(max(self.power)) if self.power else None,
                "source": self.source}


# ---------------------------------------------------------------------------- CPU (reference) path
_RHS = "scipy csr"


def _install_fast_rhs():
    """The port's right-hand side d(vec rho)/dt = L vec(rho) as a compiled CSR matvec (numba) instead of a
    scipy.sparse product: QuTiP's own RHS is compiled (Cython CSR), so this is the fairer baseline.  Same
    integrator, same tolerances, same tlist.  Falls back to scipy.sparse when numba is unavailable."""
    global _RHS
    from oracle import rydcz_oracle as orc
    if getattr(orc, "_fast_rhs_installed", False):
        return
    try:
        import numba
        import scipy.sparse as sp
        from scipy.integrate import ode

        @numba.njit(cache=True)
        def csr_mv(data, indices, indptr, y, out):
            for i in range(indptr.shape[0] - 1):
                acc = 0.0 + 0.0j
                for k in range(indptr[i], indptr[i + 1]):
                    acc += data[k] * y[indices[k]]
                out[i] = acc
            return out

        def advance(gen, y0, dt, n_tlist, nsteps):
            G = sp.csr_matrix(gen)
            data, ind, ptr = G.data.astype(np.complex128), G.indices.astype(np.int64), G.indptr.astype(np.int64)
            n = len(y0)
            solver = ode(lambda t, y: csr_mv(data, ind, ptr, y, np.empty(n, dtype=np.complex128)))
            solver.set_integrator("zvode", method="adams", atol=1e-10, rtol=1e-8, nsteps=nsteps)
            solver.set_initial_value(y0.astype(complex), 0.0)
            for t in np.linspace(0.0, dt, n_tlist)[1:]:
                solver.integrate(t)
                if not solver.successful():
                    raise RuntimeError("zvode failed")
            return solver.y

        advance(np.array([[0.0, 1.0], [-1.0, 0.0]], dtype=complex), np.array([1.0, 0.0]), 0.1, 2, 100)   # compile
        orc._zvode_advance = advance
        _RHS = "numba csr"
    except Exception:
        _RHS = "scipy csr"
    orc._fast_rhs_installed = True


def _cpu_one(args):
    """One gate evaluation the way the reference does it: zvode/Adams at atol=1e-10, rtol=1e-8,
    nsteps=10000, 51 tlist points per bang-bang segment, 4 inputs (RG/simulation.py:1890-1924)."""
    from oracle import rydcz_oracle as orc
    pt, seg = args
    rates = orc.NoiseRates(gamma_r=2 * pt[4], gamma_bbr=pt[5] - pt[4], gamma_phi_laser=pt[6],
                           gamma_scatter_intermediate=pt[7])
    segs = [orc.Segment(s[0], complex(s[1], s[2]), s[3], 51) for s in seg if s[0] >= 1e-18]
    fin = orc.propagate(3, segs, pt[0], pt[1], rates, mode="zvode", nsteps=10000)
    fid, avg, info = orc.cz_fidelity_mixed(3, fin, "compat")
    return avg


def _cpu_one_expm(args):
    """The same evaluation with the oracle's exact matrix exponentials (what the kernel is checked against):
    (avg F, F11 with phase penalty, cz phase fidelity, F01)."""
    from oracle import rydcz_oracle as orc
    pt, seg = args
    rates = orc.NoiseRates(gamma_r=2 * pt[4], gamma_bbr=pt[5] - pt[4], gamma_phi_laser=pt[6],
                           gamma_scatter_intermediate=pt[7])
    segs = [orc.Segment(s[0], complex(s[1], s[2]), s[3], 51) for s in seg if s[0] >= 1e-18]
    fid, avg, info, _ = orc.evaluate(3, segs, pt[0], pt[1], rates, "exact", "coherence")
    return avg, fid["11"], info["cz_phase_fidelity"], fid["01"]


def _single_thread_blas():
    """Pool workers run one evaluation each: BLAS / OpenMP teams inside them would oversubscribe the cores."""
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass


def _pool_map(fn, work, procs):
    import multiprocessing as mp
    if procs > 1:
        with mp.get_context("fork").Pool(procs, initializer=_single_thread_blas) as pool:
            return pool.map(fn, work, chunksize=1)
    return [fn(w) for w in work]


def cpu_path_rate(pts, segs, n_sample, procs):
    """Evaluate the first n_sample candidates with the oracle's zvode restatement on `procs`
    host processes; returns evals/s."""
    _install_fast_rhs()
    work = [(pts[i], segs[i]) for i in range(n_sample)]
    t0 = time.perf_counter()
    _pool_map(_cpu_one, work, procs)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = max(cores, 8)
    cand = make_candidates(B_PER_GPU, 1234)
    pts, segs = host_plan(cand, np.arange(min(B_PER_GPU, n_sample * (args.steps + 1))))
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_path_rate(pts, segs, min(cores, 4), cores)
    times = []
    for k in range(args.steps):
        _, dt = cpu_path_rate(pts[k * n_sample:], segs[k * n_sample:], n_sample, cores)
        times.append(dt)
    total = sum(times)
    value = args.steps * n_sample / total
    sample = (f"{n_sample} of the {B_PER_GPU} bang-bang candidates per step, oracle zvode/Adams restatement of the "
              f"reference's mesolve path (QuTiP itself is not installable), RHS = {_RHS} matvec of the 81x81 "
              f"Liouvillian, {cores} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "jp_bangbang_5seg_lindblad_3level", "candidates_per_step": n_sample,
                   "host_cores": cores, "rhs": _RHS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "evals_per_s_per_core": value / cores,
                         "reference_notebook_evals_per_s": NOTEBOOK_EVALS_PER_S},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- CUDA path
class Bench:
    """Shared state of the CUDA arm: evaluator, device, distributed plumbing."""

    def __init__(self, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        from noisyquantumsimulator_b200.engine import get_evaluator
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ev = get_evaluator(local_rank)        # the evaluator the public API uses on this device
        # one process per GPU on a multi-socket host: keep this rank's threads and pinned staging next to its GPU
        self.bound_cpus = self.ev.bind_host_to_gpu() if world > 1 else 0
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def status_histogram(ev, n):
    from noisyquantumsimulator_b200 import _lib
    st = ev.point_status(n)
    names, counts = np.unique(st & 0x0f, return_counts=True)
    h = {_lib.STATUS.get(int(k), str(int(k))): int(c) for k, c in zip(names, counts)}
    back = int(((st & _lib.STATUS_HANDED_BACK) != 0).sum())
    if back:
        h["handed_back"] = back
    return h


def measure_config(b, name, proto, n_points, kwargs, peak, n_shots=0, reps=5, note=None):
    """One BASELINE configuration on this GPU through the public API.
      value    : evals/s with the apparatus block resident in HBM (rydcz_simulate_device: prologue kernel, planner,
                 evaluate kernels; CUDA events, L2 flushed between iterations);
      e2e      : evals/s through simulate_CZ_gate_batch from NumPy arrays to NumPy result columns (the optimiser's
                 metric columns), host<->device copies inside the timed region;
      kernel_ms: device span of the evaluate kernels of the last pipelined call + its coefficient kernel;
      frac     : kernel-counted FP64 flops / kernel_ms against the measured DFMA peak."""
    torch = b.torch
    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.rydberg_gates import simulation as sim
    ev = b.ev
    kw = sim.shot_kwargs(dict(kwargs), n_shots)
    n_evals = n_points * max(n_shots, 1)
    model, cols = sim.device_prologue_inputs(proto, n_points, kw)
    app = np.empty((len(cols), n_points))
    for k, v in enumerate(cols):
        app[k] = v
    d_app = torch.from_numpy(app).to(b.dev)
    d_out = torch.empty((n_points, 16), dtype=torch.float64, device=b.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for it in range(reps + 2):
        b.flush.zero_()
        e0.record()
        ev.simulate_device(model, d_app, n_shots, 1234, out=d_out)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ms.append(e0.elapsed_time(e1))
    dev_ms = float(np.mean(ms))
    k_ms, c_ms = ev.last_kernel_ms(), ev.last_coef_ms()
    planned, back = ev.cluster_stats()
    hist = status_histogram(ev, n_evals if (n_shots and planned) else n_points)
    raw = d_out.cpu().numpy()
    flops = float(raw[:, 15].sum())
    kernel_ms = k_ms + (c_ms if planned else 0.0)

    def api():
        return rg.simulate_CZ_gate_batch(proto, B=n_points, n_shots=n_shots, seed=1234, device=b.local, warn=False,
                                         copy=False, outputs=sim.METRIC_OUTPUTS, **kwargs)
    api(); api()
    t0 = time.perf_counter()
    for _ in range(reps):
        res = api()
    api_s = (time.perf_counter() - t0) / reps
    assert np.allclose(res.avg_fidelity, raw[:, 0], rtol=0, atol=1e-12, equal_nan=True)
    d = {"name": name, "protocol": proto, "points": n_points, "shots": n_shots, "evals_per_step": n_evals,
         "value": n_evals / (dev_ms * 1e-3), "ms_per_step": dev_ms, "e2e": n_evals / api_s, "e2e_ms": api_s * 1e3,
         "kernel_ms": kernel_ms, "coef_ms": c_ms if planned else 0.0, "flops": flops,
         "tflops": flops / (kernel_ms * 1e-3) * 1e-12, "frac": flops / (kernel_ms * 1e-3) * 1e-12 / peak,
         "step_frac": flops / (dev_ms * 1e-3) * 1e-12 / peak, "paths": hist,
         "mean_fidelity": float(np.nanmean(raw[:, 0]))}
    if note:
        d["note"] = note
    return d


def measure_sharded(b, name, proto, n_points, kwargs, n_shots=0, reps=5):
    """Strong scaling through the product's multi-GPU path: sharding.simulate_sharded splits ONE batch over the
    ranks (the code behind optimize_cz_gate(population_shards=...)); every rank ends with all metric columns.
    Timed from host arrays to host arrays, barrier + synchronize on both sides, max over ranks."""
    from noisyquantumsimulator_b200 import sharding
    from noisyquantumsimulator_b200.rydberg_gates.simulation import METRIC_OUTPUTS

    def call():
        return sharding.simulate_sharded(proto, n_points, rank=b.rank, world=b.world, device=b.local, n_shots=n_shots,
                                         seed=1234, outputs=METRIC_OUTPUTS, **kwargs)
    call(); call()
    b.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        res = call()
    b.barrier()
    dt = b.max_over_ranks((time.perf_counter() - t0) / reps)
    n_evals = n_points * max(n_shots, 1)
    return {"name": name, "protocol": proto, "points": n_points, "shots": n_shots, "evals_per_step": n_evals,
            "e2e": n_evals / dt, "e2e_ms": dt * 1e3, "ranks": b.world, "scaling": "strong",
            "api": "sharding.simulate_sharded (host columns -> H2D -> kernels -> NCCL all-gather of the metric columns -> D2H)",
            "mean_fidelity": float(np.nanmean(res.avg_fidelity)), "local_failure": res.local_failure is not None}


def run_ours(args, rank, world, local_rank):
    b = Bench(rank, world, local_rank)
    torch, dist, ev, dev = b.torch, b.dist, b.ev, b.dev
    ev.set_table(args.table)

    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro
    from noisyquantumsimulator_b200.rydberg_gates.simulation import METRIC_OUTPUTS

    # N_ROT distinct candidate batches, used in turn: the apparatus block of one batch is 84 MB (40 slots x 262,144
    # doubles), so consecutive timed iterations never find their inputs in the 126 MB L2 -- no flush kernel inside
    # the step (timing rule: "flush L2 or use inputs larger than L2").  The last batch is the one checked for parity.
    N_ROT = max(2, math.ceil(3 * 126e6 / (40 * 8 * B_PER_GPU)))
    cands = [make_candidates(B_PER_GPU, 1234 + rank + 1000 * j) for j in range(N_ROT)]
    cand = cands[0]
    mk, ak = dpro.split_kwargs(dict(MEDIUM, **cand))
    model = dpro.model_constants("jp", n_segments=N_SEG, **mk)
    d_apps = []
    for cj in cands:
        _, akj = dpro.split_kwargs(dict(MEDIUM, **cj))
        d_apps.append(torch.from_numpy(dpro.pack_apparatus("jp", B_PER_GPU, **akj)).to(dev))   # [40, B] apparatus block
    d_app = d_apps[0]
    d_out = torch.empty((B_PER_GPU, 16), dtype=torch.float64, device=dev)
    step_no = [0]
    gathered = torch.empty((world * B_PER_GPU,), dtype=torch.float64, device=dev) if world > 1 else None

    dfma_peak, dmma_peak = ev.fp64_peak()
    matvec_peak = ev.fp64_matvec_peak()
    # second reference point asked for by SURVEY 8d: cuBLAS DGEMM 8192^3 (context only; the denominator is the DFMA peak)
    dgemm_peak = None
    if rank == 0:
        try:
            a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
            c = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
            torch.matmul(a, c)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for _ in range(3):
                g0.record(); r_ = torch.matmul(a, c); g1.record(); torch.cuda.synchronize()
                best = min(best, g0.elapsed_time(g1))
            dgemm_peak = 2.0 * 8192.0 ** 3 / (best * 1e-3) * 1e-12
            del a, c, r_
        except Exception:
            dgemm_peak = None
    try:     # DRAM traffic of the dominant kernel from this round's `ncu --set full` capture of this command
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f)
        traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * B_PER_GPU / tr["candidates"]
        traffic_kernel, traffic_src = tr["kernel"], tr.get("source")
    except Exception:
        traffic, traffic_kernel, traffic_src = None, None, None

    def step():
        # inputs resident in HBM: prologue kernel -> planner -> evaluate kernels -> score all-gather
        ev.simulate_device(model, d_apps[step_no[0] % N_ROT], out=d_out)
        step_no[0] += 1
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_out[:, 0].contiguous())

    for _ in range(max(args.warmup, 3)):
        step()
    step_no[0] = 1        # the timed loop ends on batch 0 when steps is a multiple of N_ROT; otherwise re-run it below
    b.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ev.launch_count()
    ev.cluster_stats()
    ev.table_stats()
    kernel_ms, coef_ms = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.barrier()
    e0.record()
    for _ in range(args.steps):
        step()
        # per-launch device time of the dominant kernel (events recorded inside the C ABI on the
        # launching stream); reading it synchronises, which is inside the timed region on purpose
        kernel_ms.append(ev.last_kernel_ms())
        coef_ms.append(ev.last_coef_ms())
    e1.record()
    b.barrier()
    total_ms = b.max_over_ranks(e0.elapsed_time(e1))
    launches = ev.launch_count() - launches0
    cl_planned, cl_back, cl_records = ev.cluster_stats(with_records=True)
    tb_entries, tb_slots = ev.table_stats()
    flags, n_unsupported = ev.call_flags()
    value = world * B_PER_GPU * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API: NumPy candidate arrays in host memory -> simulate_CZ_gate_batch
    # (apparatus slots as columns -> rydcz_simulate_host_select: per chunk, pack the varying slots into pinned
    # staging, H2D, prologue + planner + evaluate kernels, gather of the requested columns, D2H; chunks pipelined on
    # eight streams) -> the metric columns the reference's optimiser reads (extract_metrics), in host memory
    # The step's inputs sit in page-locked host memory (contract: "host->device copy of that step's inputs from pinned
    # host memory"): the candidate arrays as the optimiser would hold them, row-major [B], [B, 4], [B, 5].
    cand_host = {k: ev.pinned_copy("bench_" + k, v) for k, v in cand.items()}

    def e2e_step():
        res = rg.simulate_CZ_gate_batch("jp", B=B_PER_GPU, device=local_rank, prologue="device", copy=False,
                                        outputs=METRIC_OUTPUTS, **MEDIUM, **cand_host)
        return res.avg_fidelity
    for _ in range(2):
        e2e_step()
    b.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f_e2e = e2e_step()
    b.barrier()
    e2e_value = world * B_PER_GPU * args.steps / b.max_over_ranks(time.perf_counter() - t0)
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # kernel-counted flops per launch: mean over the batches the timed loop cycled through (re-run here, not timed);
    # batch 0 last: it is the one the end-to-end call and the parity check use
    fl = []
    for j in list(range(1, N_ROT)) + [0]:
        ev.simulate_device(model, d_apps[j], out=d_out)
        fl.append(float(d_out[:, 15].sum().item()))
    flops_per_launch = float(np.mean(fl))
    out_np = d_out.cpu().numpy()
    k_ms = float(np.mean(kernel_ms))
    c_ms = float(np.mean(coef_ms))
    achieved = flops_per_launch / (k_ms * 1e-3) * 1e-12
    step_ms = total_ms / args.steps
    clustered = cl_planned > 0
    # bytes the evaluate kernel must read / write per launch: point row, segment rows, segment count, result row
    # (<= (9 + 4S) doubles in, 16 out); the coefficient records are the builder's own scratch, not
    # algorithmic traffic, and are reported separately
    alg_bytes = float(B_PER_GPU * (12 * 8 + N_SEG * 32 + 4 + 16 * 8))
    record_bytes = 2.0 * 32.0 * cl_records / max(args.steps, 1)
    n_var = sum(isinstance(c_, np.ndarray) for c_ in dpro.apparatus_columns("jp", B_PER_GPU, **ak))
    h2d = B_PER_GPU * n_var * 8                       # only the apparatus slots that vary over the batch cross PCIe
    d2h = B_PER_GPU * len(METRIC_OUTPUTS) * 8         # the requested result columns
    assert np.allclose(f_e2e, out_np[:, 0], rtol=0, atol=1e-12)     # both paths computed the same batch

    # ---- parity of the timed batch: randomly chosen candidates against the expm oracle (rank 0, not timed)
    parity = None
    if rank == 0 and not args.no_parity:
        n_chk = args.parity_samples
        idx = np.sort(np.random.default_rng(2025).choice(B_PER_GPU, n_chk, replace=False))
        pts, segs = host_plan(cand, idx)
        t0 = time.perf_counter()
        ref = np.array(_pool_map(_cpu_one_expm, [(pts[i], segs[i]) for i in range(n_chk)], os.cpu_count() or 1))
        got = out_np[idx][:, [0, 4, 6, 2]]
        err = np.abs(got - ref)
        parity = {"checked": int(n_chk), "oracle": "oracle/rydcz_oracle.py evaluate(mode='exact'): expm of the 81x81 Liouvillian per segment",
                  "worst_abs_dF_avg": float(err[:, 0].max()), "worst_abs_dF11": float(err[:, 1].max()),
                  "worst_abs_d_cz_phase_fidelity": float(err[:, 2].max()), "worst_abs_dF01": float(err[:, 3].max()),
                  "tolerance": 1e-9, "pass": bool(err.max() < 1e-9), "oracle_s": time.perf_counter() - t0}

    # ---- the other BASELINE configurations
    configs = []
    if not args.no_configs:
        if world == 1:
            rng = np.random.default_rng(1234)
            B1 = 65536
            p2 = 10 ** rng.uniform(np.log10(0.05), np.log10(2.0), B1)
            configs.append(measure_config(b, "config1_lp_65536_distinct_second_leg_power", "lp", B1,
                                          dict(MEDIUM, laser_2_power=p2), dfma_peak))
            configs.append(measure_config(b, "config3_sweep_4096pts_x64shots", "lp", 4096, dict(MEDIUM, **sweep_grid()),
                                          dfma_peak, n_shots=64, note="16x16x16 grid over second-leg power, Delta/Omega and pulse area; "
                                          "10 uK; Philox4x32-10 shots keyed by (seed, point, shot)"))
            for proto in ("lp", "jp", "sjp"):
                configs.append(measure_config(b, f"config4_population_8192_{proto}", proto, 8192,
                                              dict(MEDIUM, **population(proto, 8192)), dfma_peak))
            B5 = 131072
            p5 = 10 ** np.random.default_rng(1234).uniform(np.log10(0.05), np.log10(2.0), B5)
            for species in ("Rb87", "Cs133"):
                configs.append(measure_config(b, f"config5_lp_5level_{species}_{B5}", "lp", B5,
                                              dict(MEDIUM, species=species, laser_2_power=p5, hilbert_space_dim=5),
                                              dfma_peak, reps=3))
        else:
            for proto in ("lp", "jp", "sjp"):
                configs.append(measure_sharded(b, f"config4_population_8192_{proto}_sharded", proto, 8192,
                                               dict(MEDIUM, **population(proto, 8192))))
            configs.append(measure_sharded(b, "config3_sweep_4096pts_x64shots_sharded", "lp", 4096,
                                           dict(MEDIUM, **sweep_grid()), n_shots=64))

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_sample = max(2 * cores, 8)
            pts, segs = host_plan(cand, np.arange(n_sample))
            rate, dt = cpu_path_rate(pts, segs, n_sample, cores)
            t0 = time.perf_counter()
            for i in range(4):                      # the exact-expm oracle, one process
                _cpu_one_expm((pts[i], segs[i]))
            expm_rate = 4 / (time.perf_counter() - t0)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {n_sample} of the {B_PER_GPU} candidates ({dt:.1f} s), oracle zvode/Adams "
                             f"restatement of the reference's mesolve path, RHS = {_RHS} matvec, {cores} processes",
                   "evals_per_s_per_core": rate / cores, "expm_oracle_evals_per_s_1core": expm_rate,
                   "reference_notebook_evals_per_s": NOTEBOOK_EVALS_PER_S}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "jp_bangbang_5seg_lindblad_3level", "candidates_per_gpu": B_PER_GPU,
                       "propagation": ("rho_11: clustered-spectrum Chebyshev (csrc/rydcz_cluster.cuh); single-atom rho "
                                       "and the |ab><00| coherences: closed forms (csrc/rydcz_closed.cuh)"
                                       if clustered else "rho_11: single-interval Chebyshev; single-atom rho and "
                                                         "coherences in closed form"),
                       "coefficient_table": {0: "off", 1: "entries generated again in every step",
                                             2: "entries kept across steps (library default): generated during warm-up, "
                                                "as an optimiser's generations find them"}[args.table],
                       "segments": N_SEG, "apparatus": "medium (RG/optimize_cz_gate.py:196-213), full noise",
                       "l2": f"{N_ROT} distinct input batches of {40 * 8 * B_PER_GPU / 1e6:.0f} MB used in turn "
                             "(inputs larger than the 126 MB L2; no flush kernel inside the step)",
                       "parallelism": f"shard{world}",
                       "host_cpus_bound_to_gpu": b.bound_cpus,
                       "mean_fidelity": float(out_np[:, 0].mean())},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "rydberg_gates.simulate_CZ_gate_batch(outputs=METRIC_OUTPUTS) -> rydcz_simulate_host_select "
                           "(C ABI): the candidates' row-major arrays in page-locked host memory -> H2D per 32768-point chunk "
                           "(transposed into apparatus slots on the device), the optimiser's 11 metric columns out into "
                           "page-locked memory, chunks pipelined on 8 streams"},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": dfma_peak, "unit": "TFLOP/s",
                         "frac": achieved / dfma_peak,
                         # the whole step against the same peak: every kernel of the step (prologue, planner,
                         # coefficient kernel, evaluate kernels, L2 flush) in the denominator's time
                         "step_achieved": flops_per_launch / (step_ms * 1e-3) * 1e-12,
                         "step_frac": flops_per_launch / (step_ms * 1e-3) * 1e-12 / dfma_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one `ncu --set full`
                         # capture of this command (profiles/ncu_traffic.json), scaled to this batch size
                         "traffic": traffic, "traffic_kernel": traffic_kernel, "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes, "coefficient_record_bytes": record_bytes,
                         # shared coefficient table (rydcz_set_table; default 2 = entries kept across calls, as an
                         # optimiser's generations find them; --table 1 regenerates them in EVERY step): slots of the
                         # step's candidates that read the table / entries generated for them inside the timed steps
                         "table_slots_per_step": tb_slots / max(args.steps, 1),
                         "table_entries_generated_per_step": tb_entries / max(args.steps, 1),
                         "kernel": ("rydcz::evaluate_cluster_kernel<3,true,true>" if clustered
                                    else "rydcz::evaluate_kernel<3,true,false,true>"),
                         # device span from the first to the last evaluate kernel of the step (several launches:
                         # one per kind of point, most of them finding no point)
                         "kernel_span_ms": k_ms, "kernel_ms": k_ms, "flops_per_launch": flops_per_launch,
                         # the clustered path's second kernel (coefficient records): its flops are not in `achieved`
                         "coef_kernel_ms": c_ms,
                         "clustered_points_per_step": cl_planned / max(args.steps, 1), "handed_back": cl_back,
                         "call_flags": flags, "unsupported_points": n_unsupported,
                         "peak_source": "DFMA microbenchmark (rydcz_fp64_peak) measured in this run; "
                                        "MEASURED_PEAKS.json has no FP64 entry", "dmma_peak": dmma_peak,
                         "cublas_dgemm_8192_tflops": dgemm_peak,
                         # the hot loop's own 78-nonzero matvec run bare at the kernel's occupancy (three
                         # distinct register operands per DFMA): the ceiling for a loop of this shape
                         "matvec_mix_peak": matvec_peak, "frac_of_matvec_mix_peak": achieved / matvec_peak},
            "parity": parity,
            "configs": configs,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global B_PER_GPU
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed batch")
    ap.add_argument("--parity-samples", type=int, default=64)
    ap.add_argument("--candidates", type=int, default=B_PER_GPU, help="bang-bang candidates per GPU per step")
    ap.add_argument("--table", type=int, default=2, choices=[0, 1, 2],
                    help="shared coefficient table (rydcz_set_table): 2 = entries kept across steps (library default), "
                         "1 = generated again in every step, 0 = off")
    args = ap.parse_args()
    B_PER_GPU = args.candidates
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
