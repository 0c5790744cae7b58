#!/usr/bin/env python3
"""bench.py -- CZ-fidelity Lindblad gate evaluations per second on N B200s.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle port of
                                                           the QuTiP mesolve path) on the host cores

One "step" = one pass of the hot path over one batch of synthetic pulse candidates.  Workload
(BASELINE.json configs[1], SURVEY.md section 8d "config 2"): Jandura-Pupillo bang-bang CZ, 5 piecewise-constant
segments, two-atom 3-level Lindblad model with the medium apparatus's noise rates; per GPU
262,144 candidates (--candidates) drawn inside the reference optimiser's bounds (RG/optimize_cz_gate.py:576-587),
seed 1234 + rank.  Every rank evaluates its own shard (no data-path collective; the only
exchange is the all-gather of per-candidate scores, timed inside the step) -> "scaling": "weak".

Prints ONE JSON line (rank 0).
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
    """Evaluator inputs of a few candidates via the NumPy prologue (for the CPU path)."""
    from noisyquantumsimulator_b200.rydberg_gates import prologue
    sub = {k: v[idx] for k, v in cand.items()}
    plan = prologue.build_plan("jp", B=len(idx), warn=False, **MEDIUM, **sub)
    return plan.points, plan.segs


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
    """The same evaluation with the oracle's exact matrix exponentials (what the kernel is checked against)."""
    from oracle import rydcz_oracle as orc
    pt, seg = args
    rates = orc.NoiseRates(gamma_r=2 * pt[4], gamma_bbr=pt[5] - pt[4], gamma_phi_laser=pt[6],
                           gamma_scatter_intermediate=pt[7])
    segs = [orc.Segment(s[0], complex(s[1], s[2]), s[3], 51) for s in seg if s[0] >= 1e-18]
    return orc.evaluate(3, segs, pt[0], pt[1], rates, "exact", "coherence")[1]


def cpu_path_rate(pts, segs, n_sample, procs):
    """Evaluate the first n_sample candidates with the oracle's zvode restatement on `procs`
    host processes; returns evals/s."""
    import multiprocessing as mp
    work = [(pts[i], segs[i]) for i in range(n_sample)]
    t0 = time.perf_counter()
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_cpu_one, work, chunksize=1)
    else:
        for w in work:
            _cpu_one(w)
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
              f"reference's mesolve path (QuTiP itself is not installable), {cores} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "jp_bangbang_5seg_lindblad_3level", "candidates_per_step": n_sample,
                   "host_cores": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- CUDA path
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from noisyquantumsimulator_b200.engine import Evaluator

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ev = Evaluator(local_rank)

    from noisyquantumsimulator_b200 import rydberg_gates as rg
    from noisyquantumsimulator_b200.rydberg_gates import device_prologue as dpro

    cand = make_candidates(B_PER_GPU, 1234 + rank)
    mk, ak = dpro.split_kwargs(dict(MEDIUM, **cand))
    model = dpro.model_constants("jp", n_segments=N_SEG, **mk)
    app = dpro.pack_apparatus("jp", B_PER_GPU, **ak)                 # [40, B] apparatus block
    d_app = torch.from_numpy(app).to(dev)
    d_out = torch.empty((B_PER_GPU, 16), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * B_PER_GPU,), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    dfma_peak, dmma_peak = ev.fp64_peak()
    matvec_peak = ev.fp64_matvec_peak()
    # second reference point asked for by SURVEY 8d: cuBLAS DGEMM 8192^3 (context only; the denominator is the DFMA peak)
    dgemm_peak = None
    if rank == 0:
        try:
            a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
            b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
            torch.matmul(a, b)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for _ in range(3):
                g0.record(); c = torch.matmul(a, b); g1.record(); torch.cuda.synchronize()
                best = min(best, g0.elapsed_time(g1))
            dgemm_peak = 2.0 * 8192.0 ** 3 / (best * 1e-3) * 1e-12
            del a, b, c
        except Exception:
            dgemm_peak = None
    try:     # DRAM traffic of the dominant kernel from the committed ncu capture (scales with the batch)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f)
        traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * B_PER_GPU / tr["candidates"]
        traffic_kernel = tr["kernel"]
    except Exception:
        traffic, traffic_kernel = None, None

    def step():
        # inputs resident in HBM: prologue kernel -> work sort -> evaluate kernels -> score all-gather
        flush.zero_()
        ev.simulate_device(model, d_app, out=d_out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_out[:, 0].contiguous())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ev.launch_count()
    ev.cluster_stats()
    kernel_ms, coef_ms = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
        # per-launch device time of the dominant kernel (events recorded inside the C ABI on the
        # launching stream); reading it synchronises, which is inside the timed region on purpose
        kernel_ms.append(ev.last_kernel_ms())
        coef_ms.append(ev.last_coef_ms())
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = ev.launch_count() - launches0
    cl_planned, cl_back, cl_records = ev.cluster_stats(with_records=True)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B_PER_GPU * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API: NumPy candidate arrays in host memory ->
    # simulate_CZ_gate_batch (apparatus slots as columns -> rydcz_simulate_host_cols: per chunk, pack the
    # varying slots into pinned staging, H2D, prologue + sort + evaluate kernels, D2H; chunks pipelined
    # on three streams) -> BatchResult arrays in host memory
    def e2e_step():
        res = rg.simulate_CZ_gate_batch("jp", B=B_PER_GPU, device=local_rank, prologue="device", copy=False,
                                        **MEDIUM, **cand)
        return res.avg_fidelity
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f_e2e = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B_PER_GPU * args.steps / float(t.item())
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    out_np = d_out.cpu().numpy()
    flops_per_launch = float(out_np[:, 15].sum())
    k_ms = float(np.mean(kernel_ms))
    c_ms = float(np.mean(coef_ms))
    achieved = flops_per_launch / (k_ms * 1e-3) * 1e-12
    clustered = cl_planned > 0
    # bytes the evaluate kernel must read / write per launch: point row, segment rows, segment count, result
    # row, and on the clustered path the 32-byte coefficient records, once for rho_11 and once for the coherence
    alg_bytes = float(B_PER_GPU * (12 * 8 + N_SEG * 32 + 4 + 16 * 8)) + 2.0 * 32.0 * cl_records / max(args.steps, 1)
    n_var = sum(isinstance(c, np.ndarray) for c in dpro.apparatus_columns("jp", B_PER_GPU, **ak))
    h2d = B_PER_GPU * n_var * 8          # only the apparatus slots that vary over the batch cross PCIe
    d2h = B_PER_GPU * (16 + 8) * 8       # result rows + reported scalars
    assert np.allclose(f_e2e, out_np[:, 0], rtol=0, atol=1e-12)     # both paths computed the same batch

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
                             f"restatement of the reference's mesolve path, {cores} processes",
                   "expm_oracle_evals_per_s_1core": expm_rate,
                   # Runtime lines saved in the reference's notebooks (author's Mac, one process)
                   "reference_notebook_evals_per_s": {"lp_noiseless": 123, "smooth_jp_noiseless": 5.4,
                                                      "lp_noisy_older_notebook": 59, "jp_noisy_older_notebook": 25}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "jp_bangbang_5seg_lindblad_3level", "candidates_per_gpu": B_PER_GPU,
                       "propagation": ("clustered-spectrum Chebyshev (csrc/rydcz_cluster.cuh)" if clustered
                                       else "single-interval Chebyshev"),
                       "segments": N_SEG, "apparatus": "medium (RG/optimize_cz_gate.py:196-213), full noise",
                       "l2": "256 MiB flush write between timed iterations", "parallelism": f"shard{world}",
                       "mean_fidelity": float(out_np[:, 0].mean())},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "rydberg_gates.simulate_CZ_gate_batch -> rydcz_simulate_host_cols (C ABI): varying apparatus "
                           "columns only, 32768-point chunks pipelined on 3 streams, pinned staging"},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": dfma_peak, "unit": "TFLOP/s",
                         "frac": achieved / dfma_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full`
                         # capture (profiles/ncu_traffic.json), scaled to this batch size
                         "traffic": traffic, "traffic_kernel": traffic_kernel,
                         "algorithmic_bytes": alg_bytes,
                         "kernel": ("rydcz::evaluate_cluster_kernel<3,true,true>" if clustered
                                    else "rydcz::evaluate_kernel<3,true,false,true>"),
                         "kernel_ms": k_ms, "flops_per_launch": flops_per_launch,
                         # the clustered path's second kernel (coefficient records, one thread per pulse segment):
                         # sincos / division heavy, its flops are not counted in `achieved`
                         "coef_kernel_ms": c_ms,
                         "clustered_points_per_step": cl_planned / max(args.steps, 1), "handed_back": cl_back,
                         "peak_source": "DFMA microbenchmark (rydcz_fp64_peak) measured in this run; "
                                        "MEASURED_PEAKS.json has no FP64 entry", "dmma_peak": dmma_peak, "cublas_dgemm_8192_tflops": dgemm_peak,
                         # the hot loop's own 78-nonzero matvec run bare at the kernel's occupancy (three
                         # distinct register operands per DFMA): the ceiling for a loop of this shape
                         "matvec_mix_peak": matvec_peak, "frac_of_matvec_mix_peak": achieved / matvec_peak},
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
    ap.add_argument("--candidates", type=int, default=B_PER_GPU, help="bang-bang candidates per GPU per step")
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
