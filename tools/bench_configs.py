#!/usr/bin/env python3
"""Throughput of the SURVEY.md section 8d configurations through the public API (one GPU).
Prints one JSON line per configuration; copy the output into profiles/."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from noisyquantumsimulator_b200 import rydberg_gates as rg  # noqa: E402
from noisyquantumsimulator_b200.engine import get_evaluator  # noqa: E402
from noisyquantumsimulator_b200.rydberg_gates import batched_objective as bo  # noqa: E402

ev = get_evaluator(0)
M = dict(bench.MEDIUM)


def timed(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


def line(name, n_evals, dt, res, extra=None):
    raw = res.raw
    flops = float(raw[:, 15].sum())
    e_ms, c_ms = ev.last_kernel_ms(), ev.last_coef_ms()
    planned, back = ev.cluster_stats()
    k_ms = e_ms + (c_ms if planned else 0.0)      # evaluate kernels + (clustered path) coefficient kernel
    d = dict(config=name, evals_per_call=n_evals, api_ms=dt * 1e3, evals_per_s_api=n_evals / dt, kernel_ms=k_ms,
             evaluate_ms=e_ms, coef_ms=c_ms if planned else 0.0, clustered=bool(planned), handed_back=back,
             evals_per_s_kernel=n_evals / (k_ms * 1e-3), flops=flops, tflops_kernel=flops / (e_ms * 1e-3) * 1e-12,
             mean_fidelity=float(res.avg_fidelity.mean()))
    d.update(extra or {})
    print(json.dumps(d), flush=True)


peak, _ = ev.fp64_peak()
print(json.dumps(dict(dfma_peak_tflops=peak)))
rng = np.random.default_rng(1234)

# config 1: single LP gate -- 65,536 distinct points, second-leg power log-uniform in [0.05, 2] W
B = 65536
p2 = 10 ** rng.uniform(np.log10(0.05), np.log10(2.0), B)
kw = dict(M, laser_2_power=p2)
dt, res = timed(lambda: rg.simulate_CZ_gate_batch("lp", B=B, warn=False, copy=False, **kw))
line("config1_lp_65536_distinct_power", B, dt, res)
kw = dict(M)
dt, res = timed(lambda: rg.simulate_CZ_gate_batch("lp", B=B, warn=False, copy=False, **kw))
line("config1_lp_65536_replicated_point", B, dt, res)

# config 2: bang-bang, 5 and 7 segments, noisy and noiseless
for nseg, ot_b in ((5, (5.0, 40.0)), (7, (3.0, 30.0))):
    ot = rng.uniform(*ot_b, B)
    fr = np.sort(rng.uniform(0.01, 0.99, (B, nseg - 1)), 1)
    ph = rng.uniform(-np.pi, np.pi, (B, nseg))
    for noise in (True, False):
        kw = dict(M, include_noise=noise, omega_tau=ot, switching_times=fr * ot[:, None], phases=ph)
        dt, res = timed(lambda: rg.simulate_CZ_gate_batch("jp", B=B, warn=False, copy=False, **kw))
        line(f"config2_bb{nseg}_65536_{'noisy' if noise else 'noiseless'}", B, dt, res)

# config 3: 16x16x16 sweep, 64 shots/point.  BASELINE.json's "Omega x Delta x T" does not say whether T is the pulse
# area or the atom temperature: both readings are run.
p2g, dg, og = np.meshgrid(np.geomspace(0.05, 2.0, 16), np.linspace(0.20, 0.50, 16), np.linspace(3.5, 5.5, 16),
                          indexing="ij")
kw = dict(M, laser_2_power=p2g.ravel(), delta_over_omega=dg.ravel(), omega_tau=og.ravel(), temperature=10e-6)
dt, res = timed(lambda: rg.simulate_CZ_gate_batch("lp", B=4096, n_shots=64, seed=1234, warn=False, copy=False, **kw), 3)
line("config3_sweep_4096pts_x64shots_T_is_pulse_area", 4096 * 64, dt, res, dict(shots=64, points=4096))
temps = np.array([2, 5, 10, 20, 50, 100]) * 1e-6
p2t, dgt, tt = np.meshgrid(np.geomspace(0.05, 2.0, 26), np.linspace(0.20, 0.50, 26), temps, indexing="ij")
kw = dict(M, laser_2_power=p2t.ravel(), delta_over_omega=dgt.ravel(), temperature=tt.ravel())
dt, res = timed(lambda: rg.simulate_CZ_gate_batch("lp", B=p2t.size, n_shots=64, seed=1234, warn=False, copy=False, **kw), 3)
line("config3_sweep_4056pts_x64shots_T_is_temperature", p2t.size * 64, dt, res, dict(shots=64, points=int(p2t.size)))

# config 4: optimiser population of 8192 candidates per iteration (one GPU evaluates all of them here;
# bench.py --gpus N shards populations across GPUs)
for proto, nseg in (("lp", 5), ("jp", 5), ("sjp", 5)):
    bounds, x0, _ = bo.bounds_and_x0(proto, nseg)
    lo, hi = np.array(bounds).T
    X = lo[:, None] + (hi - lo)[:, None] * np.random.default_rng(42).uniform(size=(len(bounds), 8192))
    holder = {}
    f = bo.make_vectorized_objective(proto, include_noise=True, n_segments=nseg,
                                     on_batch=lambda X_, r: holder.__setitem__("r", r))
    dt, cost = timed(lambda: f(X), 3)
    line(f"config4_population_8192_{proto}", 8192, dt, holder["r"], dict(best_cost=float(np.min(cost))))

# smooth JP at batch size: 65,536 candidates inside the optimiser's bounds (300 steps each)
bounds, x0, _ = bo.bounds_and_x0("sjp", 5)
lo, hi = np.array(bounds).T
X = lo[:, None] + (hi - lo)[:, None] * np.random.default_rng(43).uniform(size=(len(bounds), 65536))
holder = {}
f = bo.make_vectorized_objective("sjp", include_noise=True, n_segments=5,
                                 on_batch=lambda X_, r: holder.__setitem__("r", r))
dt, cost = timed(lambda: f(X), 3)
line("smooth_jp_65536_noisy", 65536, dt, holder["r"], dict(repeated_segment_points=ev.repeat_stats()))

# config 5: Rb vs Cs with the intermediate state kept (5 levels per atom, 25x25 rho), LP protocol, distinct
# second-leg powers; host prologue (NumPy) + rydcz_evaluate_host
B5 = 131072
p2 = 10 ** np.random.default_rng(1234).uniform(np.log10(0.05), np.log10(2.0), B5)
for species in ("Rb87", "Cs133"):
    kw = dict(M, species=species, laser_2_power=p2, hilbert_space_dim=5)
    dt, res = timed(lambda: rg.simulate_CZ_gate_batch("lp", B=B5, warn=False, copy=False, **kw), 3)
    line(f"config5_lp_5level_{species}_{B5}", B5, dt, res)
