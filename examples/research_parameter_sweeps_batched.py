#!/usr/bin/env python3
"""The reference's research parameter sweeps (examples/research_parameter_sweeps.py:644-760) as batches.

The reference walks every sweep point in a Python loop, one ``simulate_CZ_gate`` (LP) and one (JP) per
value; here every sweep is ONE ``simulate_CZ_gate_batch`` call per protocol, optionally at a much
finer resolution (``--points``) and with Monte-Carlo shots (``--shots``, BASELINE.json configs[2]).
The fixed operating point is the reference script's (``run_single_simulation`` defaults,
:81-135): lasers 1 mW / 1 um and 1 W / 10 um, 5 uK, spacing 3.0, tweezer 30 mW / 1 um, n = 70.
No plotting: prints one table per sweep and writes ``--out`` (JSON).

    python examples/research_parameter_sweeps_batched.py --points 64 --out sweeps.json
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noisyquantumsimulator_b200 import rydberg_gates as rg  # noqa: E402

FIXED = dict(laser_1_power=1e-3, laser_1_waist=1.0e-6, laser_2_power=1.0, laser_2_waist=10e-6,
             laser_1_linewidth_hz=100.0, laser_2_linewidth_hz=100.0, Delta_e=2 * np.pi * 1e9, species="Rb87",
             n_rydberg=70, temperature=5e-6, spacing_factor=3.0, tweezer_power=30e-3, tweezer_waist=1e-6,
             B_field=1e-4, NA=0.5, include_noise=True, include_motional_dephasing=True)

# name -> (keyword(s) of simulate_CZ_gate_batch, the reference's grid, transform to SI)
SWEEPS = {
    "temperature_uK": (("temperature",), [5, 10, 20, 40, 80, 150], 1e-6),
    "laser_linewidth_hz": (("laser_1_linewidth_hz", "laser_2_linewidth_hz"), [10, 50, 100, 500, 1000, 5000], 1.0),
    "Delta_e_GHz": (("Delta_e",), [0.5, 1, 2, 5, 10], 2 * np.pi * 1e9),
    "spacing_factor": (("spacing_factor",), [1.2, 1.5, 2.0, 2.5, 3.0], 1.0),
    "rydberg_power_2_W": (("laser_2_power",), [0.1, 0.3, 0.5, 1.0, 2.0, 5.0], 1.0),
    "tweezer_power_mW": (("tweezer_power",), [1, 3, 5, 10, 20, 50], 1e-3),
}
N_VALUES = [50, 60, 70, 80, 90]          # n_rydberg changes per-call model constants: one batch per value


def grid(values, points):
    if points is None or points <= len(values):
        return np.asarray(values, dtype=float)
    return np.geomspace(values[0], values[-1], points)


CAL_DIR = None


def run(proto, B, n_shots, sweep=None, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = rg.simulate_CZ_gate_batch(proto, B=B, n_shots=n_shots, seed=1234, warn=False, **kw)
    if CAL_DIR and sweep:
        from noisyquantumsimulator_b200.rydberg_gates.calibration import write_cz_calibration
        write_cz_calibration(CAL_DIR, r, {sweep[0]: sweep[1]}, protocol=proto, n_rydberg=kw["n_rydberg"],
                             species=kw["species"], tag=sweep[0], n_shots=n_shots)
    return dict(fidelity=r.avg_fidelity.tolist(), gate_time_us=r.gate_time_us.tolist(),
                V_over_Omega=r.V_over_Omega.tolist(), loss=r.loss_probability.tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=None, help="points per sweep (default: the reference's grids)")
    ap.add_argument("--shots", type=int, default=0, help="Monte-Carlo shots per point (0 = analytic dephasing rates)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--calibration-dir", default=None,
                    help="also write calibration_data-style JSON files (reference: calibration_data/README.md)")
    a = ap.parse_args()
    global CAL_DIR
    CAL_DIR = a.calibration_dir
    results, t0, n_eval = {}, time.perf_counter(), 0
    for name, (keys, values, scale) in SWEEPS.items():
        g = grid(values, a.points)
        kw = dict(FIXED, **{k: g * scale for k in keys})
        results[name] = dict(values=g.tolist(), lp=run("lp", len(g), a.shots, (name, g), **kw),
                             jp=run("jp", len(g), a.shots, (name, g), **kw))
        n_eval += 2 * len(g)
    by_n = dict(values=N_VALUES, lp=[], jp=[])
    for n in N_VALUES:
        for proto in ("lp", "jp"):
            by_n[proto].append(run(proto, 1, a.shots, **dict(FIXED, n_rydberg=n)))
        n_eval += 2
    results["n_rydberg"] = by_n
    species = {s: {p: run(p, 1, a.shots, **dict(FIXED, species=s)) for p in ("lp", "jp")} for s in ("Rb87", "Cs133")}
    results["species"] = species
    dt = time.perf_counter() - t0
    for name in SWEEPS:
        r = results[name]
        print(f"\n{name}")
        print(f"{'value':>12s} {'F_LP':>10s} {'F_JP':>10s} {'t_LP us':>9s} {'t_JP us':>9s} {'V/Omega':>9s}")
        for i, v in enumerate(r["values"]):
            print(f"{v:12.4g} {r['lp']['fidelity'][i]:10.6f} {r['jp']['fidelity'][i]:10.6f} "
                  f"{r['lp']['gate_time_us'][i]:9.4f} {r['jp']['gate_time_us'][i]:9.4f} {r['lp']['V_over_Omega'][i]:9.1f}")
    print(f"\n{n_eval} gate evaluations" + (f" x {a.shots} shots" if a.shots else "") + f" in {dt:.2f} s")
    if a.out:
        with open(a.out, "w") as f:
            json.dump(results, f)


if __name__ == "__main__":
    main()
