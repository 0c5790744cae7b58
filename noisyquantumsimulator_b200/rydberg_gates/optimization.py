"""Hardware-parameter optimisation and Pareto exploration, population-at-once
(reference: RG/optimization.py).

``explore_parameter_space`` and ``optimize_CZ_parameters`` keep the reference's arguments, bounds,
objective formulas and result records.  Both drive ``scipy.optimize.differential_evolution`` over
8-11 hardware knobs; the reference evaluates one candidate per objective call (one
``simulate_CZ_gate`` = 8-1200 QuTiP ``mesolve`` calls, RG/optimization.py:467-572, 856-943).  Here
the objective receives the whole population (``vectorized=True, updating='deferred'``) and
``evaluate_hardware_population`` turns it into a handful of ``simulate_CZ_gate_batch`` launches:
candidates are grouped by their (integer) Rydberg level, the only knob that changes per-call
model constants; every other knob is a per-point apparatus column.
"""
from __future__ import annotations

import pickle
import time
import warnings
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from .simulation import simulate_CZ_gate_batch

LP_NAMES = ("levine_pichler", "lp", "two_pulse")
JP_NAMES = ("jandura_pupillo", "jp", "smooth_jp", "single_pulse", "time_optimal")
NOISE_KEYS = ("gamma_r", "gamma_phi_laser", "gamma_phi_thermal", "gamma_phi_zeeman", "gamma_loss_antitrap",
              "gamma_loss_background", "gamma_scatter_intermediate", "gamma_leakage", "total_decay_rate",
              "total_dephasing_rate", "total_loss_rate")


# ---------------------------------------------------------------------------- records
@dataclass
class HardwareOptimizationResult:
    """Inverse problem: hardware that reaches a target fidelity and gate time (RG/optimization.py:91-127)."""

    success: bool
    target_fidelity: float
    target_gate_time_ns: float
    achieved_fidelity: float
    achieved_gate_time_ns: float
    fidelity_error_pct: float
    gate_time_error_pct: float
    optimal_parameters: Dict[str, float]
    V_over_Omega: float
    noise_breakdown: Dict[str, float]
    n_evaluations: int
    final_cost: float
    message: str

    def __repr__(self):
        return (f"HardwareOptimizationResult(\n  Target:   F={self.target_fidelity:.4f}, t={self.target_gate_time_ns:.1f} ns\n"
                f"  Achieved: F={self.achieved_fidelity:.4f}, t={self.achieved_gate_time_ns:.1f} ns\n"
                f"  Errors:   ΔF={self.fidelity_error_pct:+.2f}%, Δt={self.gate_time_error_pct:+.2f}%\n"
                f"  V/Ω={self.V_over_Omega:.1f}, Evals={self.n_evaluations}, Cost={self.final_cost:.2e}\n"
                f"  Success: {self.success}\n)")


@dataclass
class EvaluatedPoint:
    """RG/optimization.py:130-149."""

    Omega_MHz: float
    laser_linewidth_kHz: float
    V_over_Omega: float
    fidelity: float
    gate_time_ns: float
    infidelity: float
    noise_breakdown: Dict[str, float] = field(default_factory=dict)
    protocol: str = ""
    species: str = "Rb87"


@dataclass
class ExplorationResult:
    """Every evaluated point of an exploration plus its fidelity/speed Pareto front
    (RG/optimization.py:151-277)."""

    protocol: str
    species: str
    points: List[EvaluatedPoint] = field(default_factory=list)
    pareto_front: List[EvaluatedPoint] = field(default_factory=list)
    n_evaluations: int = 0
    runtime_seconds: float = 0.0
    optimizer_settings: Dict[str, Any] = field(default_factory=dict)

    def add_point(self, point: EvaluatedPoint):
        self.points.append(point)
        self.n_evaluations += 1

    def compute_pareto_front(self):
        """Sweep the points by increasing gate time; keep those that beat every faster point's fidelity."""
        front, best = [], -1.0
        for p in sorted(self.points, key=lambda q: q.gate_time_ns):
            if p.fidelity > best:
                front.append(p)
                best = p.fidelity
        self.pareto_front = front

    def get_points_above_fidelity(self, min_fidelity: float) -> List[EvaluatedPoint]:
        return [p for p in self.points if p.fidelity >= min_fidelity]

    def get_points_below_time(self, max_time_ns: float) -> List[EvaluatedPoint]:
        return [p for p in self.points if p.gate_time_ns <= max_time_ns]

    def get_best_for_target(self, target_fidelity: float = None, target_time_ns: float = None):
        if target_fidelity is not None:
            c = self.get_points_above_fidelity(target_fidelity)
            return min(c, key=lambda p: p.gate_time_ns) if c else None
        if target_time_ns is not None:
            c = self.get_points_below_time(target_time_ns)
            return max(c, key=lambda p: p.fidelity) if c else None
        return None

    def summary(self) -> str:
        if not self.points:
            return "No points evaluated."
        F = [p.fidelity for p in self.points]
        T = [p.gate_time_ns for p in self.points]
        rows = ["=" * 60, f"EXPLORATION RESULTS: {self.protocol.upper()}", "=" * 60,
                f"Total evaluations: {self.n_evaluations}",
                f"Runtime: {self.runtime_seconds:.1f}s ({self.runtime_seconds / 60:.1f} min)", "",
                f"Fidelity range: {min(F) * 100:.2f}% - {max(F) * 100:.2f}%",
                f"Gate time range: {min(T):.1f} - {max(T):.1f} ns", "", f"Pareto front: {len(self.pareto_front)} points"]
        if self.pareto_front:
            rows += ["", "Key Pareto points:"]
            rows += [f"  F={p.fidelity * 100:.2f}% @ {p.gate_time_ns:.1f}ns (V/Ω={p.V_over_Omega:.1f})"
                     for p in self.pareto_front[:5]]
            if len(self.pareto_front) > 5:
                rows.append(f"  ... and {len(self.pareto_front) - 5} more")
        return "\n".join(rows)

    def save(self, filepath: str):
        with open(filepath, "wb") as f:
            pickle.dump(self, f)
        print(f"✓ Saved {self.n_evaluations} points to {filepath}")

    @staticmethod
    def load(filepath: str) -> "ExplorationResult":
        with open(filepath, "rb") as f:
            return pickle.load(f)


# ---------------------------------------------------------------------------- population evaluator
def evaluate_hardware_population(protocol: str, params: Dict[str, np.ndarray], *, species: str = "Rb87",
                                 laser_1_waist: float, laser_2_waist: float, include_noise: bool = True,
                                 include_motional_dephasing: bool = True, include_doppler_dephasing: bool = True,
                                 include_intensity_noise: bool = True, intensity_noise_frac: float = 0.01,
                                 background_loss_rate_hz: float = None, defaults: Dict[str, float] = None,
                                 device: int = 0) -> Dict[str, np.ndarray]:
    """Evaluate P hardware configurations.  ``params`` maps the reference's knob names
    (rydberg_power_1/2, temperature, spacing_factor, n_rydberg, tweezer_power, tweezer_waist,
    laser_linewidth [Hz], Delta_e [Hz], delta_over_omega, omega_tau) to arrays of length P; missing
    knobs take ``defaults``.  Returns arrays: fidelity, gate_time_ns, V_over_Omega, Omega_MHz, ok and
    the per-point noise rates.  One ``simulate_CZ_gate_batch`` launch per distinct n_rydberg."""
    defaults = dict(defaults or {})
    P = len(next(iter(params.values())))
    is_lp = protocol.lower() in LP_NAMES
    proto = "lp" if is_lp else "jp"

    def col(name, fallback=None):
        if name in params:
            return np.asarray(params[name], dtype=float)
        v = defaults.get(name, fallback)
        return None if v is None else np.full(P, float(v))

    n_ryd = np.rint(col("n_rydberg", 70)).astype(int)
    out = {k: np.full(P, np.nan) for k in ("fidelity", "gate_time_ns", "V_over_Omega", "Omega_MHz") + NOISE_KEYS}
    out["ok"] = np.zeros(P, dtype=bool)
    lw = col("laser_linewidth", 1000.0)
    De = col("Delta_e")
    for n in np.unique(n_ryd):
        sel = np.nonzero(n_ryd == n)[0]
        pick = lambda a: None if a is None else a[sel]
        kw = dict(laser_1_power=pick(col("rydberg_power_1", 50e-6)), laser_1_waist=laser_1_waist,
                  laser_2_power=pick(col("rydberg_power_2", 0.5)), laser_2_waist=laser_2_waist,
                  Delta_e=2 * np.pi * 1e9 if De is None else 2 * np.pi * De[sel],   # TwoPhotonExcitationConfig default
                  laser_1_linewidth_hz=lw[sel], laser_2_linewidth_hz=lw[sel],
                  temperature=pick(col("temperature", 5e-6)), spacing_factor=pick(col("spacing_factor", 3.0)),
                  tweezer_power=pick(col("tweezer_power", 50e-3)), tweezer_waist=pick(col("tweezer_waist", 1e-6)),
                  species=species, n_rydberg=int(n), include_noise=include_noise,
                  include_motional_dephasing=include_motional_dephasing,
                  include_doppler_dephasing=include_doppler_dephasing,
                  include_intensity_noise=include_intensity_noise, intensity_noise_frac=intensity_noise_frac,
                  background_loss_rate_hz=background_loss_rate_hz)
        if is_lp:
            kw.update(delta_over_omega=pick(col("delta_over_omega")), omega_tau=pick(col("omega_tau")))
        else:
            kw.update(omega_tau=pick(col("omega_tau")))
        try:
            # host prologue: these explorations read the per-point noise breakdown
            res = simulate_CZ_gate_batch(proto, B=len(sel), device=device, warn=False, prologue="host", **kw)
        except Exception:
            continue
        out["fidelity"][sel] = res.avg_fidelity
        out["gate_time_ns"][sel] = res.tau_total * 1e9
        out["V_over_Omega"][sel] = res.V_over_Omega
        out["Omega_MHz"][sel] = res.Omega / (2 * np.pi * 1e6)
        nb = res.plan.noise_breakdown
        for k in NOISE_KEYS:
            if k in nb:
                out[k][sel] = np.broadcast_to(np.asarray(nb[k], dtype=float), (len(sel),))
        out["ok"][sel] = np.isfinite(res.avg_fidelity)
    return out


def _columns(names, X):
    X = np.asarray(X, dtype=float)
    single = X.ndim == 1
    if single:
        X = X[:, None]
    return {n: X[i] for i, n in enumerate(names)}, single, X.shape[1]


# ---------------------------------------------------------------------------- exploration
def explore_parameter_space(protocol: str = "levine_pichler", species: str = "Rb87", n_runs: int = 1, maxiter: int = 30,
                            popsize: int = 10, seeds: List[int] = None, verbose: bool = True,
                            bounds_rydberg_power_2: Tuple[float, float] = (0.5, 30.0),
                            bounds_rydberg_power_1: Tuple[float, float] = (0.5e-3, 10e-3),
                            bounds_temperature: Tuple[float, float] = (1e-6, 15e-6),
                            bounds_spacing_factor: Tuple[float, float] = (2.0, 5.0),
                            bounds_n_rydberg: Tuple[int, int] = (50, 90),
                            bounds_tweezer_power: Tuple[float, float] = (10e-3, 100e-3),
                            bounds_tweezer_waist: Tuple[float, float] = (0.5e-6, 2.0e-6),
                            bounds_laser_linewidth: Tuple[float, float] = (100.0, 10e3),
                            bounds_delta_over_omega: Tuple[float, float] = (0.32, 0.42),
                            bounds_omega_tau: Tuple[float, float] = (3.9, 4.8), *, device: int = 0) -> ExplorationResult:
    """All hardware knobs at once, every evaluated point kept, Pareto front at the end
    (RG/optimization.py:746-980; objective (1-F) + 0.001 t_us, DE mutation (0.5, 1), recombination 0.7)."""
    from scipy.optimize import differential_evolution

    seeds = seeds if seeds is not None else [42 + i * 111 for i in range(n_runs)]
    result = ExplorationResult(protocol=protocol, species=species,
                               optimizer_settings=dict(n_runs=n_runs, maxiter=maxiter, popsize=popsize, seeds=seeds))
    t0 = time.time()
    is_lp = protocol.lower() in LP_NAMES
    names = ["rydberg_power_2", "rydberg_power_1", "temperature", "spacing_factor", "n_rydberg", "tweezer_power",
             "tweezer_waist", "laser_linewidth"]
    bounds = [bounds_rydberg_power_2, bounds_rydberg_power_1, bounds_temperature, bounds_spacing_factor,
              bounds_n_rydberg, bounds_tweezer_power, bounds_tweezer_waist, bounds_laser_linewidth]
    if is_lp:
        names += ["delta_over_omega", "omega_tau"]
        bounds += [bounds_delta_over_omega, bounds_omega_tau]
    if verbose:
        print("=" * 60)
        print(f"FULL PARAMETER SPACE EXPLORATION: {protocol.upper()}  (population-at-once on GPU {device})")
        print("=" * 60)
        print(f"Species: {species}; runs: {n_runs} × (maxiter={maxiter}, popsize={popsize}); {len(names)} knobs")
    state = dict(best=0.0, fastest=float("inf"))

    def objective(X):
        cols, single, P = _columns(names, X)
        # the reference's fixed choices here: waists 1 µm / 10 µm, default Delta_e, full thermal noise
        ev = evaluate_hardware_population(protocol, cols, species=species, laser_1_waist=1.0e-6, laser_2_waist=10e-6,
                                          include_noise=True, device=device)
        cost = np.ones(P)
        for i in np.nonzero(ev["ok"])[0]:
            F, t = float(ev["fidelity"][i]), float(ev["gate_time_ns"][i])
            result.add_point(EvaluatedPoint(
                Omega_MHz=float(ev["Omega_MHz"][i]), laser_linewidth_kHz=float(cols["laser_linewidth"][i]) / 1e3,
                V_over_Omega=float(ev["V_over_Omega"][i]), fidelity=F, gate_time_ns=t, infidelity=1 - F,
                noise_breakdown={k: float(ev[k][i]) for k in NOISE_KEYS}, protocol=protocol, species=species))
            cost[i] = (1 - F) + 0.001 * (t / 1000)
            state["best"] = max(state["best"], F)
            if F > 0.95:
                state["fastest"] = min(state["fastest"], t)
        if verbose:
            print(f"  [{result.n_evaluations:5d}] best F={state['best'] * 100:.2f}%, "
                  f"fastest (F>95%)={state['fastest']:.0f}ns")
        return float(cost[0]) if single else cost

    for k, seed in enumerate(seeds):
        if verbose:
            print(f"\nRun {k + 1}/{n_runs} (seed={seed})...")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            differential_evolution(objective, bounds=bounds, maxiter=maxiter, popsize=popsize, seed=seed, disp=False,
                                   updating="deferred", vectorized=True, polish=False, mutation=(0.5, 1.0),
                                   recombination=0.7)
    result.compute_pareto_front()
    result.runtime_seconds = time.time() - t0
    if verbose:
        print(f"\n{result.summary()}")
    return result


def combine_explorations(*results: ExplorationResult) -> ExplorationResult:
    if not results:
        raise ValueError("No results to combine")
    out = ExplorationResult(protocol=results[0].protocol, species=results[0].species)
    for r in results:
        out.points.extend(r.points)
        out.runtime_seconds += r.runtime_seconds
    out.n_evaluations = len(out.points)
    out.compute_pareto_front()
    return out


# ---------------------------------------------------------------------------- inverse problem
def optimize_CZ_parameters(target_fidelity: float = 0.99, target_gate_time_ns: float = 300.0,
                           protocol: str = "levine_pichler", weight_fidelity: float = 1.0, weight_time: float = 0.5,
                           constraint_penalty: float = 100.0, species: str = "Rb87",
                           background_loss_rate_hz: float = 10.0, include_noise: bool = True,
                           include_motional_dephasing: bool = True,
                           bounds_rydberg_power_2: Tuple[float, float] = (0.5, 50.0),
                           bounds_rydberg_power_1: Tuple[float, float] = (0.1e-3, 20e-3),
                           bounds_temperature: Tuple[float, float] = (0.1e-6, 20e-6),
                           bounds_spacing_factor: Tuple[float, float] = (1.8, 6.0),
                           bounds_n_rydberg: Tuple[int, int] = (40, 100),
                           bounds_tweezer_power: Tuple[float, float] = (5e-3, 200e-3),
                           bounds_tweezer_waist: Tuple[float, float] = (0.4e-6, 3.0e-6),
                           bounds_Delta_e: Tuple[float, float] = (0.5e9, 15e9),
                           bounds_laser_linewidth: Tuple[float, float] = (100.0, 50e3),
                           bounds_delta_over_omega: Tuple[float, float] = (0.30, 0.45),
                           bounds_omega_tau_lp: Tuple[float, float] = (3.8, 5.0),
                           bounds_omega_tau_jp: Tuple[float, float] = (5.5, 8.5),
                           optimize_protocol_params: bool = True, couple_powers: bool = False,
                           power_ratio_780_480: float = 0.001, maxiter: int = 100, tol: float = 1e-5,
                           seed: Optional[int] = 42, polish: bool = True, workers: int = 1, popsize: int = 15,
                           fixed_params: Optional[Dict[str, float]] = None,
                           callback: Optional[Callable[[int, float, Dict], None]] = None, verbose: bool = True, *,
                           device: int = 0) -> HardwareOptimizationResult:
    """Hardware knobs that hit a target (fidelity, gate time) (RG/optimization.py:280-743).  Cost:
    w_F (1 - F/F*)^2 + w_t ((t - t*)/t*)^2 + penalty [(10 - V/Ω)^2 if V/Ω < 10, trap overlap, T floor].
    ``workers`` is accepted and ignored: a generation is already one GPU launch."""
    from scipy.optimize import differential_evolution

    fixed_params = dict(fixed_params or {})
    is_lp = protocol.lower() in LP_NAMES
    if not is_lp and protocol.lower() not in JP_NAMES:
        raise ValueError(f"Unknown protocol '{protocol}'. Use 'levine_pichler' (or 'lp') "
                         f"or 'jandura_pupillo' / 'smooth_jp' (or 'jp').")
    cfg: Dict[str, Tuple[float, float]] = {}
    if couple_powers:
        cfg["total_power"] = tuple(bounds_rydberg_power_2)
    else:
        cfg["rydberg_power_2"], cfg["rydberg_power_1"] = bounds_rydberg_power_2, bounds_rydberg_power_1
    cfg.update(temperature=bounds_temperature, spacing_factor=bounds_spacing_factor, n_rydberg=bounds_n_rydberg,
               tweezer_power=bounds_tweezer_power, tweezer_waist=bounds_tweezer_waist,
               laser_linewidth=bounds_laser_linewidth, Delta_e=bounds_Delta_e)
    if optimize_protocol_params:
        if is_lp:
            cfg["delta_over_omega"] = bounds_delta_over_omega
            cfg["omega_tau"] = bounds_omega_tau_lp
        else:
            cfg["omega_tau"] = bounds_omega_tau_jp
    names = [k for k in cfg if k not in fixed_params]
    bounds = [cfg[k] for k in names]
    defaults = dict(rydberg_power_1=50e-6, rydberg_power_2=500e-3, temperature=5e-6, spacing_factor=3.0, n_rydberg=70,
                    tweezer_power=50e-3, tweezer_waist=1e-6, laser_linewidth=1000.0, Delta_e=5e9)
    defaults.update(fixed_params)
    if verbose:
        print(f"CZ gate multi-parameter optimisation: {'LP' if is_lp else 'JP'}, target F={target_fidelity:.4f}, "
              f"t={target_gate_time_ns:.1f} ns, {species}, {len(names)} knobs, population-at-once on GPU {device}")
    st = dict(n=0, best=float("inf"), rec=None)

    def objective(X):
        cols, single, P = _columns(names, X)
        cols = dict(cols)
        if couple_powers and "total_power" in cols:
            p2 = cols.pop("total_power")
            cols["rydberg_power_2"], cols["rydberg_power_1"] = p2, p2 * power_ratio_780_480
        ev = evaluate_hardware_population(protocol, cols, species=species, laser_1_waist=50e-6, laser_2_waist=50e-6,
                                          include_noise=include_noise,
                                          include_motional_dephasing=include_motional_dephasing,
                                          background_loss_rate_hz=background_loss_rate_hz, defaults=defaults,
                                          device=device)
        get = lambda k: cols[k] if k in cols else np.full(P, float(defaults[k]))
        F, t, voo = ev["fidelity"], ev["gate_time_ns"], ev["V_over_Omega"]
        pen = np.where(voo < 10, (10 - voo) ** 2, 0.0)
        waist, T = get("tweezer_waist"), get("temperature")
        spacing = get("spacing_factor") * waist
        pen = pen + np.where(spacing < 2 * waist, ((2 * waist - spacing) / waist) ** 2, 0.0)
        pen = pen + np.where(T < 0.05e-6, ((0.05e-6 - T) / 1e-6) ** 2, 0.0)
        cost = (weight_fidelity * (1 - F / target_fidelity) ** 2
                + weight_time * ((t - target_gate_time_ns) / target_gate_time_ns) ** 2 + constraint_penalty * pen)
        cost = np.where(ev["ok"], cost, 1e6)
        for i in range(P):
            st["n"] += 1
            if cost[i] < st["best"]:
                st["best"] = float(cost[i])
                rec = {k: float(v[i]) for k, v in cols.items()}
                for k, v in defaults.items():
                    rec.setdefault(k, v)
                rec["n_rydberg"] = int(round(rec["n_rydberg"]))
                st["rec"] = (rec, float(F[i]), float(t[i]), float(voo[i]), {k: float(ev[k][i]) for k in NOISE_KEYS})
            if callback is not None:
                callback(st["n"], float(cost[i]), {k: float(v[i]) for k, v in cols.items()})
        if verbose and st["rec"] is not None:
            r = st["rec"]
            print(f"  [Eval {st['n']:5d}] best: F={r[1]:.4f}, t={r[2]:.1f}ns, V/Ω={r[3]:.1f}, cost={st['best']:.2e}")
        return float(cost[0]) if single else cost

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        de = differential_evolution(objective, bounds=bounds, maxiter=maxiter, tol=tol, seed=seed, polish=False,
                                    popsize=popsize, disp=False, updating="deferred", vectorized=True)
        if polish:
            from scipy.optimize import minimize
            minimize(objective, de.x, method="L-BFGS-B", bounds=bounds)
    if st["rec"] is None:
        raise RuntimeError("no candidate could be evaluated")
    params, F, t, voo, noise = st["rec"]
    dF, dt = (F / target_fidelity - 1) * 100, (t / target_gate_time_ns - 1) * 100
    tolF = 2.0 if target_fidelity > 0.995 else 5.0
    ok = abs(dF) < tolF and abs(dt) < 30.0 and voo >= 8
    if ok:
        msg = "Optimization converged successfully"
    else:
        issues = ([f"fidelity error {dF:+.1f}%"] if abs(dF) >= tolF else []) + \
                 ([f"time error {dt:+.1f}%"] if abs(dt) >= 30.0 else []) + ([f"V/Ω={voo:.1f} < 8"] if voo < 8 else [])
        msg = f"Optimization incomplete: {', '.join(issues)}"
    if verbose:
        print(f"Optimization complete: {st['n']} evaluations; achieved F={F:.4f}, t={t:.1f} ns, V/Ω={voo:.1f}; {msg}")
    return HardwareOptimizationResult(success=bool(ok), target_fidelity=target_fidelity,
                                      target_gate_time_ns=target_gate_time_ns, achieved_fidelity=F,
                                      achieved_gate_time_ns=t, fidelity_error_pct=dF, gate_time_error_pct=dt,
                                      optimal_parameters=params, V_over_Omega=voo, noise_breakdown=noise,
                                      n_evaluations=st["n"], final_cost=st["best"], message=msg)


__all__ = ["HardwareOptimizationResult", "optimize_CZ_parameters", "EvaluatedPoint", "ExplorationResult",
           "explore_parameter_space", "combine_explorations", "evaluate_hardware_population"]
