"""Protocol-agnostic CZ-gate optimiser, population-at-once (reference: RG/optimize_cz_gate.py).

Same entry points, arguments, result records, cost function, search boxes, cache keys and
discrete-variant logic as the reference module, so scripts and notebooks that call
``optimize_cz_gate`` / ``run_baseline`` / ``compute_cost`` keep working.  What changes is how a
generation of ``scipy.optimize.differential_evolution`` is evaluated: the reference simulates one
candidate per objective call (RG/optimize_cz_gate.py:1081-1177, one QuTiP mesolve loop each);
here SciPy hands over the whole population (``vectorized=True, updating='deferred'`` -- the
reference's own ``explore_parameter_space`` already uses deferred updating,
RG/optimization.py:959) and the candidates that miss the ``SimulationCache`` go to the GPU as ONE
``simulate_CZ_gate_batch`` call (``rydcz_simulate_host_cols``).  ``population_shards=(rank, world)``
spreads a generation over the GPUs of a box (every rank runs the same DE with the same seed and
all-gathers the per-candidate metrics).
"""
from __future__ import annotations

import hashlib
import json
import os
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, NamedTuple, Optional, Tuple

import numpy as np

from .batched_objective import ApparatusConstraints, bounds_and_x0, protocol_kwargs
from .configurations import (JPSimulationInputs, LaserParameters, LPSimulationInputs, NoiseSourceConfig,
                             SmoothJPSimulationInputs, TwoPhotonExcitationConfig)
from .simulation import METRIC_OUTPUTS, SimulationResult, simulate_CZ_gate, simulate_CZ_gate_batch

METRIC_KEYS = ("controlled_phase_deg", "phase_error_deg", "cz_phase_fidelity", "f00", "f01", "f10", "f11",
               "avg_fidelity", "gate_time_us", "V_over_Omega", "Omega_MHz")


# ---------------------------------------------------------------------------- cache
class _CacheEntry(NamedTuple):
    cost: float
    metrics: Dict[str, Any]


def _format_key(prefix: str, protocol: str, values, digits: int) -> str:
    """'<apparatus>|<protocol>|(<v0>, <v1>, ...)' with every value rounded to `digits` decimals: the text of
    the reference's keys (RG/optimize_cz_gate.py:306-312), so JSON files written by either side stay readable
    by the other.  The parameter list is printed the way Python prints a tuple of floats."""
    body = ", ".join(repr(round(float(v), digits)) for v in values)
    if len(values) == 1:
        body += ","
    return f"{prefix}|{protocol}|({body})"


class SimulationCache:
    """Memo of evaluated candidates for the optimiser's objective.  Public surface of the reference's class
    (RG/optimize_cz_gate.py:284-352): ``make_key``, ``in`` / ``[]`` access returning ``(cost, metrics)``,
    ``hits`` / ``misses`` / ``hit_rate``, ``precision``, ``save`` / ``load`` on the same JSON layout
    ``{"precision": p, "entries": {key: {"cost": c, "metrics": {...}}}}``.  A caller's own cache object with
    this surface can be passed to ``optimize_cz_gate(cache=...)`` instead."""

    def __init__(self, precision: int = 4):
        self.precision = int(precision)
        self.hits = 0          # successful lookups (counted by __getitem__, as the reference does)
        self.misses = 0        # bumped by the objective when it has to simulate
        self._entries: Dict[str, _CacheEntry] = {}

    # -- keys
    def make_key(self, protocol: str, params: list, apparatus_hash: str = "") -> str:
        return _format_key(apparatus_hash, protocol, list(params), self.precision)

    # -- mapping protocol
    def __len__(self) -> int:
        return len(self._entries)

    def __contains__(self, key: str) -> bool:
        return key in self._entries

    def __getitem__(self, key: str) -> Tuple[float, Dict]:
        entry = self._entries[key]
        self.hits += 1
        return entry.cost, entry.metrics

    def __setitem__(self, key: str, value: Tuple[float, Dict]):
        cost, metrics = value
        self._entries[key] = _CacheEntry(cost, metrics)

    @property
    def hit_rate(self) -> float:
        lookups = self.hits + self.misses
        return self.hits / lookups if lookups else 0.0

    # -- persistence
    def save(self, path: str):
        payload = {"precision": self.precision,
                   "entries": {key: {"cost": e.cost, "metrics": e.metrics} for key, e in self._entries.items()}}
        tmp = f"{path}.tmp"
        with open(tmp, "w") as fh:
            json.dump(payload, fh, indent=2, default=str)
        os.replace(tmp, path)             # never leave a half-written cache behind

    def load(self, path: str):
        """Merge the entries of a saved cache (a missing file is not an error)."""
        try:
            with open(path) as fh:
                payload = json.load(fh)
        except FileNotFoundError:
            return
        self.precision = int(payload.get("precision", self.precision))
        for key, rec in payload.get("entries", {}).items():
            self._entries[key] = _CacheEntry(rec["cost"], rec["metrics"])


_global_cache = SimulationCache(precision=4)


# ---------------------------------------------------------------------------- cost and metrics
def compute_cost(metrics: Dict[str, float], gate_time_us: float = 0.0, time_weight: float = 0.01) -> float:
    """10 (100(1-F))^2 + 5 (100(1-F11))^2 + 2 (100(1-F_phase))^2 + w t_us; 1e6 for NaN or F < 0.5
    (RG/optimize_cz_gate.py:362-431)."""
    F, f11, czf = metrics.get("avg_fidelity", 0.0), metrics.get("f11", 0.0), metrics.get("cz_phase_fidelity", 0.0)
    if any(np.isnan(x) for x in (F, f11, czf)) or F < 0.50:
        return 1e6
    return (10.0 * ((1.0 - F) * 100.0) ** 2 + 5.0 * ((1.0 - f11) * 100.0) ** 2
            + 2.0 * ((1.0 - czf) * 100.0) ** 2 + time_weight * gate_time_us)


def extract_metrics(result: SimulationResult) -> Dict[str, float]:
    """RG/optimize_cz_gate.py:434-451."""
    ph, fid = result.phase_info, result.fidelities
    return {"controlled_phase_deg": ph.get("controlled_phase_deg", np.nan),
            "phase_error_deg": ph.get("phase_error_from_pi_deg", np.nan),
            "cz_phase_fidelity": ph.get("cz_phase_fidelity", np.nan),
            "f00": fid.get("00", np.nan), "f01": fid.get("01", np.nan), "f10": fid.get("10", np.nan),
            "f11": fid.get("11", np.nan), "avg_fidelity": result.avg_fidelity,
            "gate_time_us": result.tau_total * 1e6, "V_over_Omega": result.V_over_Omega,
            "Omega_MHz": result.Omega / (2 * np.pi * 1e6)}


def batch_metric_rows(res) -> np.ndarray:
    """[B, len(METRIC_KEYS)] array of ``extract_metrics`` values for a ``BatchResult`` (any return mode)."""
    c = res.column
    omega, V = c("Omega"), c("V")
    with np.errstate(divide="ignore", invalid="ignore"):
        voo = np.where(omega > 0, V / omega, np.inf)
    cols = {"controlled_phase_deg": np.degrees(c("controlled_phase_rad")),
            "phase_error_deg": np.degrees(c("phase_error_from_pi_rad")), "cz_phase_fidelity": c("cz_phase_fidelity"),
            "f00": c("f00"), "f01": c("f01"), "f10": c("f10"), "f11": c("f11"), "avg_fidelity": c("avg_fidelity"),
            "gate_time_us": c("tau_total") * 1e6, "V_over_Omega": voo, "Omega_MHz": omega / (2 * np.pi * 1e6)}
    return np.stack([np.asarray(cols[k], dtype=float) for k in METRIC_KEYS], axis=1)


def batch_metrics(res) -> List[Dict[str, float]]:
    """``extract_metrics`` for every row of a ``BatchResult``."""
    rows = batch_metric_rows(res)
    return [dict(zip(METRIC_KEYS, map(float, r))) for r in rows]


# ---------------------------------------------------------------------------- parameter builders
def _norm(protocol: str) -> str:
    p = protocol.lower().replace("-", "_").replace(" ", "_")
    if p in ("lp", "levine_pichler"):
        return "lp"
    if p in ("jp_bangbang", "jp", "jandura_pupillo"):
        return "jp"
    if p in ("smooth_jp", "dark_state"):
        return "sjp"
    raise ValueError(f"Unknown protocol: {protocol}. Use 'lp', 'jp_bangbang', or 'smooth_jp'.")


def _build_lp_inputs(params, excitation, noise) -> LPSimulationInputs:
    return LPSimulationInputs(excitation=excitation, noise=noise, delta_over_omega=float(params[0]),
                              omega_tau=float(params[1]), pulse_shape="square")


def _build_jp_bangbang_inputs(params, excitation, noise, n_segments: int = 5) -> JPSimulationInputs:
    ns = n_segments - 1
    ot = float(params[0])
    fr = sorted(float(f) for f in params[1:1 + ns])
    return JPSimulationInputs(excitation=excitation, noise=noise, omega_tau=ot,
                              switching_times=[f * ot for f in fr],
                              phases=[float(p) for p in params[1 + ns:1 + ns + n_segments]])


def _build_smooth_jp_inputs(params, excitation, noise) -> SmoothJPSimulationInputs:
    return SmoothJPSimulationInputs(excitation=excitation, noise=noise, omega_tau=float(params[0]), A=float(params[1]),
                                    omega_mod_ratio=float(params[2]), phi_offset=float(params[3]),
                                    delta_over_omega=float(params[4]))


def _get_lp_bounds_and_x0():
    b, x0, _ = bounds_and_x0("lp")
    return b, x0


def _get_jp_bangbang_bounds_and_x0(n_segments: int = 5):
    b, x0, _ = bounds_and_x0("jp", n_segments)
    return b, x0


def _get_smooth_jp_bounds_and_x0():
    b, x0, _ = bounds_and_x0("sjp")
    return b, x0


def warm_start_bounds(opt_result: "OptimizationResult", frac: float = 0.20, original_bounds: Optional[list] = None):
    """Tight box around a previous optimum (RG/optimize_cz_gate.py:646-704): phases +-frac*pi,
    fractions +-frac inside [0.01, 0.99], everything else +-max(|v| frac, 0.01); clamped to
    ``original_bounds``."""
    x0 = opt_result.best_params.copy()
    out = []
    for i, (name, v) in enumerate(zip(opt_result.param_names, x0)):
        if "phi" in name:
            lo, hi = v - frac * np.pi, v + frac * np.pi
        elif "frac" in name:
            lo, hi = max(0.01, v - frac), min(0.99, v + frac)
        else:
            d = max(abs(v) * frac, 0.01)
            lo, hi = v - d, v + d
        if original_bounds is not None and i < len(original_bounds):
            lo, hi = max(lo, original_bounds[i][0]), min(hi, original_bounds[i][1])
        if lo >= hi:
            lo = hi - 0.01
        out.append((lo, hi))
    return out, x0


# ---------------------------------------------------------------------------- result record
@dataclass
class OptimizationResult:
    """RG/optimize_cz_gate.py:711-781 (same fields)."""

    success: bool
    protocol: str
    best_params: np.ndarray
    param_names: List[str]
    best_cost: float
    best_metrics: Dict[str, float]
    n_evaluations: int
    runtime_s: float
    discrete_variant: str = ""
    all_variants: Dict[str, Any] = field(default_factory=dict)
    cache_hits: int = 0
    n_generations: int = 0           # batched launches (extension)

    def __repr__(self) -> str:
        m = self.best_metrics
        rows = ["=" * 70, f"  CZ Gate Optimisation Result — {self.protocol}", "=" * 70,
                f"  Success:          {self.success}", f"  Variant:          {self.discrete_variant}",
                f"  Cost:             {self.best_cost:.4f}",
                f"  Avg fidelity:     {m.get('avg_fidelity', 0):.6f}  ({(1 - m.get('avg_fidelity', 0)) * 100:.4f}% error)",
                f"  F(|11⟩):          {m.get('f11', 0):.6f}", f"  CZ phase fid:     {m.get('cz_phase_fidelity', 0):.6f}",
                f"  Phase error:      {m.get('phase_error_deg', 999):.2f}°",
                f"  Controlled phase: {m.get('controlled_phase_deg', 0):.2f}°",
                f"  Gate time:        {m.get('gate_time_us', 0):.3f} µs", f"  V/Ω:              {m.get('V_over_Omega', 0):.1f}",
                f"  Ω/2π:             {m.get('Omega_MHz', 0):.3f} MHz", f"  Evaluations:      {self.n_evaluations}",
                f"  Runtime:          {self.runtime_s:.1f} s", f"  Cache hits:       {self.cache_hits}", "-" * 70,
                "  Optimal parameters:"]
        rows += [f"    {n:25s} = {v:.6f}" for n, v in zip(self.param_names, self.best_params)]
        rows.append("=" * 70)
        return "\n".join(rows)


# ---------------------------------------------------------------------------- population objective
def _noise_hash(noise: NoiseSourceConfig) -> str:
    """RG/optimize_cz_gate.py:1066-1079."""
    flags = (f"se{int(noise.include_spontaneous_emission)}is{int(noise.include_intermediate_scattering)}"
             f"md{int(noise.include_motional_dephasing)}dd{int(noise.include_doppler_dephasing)}"
             f"in{int(noise.include_intensity_noise)}ld{int(noise.include_laser_dephasing)}"
             f"mg{int(noise.include_magnetic_dephasing)}")
    if noise.include_intensity_noise and noise.intensity_noise_frac:
        flags += f"_inf{noise.intensity_noise_frac:.4f}"
    return hashlib.md5(flags.encode()).hexdigest()[:8]


class PopulationObjective:
    """f(X), X of shape (n_params, pop) or (n_params,): the reference's objective closure
    (RG/optimize_cz_gate.py:1081-1177) for a whole population.  Cache hits are served from the
    ``SimulationCache``; all misses of a generation are ONE batched GPU evaluation."""

    def __init__(self, proto: str, apparatus: ApparatusConstraints, noise: NoiseSourceConfig, include_noise: bool,
                 time_weight: float, n_segments: int, cache: SimulationCache, optimize_spacing: bool = False,
                 device: int = 0, verbose: bool = False, population_shards: Tuple[int, int] = (0, 1)):
        self.proto, self.apparatus, self.noise = proto, apparatus, noise
        self.include_noise, self.time_weight, self.n_segments = include_noise, time_weight, n_segments
        self.cache, self.optimize_spacing, self.device, self.verbose = cache, optimize_spacing, device, verbose
        self.rank, self.world = population_shards
        self.base = apparatus.plan_kwargs(include_noise, noise)
        self.cache_protocol = {"lp": "lp", "jp": f"jp_bangbang_{n_segments}seg", "sjp": "smooth_jp"}[proto]
        self.apparatus_hash, self.noise_hash = apparatus.fingerprint(), _noise_hash(noise)
        self.eval_count = self.cache_hits = self.generations = 0
        self.best_so_far = 1e9

    @staticmethod
    def _outputs(kw):
        """Only the columns ``extract_metrics`` reads cross PCIe when the device prologue serves the call."""
        host_only = kw.get("pulse_shape", "square") not in ("square", "gaussian", "cosine", "blackman")
        return None if host_only else METRIC_OUTPUTS

    def _key(self, x: np.ndarray) -> str:
        if self.optimize_spacing:
            h = f"{self.apparatus_hash}_sf{float(x[-1]):.6f}"
            x = x[:-1]
        else:
            h = self.apparatus_hash
        return self.cache.make_key(self.cache_protocol, x.tolist(), f"{h}_n{self.noise_hash}")

    def _evaluate(self, X: np.ndarray) -> List[Dict[str, float]]:
        """Metrics of every column of X through one batched launch; with ``population_shards`` every rank evaluates
        its block of the generation and the metric columns are all-gathered (``sharding.simulate_sharded``, which
        keeps a failing rank inside the collective: its block comes back NaN -> the reference's failure cost 1e6
        on every rank alike)."""
        from .. import sharding

        n = X.shape[1]
        kw = dict(self.base)
        n_proto = X.shape[0] - (1 if self.optimize_spacing else 0)
        if self.optimize_spacing:
            kw["spacing_factor"] = np.ascontiguousarray(X[-1])
        kw.update(protocol_kwargs(self.proto, X[:n_proto], self.n_segments))
        if self.world > 1:
            res = sharding.simulate_sharded(self.proto, n, rank=self.rank, world=self.world, device=self.device,
                                            outputs=METRIC_OUTPUTS, **kw)
            if res.local_failure is not None and self.verbose:
                print(f"    rank {self.rank}: shard evaluation failed ({res.local_failure}); contributed NaN rows")
        else:
            res = simulate_CZ_gate_batch(self.proto, B=n, device=self.device, warn=False, outputs=self._outputs(kw), **kw)
        rows = batch_metric_rows(res)
        return [dict(zip(METRIC_KEYS, map(float, r))) for r in rows]

    def __call__(self, X: np.ndarray):
        X = np.asarray(X, dtype=float)
        single = X.ndim == 1
        if single:
            X = X[:, None]
        n = X.shape[1]
        cost = np.empty(n)
        keys = [self._key(X[:, i]) for i in range(n)]
        miss, first = [], {}
        for i, k in enumerate(keys):
            self.eval_count += 1
            if k in self.cache:
                self.cache_hits += 1
                cost[i] = self.cache[k][0]
            elif k in first:                      # duplicate inside the generation: evaluated once
                miss.append(i)
            else:
                first[k] = i
                self.cache.misses += 1
                miss.append(i)
        todo = sorted(first.values())
        if todo:
            self.generations += 1
            try:
                metrics = self._evaluate(X[:, todo])
            except Exception as e:               # the reference maps any failure to 1e6 (:1174-1177)
                if self.verbose:
                    print(f"    [{self.eval_count:4d}] ERROR: {e}")
                metrics = None
            for j, i in enumerate(todo):
                if metrics is None:
                    c = 1e6
                else:
                    m = metrics[j]
                    c = compute_cost(m, gate_time_us=m.get("gate_time_us", 0), time_weight=self.time_weight)
                    self.cache[keys[i]] = (c, m)
                    if c < self.best_so_far:
                        self.best_so_far = c
                        if self.verbose:
                            print(f"    [{self.eval_count:4d}] cost={c:10.4f}  F={m['avg_fidelity']:.6f}  "
                                  f"F11={m['f11']:.6f}  CZφ={m['cz_phase_fidelity']:.4f}  "
                                  f"φ_err={m['phase_error_deg']:.2f}°  t={m['gate_time_us']:.3f}µs")
                first[keys[i]] = c
            for i in miss:
                cost[i] = first[keys[i]]
        return float(cost[0]) if single else cost


# ---------------------------------------------------------------------------- optimiser
def optimize_cz_gate(protocol: str, apparatus: ApparatusConstraints, include_noise: bool = False,
                     noise_config: Optional[NoiseSourceConfig] = None, time_weight: float = 0.01,
                     optimize_spacing: bool = False, spacing_bounds: Optional[Tuple[float, float]] = None,
                     maxiter: int = 80, popsize: int = 15, tol: float = 1e-6, seed: int = 42,
                     bounds: Optional[list] = None, x0: Optional[np.ndarray] = None,
                     cache: Optional[SimulationCache] = None, cache_path: Optional[str] = None,
                     strategy: str = "standard", variant: Optional[str] = None, verbose: bool = True, *,
                     device: int = 0, population_shards: Tuple[int, int] = (0, 1),
                     polish: bool = True) -> OptimizationResult:
    """Same contract as the reference's ``optimize_cz_gate`` (RG/optimize_cz_gate.py:786-990):
    differential evolution over the protocol parameters (optionally + spacing_factor), bang-bang
    tried as 5- and 7-segment variants, best variant returned.  Keyword-only extensions: ``device``,
    ``population_shards=(rank, world)``, ``polish`` (L-BFGS-B polish of the winner, one candidate
    per call like the reference)."""
    proto = _norm(protocol)
    protocol_norm = protocol.lower().replace("-", "_").replace(" ", "_")
    noise = noise_config if noise_config is not None else (
        apparatus.make_full_noise() if include_noise else apparatus.make_noiseless())
    cache = cache if cache is not None else _global_cache
    if cache_path:
        cache.load(cache_path)
    if proto == "jp":
        allowed = {"5-segment": 5, "7-segment": 7}
        if variant is not None:
            if variant not in allowed:
                raise ValueError(f"Unknown variant '{variant}'. Use '5-segment' or '7-segment'.")
            variants = {variant: allowed[variant]}
        else:
            variants = dict(allowed)
    else:
        variants = {"default": None}
    if verbose:
        print("=" * 70)
        print(f"  CZ Gate Optimisation — {protocol}  (population-at-once on GPU {device})")
        print("=" * 70)
        print(f"  Species: {apparatus.species}, n_rydberg: {apparatus.n_rydberg}, spacing factor: "
              f"{apparatus.spacing_factor}{' (will be optimised)' if optimize_spacing else ''}, "
              f"noise: {'ON' if include_noise else 'OFF'}")
    results: Dict[str, OptimizationResult] = {}
    for name, n_seg in variants.items():
        r = _optimize_single_variant(proto, protocol_norm, n_seg, apparatus, noise, include_noise, time_weight, maxiter,
                                     popsize, tol, seed, bounds, x0, cache, strategy, verbose, optimize_spacing,
                                     spacing_bounds, device, population_shards, polish)
        r.discrete_variant = name
        results[name] = r
        if verbose:
            print(f"    → {name}: cost={r.best_cost:.4f}, F={r.best_metrics.get('avg_fidelity', 0):.6f}, "
                  f"phase_err={r.best_metrics.get('phase_error_deg', 999):.2f}°")
    best_name = min(results, key=lambda k: results[k].best_cost)
    best = results[best_name]
    best.all_variants = {k: {"cost": v.best_cost, "avg_fidelity": v.best_metrics.get("avg_fidelity", 0),
                             "phase_error_deg": v.best_metrics.get("phase_error_deg", 999),
                             "params": v.best_params.tolist()} for k, v in results.items()}
    if cache_path:
        cache.save(cache_path)
    if verbose:
        print(f"\n  Best variant: {best_name}")
        print(best)
    return best


def _optimize_single_variant(proto, protocol_norm, n_segments, apparatus, noise, include_noise, time_weight, maxiter,
                             popsize, tol, seed, bounds, x0, cache, strategy, verbose, optimize_spacing,
                             spacing_bounds, device, population_shards, polish) -> OptimizationResult:
    from scipy.optimize import differential_evolution

    n_seg = n_segments if n_segments is not None else 5
    d_bounds, d_x0, names = bounds_and_x0(proto, n_seg)
    names = list(names)
    bounds = list(bounds) if bounds is not None else list(d_bounds)
    x0 = np.array(x0 if x0 is not None else d_x0, dtype=float)
    if optimize_spacing:
        bounds = bounds + [spacing_bounds or (1.5, 5.0)]
        x0 = np.append(x0, apparatus.spacing_factor)
        names.append("spacing_factor")
    f = PopulationObjective(proto, apparatus, noise, include_noise, time_weight, n_seg, cache, optimize_spacing,
                            device, verbose, population_shards)
    t0 = time.time()
    de_kw = dict(tol=tol, seed=seed, polish=False, disp=False, vectorized=True, updating="deferred")
    if strategy == "two_phase" and proto == "sjp":
        # phase 1: omega_tau (and spacing) only, the other parameters at their start values
        # (RG/optimize_cz_gate.py:1184-1243)
        fixed = x0.copy()
        idx = [0] + ([len(x0) - 1] if optimize_spacing else [])

        def phase1(S):
            S = np.asarray(S, dtype=float)
            cols = S[:, None] if S.ndim == 1 else S
            X = np.repeat(fixed[:, None], cols.shape[1], axis=1)
            X[idx] = cols
            c = f(X)
            return float(c[0]) if S.ndim == 1 else c

        p1 = differential_evolution(phase1, bounds=[bounds[i] for i in idx], x0=x0[idx],
                                    maxiter=max(20, maxiter // 4), popsize=10, **de_kw)
        x0[idx] = p1.x
        ot = p1.x[0]
        bounds[0] = (max(bounds[0][0], ot * 0.7), min(bounds[0][1], ot * 1.3))
        if verbose:
            print(f"    Phase 1 best Ωτ = {ot:.3f}, cost = {p1.fun:.4f}")
    x0 = np.clip(x0, [b[0] for b in bounds], [b[1] for b in bounds])
    if verbose:
        print(f"    Starting differential_evolution: maxiter={maxiter}, popsize={popsize}, dim={len(bounds)}, "
              f"strategy={strategy}, one launch per generation")
    de = differential_evolution(f, bounds=bounds, x0=x0, maxiter=maxiter, popsize=popsize, **de_kw)
    best_x, best_c = de.x, float(de.fun)
    if polish:
        # the reference polishes with L-BFGS-B (SciPy's polish=True); finite differences go through the
        # same objective one column at a time
        from scipy.optimize import minimize
        pol = minimize(f, best_x, method="L-BFGS-B", bounds=bounds)
        if pol.fun < best_c:
            best_x, best_c = pol.x, float(pol.fun)
    key = f._key(np.asarray(best_x, dtype=float))
    if key in cache:
        _, metrics = cache[key]
    else:
        f(np.asarray(best_x, dtype=float))
        metrics = cache[key][1] if key in cache else {"avg_fidelity": 0, "phase_error_deg": 999, "f11": 0}
    ok = (metrics.get("avg_fidelity", 0) >= 0.99 and metrics.get("cz_phase_fidelity", 0) >= 0.99
          and metrics.get("phase_error_deg", 999) < 10.0)
    return OptimizationResult(success=bool(ok), protocol=protocol_norm, best_params=np.asarray(best_x), param_names=names,
                              best_cost=best_c, best_metrics=metrics, n_evaluations=f.eval_count,
                              runtime_s=time.time() - t0, cache_hits=f.cache_hits, n_generations=f.generations)


def run_baseline(protocol: str, apparatus: ApparatusConstraints, include_noise: bool = False,
                 verbose: bool = True) -> Dict[str, float]:
    """One evaluation at the protocol's default parameters (RG/optimize_cz_gate.py:1331-1407)."""
    proto = _norm(protocol)
    exc = apparatus.make_excitation_config(pol_purity=0.99 if include_noise else 1.0)
    noise = apparatus.make_full_noise() if include_noise else apparatus.make_noiseless()
    cls = {"lp": LPSimulationInputs, "jp": JPSimulationInputs, "sjp": SmoothJPSimulationInputs}[proto]
    res = simulate_CZ_gate(simulation_inputs=cls(excitation=exc, noise=noise), species=apparatus.species,
                           n_rydberg=apparatus.n_rydberg, spacing_factor=apparatus.spacing_factor,
                           temperature=apparatus.temperature, tweezer_power=apparatus.tweezer_power,
                           tweezer_waist=apparatus.tweezer_waist, B_field=apparatus.B_field, NA=apparatus.NA,
                           include_noise=include_noise, verbose=verbose, return_dataclass=True)
    m = extract_metrics(res)
    if verbose:
        print(f"\n  Baseline — {protocol}")
        for label, k, fmt in (("Avg fidelity", "avg_fidelity", ".6f"), ("F(|11⟩)", "f11", ".6f"),
                              ("CZ phase fid", "cz_phase_fidelity", ".6f"), ("Phase error", "phase_error_deg", ".2f"),
                              ("Ctrl phase", "controlled_phase_deg", ".2f"), ("Gate time", "gate_time_us", ".3f"),
                              ("V/Ω", "V_over_Omega", ".1f"), ("Ω/2π", "Omega_MHz", ".3f")):
            print(f"  {label + ':':18s}{format(m[k], fmt)}")
        print(f"  Cost (ref):       {compute_cost(m, m['gate_time_us']):.4f}")
    return m


__all__ = ["ApparatusConstraints", "SimulationCache", "compute_cost", "extract_metrics", "batch_metrics",
           "warm_start_bounds", "OptimizationResult", "PopulationObjective", "optimize_cz_gate", "run_baseline"]
