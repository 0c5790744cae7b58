"""Population-at-once objective for the reference's optimisers.

The reference's ``optimize_cz_gate`` drives ``scipy.optimize.differential_evolution`` with a
closure that simulates ONE candidate per call (RG/optimize_cz_gate.py:1081-1177, 1250-1260).
SciPy can hand the objective the whole population instead (``vectorized=True,
updating='deferred'``: x has shape (n_params, pop)); this module provides that objective on top of
``simulate_CZ_gate_batch`` so a generation is one kernel launch.  Parameter vectors, bounds and
starting points are the reference's (RG/optimize_cz_gate.py:458-643); the cost is its
``compute_cost`` (:362-431).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .configurations import LaserParameters, NoiseSourceConfig, TwoPhotonExcitationConfig
from .protocols import (JP_OMEGA_TAU_VALIDATED, JP_PHASES_DEFAULT, JP_PHASES_VALIDATED,
                        JP_SWITCHING_TIMES_DEFAULT, JP_SWITCHING_TIMES_VALIDATED, LEVINE_PICHLER_PARAMS,
                        SMOOTH_JP_PARAMS)
from .simulation import BatchResult, simulate_CZ_gate_batch


@dataclass
class ApparatusConstraints:
    """Hardware the optimiser may not change (RG/optimize_cz_gate.py:153-281; same defaults)."""

    laser_1_power: float = 50e-6
    laser_1_waist: float = 50e-6
    laser_2_power: float = 0.3
    laser_2_waist: float = 50e-6
    Delta_e: float = 2 * np.pi * 1e9
    laser_1_linewidth_hz: float = 100.0
    laser_2_linewidth_hz: float = 100.0
    temperature: float = 2e-6
    spacing_factor: float = 2.8
    n_rydberg: int = 70
    species: str = "Rb87"
    tweezer_power: float = 0.020
    tweezer_waist: float = 0.8e-6
    B_field: float = 1e-4
    NA: float = 0.5
    counter_propagating: bool = True

    def fingerprint(self) -> str:
        """12 hex digits identifying the apparatus in cache keys (RG/optimize_cz_gate.py:215-225)."""
        vals = (round(self.laser_1_power, 8), round(self.laser_1_waist, 8), round(self.laser_2_power, 8),
                round(self.laser_2_waist, 8), round(self.Delta_e, 2), self.n_rydberg, round(self.spacing_factor, 4),
                round(self.temperature, 10), self.species, round(self.tweezer_power, 6),
                round(self.tweezer_waist, 8), round(self.NA, 3))
        return hashlib.md5(str(vals).encode()).hexdigest()[:12]

    def make_excitation_config(self, pol_purity: float = 1.0) -> TwoPhotonExcitationConfig:
        mk = lambda p, w, pol, lw: LaserParameters(power=p, waist=w, polarization=pol,
                                                   polarization_purity=pol_purity, linewidth_hz=lw)
        return TwoPhotonExcitationConfig(
            laser_1=mk(self.laser_1_power, self.laser_1_waist, "pi", self.laser_1_linewidth_hz),
            laser_2=mk(self.laser_2_power, self.laser_2_waist, "sigma+", self.laser_2_linewidth_hz),
            Delta_e=self.Delta_e, counter_propagating=self.counter_propagating)

    @staticmethod
    def make_noiseless() -> NoiseSourceConfig:
        return NoiseSourceConfig(include_spontaneous_emission=False, include_intermediate_scattering=False,
                                 include_motional_dephasing=False, include_doppler_dephasing=False,
                                 include_intensity_noise=False, intensity_noise_frac=0.0,
                                 include_laser_dephasing=False, include_magnetic_dephasing=False)

    @staticmethod
    def make_full_noise() -> NoiseSourceConfig:
        return NoiseSourceConfig(include_spontaneous_emission=True, include_intermediate_scattering=True,
                                 include_motional_dephasing=True, include_doppler_dephasing=True,
                                 include_intensity_noise=True, intensity_noise_frac=0.01,
                                 include_laser_dephasing=True, include_magnetic_dephasing=True)

    def plan_kwargs(self, include_noise: bool, noise: NoiseSourceConfig = None) -> Dict:
        """Keywords of ``simulate_CZ_gate_batch`` for this apparatus (what ``simulate_CZ_gate`` would
        receive through ``make_excitation_config`` + the noise record)."""
        pur = 0.99 if include_noise else 1.0          # RG/optimize_cz_gate.py:885-887
        kw = dict(laser_1_power=self.laser_1_power, laser_1_waist=self.laser_1_waist,
                  laser_2_power=self.laser_2_power, laser_2_waist=self.laser_2_waist, Delta_e=self.Delta_e,
                  laser_1_linewidth_hz=self.laser_1_linewidth_hz, laser_2_linewidth_hz=self.laser_2_linewidth_hz,
                  polarization_purity_1=pur, polarization_purity_2=pur,
                  counter_propagating=self.counter_propagating, temperature=self.temperature,
                  spacing_factor=self.spacing_factor, n_rydberg=self.n_rydberg, species=self.species,
                  tweezer_power=self.tweezer_power, tweezer_waist=self.tweezer_waist, B_field=self.B_field,
                  NA=self.NA, include_noise=include_noise)
        if noise is not None:
            kw.update(include_motional_dephasing=noise.include_motional_dephasing,
                      include_doppler_dephasing=noise.include_doppler_dephasing,
                      include_intensity_noise=noise.include_intensity_noise,
                      intensity_noise_frac=noise.intensity_noise_frac)
        return kw


def bounds_and_x0(protocol: str, n_segments: int = 5) -> Tuple[List[Tuple[float, float]], np.ndarray, List[str]]:
    """Search box, start point and parameter names of the reference (RG/optimize_cz_gate.py:544-643)."""
    if protocol == "lp":
        return ([(0.20, 0.50), (3.5, 5.5)],
                np.array([LEVINE_PICHLER_PARAMS["delta_over_omega"], LEVINE_PICHLER_PARAMS["omega_tau"]]),
                ["delta_over_omega", "omega_tau"])
    if protocol == "sjp":
        return ([(5.0, 25.0), (0.05 * np.pi, np.pi), (0.5, 3.0), (0.0, 2 * np.pi), (0.001, 0.10)],
                np.array([SMOOTH_JP_PARAMS["omega_tau"], SMOOTH_JP_PARAMS["A"], SMOOTH_JP_PARAMS["omega_mod_ratio"],
                          SMOOTH_JP_PARAMS["phi_offset"], abs(SMOOTH_JP_PARAMS["delta_over_omega"])]),
                ["omega_tau", "A", "omega_mod_ratio", "phi_offset", "delta_over_omega"])
    if protocol == "jp":
        if n_segments == 5:
            ot0, sw, ph, ot_b = JP_OMEGA_TAU_VALIDATED, JP_SWITCHING_TIMES_VALIDATED, JP_PHASES_VALIDATED, (5.0, 40.0)
        elif n_segments == 7:
            ot0, sw, ph, ot_b = 7.0, JP_SWITCHING_TIMES_DEFAULT, JP_PHASES_DEFAULT, (3.0, 30.0)
        else:
            raise ValueError(f"Unsupported n_segments: {n_segments}. Use 5 or 7.")
        ns = n_segments - 1
        return ([ot_b] + [(0.01, 0.99)] * ns + [(-np.pi, np.pi)] * n_segments,
                np.array([ot0] + [t / ot0 for t in sw] + list(ph)),
                ["omega_tau"] + [f"frac{i + 1}" for i in range(ns)] + [f"phi{i}" for i in range(n_segments)])
    raise ValueError(f"Unknown protocol: {protocol}. Use 'lp', 'jp' or 'sjp'.")


def protocol_kwargs(protocol: str, X: np.ndarray, n_segments: int = 5) -> Dict:
    """Rows of X (n_params, pop) -> build_plan keywords, exactly as _build_*_inputs do
    (RG/optimize_cz_gate.py:458-537): bang-bang fractions are sorted, then scaled by omega_tau."""
    if protocol == "lp":
        return dict(delta_over_omega=X[0], omega_tau=X[1], pulse_shape="square")
    if protocol == "sjp":
        return dict(omega_tau=X[0], A=X[1], omega_mod_ratio=X[2], phi_offset=X[3], delta_over_omega=X[4])
    ns = n_segments - 1
    fr = np.sort(X[1:1 + ns], axis=0)
    return dict(omega_tau=X[0], switching_times=(fr * X[0][None, :]).T, phases=X[1 + ns:1 + ns + n_segments].T)


def make_vectorized_objective(protocol: str, apparatus: ApparatusConstraints = None, include_noise: bool = False,
                              time_weight: float = 0.01, n_segments: int = 5, optimize_spacing: bool = False,
                              device: int = 0, on_batch: Optional[Callable[[np.ndarray, BatchResult], None]] = None):
    """f(X) with X of shape (n_params, pop) [or (n_params,)] -> cost array of length pop.
    Any failure of the batch maps to the reference's penalty 1e6 (RG/optimize_cz_gate.py:1174-1177)."""
    apparatus = apparatus or ApparatusConstraints()
    base = apparatus.plan_kwargs(include_noise)

    def objective(X: np.ndarray) -> np.ndarray:
        X = np.asarray(X, dtype=float)
        single = X.ndim == 1
        if single:
            X = X[:, None]
        kw = dict(base)
        n_proto = X.shape[0] - (1 if optimize_spacing else 0)
        if optimize_spacing:
            kw["spacing_factor"] = X[-1]
        kw.update(protocol_kwargs(protocol, X[:n_proto], n_segments))
        try:
            res = simulate_CZ_gate_batch(protocol, B=X.shape[1], device=device, warn=False, **kw)
            cost = res.cost(time_weight)
            if on_batch is not None:
                on_batch(X, res)
        except Exception:
            cost = np.full(X.shape[1], 1e6)
        return float(cost[0]) if single else cost

    return objective


def optimize_cz_gate_batched(protocol: str, apparatus: ApparatusConstraints = None, include_noise: bool = False,
                             time_weight: float = 0.01, n_segments: int = 5, maxiter: int = 80, popsize: int = 15,
                             tol: float = 1e-6, seed: int = 42, optimize_spacing: bool = False,
                             spacing_bounds: Tuple[float, float] = (1.5, 5.0), device: int = 0):
    """differential_evolution with the reference's settings (RG/optimize_cz_gate.py:1250-1260) but
    one launch per generation.  Returns SciPy's OptimizeResult plus ``param_names``."""
    from scipy.optimize import differential_evolution

    apparatus = apparatus or ApparatusConstraints()
    bounds, x0, names = bounds_and_x0(protocol, n_segments)
    if optimize_spacing:
        bounds = list(bounds) + [spacing_bounds]
        x0 = np.append(x0, apparatus.spacing_factor)
        names = names + ["spacing_factor"]
    f = make_vectorized_objective(protocol, apparatus, include_noise, time_weight, n_segments, optimize_spacing, device)
    res = differential_evolution(f, bounds, x0=x0, maxiter=maxiter, popsize=popsize, tol=tol, seed=seed,
                                 polish=False, vectorized=True, updating="deferred")
    res.param_names = names
    return res
