"""Noise rates of the two-atom Lindblad model (reference: RG/noise_models.py).

Part 1 (scalar rates, RG/noise_models.py:310-959) is restated elementwise.  Part 2 of the
reference builds ~14 QuTiP collapse operators; Lindblad terms with the same operator add in
rate, so on the B200 path the operator list is replaced by the handful of rates the kernel
needs (``collapse_rates``) plus the same ``noise_breakdown`` dictionary the reference returns.
NumPy versions of the operator builders are kept for callers that want the matrices.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from .constants import HBAR, KB, MU_B


@dataclass
class NoiseRates:
    """Bundle of rates (RG/noise_models.py:190-303)."""

    gamma_r: float = 0.0
    gamma_bbr: float = 0.0
    gamma_phi_laser: float = 0.0
    gamma_phi_thermal: float = 0.0
    gamma_phi_zeeman: float = 0.0
    gamma_loss_antitrap: float = 0.0
    gamma_loss_bg: float = 0.0
    gamma_scatter: float = 0.0
    gamma_leakage: float = 0.0
    gamma_mJ: float = 0.0

    @property
    def total_decay_rate(self) -> float:
        return (self.gamma_r + self.gamma_bbr + self.gamma_loss_antitrap + self.gamma_loss_bg
                + self.gamma_scatter + self.gamma_leakage)

    @property
    def total_dephasing_rate(self) -> float:
        return self.gamma_phi_laser + self.gamma_phi_thermal + self.gamma_phi_zeeman + self.gamma_mJ

    @property
    def total_T2_rate(self) -> float:
        return 0.5 * self.total_decay_rate + self.total_dephasing_rate

    def to_dict(self) -> dict:
        d = {k: getattr(self, k) for k in self.__dataclass_fields__}
        d["total_decay"] = self.total_decay_rate
        d["total_dephasing"] = self.total_dephasing_rate
        return d

    def summary_table(self, gate_time: float = 1e-6) -> str:
        names = [("Rydberg decay", self.gamma_r), ("BBR decay", self.gamma_bbr),
                 ("Laser dephasing", self.gamma_phi_laser), ("Thermal dephasing", self.gamma_phi_thermal),
                 ("Zeeman dephasing", self.gamma_phi_zeeman), ("Anti-trap loss", self.gamma_loss_antitrap),
                 ("Background loss", self.gamma_loss_bg), ("Int. state scatter", self.gamma_scatter),
                 ("Rydberg leakage", self.gamma_leakage), ("mJ mixing", self.gamma_mJ)]
        bar = "=" * 60
        out = [bar, "NOISE RATE SUMMARY", bar, f"{'Source':<30} {'Rate (kHz)':<15} {'Error/gate (%)'}", "-" * 60]
        for name, rate in names:
            out.append(f"{name:<30} {rate / (2 * np.pi * 1e3):<15.2f} {rate * gate_time * 100:.3f}")
        for name, rate in (("TOTAL DECAY", self.total_decay_rate), ("TOTAL DEPHASING", self.total_dephasing_rate)):
            out.append(f"{name:<30} {rate / (2 * np.pi * 1e3):<15.2f} {rate * gate_time * 100:.3f}")
        out.append(bar)
        return "\n".join(out)


# ------------------------------------------------------------------ scalar rates (elementwise)
def rydberg_decay_rate(lifetime):
    return 1.0 / lifetime if lifetime > 0 else 0.0


def bbr_decay_rate(lifetime_0K, lifetime_T):
    if lifetime_0K <= 0 or lifetime_T <= 0:
        return 0.0
    return max(0.0, 1.0 / lifetime_T - 1.0 / lifetime_0K)


def laser_dephasing_rate(linewidth_hz):
    return 2 * np.pi * linewidth_hz


def zeeman_dephasing_rate(B_noise_gauss, qubit_type: str = "clock", K_quad: float = 575.0):
    """2π·(2 K B_bias δB) for clock states at a 1 G bias; 2π·700 kHz/G otherwise
    (RG/noise_models.py:540-554)."""
    if qubit_type == "clock":
        return 2 * np.pi * (2 * K_quad * 1.0 * B_noise_gauss)
    if qubit_type == "stretched":
        return 2 * np.pi * (700e3 * B_noise_gauss)
    raise ValueError(f"Unknown qubit_type: {qubit_type}")


def intermediate_state_scattering_rate(Omega_1, Delta_e, Gamma_e):
    """Γ_e (Ω1/2)² / (Δe² + (Γ_e/2)²)  (RG/noise_models.py:614-617)."""
    return Gamma_e * (Omega_1 / 2) ** 2 / (Delta_e ** 2 + (Gamma_e / 2) ** 2)


def dark_state_suppression_factor(Delta_e, delta, Omega_1, Omega_2):
    Omega_eff = Omega_1 * Omega_2 / (2 * np.abs(Delta_e))
    if np.sign(Delta_e) == np.sign(delta) and np.abs(delta) > 1e-6:
        ratio = np.abs(delta) / (np.abs(Omega_eff) + 1e-10)
        s = 0.42 if 0.1 < ratio < 2.0 else 0.42 + 0.58 * (1 - np.exp(-np.abs(ratio - 0.377) / 0.5))
    else:
        s = 1.0
    return min(s, 1.5)


def enhanced_scattering_rate(Omega_1, Omega_2, Delta_e, delta, Gamma_e, use_dark_state=True):
    base = intermediate_state_scattering_rate(Omega_1, Delta_e, Gamma_e)
    return base * dark_state_suppression_factor(Delta_e, delta, Omega_1, Omega_2) if use_dark_state else base


def leakage_rate_to_adjacent_states(Omega, Delta_leak, pulse_shape: str = "square", tau=1e-6,
                                    gamma_rydberg=7143.0):
    """(Ω/Δ_leak)² γ_r S(Δ_leak τ/2π) with a per-shape spectral factor (RG/noise_models.py:807-849)."""
    Omega = np.asarray(Omega, dtype=float)
    tau = np.asarray(tau, dtype=float)
    x = Delta_leak * tau / (2 * np.pi)
    with np.errstate(divide="ignore", invalid="ignore"):
        if pulse_shape == "square":
            S = (np.sin(np.pi * x) / (np.pi * x)) ** 2
        elif pulse_shape == "gaussian":
            S = np.exp(-(Delta_leak * tau / 8) ** 2)
        elif pulse_shape == "cosine":
            S = np.where(np.abs(np.abs(x) - 0.5) < 1e-10, 0.25,
                         (np.sin(np.pi * x) / (np.pi * x * (1 - x ** 2))) ** 2)
        elif pulse_shape == "blackman":
            S = np.exp(-3 * np.abs(x)) * 0.1
        elif pulse_shape == "drag":
            S = np.exp(-(Delta_leak * tau / 8) ** 2) * 0.1
        else:
            S = (np.sin(np.pi * x) / (np.pi * x + 1e-10)) ** 2
        S = np.where(np.abs(x) < 1e-10, 1.0, S)
        g = (Omega / Delta_leak) ** 2 * gamma_rydberg * np.clip(S, 0, 1)
    g = np.where((np.abs(Delta_leak) < 1e-6) | (np.abs(Omega) < 1e-6), 0.0, g)
    return float(g) if g.ndim == 0 else g


def mJ_mixing_rate(Omega_eff, polarization_purity, Delta_zeeman):
    eps = 1.0 - polarization_purity
    if np.abs(Delta_zeeman) < 1e-10:
        return eps ** 2 * np.abs(Omega_eff)
    return eps ** 2 * Omega_eff ** 2 / np.abs(Delta_zeeman)


def rydberg_zeeman_splitting(B_field, L: int = 0, J: float = 0.5):
    S = 0.5
    g_J = 0 if J == 0 else 1 + (J * (J + 1) + S * (S + 1) - L * (L + 1)) / (2 * J * (J + 1))
    if L == 0:
        g_J += 0.002
    return g_J * MU_B * B_field / HBAR


def compute_noise_rates(species="Rb87", n_rydberg=70, rydberg_lifetime=140e-6, rydberg_lifetime_0K=280e-6,
                        Omega_eff=2 * np.pi * 5e6, Omega_1=None, Omega_2=None, Delta_e=2 * np.pi * 1e9,
                        delta=0, Gamma_e=2 * np.pi * 6.065e6, laser_linewidth_hz=1e3, temperature=20e-6,
                        omega_trap=2 * np.pi * 100e3, V=2 * np.pi * 200e6, R=3e-6, B_field=0.5e-4,
                        B_noise_gauss=0.001, qubit_type="clock", polarization_purity=0.99, U0=None,
                        alpha_ratio=300, mass=1.44e-25, waist=1e-6, gate_time=1e-6, rydberg_fraction=0.3,
                        pulse_shape="square", use_dark_state=True) -> NoiseRates:
    """Stand-alone rate estimate (RG/noise_models.py:1077-1150); not used by simulate_CZ_gate."""
    if Omega_1 is None:
        Omega_1 = np.sqrt(2 * np.abs(Delta_e) * np.abs(Omega_eff))
    if Omega_2 is None:
        Omega_2 = Omega_1
    sigma_r = np.sqrt(KB * temperature / (mass * omega_trap ** 2))
    dV = 6 * (np.sqrt(2) * sigma_r) / R
    vo = np.abs(V) / np.abs(Omega_eff)
    infid = dV ** 2 * vo ** 2 if vo < 3 else (dV ** 2 * (Omega_eff / V) ** 2 if vo > 10 else dV ** 2)
    g_anti = 0.0
    if U0 is not None and U0 > 0:
        from .trap_physics import effective_loss_rate
        g_anti = effective_loss_rate(gate_time, U0, alpha_ratio, mass, waist, temperature, rydberg_fraction)
    return NoiseRates(
        gamma_r=rydberg_decay_rate(rydberg_lifetime_0K),
        gamma_bbr=bbr_decay_rate(rydberg_lifetime_0K, rydberg_lifetime),
        gamma_phi_laser=laser_dephasing_rate(laser_linewidth_hz),
        gamma_phi_thermal=infid * np.abs(Omega_eff) / (2 * np.pi),
        gamma_phi_zeeman=zeeman_dephasing_rate(B_noise_gauss, qubit_type),
        gamma_loss_antitrap=g_anti, gamma_loss_bg=2 * np.pi * 0.1,
        gamma_scatter=enhanced_scattering_rate(Omega_1, Omega_2, Delta_e, delta, Gamma_e, use_dark_state),
        gamma_leakage=leakage_rate_to_adjacent_states(Omega_eff, 2 * np.pi * 50e6, pulse_shape, gate_time),
        gamma_mJ=mJ_mixing_rate(Omega_eff, polarization_purity, rydberg_zeeman_splitting(B_field, 0, 0.5)))


# ------------------------------------------------------------------ collapse channels
def collapse_channels(dim: int, gamma_r=0.0, gamma_bbr=0.0, gamma_phi_total=0.0, gamma_loss_antitrap=0.0,
                      gamma_loss_background=0.0, gamma_scatter_intermediate=0.0, gamma_leakage=0.0,
                      branching_1=0.5, mJ_leakage_rate=0.0) -> List[Tuple[str, float, Tuple[int, int]]]:
    """The reference's operator list as (name, rate, (to, from) single-atom levels), one entry per
    operator *pair* (atom A and atom B), in the order and with the ``> 0`` guards of
    build_all_noise_operators (RG/noise_models.py:1573-1592).  len()*2 == n_collapse_ops."""
    ryd = [2] if dim == 3 else [2, 3]
    ch: List[Tuple[str, float, Tuple[int, int]]] = []
    if gamma_r > 0:
        for r in ryd:
            ch.append(("decay_r_to_1", gamma_r * branching_1, (1, r)))
            ch.append(("decay_r_to_0", gamma_r * (1 - branching_1), (0, r)))
    if gamma_bbr > 0:
        ch += [("bbr", gamma_bbr, (0, r)) for r in ryd]
    if dim == 4 and mJ_leakage_rate > 0:
        ch += [("mJ_mixing", mJ_leakage_rate, (3, 2)), ("mJ_mixing", mJ_leakage_rate, (2, 3))]
    if gamma_phi_total > 0:
        ch += [("dephasing", gamma_phi_total, (r, r)) for r in ryd]
    for name, g in (("antitrap_loss", gamma_loss_antitrap), ("background_loss", gamma_loss_background)):
        if g > 0:
            ch += [(name, g, (0, r)) for r in ryd]
    if gamma_scatter_intermediate > 0:
        ch.append(("scatter", gamma_scatter_intermediate, (1, 1)))
    if gamma_leakage > 0:
        ch += [("leakage", gamma_leakage, (0, r)) for r in ryd]
    return ch


def _two(a, b):
    return np.kron(a, b)


def _sigma(dim, to, frm):
    m = np.zeros((dim, dim), dtype=complex)
    m[to, frm] = 1.0
    return m


def build_collapse_matrices(dim: int, channels) -> List[np.ndarray]:
    """NumPy (d²×d²) collapse operators, atom A then atom B per channel, like the reference's list."""
    I = np.eye(dim, dtype=complex)
    ops = []
    for _, g, (to, frm) in channels:
        s = np.sqrt(g) * _sigma(dim, to, frm)
        ops += [_two(s, I), _two(I, s)]
    return ops


def _rydberg_levels(hs) -> List[str]:
    return ["r"] if hs.dim == 3 else ["r+", "r-"]


def _on_each_atom(rate: float, single_atom_op, hs) -> List[np.ndarray]:
    """sqrt(rate) * (op x 1) and sqrt(rate) * (1 x op): atom A first, as the reference orders its lists."""
    amp = np.sqrt(rate)
    return [amp * np.kron(single_atom_op, hs.identity), amp * np.kron(hs.identity, single_atom_op)]


def build_decay_operators(gamma_optical: float, hs, gamma_bbr: float = 0, branching_1: float = 0.5,
                          leakage_rate: float = 0) -> List[np.ndarray]:
    """|r> -> |1> (branching_1) and |r> -> |0> radiative decay, black-body decay to |0>, and for the 4-level model
    the r+ <-> r- mixing (RG/noise_models.py:1199-1297).  NumPy matrices in the reference's order."""
    ops: List[np.ndarray] = []
    levels = _rydberg_levels(hs)
    if gamma_optical > 0:
        for r in levels:
            ops += _on_each_atom(gamma_optical * branching_1, hs.transitions[f"{r}->1"], hs)
            ops += _on_each_atom(gamma_optical * (1 - branching_1), hs.transitions[f"{r}->0"], hs)
    if gamma_bbr > 0:
        for r in levels:
            ops += _on_each_atom(gamma_bbr, hs.transitions[f"{r}->0"], hs)
    if hs.dim == 4 and leakage_rate > 0:
        for key in ("r+->r-", "r-->r+"):
            ops += _on_each_atom(leakage_rate, hs.transitions[key], hs)
    return ops


def build_dephasing_operators(gamma_phi: float, hs, gamma_phi_minus: float = None) -> List[np.ndarray]:
    """sqrt(gamma) |r><r| on each atom (r+ with gamma_phi, r- with gamma_phi_minus for the 4-level model;
    RG/noise_models.py:1300-1356)."""
    if gamma_phi <= 0:
        return []
    rates = [gamma_phi] if hs.dim == 3 else [gamma_phi, gamma_phi if gamma_phi_minus is None else gamma_phi_minus]
    ops: List[np.ndarray] = []
    for r, g in zip(_rydberg_levels(hs), rates):
        ops += _on_each_atom(g, hs.projectors[r], hs)
    return ops


def build_loss_operators(gamma_loss: float, hs, loss_source: str = "rydberg") -> List[np.ndarray]:
    """Loss channels land in |0>: from the Rydberg level(s) or from |1> (RG/noise_models.py:1359-1412)."""
    if gamma_loss <= 0:
        return []
    ops: List[np.ndarray] = []
    if loss_source == "rydberg":
        for r in _rydberg_levels(hs):
            ops += _on_each_atom(gamma_loss, hs.transitions[f"{r}->0"], hs)
    elif loss_source == "qubit":
        sigma = hs.transitions.get("1->0")
        if sigma is None:
            sigma = hs.basis["0"] @ hs.basis["1"].conj().T
        ops += _on_each_atom(gamma_loss, sigma, hs)
    return ops


def build_scatter_operators(gamma_scatter: float, hs) -> List[np.ndarray]:
    """Intermediate-state scattering as dephasing of |1>: sqrt(gamma) |1><1| on each atom
    (RG/noise_models.py:1415-1446)."""
    if gamma_scatter <= 0:
        return []
    return _on_each_atom(gamma_scatter, hs.projectors["1"], hs)


def build_all_noise_operators(params: dict = None, hs=None, dim: int = 3, gamma_r=None, gamma_bbr=None,
                              gamma_phi_laser=None, gamma_phi_thermal=None, gamma_phi_zeeman=None,
                              gamma_loss_antitrap=None, gamma_loss_background=None,
                              gamma_scatter_intermediate=None, gamma_leakage=None, branching_1: float = 0.5,
                              mJ_leakage_rate: float = 0):
    """Same call signature and ``noise_breakdown`` dictionary as RG/noise_models.py:1449-1620.
    Returns NumPy matrices instead of QuTiP Qobj."""
    if hs is not None:
        dim = hs.dim
    if params is not None:
        gamma_r = params.get("gamma_r", params.get("gamma_optical", 0))
        if gamma_r == 0 and params.get("T1", 0) > 0:
            gamma_r = 1.0 / params["T1"]
        gamma_bbr = params.get("gamma_bbr", 0)
        gamma_phi_laser = params.get("gamma_phi_laser", 0)
        gamma_phi_thermal = params.get("gamma_phi_thermal", 0)
        gamma_phi_zeeman = params.get("gamma_phi_zeeman", 0)
        if gamma_phi_laser == 0 and gamma_phi_thermal == 0 and gamma_phi_zeeman == 0:
            g = params.get("gamma_dephasing", 0)
            if g == 0 and params.get("T2_star", 0) > 0:
                g = 1.0 / params["T2_star"]
            gamma_phi_laser = g
        gamma_loss_antitrap = params.get("gamma_loss_antitrap", 0)
        gamma_loss_background = params.get("gamma_loss_background", 0)
        gamma_scatter_intermediate = params.get("gamma_scatter_intermediate", 0)
        gamma_leakage = params.get("gamma_leakage", 0)
        branching_1 = params.get("branching_1", 0.5)
        mJ_leakage_rate = params.get("mJ_leakage_rate", 0)
    (gamma_r, gamma_bbr, gamma_phi_laser, gamma_phi_thermal, gamma_phi_zeeman, gamma_loss_antitrap,
     gamma_loss_background, gamma_scatter_intermediate, gamma_leakage) = (
        v or 0 for v in (gamma_r, gamma_bbr, gamma_phi_laser, gamma_phi_thermal, gamma_phi_zeeman,
                         gamma_loss_antitrap, gamma_loss_background, gamma_scatter_intermediate,
                         gamma_leakage))
    gamma_phi_total = gamma_phi_laser + gamma_phi_thermal + gamma_phi_zeeman
    channels = collapse_channels(dim, gamma_r, gamma_bbr, gamma_phi_total, gamma_loss_antitrap,
                                 gamma_loss_background, gamma_scatter_intermediate, gamma_leakage,
                                 branching_1, mJ_leakage_rate)
    c_ops = build_collapse_matrices(dim, channels)
    breakdown = {
        "gamma_r": gamma_r, "gamma_bbr": gamma_bbr, "gamma_phi_laser": gamma_phi_laser,
        "gamma_phi_thermal": gamma_phi_thermal, "gamma_phi_zeeman": gamma_phi_zeeman,
        "gamma_loss_antitrap": gamma_loss_antitrap, "gamma_loss_background": gamma_loss_background,
        "gamma_scatter_intermediate": gamma_scatter_intermediate, "gamma_leakage": gamma_leakage,
        "mJ_leakage_rate": mJ_leakage_rate, "branching_1": branching_1,
        "gamma_phi_total": gamma_phi_total, "total_decay_rate": gamma_r + gamma_bbr,
        "total_dephasing_rate": gamma_phi_total,
        "total_loss_rate": gamma_loss_antitrap + gamma_loss_background + gamma_leakage,
        "dim": dim, "n_collapse_ops": len(c_ops),
    }
    return c_ops, breakdown


def kernel_rates(breakdown: dict) -> Tuple[float, float, float, float, float]:
    """(Γ_r→1, Γ_r→0, γ_φ, γ_sc, γ_mJ) from a ``noise_breakdown`` dictionary.  Elementwise, so the
    dictionary may hold arrays (batch prologue)."""
    b = breakdown.get("branching_1", 0.5)
    g_r = breakdown["gamma_r"]
    return (g_r * b,
            g_r * (1 - b) + breakdown["gamma_bbr"] + breakdown["gamma_loss_antitrap"]
            + breakdown["gamma_loss_background"] + breakdown["gamma_leakage"],
            breakdown["gamma_phi_total"], breakdown["gamma_scatter_intermediate"],
            breakdown["mJ_leakage_rate"])


__all__ = ["NoiseRates", "rydberg_decay_rate", "bbr_decay_rate", "laser_dephasing_rate",
           "zeeman_dephasing_rate", "intermediate_state_scattering_rate", "dark_state_suppression_factor",
           "enhanced_scattering_rate", "leakage_rate_to_adjacent_states", "mJ_mixing_rate",
           "rydberg_zeeman_splitting", "compute_noise_rates", "collapse_channels",
           "build_collapse_matrices", "build_decay_operators", "build_dephasing_operators", "build_loss_operators",
           "build_scatter_operators", "build_all_noise_operators", "kernel_rates"]
