"""Pulse envelopes and leakage helpers (reference: RG/pulse_shaping.py)."""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np

from .constants import HBAR, MU_B, RY_JOULES


def pulse_envelope_square(t, tau, **kwargs):
    return np.ones_like(t, dtype=float)


def _peak_normalised(env):
    # quirk: normalising by env.max() means a *scalar* t always returns 1.0, so gaussian/blackman
    # pulses degenerate to square amplitude inside evolve_shaped_pulse (RG/pulse_shaping.py:184-186)
    m = env.max()
    return env / m if m > 0 else env


def pulse_envelope_gaussian(t, tau, sigma_factor: float = 3.0, **kwargs):
    sigma = tau / sigma_factor
    return _peak_normalised(np.exp(-(t - tau / 2) ** 2 / (2 * sigma ** 2)))


def pulse_envelope_cosine(t, tau, **kwargs):
    return np.sin(np.pi * t / tau) ** 2


def pulse_envelope_blackman(t, tau, **kwargs):
    return _peak_normalised(0.42 - 0.5 * np.cos(2 * np.pi * t / tau) + 0.08 * np.cos(4 * np.pi * t / tau))


def pulse_envelope_drag(t, tau, Delta_leak, lambda_drag: float = 1.0, base_shape: str = "gaussian",
                        sigma_factor: float = 4.0, **kwargs):
    if base_shape == "gaussian":
        base = pulse_envelope_gaussian(t, tau, sigma_factor=sigma_factor)
    elif base_shape == "cosine":
        base = pulse_envelope_cosine(t, tau)
    elif base_shape == "blackman":
        base = pulse_envelope_blackman(t, tau)
    else:
        raise ValueError(f"Unknown base_shape: {base_shape}. Use 'gaussian', 'cosine', or 'blackman'.")
    dt = t[1] - t[0] if len(t) > 1 else tau / 100
    return base + 1j * (lambda_drag / np.abs(Delta_leak) * np.gradient(base, dt))


PULSE_SHAPES: Dict[str, Callable] = {"square": pulse_envelope_square, "gaussian": pulse_envelope_gaussian,
                                     "cosine": pulse_envelope_cosine, "blackman": pulse_envelope_blackman,
                                     "drag": pulse_envelope_drag}


def get_pulse_envelope(shape: str, t, tau, **kwargs):
    key = shape.lower()
    if key not in PULSE_SHAPES:
        raise ValueError(f"Unknown pulse shape: {shape}. Available shapes: {list(PULSE_SHAPES.keys())}")
    return PULSE_SHAPES[key](t, tau, **kwargs)


def list_available_shapes() -> list:
    return list(PULSE_SHAPES.keys())


def spectral_leakage_factor(pulse_shape: str, tau: float, Delta_leak: float) -> float:
    x = Delta_leak * tau / (2 * np.pi)
    if abs(x) < 1e-10:
        return 1.0
    sinc2 = (np.sin(np.pi * x) / (np.pi * x)) ** 2
    if pulse_shape == "gaussian":
        S = np.exp(-(Delta_leak * tau / 3.0) ** 2)
    elif pulse_shape == "cosine":
        S = 0.25 if (abs(x - 0.5) < 1e-10 or abs(x + 0.5) < 1e-10) else (np.sqrt(sinc2) / (1 - x ** 2)) ** 2
    elif pulse_shape == "blackman":
        S = min(np.exp(-3 * abs(x)) * sinc2, sinc2 * 0.1)
    elif pulse_shape == "drag":
        S = np.exp(-(Delta_leak * tau / 12) ** 2) * 0.1
    else:
        S = sinc2
    return float(np.clip(S, 0, 1))


def compute_leakage_detuning(species: str, n_rydberg: int, L: int = 0, leakage_target: str = "fine_structure",
                             atom_database: dict = None) -> float:
    """simulate_CZ_gate only uses the default: 2π·50 MHz (RG/pulse_shaping.py:633-638)."""
    dq_S = 3.13
    if atom_database is not None and species in atom_database:
        dq_S = atom_database[species].get("quantum_defects", {}).get("S", dq_S)
    if leakage_target == "fine_structure":
        return 2 * np.pi * 50e6
    if leakage_target == "adjacent_n":
        return 2 * RY_JOULES / HBAR / (n_rydberg - dq_S) ** 3
    if leakage_target == "zeeman":
        return 2.002 * MU_B * 1e-4 / HBAR
    raise ValueError(f"Unknown leakage_target: {leakage_target}. Use 'fine_structure', 'adjacent_n', or 'zeeman'.")


def leakage_rate_to_adjacent_states(Omega: float, Delta_leak: float, pulse_shape: str, tau: float) -> float:
    S = spectral_leakage_factor(pulse_shape, tau, Delta_leak)
    return 0.5 * Delta_leak * (Omega / Delta_leak) ** 2 * S


def compute_pulse_area(envelope, t) -> float:
    trap = getattr(np, "trapezoid", None) or np.trapz
    return trap(np.abs(envelope), t)


def normalize_pulse_area(envelope, t, target_area: float):
    area = compute_pulse_area(envelope, t)
    return envelope if area < 1e-15 else envelope * (target_area / area)


def area_correction_factor(pulse_shape: str, tau: float = 1.0) -> float:
    """square area / shaped area on a 1000-point grid (RG/pulse_shaping.py:821-842)."""
    if pulse_shape == "square":
        return 1.0
    t = np.linspace(0, tau, 1000)
    kwargs = {"Delta_leak": 2 * np.pi * 50e6} if pulse_shape == "drag" else {}
    area = compute_pulse_area(get_pulse_envelope(pulse_shape, t, tau, **kwargs), t)
    return 1.0 if area < 1e-15 else tau / area


def prepare_pulse_for_evolution(pulse_shape: str, t_pulse, tau: float, preserve_area: bool = True,
                                **kwargs) -> Tuple[np.ndarray, dict]:
    env = get_pulse_envelope(pulse_shape, t_pulse, tau, **kwargs)
    trap = getattr(np, "trapezoid", None) or np.trapz
    area_factor = (trap(np.abs(env), t_pulse) / tau if tau > 0 else 1.0) if preserve_area else 1.0
    normed = env / area_factor
    safe = np.abs(normed) + 1e-6
    a = trap(safe, t_pulse)
    if a > 0:
        safe = safe * tau / a
    info = {"pulse_shape": pulse_shape, "area_factor": area_factor,
            "peak_scaling": np.max(np.abs(normed)), "envelope_floor": 1e-6,
            "is_complex": np.any(np.iscomplex(env))}
    return np.real(safe), info


__all__ = ["pulse_envelope_square", "pulse_envelope_gaussian", "pulse_envelope_cosine",
           "pulse_envelope_blackman", "pulse_envelope_drag", "PULSE_SHAPES", "get_pulse_envelope",
           "list_available_shapes", "spectral_leakage_factor", "compute_leakage_detuning",
           "leakage_rate_to_adjacent_states", "compute_pulse_area", "normalize_pulse_area",
           "area_correction_factor", "prepare_pulse_for_evolution"]
