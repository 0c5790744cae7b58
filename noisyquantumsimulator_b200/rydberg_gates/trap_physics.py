"""Tweezer physics -> noise rates and coherent shifts (reference: RG/trap_physics.py).

Written branch-free (``np.where``) so the same functions serve one point (Python floats in,
floats out) and the struct-of-arrays batch prologue (``prologue.py``).  Quirks of the reference
are kept and marked ``# quirk``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .constants import C, EPS0, HBAR, KB, MU_B

_AU_REF = 4 * np.pi * EPS0 * (0.529e-10) ** 3   # the rounded Bohr radius RG/trap_physics.py:1843 uses


def _out(x):
    """0-d results become Python floats, arrays stay arrays."""
    return float(x) if np.ndim(x) == 0 else x


# ------------------------------------------------------------------ polarizabilities
def get_polarizability_at_wavelength(species: str, state: str, wavelength_nm, n_rydberg=70,
                                     L_rydberg: int = 0, F: int = None):
    """RG/trap_physics.py:131-207."""
    from .atom_database import ATOM_DB

    atom = ATOM_DB[species]
    wavelength_m = np.asarray(wavelength_nm, dtype=float) * 1e-9
    omega = 2 * np.pi * C / wavelength_m
    st = state.lower()
    if st in ("ground", "5s", "6s"):
        a0 = atom["alpha_ground"]
        omega_D1 = 2 * np.pi * (377.107e12 if species == "Rb87" else 335.116e12)
        red = omega < omega_D1
        with np.errstate(divide="ignore", invalid="ignore"):
            corr = 1.0 + 0.3 * omega ** 2 / (omega_D1 ** 2 - omega ** 2)
        alpha = np.where(red, a0 * corr, a0)
        if F is not None and F in atom.get("alpha_hyperfine", {}):
            alpha = np.full_like(alpha, atom["alpha_hyperfine"][F])
        return _out(alpha)
    if st == "rydberg" or state[0].isdigit():
        if state[0].isdigit():
            n_rydberg = int("".join(filter(str.isdigit, state)))
        label = {0: "S", 1: "P", 2: "D", 3: "F"}.get(L_rydberg, "S")
        dq = atom["quantum_defects"].get(label, 3.0)
        x = (n_rydberg - dq) / (atom["n_ref"] - dq)
        static = atom["alpha_rydberg_ref"] * x ** atom["scaling_exponents"]["polarizability"]
        return _out(static * (wavelength_m / 1064e-9) ** 2)
    raise ValueError(f"Unknown state: {state}. Use 'ground', 'rydberg', or 'nS/nP'")


def magic_trap_enhancement(species: str, wavelength_nm, n_rydberg=70):
    ag = get_polarizability_at_wavelength(species, "ground", wavelength_nm)
    ar = get_polarizability_at_wavelength(species, "rydberg", wavelength_nm, n_rydberg=n_rydberg)
    ratio = np.where(np.abs(ag) > 1e-50, ar / ag, 0.0)
    return _out(1.0 / (1.0 + np.abs(1.0 - ratio)))


# ------------------------------------------------------------------ geometry and trap
def tweezer_spacing(wavelength, NA, factor=1.0):
    """R = factor · λ/(2 NA)."""
    return factor * (wavelength / (2 * NA))


def diffraction_limited_spot(wavelength, NA):
    return 0.82 * wavelength / NA


def trap_depth(power, waist, alpha):
    """U0 = |α| I0 / (2 ε0 c), I0 = 2P/(π w²)."""
    I0 = 2 * power / (np.pi * waist ** 2)
    return np.abs(alpha) * I0 / (2 * EPS0 * C)


def trap_depth_from_temperature(temperature_K):
    return KB * temperature_K


def trap_temperature(U0):
    return U0 / KB


def trap_frequencies(U0, mass, waist, wavelength) -> Tuple[float, float]:
    z_R = np.pi * waist ** 2 / wavelength
    return np.sqrt(4 * U0 / (mass * waist ** 2)), np.sqrt(2 * U0 / (mass * z_R ** 2))


def trap_frequency_from_depth_and_waist(U0, mass, waist):
    return np.sqrt(4 * U0 / (mass * waist ** 2))


def position_uncertainty(temperature, mass, omega):
    return np.sqrt(KB * temperature / (mass * omega ** 2))


def quantum_ground_state_size(mass, omega):
    return np.sqrt(HBAR / (2 * mass * omega))


def thermal_de_broglie_wavelength(temperature, mass):
    return 2 * np.pi * HBAR / np.sqrt(2 * np.pi * mass * KB * temperature)


def anti_trap_potential(U0_ground, alpha_ratio):
    return U0_ground * alpha_ratio


def anti_trap_frequency(U0_ground, alpha_ratio, mass, waist):
    return np.sqrt(4 * (alpha_ratio * U0_ground) / (mass * waist ** 2))


def thermal_velocity(temperature, mass):
    return np.sqrt(KB * temperature / mass)


# ------------------------------------------------------------------ anti-trapping loss
def atom_loss_probability(gate_time, U0, alpha_ratio, mass, waist, temperature, rydberg_fraction=0.3,
                          trap_on_during_rydberg=True):
    """Ballistic / anti-trapped expansion out of a 2w capture range (RG/trap_physics.py:944-995)."""
    t_ryd = rydberg_fraction * np.asarray(gate_time, dtype=float)
    omega_trap = np.sqrt(4 * U0 / (mass * waist ** 2))
    v_th = np.sqrt(KB * temperature / mass)
    capture = 2.0 * waist
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if trap_on_during_rydberg:
            omega_anti = np.sqrt(4 * alpha_ratio * U0 / (mass * waist ** 2))
            sigma0 = np.sqrt(KB * temperature / (mass * omega_trap ** 2))
            s_pos = sigma0 * np.cosh(omega_anti * t_ryd)
            s_vel = (v_th / omega_anti) * np.sinh(omega_anti * t_ryd)
            sigma = np.sqrt(s_pos ** 2 + s_vel ** 2)
            P = np.where(sigma > 0, 1.0 - np.exp(-(capture / sigma) ** 2 / 2), 0.0)
            P = np.where((omega_anti > 0) & (t_ryd > 0), P, 0.0)
        else:
            drift = v_th * t_ryd
            P = np.where(drift > 0, 1.0 - np.exp(-(drift / capture) ** 2 / 2), 0.0)
    return _out(np.clip(P, 0.0, 1.0))


def effective_loss_rate(gate_time, U0, alpha_ratio, mass, waist, temperature, rydberg_fraction=0.3):
    """γ = -ln(1-P)/t, 5/t when P >= 0.99, capped at 1/t (RG/trap_physics.py:1040-1057)."""
    gate_time = np.asarray(gate_time, dtype=float)
    P = np.asarray(atom_loss_probability(gate_time, U0, alpha_ratio, mass, waist, temperature,
                                         rydberg_fraction))
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(P >= 0.99, 5.0 / gate_time,
                     np.where(P > 0, -np.log(1 - np.minimum(P, 0.99)) / gate_time, 0.0))
        cap = np.where(gate_time > 0, 1.0 / gate_time, 1e6)
    return _out(np.minimum(g, cap))


# ------------------------------------------------------------------ dephasing rates
def blockade_fluctuation(R0, sigma_r):
    """δV/V = 6 (√2 σ_r)/R0."""
    return 6 * (np.sqrt(2) * sigma_r) / R0


def thermal_dephasing_rate(delta_V_over_V, V0, Omega=None):
    """Piecewise in V/Ω with a smoothstep between 3 and 10, capped at 10 MHz
    (RG/trap_physics.py:1169-1203)."""
    if Omega is None:
        Omega = 2 * np.pi * 5e6
    Omega = np.where(np.asarray(Omega) <= 0, 2 * np.pi * 5e6, Omega)
    aO, aV = np.abs(Omega), np.abs(V0)
    r = aV / aO
    dv2 = np.asarray(delta_V_over_V) ** 2
    g_weak = (dv2 * r ** 2) * aO / (2 * np.pi)
    g_strong = (dv2 * (aO / aV) ** 2) * aO / (2 * np.pi)
    x = np.clip((r - 3) / 7, 0, 1)
    smooth = 3 * x ** 2 - 2 * x ** 3
    g_mid = g_weak * (1 - smooth) + g_strong * smooth
    g = np.where(r < 3, g_weak, np.where(r > 10, g_strong, g_mid))
    return _out(np.minimum(g, 10e6))


def doppler_dephasing_rate(temperature, mass, k_eff, gate_time):
    """γ = (k_eff v_rms)² τ  (RG/trap_physics.py:1278-1288)."""
    v_rms = np.sqrt(KB * temperature / mass)
    return (k_eff * v_rms) ** 2 * gate_time


def intensity_noise_dephasing_rate(trap_depth_J, intensity_noise_frac, gate_time=200e-9,
                                   differential_stark_fraction=0.01):
    return (trap_depth_J / HBAR) * intensity_noise_frac * differential_stark_fraction


def thermal_infidelity_estimate(R0, sigma_r, V0, Omega, gate_time):
    return (blockade_fluctuation(R0, sigma_r) * V0 * gate_time) ** 2


# ------------------------------------------------------------------ recoil / heating (not on the hot path)
def photon_recoil_energy(mass, wavelength):
    return (HBAR * 2 * np.pi / wavelength) ** 2 / (2 * mass)


def recoil_temperature(mass, wavelength):
    return photon_recoil_energy(mass, wavelength) / KB


def trap_heating_rate_intensity_noise(omega_trap, relative_intensity_noise=1e-4, noise_bandwidth=1e6):
    return (omega_trap / 4) * relative_intensity_noise ** 2 * noise_bandwidth


def trap_heating_rate_pointing_noise(omega_trap, waist, pointing_noise_rad=1e-6):
    return omega_trap * (pointing_noise_rad * 1.0 / waist) ** 2


def compute_trap_properties(power, waist, wavelength, alpha, mass, temperature) -> dict:
    U0 = trap_depth(power, waist, alpha)
    w_r, w_z = trap_frequencies(U0, mass, waist, wavelength)
    s_r, s_z = position_uncertainty(temperature, mass, w_r), position_uncertainty(temperature, mass, w_z)
    return {"U0": U0, "omega_r": w_r, "omega_z": w_z, "sigma_r": s_r, "sigma_z": s_z,
            "v_thermal": thermal_velocity(temperature, mass),
            "sigma_quantum": quantum_ground_state_size(mass, w_r), "U0_mK": U0 / KB * 1e3,
            "omega_r_kHz": w_r / (2 * np.pi) / 1e3, "omega_z_kHz": w_z / (2 * np.pi) / 1e3,
            "sigma_r_nm": s_r * 1e9, "sigma_z_nm": s_z * 1e9}


# ------------------------------------------------------------------ the bundle simulate_CZ_gate uses
def compute_trap_dependent_noise(species, tweezer_power, tweezer_waist, temperature, spacing, gate_time,
                                 n_rydberg=70, gamma_phi_laser=1e4, Omega_1=0.0, Delta_e=2 * np.pi * 5e9,
                                 intermediate_state=None, Omega_eff=None, tweezer_wavelength_nm=1064.0,
                                 include_doppler=True, include_intensity_noise=True,
                                 intensity_noise_frac=0.01, rydberg_wavelength_1_nm=780.0,
                                 rydberg_wavelength_2_nm=480.0, counter_propagating=True) -> dict:
    """RG/trap_physics.py:1704-1848.  Species, n and the flags are per call; every other argument
    may be an array of points."""
    from .atom_database import ATOM_DB, get_atom_properties, get_C6, get_rydberg_lifetime
    from .laser_physics import rydberg_blockade
    from .noise_models import intermediate_state_scattering_rate

    atom = get_atom_properties(species)
    mass = atom["mass"]
    # quirk: depth and frequency always use the 1064 nm ground polarizability and trap wavelength,
    # even when tweezer_wavelength_nm is given (RG/trap_physics.py:1713-1715)
    U0 = trap_depth(tweezer_power, tweezer_waist, atom["alpha_ground"])
    omega_r, _ = trap_frequencies(U0, mass, tweezer_waist, atom["trap_wavelength"])
    sigma_r = position_uncertainty(temperature, mass, omega_r)
    V = rydberg_blockade(get_C6(n_rydberg, species), spacing)
    dV = blockade_fluctuation(spacing, sigma_r)
    g_thermal = thermal_dephasing_rate(dV, V, Omega_eff)

    if intermediate_state is None:
        intermediate_state = "5P3/2" if species == "Rb87" else "6P3/2"
    Gamma_e = ATOM_DB[species]["intermediate_states"][intermediate_state]["linewidth"]
    ok = (np.asarray(Omega_1) > 0) & (np.asarray(Delta_e) > 0)   # quirk: Δe < 0 switches scattering off
    g_scatter = _out(np.where(ok, intermediate_state_scattering_rate(Omega_1, Delta_e, Gamma_e), 0.0))

    a_g = get_polarizability_at_wavelength(species, "ground", tweezer_wavelength_nm)
    a_r = get_polarizability_at_wavelength(species, "rydberg", tweezer_wavelength_nm, n_rydberg=n_rydberg)
    alpha_ratio = _out(np.where(np.abs(a_g) > 1e-50, np.abs(np.asarray(a_r) / a_g), 0.0))
    g_anti = effective_loss_rate(gate_time, U0, alpha_ratio, mass, tweezer_waist, temperature)
    g_anti = _out(np.where((np.asarray(alpha_ratio) > 0) & (np.asarray(gate_time) > 0), g_anti, 0.0))

    I_center = np.where(np.abs(a_g) > 1e-50, 2 * EPS0 * C * np.abs(U0) / np.abs(a_g), 0.0)
    diff_shift = _out(np.abs(np.asarray(a_r) - a_g) * I_center / (2 * EPS0 * C * HBAR * 2 * np.pi))
    enhancement = magic_trap_enhancement(species, tweezer_wavelength_nm, n_rydberg)

    k_eff, g_doppler = 0.0, 0.0
    if include_doppler:
        k1 = 2 * np.pi / (rydberg_wavelength_1_nm * 1e-9)
        k2 = 2 * np.pi / (rydberg_wavelength_2_nm * 1e-9)
        k_eff_on = np.abs(k1 - k2) if counter_propagating else k1 + k2
        on = np.asarray(gate_time) > 0
        g_doppler = _out(np.where(on, doppler_dephasing_rate(temperature, mass, k_eff_on, gate_time), 0.0))
        k_eff = _out(np.where(on, k_eff_on, 0.0))

    g_intensity = 0.0
    if include_intensity_noise:
        frac = np.asarray(intensity_noise_frac, dtype=float)
        g_int = intensity_noise_dephasing_rate(U0, frac, gate_time, np.minimum(enhancement, 0.1))
        g_intensity = _out(np.where(frac > 0, g_int, 0.0))

    return {
        "trap_depth_uK": U0 / KB * 1e6, "trap_freq_radial_kHz": omega_r / (2 * np.pi) / 1e3,
        "position_uncertainty_nm": sigma_r * 1e9, "V_over_2pi_MHz": V / (2 * np.pi) / 1e6,
        "gamma_r": 1.0 / get_rydberg_lifetime(n_rydberg, species, temperature=300.0),
        "gamma_scatter_intermediate": g_scatter, "gamma_phi_laser": gamma_phi_laser,
        "gamma_phi_thermal": g_thermal, "gamma_phi_doppler": g_doppler,
        "gamma_phi_intensity": g_intensity, "gamma_loss_antitrap": g_anti,
        "gamma_loss_background": 1e3,
        "blockade_fluctuation_percent": dV * 100,
        "intermediate_linewidth_MHz": Gamma_e / (2 * np.pi) / 1e6,
        "k_eff_rad_per_m": k_eff, "v_thermal_m_per_s": np.sqrt(KB * temperature / mass),
        "alpha_ratio": alpha_ratio, "alpha_ground_au": a_g / _AU_REF, "alpha_rydberg_au": a_r / _AU_REF,
        "differential_shift_Hz": diff_shift, "magic_enhancement": enhancement,
        "wavelength_nm": tweezer_wavelength_nm,
    }


# ------------------------------------------------------------------ coherent shifts on |1>
def calculate_zeeman_shift(B_field, qubit_0, qubit_1, species):
    """Clock states: 2π K B²; otherwise linear (RG/trap_physics.py:1908-1965)."""
    (F0, mF0), (F1, mF1) = qubit_0, qubit_1
    if mF0 == 0 and mF1 == 0:
        K = {"Rb87": 575.0, "Cs133": 2000.0}.get(species, 1000.0)   # quirk: 2000 here, 427 in the dephasing rate
        return K * (B_field * 1e4) ** 2 * 2 * np.pi
    if species == "Cs133":
        g0, g1 = (-0.25 if F0 == 3 else 0.25), (-0.25 if F1 == 3 else 0.25)
    elif species == "Rb87":
        g0, g1 = (-0.5 if F0 == 1 else 0.5), (-0.5 if F1 == 1 else 0.5)
    else:
        g0, g1 = (0.5 if F0 == 2 else -0.5), (0.5 if F1 == 2 else -0.5)
    return (g1 * mF1 - g0 * mF0) * MU_B * B_field / HBAR


def calculate_stark_shift(tweezer_power, tweezer_waist, tweezer_wavelength, alpha_ground, alpha_excited):
    I_center = 2 * tweezer_power / (np.pi * tweezer_waist ** 2)
    return np.abs(alpha_excited - alpha_ground) * I_center / (4 * np.pi * EPS0 * C * HBAR) * 2 * np.pi


def calculate_qubit_stark_shift(tweezer_power, tweezer_waist, species="Rb87", trap_depth_mK=None):
    """70 kHz/mK (Rb) or 200 kHz/mK (Cs) of trap depth (RG/trap_physics.py:2112-2142)."""
    if trap_depth_mK is not None:
        hz = {"Rb87": 70e3, "Cs133": 200e3}.get(species, 70e3) * trap_depth_mK
    else:
        d_alpha = {"Rb87": 2.4, "Cs133": 7.0}.get(species, 2.4) * 1.6488e-41
        I_center = 2 * tweezer_power / (np.pi * tweezer_waist ** 2)
        hz = d_alpha * I_center / (4 * np.pi * EPS0 * C * HBAR)
    return hz * 2 * np.pi


__all__ = ["tweezer_spacing", "diffraction_limited_spot", "trap_depth", "trap_depth_from_temperature",
           "trap_temperature", "trap_frequencies", "trap_frequency_from_depth_and_waist",
           "position_uncertainty", "quantum_ground_state_size", "thermal_de_broglie_wavelength",
           "thermal_velocity", "anti_trap_potential", "anti_trap_frequency", "atom_loss_probability",
           "effective_loss_rate", "blockade_fluctuation", "thermal_dephasing_rate",
           "thermal_infidelity_estimate", "photon_recoil_energy", "recoil_temperature",
           "trap_heating_rate_intensity_noise", "trap_heating_rate_pointing_noise",
           "compute_trap_properties", "compute_trap_dependent_noise", "calculate_zeeman_shift",
           "calculate_stark_shift", "calculate_qubit_stark_shift", "get_polarizability_at_wavelength",
           "magic_trap_enhancement", "doppler_dephasing_rate", "intensity_noise_dephasing_rate"]
