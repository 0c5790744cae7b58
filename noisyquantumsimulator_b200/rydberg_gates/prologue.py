"""Struct-of-arrays scalar prologue of simulate_CZ_gate (reference: RG/simulation.py:2796-3355).

Everything the reference computes per gate evaluation *before* it touches QuTiP -- Rabi
frequencies, blockade, protocol parameters, trap-dependent noise rates, coherent shifts, the LP
phase jump, the piecewise-constant pulse schedule -- for B parameter points at once.  The
reference does this in ~100 µs of Python per point, which would cap a GPU that evaluates
millions of points per second at ~10^4/s; here it is ~50 NumPy array operations per batch.
``simulate_CZ_gate`` (one point) runs through the same code with B = 1.

The result is a ``GatePlan``: the kernel inputs (point rows, segment rows) plus every scalar the
reference reports in ``SimulationResult``.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from .atom_database import get_atom_properties, get_C6, get_default_intermediate_state
from .configurations import AtomicConfiguration
from .constants import KB
from .laser_physics import laser_E0, rydberg_blockade, single_photon_rabi, two_photon_rabi
from .noise_models import (leakage_rate_to_adjacent_states, mJ_mixing_rate, rydberg_zeeman_splitting,
                           zeeman_dephasing_rate)
from .protocols import (JP_DEFAULT, SMOOTH_JP_PARAMS, compute_phase_shift_xi, lp_table_lookup)
from .pulse_shaping import area_correction_factor, compute_leakage_detuning, get_pulse_envelope
from .trap_physics import (calculate_qubit_stark_shift, calculate_zeeman_shift, compute_trap_dependent_noise,
                           get_polarizability_at_wavelength, tweezer_spacing)

POINT_STRIDE, POINT_STRIDE5, SEG_STRIDE = 12, 20, 4
PROTOCOL_NAMES = {"lp": "levine_pichler", "jp": "jandura_pupillo", "sjp": "smooth_jp"}


@dataclass
class GatePlan:
    """Kernel inputs and reported scalars for B points (arrays of length B unless noted)."""

    protocol: str                      # "lp" | "jp" | "sjp"
    dim: int
    B: int
    include_noise: bool
    points: np.ndarray                 # [B, 12]  C-ABI point rows ([B, 20] for dim 5)
    segs: np.ndarray                   # [B, S, 4]
    n_seg: np.ndarray                  # [B] int32
    scalars: Dict[str, np.ndarray] = field(default_factory=dict)
    noise_breakdown: Dict[str, object] = field(default_factory=dict)
    trap_noise: Dict[str, object] = field(default_factory=dict)
    pulse_info: Dict[str, object] = field(default_factory=dict)
    config: AtomicConfiguration = None
    n_pulses: int = 1
    pulse_shape: str = "square"
    trap_laser_on: bool = True


def _arr(x, B):
    return np.broadcast_to(np.asarray(x, dtype=float), (B,)).copy()


def _opt(x, default, B):
    """None -> default (the reference's ``x if x is not None else default``)."""
    return _arr(default if x is None else x, B)


def _opt_or(x, default, B):
    """None or 0.0 -> default (the reference's ``getattr(...) or default`` quirk,
    RG/simulation.py:3462-3464)."""
    a = _arr(default if x is None else x, B)
    return np.where(a == 0.0, default, a)


def build_plan(protocol: str, *, B: int,
               # excitation (TwoPhotonExcitationConfig)
               laser_1_power, laser_1_waist, laser_2_power, laser_2_waist, Delta_e,
               laser_1_linewidth_hz=100.0, laser_2_linewidth_hz=100.0,
               polarization_purity_1=0.99, polarization_purity_2=0.99,
               polarization_1="pi", polarization_2="sigma+", counter_propagating=True,
               # noise switches (NoiseSourceConfig)
               include_motional_dephasing=True, include_doppler_dephasing=True,
               include_intensity_noise=True, intensity_noise_frac=0.01,
               # atom
               config: AtomicConfiguration = None, species="Rb87", n_rydberg=70, qubit_0=(1, 0),
               qubit_1=(2, 0), hilbert_space_dim=3,
               # tweezer and environment
               tweezer_power=30e-3, tweezer_waist=1.0e-6, tweezer_wavelength_nm=None, temperature=2e-6,
               B_field=1e-4, NA=0.5, spacing_factor=2.8, include_noise=True,
               background_loss_rate_hz=None, trap_laser_on=True,
               # protocol parameters (None = reference defaults)
               delta_over_omega=None, omega_tau=None, pulse_shape="square",
               switching_times=None, phases=None,
               A=None, omega_mod_ratio=None, phi_offset=None,
               n_smooth_steps=300, n_shaped_steps=500, warn=True,
               # 5-level model only (hilbert_space_dim=5, extension)
               intermediate_branching_1=0.5, light_shift_compensation=True) -> GatePlan:
    """Vectorised steps 1-10 of simulate_CZ_gate plus the segment schedule of step 12."""
    if protocol not in PROTOCOL_NAMES:
        raise ValueError(f"Unknown protocol: {protocol}")
    if hilbert_space_dim not in (3, 4, 5):
        raise ValueError(f"Unsupported Hilbert space dimension: {hilbert_space_dim}. Use 3, 4 or 5.")
    if config is None:
        config = AtomicConfiguration(species=species, qubit_0=qubit_0, qubit_1=qubit_1,
                                     n_rydberg=n_rydberg, L_rydberg="S")
    atom = get_atom_properties(config.species)
    n_ryd = config.n_rydberg

    P1, w1, P2, w2 = (_arr(v, B) for v in (laser_1_power, laser_1_waist, laser_2_power, laser_2_waist))
    De = _arr(Delta_e, B)
    # --- laser linewidth (RG/simulation.py:2853-2860)
    if laser_1_linewidth_hz is not None and laser_2_linewidth_hz is not None:
        lw = np.sqrt(_arr(laser_1_linewidth_hz, B) ** 2 + _arr(laser_2_linewidth_hz, B) ** 2)
    elif laser_1_linewidth_hz is not None:
        lw = _arr(laser_1_linewidth_hz, B)
    elif laser_2_linewidth_hz is not None:
        lw = _arr(laser_2_linewidth_hz, B)
    else:
        lw = _arr(1000.0, B)
    # --- geometry (2869-2887)
    if tweezer_wavelength_nm is not None:
        trap_wl = _arr(tweezer_wavelength_nm, B) * 1e-9
    else:
        trap_wl = _arr(atom["trap_wavelength"], B)
    wl_nm = trap_wl * 1e9
    R = tweezer_spacing(trap_wl, _arr(NA, B), _arr(spacing_factor, B))
    # --- Rabi chain (2896-2912); quirk: dipole_er scales with plain n, not n*
    inter = get_default_intermediate_state(config.species)
    d_1e = atom["intermediate_states"][inter]["dipole_from_ground"]
    d_er = atom["dipole_intermediate_to_rydberg_ref"] * (n_ryd / atom["n_ref"]) ** (-1.5)
    Om1 = single_photon_rabi(d_1e, laser_E0(P1, w1))
    Om2 = single_photon_rabi(d_er, laser_E0(P2, w2))
    Om = two_photon_rabi(Om1, Om2, De)
    # --- blockade (2958-2968)
    V = rydberg_blockade(get_C6(n_ryd, config.species), R)
    with np.errstate(divide="ignore", invalid="ignore"):
        VoO = np.where(Om > 0, V / Om, np.inf)

    # --- protocol parameters (2989-3056)
    if protocol == "lp":
        tab_d, tab_o = lp_table_lookup(VoO)
        d_o = _opt(delta_over_omega, tab_d, B)
        o_t = _opt(omega_tau, tab_o, B)
        tau_single = o_t / Om
        tau_total = 2 * tau_single
        Delta_gate = d_o * Om
        n_pulses = 2
        leak_shape = pulse_shape
    elif protocol == "jp":
        o_t = _opt(omega_tau, JP_DEFAULT.omega_tau, B)
        tau_single = o_t / Om
        tau_total = tau_single.copy()
        Delta_gate = np.zeros(B)
        d_o = np.zeros(B)
        n_pulses = 1
        leak_shape = "bangbang"
        sw = JP_DEFAULT.switching_times if switching_times is None or np.size(switching_times) == 0 \
            else switching_times
        ph = JP_DEFAULT.phases if phases is None or np.size(phases) == 0 else phases
        sw = np.broadcast_to(np.asarray(sw, dtype=float), (B, np.shape(sw)[-1])).copy()
        ph = np.broadcast_to(np.asarray(ph, dtype=float), (B, np.shape(ph)[-1])).copy()
        assert ph.shape[1] == sw.shape[1] + 1, (
            f"Need len(phases) = len(switching_times) + 1, got {ph.shape[1]} phases and "
            f"{sw.shape[1]} switching times")
    else:
        o_t = _opt(omega_tau, SMOOTH_JP_PARAMS["omega_tau"], B)
        tau_single = o_t / Om
        tau_total = tau_single.copy()
        Delta_gate = np.zeros(B)          # reported Delta; the evolution's static detuning is below
        d_o = _opt(delta_over_omega, SMOOTH_JP_PARAMS["delta_over_omega"], B)
        n_pulses = 1
        leak_shape = "smooth_sinusoidal"

    # --- trap-dependent noise (3076-3097)
    tn = compute_trap_dependent_noise(
        species=config.species, tweezer_power=_arr(tweezer_power, B), tweezer_waist=_arr(tweezer_waist, B),
        temperature=_arr(temperature, B), spacing=R, gate_time=tau_total, n_rydberg=n_ryd,
        gamma_phi_laser=np.pi * lw, Omega_1=Om1, Delta_e=De, intermediate_state=config.intermediate_state,
        Omega_eff=Om, tweezer_wavelength_nm=wl_nm, include_doppler=include_doppler_dephasing,
        include_intensity_noise=include_intensity_noise, intensity_noise_frac=intensity_noise_frac,
        rydberg_wavelength_1_nm=config.excitation_wavelength_1_nm,
        rydberg_wavelength_2_nm=config.excitation_wavelength_2_nm, counter_propagating=counter_propagating)
    U0 = tn["trap_depth_uK"] * KB / 1e6
    # --- coherent shifts (3129-3159)
    d_zee = _arr(calculate_zeeman_shift(_arr(B_field, B), config.qubit_0, config.qubit_1, config.species), B)
    if trap_laser_on:
        depth_mK = tn["trap_depth_uK"] / 1000
        d_stark = np.where(depth_mK > 0,
                           calculate_qubit_stark_shift(_arr(tweezer_power, B), _arr(tweezer_waist, B),
                                                       config.species, np.where(depth_mK > 0, depth_mK, 1.0)),
                           calculate_qubit_stark_shift(_arr(tweezer_power, B), _arr(tweezer_waist, B),
                                                       config.species, None))
    else:
        d_stark = np.zeros(B)
    delta_1 = d_zee + d_stark
    # --- LP phase jump (3192)
    xi = compute_phase_shift_xi(Delta_gate, Om, tau_single) if protocol == "lp" else np.ones(B, dtype=complex)
    xi = np.broadcast_to(np.asarray(xi, dtype=complex), (B,)).copy()

    # --- rates (3219-3355)
    nb: Dict[str, object] = {
        "total_decay_rate": 0.0, "total_dephasing_rate": 0.0, "total_loss_rate": 0.0, "n_collapse_ops": 0,
        "motional_dephasing_included": include_motional_dephasing,
        "gamma_scatter_intermediate": tn["gamma_scatter_intermediate"],
        "Omega1_MHz": Om1 / (2 * np.pi * 1e6),
    }
    rates5 = [np.zeros(B) for _ in range(5)]
    if include_noise:
        g_phi_laser = np.pi * lw
        g_bg = _arr(tn["gamma_loss_background"], B) if background_loss_rate_hz is None \
            else _arr(background_loss_rate_hz, B)
        g_motional = _arr(tn["gamma_phi_thermal"], B) if include_motional_dephasing else np.zeros(B)
        g_doppler = _arr(tn.get("gamma_phi_doppler", 0.0), B)
        g_intensity = _arr(tn.get("gamma_phi_intensity", 0.0), B)
        B_rms = np.maximum(0.01 * _arr(B_field, B) * 1e4, 1e-3)
        K_quad = 575.0 if config.species == "Rb87" else 427.0
        g_zee = zeeman_dephasing_rate(B_rms, "clock" if config.is_clock_transition else "stretched", K_quad)
        time_factor = np.minimum(1.0, (tau_total / 1e-6) ** 2)
        g_anti = _arr(tn["gamma_loss_antitrap"], B) * 0.3 * time_factor
        g_leak = _arr(leakage_rate_to_adjacent_states(
            Omega=Om, Delta_leak=compute_leakage_detuning(config.species, n_ryd), pulse_shape=leak_shape,
            tau=tau_single, gamma_rydberg=tn["gamma_r"]), B)
        g_thermal_total = g_motional + g_doppler + g_intensity
        pur = np.minimum(_arr(polarization_purity_1, B), _arr(polarization_purity_2, B))
        if hilbert_space_dim == 4:
            dz_ryd = rydberg_zeeman_splitting(_arr(B_field, B), L=0, J=0.5)
            eps = 1.0 - pur
            with np.errstate(divide="ignore", invalid="ignore"):
                g_mj = np.where(np.abs(dz_ryd) < 1e-10, eps ** 2 * np.abs(Om), eps ** 2 * Om ** 2 / np.abs(dz_ryd))
        else:
            g_mj, dz_ryd, pur = np.zeros(B), np.zeros(B), np.ones(B)
        g_r = _arr(tn["gamma_r"], B)
        g_bbr = _arr(tn.get("gamma_bbr", 0), B)
        g_sc = _arr(tn["gamma_scatter_intermediate"], B)
        g_phi_tot = g_phi_laser + g_thermal_total + g_zee
        b1 = 0.5
        nd = hilbert_space_dim
        nryd = 2 if nd == 4 else 1
        # number of collapse operators under the reference's `> 0` guards (RG/noise_models.py:1199-1446)
        n_ops = (2 * nryd * (2 * (g_r > 0) + (g_bbr > 0) + (g_phi_tot > 0) + (g_anti > 0) + (g_bg > 0)
                             + (g_leak > 0)) + 2 * (g_sc > 0) + (4 * (g_mj > 0) if nd == 4 else 0))
        nb.update({
            "gamma_r": g_r, "gamma_bbr": g_bbr, "gamma_phi_laser": g_phi_laser,
            "gamma_phi_thermal": g_thermal_total, "gamma_phi_zeeman": g_zee,
            "gamma_loss_antitrap": g_anti, "gamma_loss_background": g_bg,
            "gamma_scatter_intermediate": g_sc, "gamma_leakage": g_leak, "mJ_leakage_rate": g_mj,
            "branching_1": b1, "gamma_phi_total": g_phi_tot, "total_decay_rate": g_r + g_bbr,
            "total_dephasing_rate": g_phi_tot, "total_loss_rate": g_anti + g_bg + g_leak, "dim": nd,
            "n_collapse_ops": n_ops.astype(int),
            "gamma_blockade_fluct": g_motional, "gamma_doppler": g_doppler,
            "gamma_intensity_noise": g_intensity, "gamma_thermal_total": g_thermal_total,
            "delta_V_over_V_percent": tn["blockade_fluctuation_percent"],
            "anti_trap_time_factor": time_factor, "magic_enhancement": tn["magic_enhancement"],
            "alpha_ratio": tn["alpha_ratio"], "k_eff_rad_per_m": tn.get("k_eff_rad_per_m", 0),
            "v_thermal_m_per_s": tn.get("v_thermal_m_per_s", 0), "gamma_mJ_leakage": g_mj,
            "polarization_1": polarization_1, "polarization_2": polarization_2,
            "polarization_purity_1": polarization_purity_1, "polarization_purity_2": polarization_purity_2,
            "combined_polarization_purity": pur, "Delta_zeeman_rydberg_Hz": dz_ryd / (2 * np.pi),
        })
        rates5 = [g_r * b1, g_r * (1 - b1) + g_bbr + g_anti + g_bg + g_leak, g_phi_tot, g_sc,
                  g_mj if nd == 4 else np.zeros(B)]

    # --- pulse schedule (step 12)
    pulse_info: Dict[str, object] = {}
    if protocol == "lp" and pulse_shape == "square":
        segs = np.zeros((B, 2, SEG_STRIDE))
        segs[:, :, 0] = tau_single[:, None]
        segs[:, 0, 1] = Om
        segs[:, 1, 1], segs[:, 1, 2] = (Om * xi).real, (Om * xi).imag
        segs[:, :, 3] = Delta_gate[:, None]
    elif protocol == "lp":
        # shaped LP (RG/simulation.py:2170-2220): 499 midpoints per pulse; the envelope is evaluated
        # on a scalar t_mid, one point at a time, exactly as the reference does (quirk above)
        n = n_shaped_steps
        segs = np.zeros((B, 2 * (n - 1), SEG_STRIDE))
        for i in range(B):
            c_i = area_correction_factor(pulse_shape, tau_single[i])
            t = np.linspace(0, tau_single[i], n)
            env = np.array([get_pulse_envelope(pulse_shape, (t[k] + t[k + 1]) / 2, tau_single[i])
                            for k in range(n - 1)])
            om_t = Om[i] * c_i * env
            segs[i, :, 0] = tau_single[i] / n
            segs[i, :n - 1, 1], segs[i, :n - 1, 2] = om_t.real, om_t.imag
            om2 = om_t * xi[i]
            segs[i, n - 1:, 1], segs[i, n - 1:, 2] = om2.real, om2.imag
            segs[i, :, 3] = Delta_gate[i]
            pulse_info.setdefault("area_correction", []).append(c_i)
    elif protocol == "jp":
        S = ph.shape[1]
        bounds = np.concatenate([np.zeros((B, 1)), sw, o_t[:, None]], axis=1) / Om[:, None]
        dt = np.diff(bounds, axis=1)
        dt = np.where(dt < 1e-18, 0.0, dt)             # zero / negative pieces are skipped (1902)
        segs = np.zeros((B, S, SEG_STRIDE))
        segs[:, :, 0] = dt
        segs[:, :, 1] = Om[:, None] * np.cos(ph)
        segs[:, :, 2] = Om[:, None] * np.sin(ph)
        pulse_info.update(switching_times=sw, phases=ph, n_segments=S, omega_tau=o_t)
    else:
        sA = _opt_or(A, SMOOTH_JP_PARAMS["A"], B)
        s_ratio = _opt_or(omega_mod_ratio, SMOOTH_JP_PARAMS["omega_mod_ratio"], B)
        s_phi = _opt_or(phi_offset, SMOOTH_JP_PARAMS["phi_offset"], B)
        mag = np.abs(d_o)
        s_doo = np.where(De > 0, -mag, mag)            # dark-state sign rule (3477-3483)
        n = n_smooth_steps
        dts = tau_total / n
        # t_mid = linspace(0, tau, n+1)[i] + dt/2  (RG/simulation.py:1698-1711)
        grid = np.arange(n)[None, :] * dts[:, None]     # == np.linspace(0, tau, n+1)[:n]
        t_mid = grid + dts[:, None] / 2
        phase = sA[:, None] * np.cos((s_ratio * Om)[:, None] * t_mid - s_phi[:, None])
        segs = np.zeros((B, n, SEG_STRIDE))
        segs[:, :, 0] = dts[:, None]
        segs[:, :, 1] = Om[:, None] * np.cos(phase)
        segs[:, :, 2] = Om[:, None] * np.sin(phase)
        segs[:, :, 3] = (s_doo * Om)[:, None]
        pulse_info.update(A=sA, omega_mod_ratio=s_ratio, phi_offset=s_phi, delta_over_omega=s_doo,
                          dark_state_valid=np.where(De > 0, s_doo < 0, s_doo > 0))
    n_seg = np.full(B, segs.shape[1], dtype=np.int32)

    points = np.zeros((B, POINT_STRIDE5 if hilbert_space_dim == 5 else POINT_STRIDE))
    points[:, 0], points[:, 1] = V, delta_1
    for k in range(5):
        points[:, 4 + k] = rates5[k]
    if hilbert_space_dim == 5:
        # The intermediate state is explicit: scattering is the decay of |e> (rate Γe, split between
        # |1> and |0>), not a dephasing rate on |1>; the loss channels of |r> that remove the atom
        # (anti-trapping, background gas, leakage to neighbouring Rydberg states) end in |l>, real
        # decay to other ground states in |0>.
        Ge = atom["intermediate_states"][config.intermediate_state]["linewidth"]
        points[:, 7] = 0.0
        points[:, 12], points[:, 13] = Om1, De
        points[:, 17] = 1.0 if light_shift_compensation else 0.0
        if include_noise:
            points[:, 14] = Ge * intermediate_branching_1
            points[:, 15] = Ge * (1.0 - intermediate_branching_1)
            points[:, 5] = g_r * (1 - b1) + g_bbr
            points[:, 16] = g_anti + g_bg + g_leak
            # one jump operator per atom and non-zero collapsed rate
            n5 = 2 * sum((points[:, k] > 0).astype(int) for k in (4, 5, 6, 14, 15, 16))
            nb.update({"gamma_e": Ge, "intermediate_branching_1": intermediate_branching_1, "n_collapse_ops": n5})
    # Monte-Carlo shot widths (extension; used only when shots are requested): per-atom Doppler
    # detuning std k_eff v_rms and relative distance std sqrt(2) sigma_r / R
    k1 = 2 * np.pi / (config.excitation_wavelength_1_nm * 1e-9)
    k2 = 2 * np.pi / (config.excitation_wavelength_2_nm * 1e-9)
    k_eff = abs(k1 - k2) if counter_propagating else k1 + k2
    points[:, 9] = k_eff * np.sqrt(KB * _arr(temperature, B) / atom["mass"])
    points[:, 10] = np.sqrt(2) * (tn["position_uncertainty_nm"] * 1e-9) / R

    if warn:
        _emit_warnings(protocol, Om, VoO, delta_over_omega, omega_tau, pulse_info, De)

    scalars = dict(Omega=Om, Omega1=Om1, Omega2=Om2, V=V, Delta=Delta_gate, V_over_Omega=VoO,
                   Delta_over_Omega=d_o, tau_single=tau_single, tau_total=tau_total, R=R, xi=xi,
                   spacing_factor=_arr(spacing_factor, B), U0_mK=U0 / KB * 1e3,
                   omega_r_kHz=_arr(tn["trap_freq_radial_kHz"], B),
                   sigma_r_nm=_arr(tn["position_uncertainty_nm"], B), trap_wavelength_nm=wl_nm,
                   delta_zeeman=d_zee, delta_stark=d_stark, delta_1=delta_1, omega_tau=o_t,
                   temperature_K=_arr(temperature, B), B_field_T=_arr(B_field, B))
    return GatePlan(protocol=protocol, dim=hilbert_space_dim, B=B, include_noise=include_noise,
                    points=points, segs=segs, n_seg=n_seg, scalars=scalars, noise_breakdown=nb,
                    trap_noise=tn, pulse_info=pulse_info, config=config, n_pulses=n_pulses,
                    pulse_shape=pulse_shape if protocol == "lp" else leak_shape, trap_laser_on=trap_laser_on)


def _emit_warnings(protocol, Om, VoO, delta_over_omega, omega_tau, pulse_info, De):
    """The reference's UserWarnings (RG/simulation.py:2930-2946, 1640-1647, 1670-1676;
    RG/protocols.py:615-621), once per batch."""
    mhz = Om / (2 * np.pi * 1e6)
    if np.any(Om > 2 * np.pi * 100e6):
        warnings.warn(f"Ω/2π = {float(np.max(mhz)):.1f} MHz exceeds physical limit (~100 MHz). "
                      f"Results may be unphysical. Check laser powers.", UserWarning)
    if np.any(Om < 2 * np.pi * 0.1e6):
        warnings.warn(f"Ω/2π = {float(np.min(mhz)) * 1e3:.1f} kHz is very low. Gate will be very slow "
                      f"and susceptible to decoherence.", UserWarning)
    if protocol == "lp" and np.any(VoO < 10):
        warnings.warn(f"V/Ω = {float(np.min(VoO)):.1f} < 10. Blockade too weak for reliable CZ gate!", UserWarning)
    if protocol == "sjp":
        if np.any(~np.asarray(pulse_info["dark_state_valid"]) & (pulse_info["delta_over_omega"] != 0)):
            warnings.warn("Dark state condition violated! This will increase scattering error.", UserWarning)
        if np.any(VoO < 5):
            warnings.warn(f"V/Ω = {float(np.min(VoO)):.1f} may be too weak for reliable CZ operation. "
                          f"Recommend V/Ω > 10 for high-fidelity gates.", UserWarning)
