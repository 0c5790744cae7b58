"""Host half of the device prologue (csrc/rydcz_prologue.cuh, include/rydcz.h "simulate").

``model_constants`` evaluates everything that depends only on species, n, qubit states and flags
-- once per call, with the same mirror functions the NumPy prologue uses -- into the flat
``model`` array; ``pack_apparatus`` lays the per-point physical inputs out as apparatus rows.
The per-point physics then runs on the GPU (``rydcz_simulate_host`` / ``_device``).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .atom_database import get_atom_properties, get_C6, get_default_intermediate_state, get_rydberg_lifetime
from .configurations import AtomicConfiguration
from .constants import C, EPS0, HBAR, MU_B
from .noise_models import rydberg_zeeman_splitting
from .pulse_shaping import compute_leakage_detuning
from .trap_physics import get_polarizability_at_wavelength

APP_STRIDE, MODEL_STRIDE, SCALAR_STRIDE, APP_PROTO = 40, 32, 8, 20
PROTOCOL_CODE = {"lp": 0, "jp": 1, "sjp": 2}
LEAK_SHAPE_CODE = {"square": 0, "gaussian": 1, "cosine": 2, "blackman": 3, "drag": 4}
SCALARS = ("Omega", "V", "tau_total", "tau_single", "Delta", "xi_re", "xi_im", "omega_tau")


_MODEL_CACHE: Dict = {}


def model_constants(protocol: str, **kwargs) -> np.ndarray:
    """The 32 model constants of a batch (``_model_constants``), kept per distinct keyword set: an optimiser calls
    with the same model thousands of times, and the atom-database lookups behind the constants took a third of the
    Python time of a small generation.  (An ``AtomicConfiguration`` object, or any unhashable value, is not cached.)"""
    try:
        key = (protocol, tuple(sorted(kwargs.items())))
        hit = _MODEL_CACHE.get(key)
    except TypeError:
        return _model_constants(protocol, **kwargs)
    if hit is None:
        if kwargs.get("config") is not None:
            return _model_constants(protocol, **kwargs)
        if len(_MODEL_CACHE) > 256:
            _MODEL_CACHE.clear()
        hit = _MODEL_CACHE[key] = _model_constants(protocol, **kwargs)
        hit.setflags(write=False)
    return hit


def _model_constants(protocol: str, *, config: AtomicConfiguration = None, species="Rb87", n_rydberg=70,
                    qubit_0=(1, 0), qubit_1=(2, 0), hilbert_space_dim=3, include_noise=True,
                    include_motional_dephasing=True, include_doppler_dephasing=True,
                    include_intensity_noise=True, counter_propagating=True, trap_laser_on=True,
                    n_segments: int = None, n_smooth_steps: int = 300, intermediate_branching_1: float = 0.5,
                    light_shift_compensation: bool = True, pulse_shape: str = "square",
                    n_shaped_steps: int = 500) -> np.ndarray:
    if config is None:
        config = AtomicConfiguration(species=species, qubit_0=qubit_0, qubit_1=qubit_1, n_rydberg=n_rydberg,
                                     L_rydberg="S")
    atom = get_atom_properties(config.species)
    n = config.n_rydberg
    m = np.zeros(MODEL_STRIDE)
    m[0] = hilbert_space_dim
    m[1] = PROTOCOL_CODE[protocol]
    shaped = protocol == "lp" and pulse_shape != "square"
    if shaped and pulse_shape not in ("gaussian", "cosine", "blackman"):
        # 'drag' needs a Delta_leak the reference never passes (RG/simulation.py:2205): same failure as there
        from .pulse_shaping import get_pulse_envelope
        get_pulse_envelope(pulse_shape, 0.5, 1.0)
        raise ValueError(f"pulse shape '{pulse_shape}' is not available in the device prologue")
    m[2] = {"lp": 2 * (n_shaped_steps - 1) if shaped else 2, "jp": n_segments or 5, "sjp": n_smooth_steps}[protocol]
    m[3:9] = [include_noise, include_motional_dephasing, include_doppler_dephasing, include_intensity_noise,
              trap_laser_on, config.is_clock_transition]
    # leakage spectral factor: the LP pulse shape; 'bangbang' / 'smooth_sinusoidal' take the generic sinc²
    m[9] = {"lp": LEAK_SHAPE_CODE[pulse_shape], "jp": 5, "sjp": 5}[protocol]
    inter = get_default_intermediate_state(config.species)
    m[10] = atom["mass"]
    m[11] = atom["alpha_ground"]
    m[12] = atom["trap_wavelength"]
    m[13] = get_C6(n, config.species)
    m[14] = atom["intermediate_states"][inter]["dipole_from_ground"]
    m[15] = atom["dipole_intermediate_to_rydberg_ref"] * (n / atom["n_ref"]) ** (-1.5)   # plain n (quirk)
    m[16] = atom["intermediate_states"][config.intermediate_state]["linewidth"]
    m[17] = 1.0 / get_rydberg_lifetime(n, config.species, temperature=300.0)
    # static Rydberg polarizability at n; the kernel applies the (λ/1064 nm)² factor
    m[18] = get_polarizability_at_wavelength(config.species, "rydberg", 1064.0, n_rydberg=n)
    m[19] = 2 * np.pi * (377.107e12 if config.species == "Rb87" else 335.116e12)
    m[20] = 575.0 if config.species == "Rb87" else 427.0
    m[21] = {"Rb87": 575.0, "Cs133": 2000.0}.get(config.species, 1000.0)
    if not config.is_clock_transition:
        from .trap_physics import calculate_zeeman_shift
        m[22] = calculate_zeeman_shift(1.0, config.qubit_0, config.qubit_1, config.species)   # per tesla
    m[23] = {"Rb87": 70e3, "Cs133": 200e3}.get(config.species, 70e3)
    m[24] = {"Rb87": 2.4, "Cs133": 7.0}.get(config.species, 2.4) * 1.6488e-41 / (4 * np.pi * EPS0 * C * HBAR)
    k1 = 2 * np.pi / (config.excitation_wavelength_1_nm * 1e-9)
    k2 = 2 * np.pi / (config.excitation_wavelength_2_nm * 1e-9)
    m[25] = abs(k1 - k2) if counter_propagating else k1 + k2
    m[26] = compute_leakage_detuning(config.species, n)
    m[27] = rydberg_zeeman_splitting(1.0, L=0, J=0.5)
    m[28] = 1e3
    m[29] = intermediate_branching_1          # dim 5 only
    m[30] = 1.0 if light_shift_compensation else 0.0
    # shaped LP: square area / envelope area on the reference's 1000-point grid (independent of tau)
    from .pulse_shaping import area_correction_factor
    m[31] = area_correction_factor(pulse_shape, 1.0) if shaped else 1.0
    return m


def apparatus_columns(protocol: str, B: int, *, laser_1_power, laser_1_waist, laser_2_power, laser_2_waist, Delta_e,
                      laser_1_linewidth_hz=100.0, laser_2_linewidth_hz=100.0, polarization_purity_1=0.99,
                      polarization_purity_2=0.99, tweezer_power=30e-3, tweezer_waist=1.0e-6,
                      tweezer_wavelength_nm=None, temperature=2e-6, B_field=1e-4, NA=0.5, spacing_factor=2.8,
                      background_loss_rate_hz=None, intensity_noise_frac=0.01, delta_over_omega=None,
                      omega_tau=None, switching_times=None, phases=None, A=None, omega_mod_ratio=None,
                      phi_offset=None) -> List:
    """The 40 apparatus slots as a list: a float (same for every point; None -> NaN = 'use the
    reference default') or a 1-D float64 array view of length B (possibly strided: column k of a
    row-major [B, m] array is passed as is).  Nothing is copied for float64 inputs."""
    cols: List = [0.0] * APP_STRIDE

    def put(slot, v):
        if v is None:
            cols[slot] = np.nan
        elif type(v) is float or type(v) is int:
            cols[slot] = float(v)
        elif np.ndim(v) == 0:
            cols[slot] = float(v)
        else:
            v = np.asarray(v, dtype=np.float64)
            assert v.shape == (B,), f"apparatus slot {slot}: expected a scalar or {B} values, got shape {v.shape}"
            cols[slot] = v

    for slot, v in enumerate((laser_1_power, laser_1_waist, laser_2_power, laser_2_waist, Delta_e,
                              laser_1_linewidth_hz, laser_2_linewidth_hz)):
        put(slot, v)
    put(7, min(polarization_purity_1, polarization_purity_2)
        if type(polarization_purity_1) is float and type(polarization_purity_2) is float
        else np.minimum(polarization_purity_1, polarization_purity_2))
    for slot, v in ((8, tweezer_power), (9, tweezer_waist), (10, tweezer_wavelength_nm), (11, temperature),
                    (12, B_field), (13, NA), (14, spacing_factor), (15, background_loss_rate_hz),
                    (16, intensity_noise_frac)):
        put(slot, v)
    P = APP_PROTO
    if protocol == "lp":
        put(P, delta_over_omega)
        put(P + 1, omega_tau)
    elif protocol == "sjp":
        for k, v in enumerate((omega_tau, A, omega_mod_ratio, phi_offset, delta_over_omega)):
            put(P + k, v)
    else:
        from .protocols import JP_DEFAULT
        sw = JP_DEFAULT.switching_times if switching_times is None or np.size(switching_times) == 0 else switching_times
        ph = JP_DEFAULT.phases if phases is None or np.size(phases) == 0 else phases
        sw, ph = np.asarray(sw, dtype=np.float64), np.asarray(ph, dtype=np.float64)
        S = ph.shape[-1]
        assert S == sw.shape[-1] + 1, (f"Need len(phases) = len(switching_times) + 1, got {S} phases and "
                                       f"{sw.shape[-1]} switching times")
        assert 2 * S <= APP_STRIDE - APP_PROTO, "too many bang-bang segments for the device prologue"
        put(P, omega_tau)
        for k in range(S - 1):
            put(P + 1 + k, sw[:, k] if sw.ndim == 2 else sw[k])
        for k in range(S):
            put(P + S + k, ph[:, k] if ph.ndim == 2 else ph[k])
    return cols


def pack_apparatus(protocol: str, B: int, *, out: np.ndarray = None, **kwargs) -> np.ndarray:
    """[40, B] apparatus block (slot-major: one contiguous column per slot) from the same keywords as
    ``apparatus_columns``; scalars broadcast."""
    app = out if out is not None else np.empty((APP_STRIDE, B))
    assert app.shape == (APP_STRIDE, B)
    for slot, v in enumerate(apparatus_columns(protocol, B, **kwargs)):
        app[slot] = v
    return app


def split_kwargs(kwargs: Dict):
    """Separate build_plan-style keywords into (model keywords, apparatus keywords)."""
    model_keys = {"config", "species", "n_rydberg", "qubit_0", "qubit_1", "hilbert_space_dim", "include_noise",
                  "include_motional_dephasing", "include_doppler_dephasing", "include_intensity_noise",
                  "counter_propagating", "trap_laser_on", "n_smooth_steps", "intermediate_branching_1",
                  "light_shift_compensation", "pulse_shape", "n_shaped_steps"}
    ignored = {"polarization_1", "polarization_2", "warn"}
    mk = {k: v for k, v in kwargs.items() if k in model_keys}
    ak = {k: v for k, v in kwargs.items() if k not in model_keys and k not in ignored}
    return mk, ak
