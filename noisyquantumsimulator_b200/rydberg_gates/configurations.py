"""Input records of simulate_CZ_gate (reference: RG/configurations.py).  Field names, defaults and
the derived properties callers read are the reference's."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .atom_database import (ATOM_DB, effective_n, get_C6, get_hyperfine_splitting,
                            get_intermediate_state_linewidth, get_quantum_defect, get_rydberg_energy,
                            get_rydberg_lifetime, get_rydberg_polarizability)
from .constants import C, EPS0, HBAR, KB


@dataclass
class LaserParameters:
    power: float = 1e-3
    waist: float = 50e-6
    polarization: str = "sigma+"
    polarization_purity: float = 0.99
    linewidth_hz: float = 100.0

    def peak_intensity(self) -> float:
        return 2 * self.power / (np.pi * self.waist ** 2)

    def peak_electric_field(self) -> float:
        return np.sqrt(2 * self.peak_intensity() / (EPS0 * C))


@dataclass
class TwoPhotonExcitationConfig:
    laser_1: LaserParameters = field(default_factory=lambda: LaserParameters(
        power=50e-6, waist=50e-6, polarization="pi", linewidth_hz=1000))
    laser_2: LaserParameters = field(default_factory=lambda: LaserParameters(
        power=500e-3, waist=50e-6, polarization="sigma+", linewidth_hz=1000))
    Delta_e: float = 2 * np.pi * 1e9
    counter_propagating: bool = True


@dataclass
class NoiseSourceConfig:
    """Only include_motional_dephasing, include_doppler_dephasing, include_intensity_noise and
    intensity_noise_frac change the result; the other flags are read and ignored by the
    reference (RG/simulation.py:2804-2811) and therefore here."""
    include_spontaneous_emission: bool = True
    include_intermediate_scattering: bool = True
    include_motional_dephasing: bool = True
    include_doppler_dephasing: bool = True
    include_intensity_noise: bool = True
    intensity_noise_frac: float = 0.01
    include_laser_dephasing: bool = True
    include_magnetic_dephasing: bool = True


@dataclass
class LPSimulationInputs:
    excitation: TwoPhotonExcitationConfig = field(default_factory=TwoPhotonExcitationConfig)
    noise: NoiseSourceConfig = field(default_factory=NoiseSourceConfig)
    delta_over_omega: Optional[float] = None
    omega_tau: Optional[float] = None
    pulse_shape: str = "square"
    drag_lambda: float = 1.0

    @property
    def protocol_name(self) -> str:
        return "levine_pichler"

    @property
    def n_pulses(self) -> int:
        return 2


@dataclass
class JPSimulationInputs:
    excitation: TwoPhotonExcitationConfig = field(default_factory=TwoPhotonExcitationConfig)
    noise: NoiseSourceConfig = field(default_factory=NoiseSourceConfig)
    omega_tau: Optional[float] = None
    switching_times: Optional[List[float]] = None
    phases: Optional[List[float]] = None

    @property
    def protocol_name(self) -> str:
        return "jandura_pupillo"

    @property
    def pulse_shape(self) -> str:
        return "bangbang"

    @property
    def n_pulses(self) -> int:
        return 1


@dataclass
class SmoothJPSimulationInputs:
    excitation: TwoPhotonExcitationConfig = field(default_factory=TwoPhotonExcitationConfig)
    noise: NoiseSourceConfig = field(default_factory=NoiseSourceConfig)
    omega_tau: Optional[float] = None
    A: Optional[float] = None
    omega_mod_ratio: Optional[float] = None
    phi_offset: Optional[float] = None
    delta_over_omega: Optional[float] = None

    @property
    def protocol_name(self) -> str:
        return "smooth_jp"

    @property
    def pulse_shape(self) -> str:
        return "smooth_sinusoidal"

    @property
    def n_pulses(self) -> int:
        return 1


@dataclass
class TweezerParameters:
    power: float = 10e-3
    waist: float = 0.9e-6
    wavelength_nm: float = 820.0
    NA: float = 0.5

    def diffraction_limited_waist(self) -> float:
        return 0.64 * self.wavelength_nm * 1e-9 / self.NA

    def rayleigh_range(self) -> float:
        return np.pi * self.waist ** 2 / (self.wavelength_nm * 1e-9)

    def peak_intensity(self) -> float:
        return 2 * self.power / (np.pi * self.waist ** 2)


@dataclass
class EnvironmentParameters:
    temperature: float = 20e-6
    B_field: float = 0.5e-4
    B_field_angle: float = 0.0
    spacing_factor: float = 2.8

    def thermal_velocity(self, mass: float) -> float:
        return np.sqrt(KB * self.temperature / mass)

    def B_field_gauss(self) -> float:
        return self.B_field / 1e-4


_D_LINES = {"Rb87": {"5P3/2": 384.230484468e12, "5P1/2": 377.107385690e12},
            "Cs133": {"6P3/2": 351.725718509e12, "6P1/2": 335.116048807e12}}


@dataclass
class AtomicConfiguration:
    species: str = "Rb87"
    n_rydberg: int = 70
    L_rydberg: str = "S"
    qubit_0: Tuple[int, int] = (1, 0)
    qubit_1: Tuple[int, int] = (2, 0)
    intermediate_state: str = None

    def __post_init__(self):
        if self.species not in ATOM_DB:
            raise ValueError(f"Unknown species: {self.species}. Available: {list(ATOM_DB.keys())}")
        self._atom_props = ATOM_DB[self.species]
        if self.intermediate_state is None:
            self.intermediate_state = {"Rb87": "5P3/2", "Cs133": "6P3/2"}[self.species]

    # --- Rydberg state ---
    @property
    def n_star(self) -> float:
        return effective_n(self.n_rydberg, self.species, self.L_rydberg)

    @property
    def quantum_defect(self) -> float:
        return get_quantum_defect(self.species, self.L_rydberg)

    @property
    def rydberg_energy(self) -> float:
        return get_rydberg_energy(self.n_rydberg, self.species, self.L_rydberg)

    @property
    def C6(self) -> float:
        return get_C6(self.n_rydberg, self.species)

    @property
    def rydberg_lifetime_300K(self) -> float:
        return get_rydberg_lifetime(self.n_rydberg, self.species, 300.0)

    @property
    def rydberg_lifetime_0K(self) -> float:
        return get_rydberg_lifetime(self.n_rydberg, self.species, 0.0)

    @property
    def rydberg_polarizability(self) -> float:
        return get_rydberg_polarizability(self.n_rydberg, self.species)

    # --- ground state / qubit ---
    @property
    def mass(self) -> float:
        return self._atom_props["mass"]

    @property
    def hyperfine_splitting(self) -> float:
        return get_hyperfine_splitting(self.species)

    @property
    def ground_polarizability(self) -> float:
        return self._atom_props["alpha_ground"]

    @property
    def is_clock_transition(self) -> bool:
        return self.qubit_0[1] == 0 and self.qubit_1[1] == 0

    @property
    def delta_mF(self) -> int:
        return abs(self.qubit_1[1] - self.qubit_0[1])

    @property
    def delta_F(self) -> int:
        return abs(self.qubit_1[0] - self.qubit_0[0])

    @property
    def intermediate_linewidth(self) -> float:
        return get_intermediate_state_linewidth(self.species, self.intermediate_state)

    # --- excitation wavelengths (feed k_eff of the Doppler rate) ---
    def _first_leg_hz(self) -> float:
        lines = _D_LINES["Rb87" if self.species == "Rb87" else "Cs133"]
        strong = "5P3/2" if self.species == "Rb87" else "6P3/2"
        weak = "5P1/2" if self.species == "Rb87" else "6P1/2"
        key = strong if self.intermediate_state == strong else weak
        return self._atom_props.get("transitions", {}).get(f"ground_to_{key}", lines[key])

    @property
    def excitation_wavelength_1_nm(self) -> float:
        return C / self._first_leg_hz() * 1e9

    @property
    def excitation_wavelength_2_nm(self) -> float:
        """Second leg from energy conservation: E_ion + E_binding(n*) - h f1
        (RG/configurations.py:855-884)."""
        h = HBAR * 2 * np.pi
        E2 = (self._atom_props["E_ionization"] + self.rydberg_energy) - h * self._first_leg_hz()
        return C / (E2 / h) * 1e9

    def get_g_F(self, state: Tuple[int, int]) -> float:
        return self._atom_props["g_F"][state[0]]

    def blockade_radius(self, Omega: float) -> float:
        return (self.C6 / (HBAR * Omega)) ** (1 / 6)

    def blockade_shift(self, R: float) -> float:
        return self.C6 / R ** 6

    def summary(self) -> str:
        return "\n".join([
            "AtomicConfiguration Summary", "=" * 40, f"Species: {self.species}",
            f"Rydberg state: {self.n_rydberg}{self.L_rydberg}₁/₂", f"  n* = {self.n_star:.3f}",
            f"  C₆/2π = {self.C6 / (2 * np.pi):.2e} Hz·m⁶",
            f"        = {self.C6 / (2 * np.pi * 1e9 * (1e-6) ** 6):.1f} GHz·μm⁶",
            f"  τ(300K) = {self.rydberg_lifetime_300K * 1e6:.1f} μs", "Qubit encoding:",
            f"  |0⟩ = |F={self.qubit_0[0]}, mF={self.qubit_0[1]}⟩",
            f"  |1⟩ = |F={self.qubit_1[0]}, mF={self.qubit_1[1]}⟩",
            f"  Clock transition: {self.is_clock_transition}", f"Intermediate: {self.intermediate_state}"])


def get_standard_rb87_config(n_rydberg: int = 70) -> AtomicConfiguration:
    return AtomicConfiguration(species="Rb87", n_rydberg=n_rydberg, L_rydberg="S", qubit_0=(1, 0),
                               qubit_1=(2, 0), intermediate_state="5P3/2")


def get_standard_cs133_config(n_rydberg: int = 70) -> AtomicConfiguration:
    return AtomicConfiguration(species="Cs133", n_rydberg=n_rydberg, L_rydberg="S", qubit_0=(3, 0),
                               qubit_1=(4, 0), intermediate_state="6P3/2")


__all__ = ["LaserParameters", "TweezerParameters", "EnvironmentParameters", "AtomicConfiguration",
           "TwoPhotonExcitationConfig", "NoiseSourceConfig", "LPSimulationInputs", "JPSimulationInputs",
           "SmoothJPSimulationInputs", "get_standard_rb87_config", "get_standard_cs133_config"]
