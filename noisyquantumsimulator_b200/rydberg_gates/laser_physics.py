"""Laser -> Rabi frequency chain and the blockade shift (reference: RG/laser_physics.py:111-669).
Every function is elementwise, so NumPy arrays of parameter points pass straight through."""
from __future__ import annotations

import numpy as np

from .atom_database import ATOM_DB, effective_n
from .constants import C, EPS0, HBAR


def laser_intensity(power, waist):
    """Peak intensity of a Gaussian beam, 2P/(π w²)."""
    return 2 * power / (np.pi * waist ** 2)


def laser_E0(power, waist):
    """Peak field E0 = sqrt(2 I / (ε0 c))  (RG/laser_physics.py:155-158)."""
    return np.sqrt(2 * laser_intensity(power, waist) / (EPS0 * C))


def single_photon_rabi(dipole, E0):
    return dipole * E0 / HBAR


def single_photon_rabi_from_power(dipole, power, waist):
    return single_photon_rabi(dipole, laser_E0(power, waist))


def two_photon_rabi(Omega1, Omega2, Delta_e):
    """Ω = Ω1 Ω2 / (2 Δe)  (RG/laser_physics.py:359)."""
    return Omega1 * Omega2 / (2 * Delta_e)


def required_powers_for_two_photon_rabi(Omega_eff_target, Delta_e, dipole_1, dipole_2, waist_1, waist_2,
                                        power_ratio=1.0):
    prefactor = (HBAR ** 2 * np.pi * EPS0 * C) / (4 * dipole_1 * dipole_2)
    P_product = (Omega_eff_target * 2 * Delta_e * prefactor) ** 2 * waist_1 ** 2 * waist_2 ** 2
    P1 = np.sqrt(P_product / power_ratio)
    return P1, power_ratio * P1


def rydberg_blockade(C6, R):
    """V = C6 / R⁶ (rad/s, since C6_ref carries the 2π)."""
    return C6 / R ** 6


def blockade_shift_MHz(C6_GHz_um6, R_um):
    return C6_GHz_um6 * 1000 / R_um ** 6


def blockade_radius(C6, Omega):
    return (C6 / (HBAR * Omega)) ** (1 / 6)


def blockade_radius_um(C6_GHz_um6, Omega_MHz):
    return (C6_GHz_um6 * 1000 / Omega_MHz) ** (1 / 6)


def V_over_Omega(C6, R, Omega):
    return rydberg_blockade(C6, R) / (HBAR * Omega)


def spacing_for_target_V_over_Omega(C6, Omega, target_ratio):
    return (C6 / (HBAR * Omega * target_ratio)) ** (1 / 6)


def intermediate_state_scattering_rate(Omega1, Delta_e, Gamma_e):
    return Omega1 ** 2 * Gamma_e / (4 * Delta_e ** 2)


def dark_state_suppression_factor(Omega1, Omega2):
    return Omega2 ** 2 / (Omega1 ** 2 + Omega2 ** 2)


def _cg_table():
    s = np.sqrt
    rows = {
        (1, -1, 0): (0, 0, 1 / s(3)), (1, 0, 0): (1 / s(3), 0, 0), (1, 1, 0): (0, 1 / s(3), 0),
        (1, -1, 1): (1 / s(6), -1 / s(2), 0), (1, 0, 1): (0, 1 / s(6), -1 / s(6)), (1, 1, 1): (-1 / s(6), 0, 1 / s(2)),
        (1, -1, 2): (-1 / s(2), 0, -1 / s(10)), (1, 0, 2): (-s(2 / 5), -1 / s(2), 1 / s(2)),
        (1, 1, 2): (-1 / s(2), 1 / s(10), 0),
        (2, -2, 1): (0, 0, 1 / s(2)), (2, -1, 1): (1 / s(6), 0, 1 / s(6)), (2, 0, 1): (s(2 / 5), 1 / s(6), 1 / s(6)),
        (2, 1, 1): (1 / s(6), 1 / s(6), 0), (2, 2, 1): (0, 1 / s(2), 0),
        (2, -2, 2): (-1 / s(3), 0, -1 / s(30)), (2, -1, 2): (-1 / s(12), -1 / s(3), 1 / s(20)),
        (2, 0, 2): (0, -1 / s(12), 1 / s(12)), (2, 1, 2): (1 / s(12), -1 / s(20), 1 / s(3)),
        (2, 2, 2): (1 / s(3), 1 / s(30), 0),
        (2, -2, 3): (-s(2 / 5), 0, -1 / s(15)), (2, -1, 3): (-s(3 / 10), -s(2 / 5), 1 / (2 * s(5))),
        (2, 0, 3): (-s(2 / 5), -s(3 / 10), s(3 / 10)), (2, 1, 3): (-s(3 / 10), -1 / (2 * s(5)), s(2 / 5)),
        (2, 2, 3): (-s(2 / 5), 1 / s(15), 0),
    }
    return {k: dict(zip(("pi", "sigma+", "sigma-"), v)) for k, v in rows.items()}


CLEBSCH_GORDAN_D2 = _cg_table()


def get_clebsch_gordan(F_i, mF_i, F_f, polarization):
    return CLEBSCH_GORDAN_D2.get((F_i, mF_i, F_f), {}).get(polarization, 0) or 0.0


def compute_all_rabi_frequencies(species, n_rydberg, power_1, power_2, waist_1, waist_2, Delta_e,
                                 intermediate_state="5P3/2"):
    """Note: this helper scales ⟨nS|er|P⟩ with n*, unlike simulate_CZ_gate which uses plain n
    (RG/laser_physics.py:923-925 vs RG/simulation.py:2907)."""
    atom = ATOM_DB[species]
    d1 = atom["intermediate_states"][intermediate_state]["dipole_from_ground"]
    ratio = effective_n(n_rydberg, species, "S") / effective_n(atom["n_ref"], species, "S")
    d2 = atom["dipole_intermediate_to_rydberg_ref"] * ratio ** (-1.5)
    O1 = single_photon_rabi(d1, laser_E0(power_1, waist_1))
    O2 = single_photon_rabi(d2, laser_E0(power_2, waist_2))
    Oe = two_photon_rabi(O1, O2, Delta_e)
    two_pi = 2 * np.pi
    return {"Omega1": O1, "Omega2": O2, "Omega_eff": Oe, "Omega1_MHz": O1 / two_pi / 1e6,
            "Omega2_MHz": O2 / two_pi / 1e6, "Omega_eff_MHz": Oe / two_pi / 1e6,
            "Delta_e_GHz": Delta_e / two_pi / 1e9}


__all__ = ["laser_E0", "laser_intensity", "single_photon_rabi", "single_photon_rabi_from_power",
           "two_photon_rabi", "required_powers_for_two_photon_rabi", "rydberg_blockade",
           "blockade_shift_MHz", "blockade_radius", "blockade_radius_um", "V_over_Omega",
           "spacing_for_target_V_over_Omega", "intermediate_state_scattering_rate",
           "dark_state_suppression_factor", "CLEBSCH_GORDAN_D2", "get_clebsch_gordan",
           "compute_all_rabi_frequencies"]
