"""SI constants, same values as the reference (RG/constants.py:95-302) so every derived scalar
(Ω, V, rates) agrees to rounding."""
import numpy as np

HBAR = 1.054571817e-34
EPS0 = 8.8541878128e-12
C = 299792458.0
E_CHARGE = 1.602176634e-19
A0 = 5.29177210903e-11
KB = 1.380649e-23
MU_B = 9.2740100783e-24
RY_JOULES = 2.1798723611035e-18
RY_EV = RY_JOULES / E_CHARGE
G_I_RB87 = -0.0009951414
G_I_CS133 = -0.00039885395
G_E = 2.00231930436256

AU_POLARIZABILITY = 4 * np.pi * EPS0 * A0 ** 3   # one atomic unit of polarizability in SI


def frequency_to_energy(freq_hz):
    return 2 * np.pi * HBAR * freq_hz


def energy_to_frequency(energy_joules):
    return energy_joules / (2 * np.pi * HBAR)


def temperature_to_energy(temp_kelvin):
    return KB * temp_kelvin


def wavelength_to_frequency(wavelength_m):
    return C / wavelength_m


def au_to_si_polarizability(alpha_au):
    return alpha_au * AU_POLARIZABILITY


def si_to_au_polarizability(alpha_si):
    return alpha_si / AU_POLARIZABILITY


__all__ = ["HBAR", "EPS0", "C", "E_CHARGE", "A0", "KB", "MU_B", "RY_JOULES", "RY_EV", "G_I_RB87",
           "G_I_CS133", "G_E", "frequency_to_energy", "energy_to_frequency",
           "temperature_to_energy", "wavelength_to_frequency", "au_to_si_polarizability",
           "si_to_au_polarizability"]
