"""Species table and n-scaling laws (reference: RG/atom_database.py:104-456, 662-914).

Only Rb87 and Cs133 exist in the reference.  The table is generated from a compact
per-species spec; keys and values match the reference's ATOM_DB so callers that index it
(``ATOM_DB[species]["mass"]`` ...) keep working.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .constants import A0, E_CHARGE, EPS0, G_I_CS133, G_I_RB87, HBAR, RY_JOULES

_AU = 4 * np.pi * EPS0 * A0 ** 3
_EA0 = E_CHARGE * A0
_H = HBAR * 2 * np.pi


def _hyperfine(F_lo, e_lo_hz, F_hi, e_hi_hz, term):
    out = {}
    for F, e in ((F_lo, e_lo_hz), (F_hi, e_hi_hz)):
        for mF in range(-F, F + 1):
            out[(F, mF)] = {"energy_B0": e * _H,
                            "label": f"|{term}, F={F}, mF={mF:+d}⟩" if mF else f"|{term}, F={F}, mF=0⟩"}
    return out


def _species(*, mass, I, g_I, E_ion_eV, defects, hf, gF, K_quad, P, lines, n_ref, C6_GHz_um6, tau_ref,
             tau_0K_ref, alpha_g, alpha_ryd, alpha_hf, dip_er, magic):
    (p12, f12, lw12, d12), (p32, f32, lw32, d32) = lines
    return {
        "mass": mass, "nuclear_spin": I, "g_I": g_I, "E_ionization": E_ion_eV * E_CHARGE,
        "quantum_defects": dict(zip("SPDF", defects)),
        "hyperfine_ground": hf, "g_F": gF, "K_quad_clock": K_quad,
        "intermediate_states": {
            p12: {"energy": f12 * _H, "linewidth": 2 * np.pi * lw12, "dipole_from_ground": d12 * _EA0, "g_J": 2 / 3},
            p32: {"energy": f32 * _H, "linewidth": 2 * np.pi * lw32, "dipole_from_ground": d32 * _EA0, "g_J": 4 / 3},
        },
        "n_ref": n_ref, "C6_ref": 2 * np.pi * C6_GHz_um6 * 1e9 * (1e-6) ** 6,
        "tau_ref": tau_ref, "tau_0K_ref": tau_0K_ref,
        "alpha_ground": alpha_g * _AU, "alpha_rydberg_ref": alpha_ryd * _AU,
        "alpha_hyperfine": {F: a * _AU for F, a in alpha_hf.items()},
        "dipole_intermediate_to_rydberg_ref": dip_er * _EA0,
        "scaling_exponents": {"C6": 11, "lifetime_0K": 3, "lifetime_BBR": 2, "polarizability": 7,
                              "dipole_to_rydberg": -1.5},
        "transitions": {f"ground_to_{p32}": f32, f"ground_to_{p12}": f12},
        "trap_wavelength": 1064e-9,
        "magic_wavelengths": {"hyperfine": {"scalar_magic_nm": magic[0],
                                            "1064nm_differential_Hz_per_mK": magic[1]},
                              "ground_rydberg": {"near_magic_nm": magic[2]}},
    }


ATOM_DB = {
    "Rb87": _species(
        mass=1.443160648e-25, I=1.5, g_I=G_I_RB87, E_ion_eV=4.177128,
        defects=(3.1311807, 2.6548849, 1.3480917, 0.0165192),
        hf=_hyperfine(1, -4.271676631815181e9, 2, 2.563005979089109e9, "5S₁/₂"),
        gF={1: -0.5, 2: +0.5}, K_quad=575.0, P="5P",
        lines=(("5P1/2", 377.107385690e12, 5.746e6, 2.99), ("5P3/2", 384.230484468e12, 6.065e6, 4.23)),
        n_ref=70, C6_GHz_um6=862.69, tau_ref=140e-6, tau_0K_ref=280e-6,
        alpha_g=687.3, alpha_ryd=-200000, alpha_hf={1: 686.1, 2: 688.5}, dip_er=0.014,
        magic=(790.0, 70e3, 1004)),
    "Cs133": _species(
        mass=2.20694657e-25, I=3.5, g_I=G_I_CS133, E_ion_eV=3.8939,
        defects=(4.0493532, 3.5915871, 2.4754562, 0.0334),
        hf=_hyperfine(3, -4.021776399375e9, 4, 5.170855370625e9, "6S₁/₂"),
        gF={3: -0.25, 4: +0.25}, K_quad=427.0, P="6P",
        lines=(("6P1/2", 335.116048807e12, 4.575e6, 3.18), ("6P3/2", 351.725718509e12, 5.234e6, 4.49)),
        n_ref=70, C6_GHz_um6=1400.0, tau_ref=160e-6, tau_0K_ref=320e-6,
        alpha_g=1000, alpha_ryd=-300000, alpha_hf={3: 998, 4: 1002}, dip_er=0.012,
        magic=(866.0, 50e3, 1064)),
}

_DEFAULT_INTERMEDIATE = {"Rb87": "5P3/2", "Cs133": "6P3/2"}


def _check(species):
    if species not in ATOM_DB:
        raise ValueError(f"Unknown species: {species}. Available: {list(ATOM_DB.keys())}")


def get_default_intermediate_state(species: str) -> str:
    if species not in _DEFAULT_INTERMEDIATE:
        raise ValueError(f"Unknown species: {species}. Available: {list(_DEFAULT_INTERMEDIATE.keys())}")
    return _DEFAULT_INTERMEDIATE[species]


def get_atom_properties(species: str) -> dict:
    _check(species)
    return ATOM_DB[species]


def get_quantum_defect(species: str, orbital: str = "S") -> float:
    return ATOM_DB[species]["quantum_defects"].get(orbital, 0)


def effective_n(n, species: str, orbital: str = "S"):
    """n* = n - δ_l."""
    return n - get_quantum_defect(species, orbital)


def _ratio(n, species):
    props = ATOM_DB[species]
    return effective_n(n, species, "S") / effective_n(props["n_ref"], species, "S")


def get_rydberg_energy(n, species: str, orbital: str = "S"):
    return -RY_JOULES / effective_n(n, species, orbital) ** 2


def get_C6(n, species: str):
    """C6(n) = C6_ref (n*/n*_ref)^11 in rad/s·m^6 (RG/atom_database.py:713-719)."""
    props = ATOM_DB[species]
    return props["C6_ref"] * _ratio(n, species) ** props["scaling_exponents"]["C6"]


def get_rydberg_lifetime(n, species: str, temperature: float = 300.0):
    """1/τ = 1/τ_0K(n*³) + 1/τ_BBR(n*², T⁻⁴)  (RG/atom_database.py:768-789)."""
    props = ATOM_DB[species]
    x = _ratio(n, species)
    ex = props["scaling_exponents"]
    tau_0K = props["tau_0K_ref"] * x ** ex["lifetime_0K"]
    if temperature < 1:
        return tau_0K
    tau_bbr = props["tau_ref"] * props["tau_0K_ref"] / (props["tau_0K_ref"] - props["tau_ref"])
    tau_bbr = tau_bbr * x ** ex["lifetime_BBR"]
    tau_bbr = tau_bbr * (300.0 / temperature) ** 4
    return 1.0 / (1.0 / tau_0K + 1.0 / tau_bbr)


def get_rydberg_polarizability(n, species: str):
    props = ATOM_DB[species]
    return props["alpha_rydberg_ref"] * _ratio(n, species) ** props["scaling_exponents"]["polarizability"]


def get_dipole_to_rydberg(n, species: str):
    props = ATOM_DB[species]
    return (props["dipole_intermediate_to_rydberg_ref"]
            * _ratio(n, species) ** props["scaling_exponents"]["dipole_to_rydberg"])


def get_intermediate_state_linewidth(species: str, intermediate_state: str = "5P3/2") -> float:
    key = intermediate_state
    if species != "Rb87" and "5P" in key:
        key = key.replace("5P", "6P")
    return ATOM_DB[species]["intermediate_states"][key]["linewidth"]


def get_hyperfine_splitting(species: str) -> float:
    if species == "Rb87":
        return 6.834682610904e9
    if species == "Cs133":
        return 9.192631770e9
    raise ValueError(f"Unknown species: {species}")


def list_available_species() -> List[str]:
    return list(ATOM_DB.keys())


def get_mass(species: str) -> float:
    return ATOM_DB[species]["mass"]


def get_ionization_energy(species: str) -> float:
    return ATOM_DB[species]["E_ionization"]


def get_ground_state_polarizability(species: str, F: Optional[int] = None) -> float:
    if F is not None and "alpha_hyperfine" in ATOM_DB[species]:
        return ATOM_DB[species]["alpha_hyperfine"][F]
    return ATOM_DB[species]["alpha_ground"]


__all__ = ["ATOM_DB", "get_atom_properties", "get_quantum_defect", "effective_n", "get_rydberg_energy",
           "get_C6", "get_rydberg_lifetime", "get_rydberg_polarizability", "get_dipole_to_rydberg",
           "get_intermediate_state_linewidth", "get_hyperfine_splitting",
           "get_ground_state_polarizability", "get_mass", "get_ionization_energy",
           "list_available_species", "get_default_intermediate_state"]
