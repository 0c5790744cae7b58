"""NumPy versions of the reference's Hamiltonian builders (RG/hamiltonians.py:424-1274).

On the B200 path the Hamiltonian is never materialised: the kernel assembles the Liouvillian
from (Re Ω, Im Ω, Δ, V, δ1) directly (csrc/rydcz_systems_d*.inc).  These builders exist so that
callers who read ``result.H1`` / ``result.H2`` or build operators themselves keep working; they
return plain ``ndarray`` (d²×d²) instead of QuTiP ``Qobj``.  Basis index = d*a + b.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple

import numpy as np

from .constants import HBAR, MU_B


@dataclass
class HilbertSpace:
    dim: int
    basis: Dict[str, np.ndarray] = field(default_factory=dict)
    projectors: Dict[str, np.ndarray] = field(default_factory=dict)
    transitions: Dict[str, np.ndarray] = field(default_factory=dict)
    identity: np.ndarray = None


def build_hilbert_space(dim: int = 3) -> HilbertSpace:
    if dim not in (3, 4, 5):
        raise ValueError(f"Unsupported Hilbert space dimension: {dim}. Use 3, 4 or 5.")
    # dim 5 (extension): the intermediate state |e> of the two-photon excitation and a loss level |l>
    names = {3: ["0", "1", "r"], 4: ["0", "1", "r+", "r-"], 5: ["0", "1", "e", "r", "l"]}[dim]
    hs = HilbertSpace(dim=dim)
    for i, nm in enumerate(names):
        v = np.zeros((dim, 1), dtype=complex)
        v[i] = 1.0
        hs.basis[nm] = v
        hs.projectors[nm] = v @ v.conj().T
    for frm in names:
        for to in names:
            if frm != to and (frm.startswith("r") or frm in ("1", "e")):
                hs.transitions[f"{frm}->{to}"] = hs.basis[to] @ hs.basis[frm].conj().T
    hs.identity = np.eye(dim, dtype=complex)
    return hs


HS3 = build_hilbert_space(3)
HS4 = build_hilbert_space(4)
b0, b1, br = HS3.basis["0"], HS3.basis["1"], HS3.basis["r"]


def op_two_atom(op1, op2):
    return np.kron(op1, op2)


def _both(op, hs):
    return op_two_atom(op, hs.identity) + op_two_atom(hs.identity, op)


def build_laser_hamiltonian(Omega: complex, hs: HilbertSpace, polarization: str = "sigma+",
                            Omega_minus: float = None):
    """(Ω/2)|r><1| + h.c. on each atom; for d=4 the polarization picks r+ / r- (RG/hamiltonians.py:642-681)."""
    def arm(Om, key):
        s = hs.transitions[key]
        return 0.5 * (Om * s + np.conj(Om) * s.conj().T)

    if hs.dim == 3:
        return _both(arm(Omega, "1->r"), hs)
    zero = 0.0 * hs.identity
    if polarization == "sigma+":
        plus, minus = arm(Omega, "1->r+"), zero
    elif polarization == "sigma-":
        plus, minus = zero, arm(Omega, "1->r-")
    elif polarization == "pi":
        plus, minus = arm(Omega / np.sqrt(2), "1->r+"), arm(Omega / np.sqrt(2), "1->r-")
    else:
        plus = arm(Omega, "1->r+")
        minus = arm(Omega_minus if Omega_minus is not None else 0.02 * np.abs(Omega), "1->r-")
    return _both(plus + minus, hs)


def build_detuning_hamiltonian(Delta: float, hs: HilbertSpace, zeeman_splitting: float = 0,
                               Delta_minus: float = None):
    if hs.dim == 3:
        return -Delta * _both(hs.projectors["r"], hs)
    Dm = Delta_minus if Delta_minus is not None else Delta + zeeman_splitting
    return -Delta * _both(hs.projectors["r+"], hs) - Dm * _both(hs.projectors["r-"], hs)


def build_interaction_hamiltonian(V: float, hs: HilbertSpace, V_pm: float = None, V_mm: float = None):
    if hs.dim == 3:
        P = hs.projectors["r"]
        return V * op_two_atom(P, P)
    Pp, Pm = hs.projectors["r+"], hs.projectors["r-"]
    Vx = V_pm if V_pm is not None else V
    Vm = V_mm if V_mm is not None else V
    return (V * op_two_atom(Pp, Pp) + Vx * (op_two_atom(Pp, Pm) + op_two_atom(Pm, Pp))
            + Vm * op_two_atom(Pm, Pm))


def build_qubit_hamiltonian(omega_qubit: float, hs: HilbertSpace):
    return omega_qubit * _both(hs.projectors["1"], hs)


def build_zeeman_hamiltonian(delta_zeeman: float, hs: HilbertSpace):
    return delta_zeeman * _both(hs.projectors["1"], hs)


def build_stark_hamiltonian(delta_stark: float, hs: HilbertSpace, trap_laser_on: bool = True):
    if not trap_laser_on:
        return 0 * op_two_atom(hs.identity, hs.identity)
    return delta_stark * _both(hs.projectors["1"], hs)


def check_rwa_validity(Omega: float, omega_optical: float = 2 * np.pi * 384e12, threshold: float = 0.01,
                       verbose: bool = True) -> bool:
    ratio = abs(Omega) / omega_optical
    ok = ratio < threshold
    if verbose and not ok:
        print(f"⚠️  RWA WARNING: Ω/ω_optical = {ratio:.2e} > {threshold}")
    return ok


def build_full_hamiltonian(Omega: complex, Delta: float, V: float, hs: HilbertSpace = None, dim: int = 3,
                           polarization: str = "sigma+", zeeman_splitting: float = 0, omega_qubit: float = 0.0,
                           delta_zeeman: float = 0.0, delta_stark: float = 0.0, trap_laser_on: bool = True,
                           check_rwa: bool = False, **kwargs):
    if hs is None:
        hs = HS3 if dim == 3 else HS4
    if check_rwa and abs(Omega) > 0:
        check_rwa_validity(abs(Omega), verbose=True)
    H = (build_laser_hamiltonian(Omega, hs, polarization=polarization, Omega_minus=kwargs.get("Omega_minus"))
         + build_detuning_hamiltonian(Delta, hs, zeeman_splitting=zeeman_splitting,
                                      Delta_minus=kwargs.get("Delta_minus"))
         + build_interaction_hamiltonian(V, hs, V_pm=kwargs.get("V_pm"), V_mm=kwargs.get("V_mm")))
    if omega_qubit != 0:
        H = H + build_qubit_hamiltonian(omega_qubit, hs)
    if delta_zeeman != 0:
        H = H + build_zeeman_hamiltonian(delta_zeeman, hs)
    if delta_stark != 0:
        H = H + build_stark_hamiltonian(delta_stark, hs, trap_laser_on)
    return H


def build_five_level_hamiltonian(Omega: complex, Delta: float, V: float, Omega_1: float, Delta_e: float,
                                 hs: HilbertSpace = None, delta_zeeman: float = 0.0, delta_stark: float = 0.0,
                                 trap_laser_on: bool = True, light_shift_compensation: bool = True):
    """Two-atom Hamiltonian of the 5-level model (extension: no reference counterpart; the reference
    eliminates |e> and keeps only Ω = Ω1Ω2/2Δe, RG/laser_physics.py:265-359).  ``Omega`` is the
    effective two-photon Rabi frequency (complex: carries the pulse phase); the second leg is
    Ω2 = Ω·2Δe/Ω1.  With ``light_shift_compensation`` the light shifts Ω1²/4Δe (on |1>) and |Ω2|²/4Δe
    (on |r>) are cancelled, so that Δe -> ∞ reproduces ``build_full_hamiltonian``."""
    if hs is None:
        hs = build_hilbert_space(5)
    om2 = Omega * 2.0 * Delta_e / Omega_1
    ls = (1.0 if light_shift_compensation else 0.0) / (4.0 * Delta_e)
    s1, s2 = hs.transitions["1->e"], hs.transitions["e->r"]
    h = 0.5 * Omega_1 * (s1 + s1.conj().T) + 0.5 * (om2 * s2 + np.conj(om2) * s2.conj().T)
    h = h - Delta_e * hs.projectors["e"] - (Delta + ls * abs(om2) ** 2) * hs.projectors["r"]
    d1 = delta_zeeman + (delta_stark if trap_laser_on else 0.0) - ls * Omega_1 ** 2
    h = h + d1 * hs.projectors["1"]
    return _both(h, hs) + V * op_two_atom(hs.projectors["r"], hs.projectors["r"])


def build_phase_modulated_hamiltonian(Omega: float, phase: float, V: float, hs: HilbertSpace, Delta: float = 0.0,
                                      delta_zeeman: float = 0.0, delta_stark: float = 0.0,
                                      trap_laser_on: bool = True):
    H = build_laser_hamiltonian(Omega * np.exp(1j * phase), hs) + build_interaction_hamiltonian(V, hs)
    if Delta != 0:
        H = H + build_detuning_hamiltonian(Delta, hs)
    if delta_zeeman != 0:
        H = H + build_zeeman_hamiltonian(delta_zeeman, hs)
    if delta_stark != 0:
        H = H + build_stark_hamiltonian(delta_stark, hs, trap_laser_on)
    return H


def compute_mJ_zeeman_splitting(B_field: float, g_J: float = 2.002) -> float:
    return g_J * MU_B * B_field / HBAR


def compute_mJ_coupling_ratio(polarization: str) -> Tuple[float, float]:
    if polarization == "sigma+":
        return (1.0, 0.0)
    if polarization == "sigma-":
        return (0.0, 1.0)
    if polarization == "pi":
        return (1 / np.sqrt(2), 1 / np.sqrt(2))
    return (np.sqrt(1 - 0.02), np.sqrt(0.02))


__all__ = ["HilbertSpace", "build_hilbert_space", "HS3", "HS4", "b0", "b1", "br", "op_two_atom",
           "build_laser_hamiltonian", "build_detuning_hamiltonian", "build_interaction_hamiltonian",
           "build_qubit_hamiltonian", "build_zeeman_hamiltonian", "build_stark_hamiltonian",
           "build_full_hamiltonian", "build_five_level_hamiltonian", "build_phase_modulated_hamiltonian", "check_rwa_validity",
           "compute_mJ_zeeman_splitting", "compute_mJ_coupling_ratio"]
