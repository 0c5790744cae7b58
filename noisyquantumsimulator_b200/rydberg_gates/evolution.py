"""The reference's ``evolve_*`` helpers (RG/simulation.py:647-776, 1502-1760, 1795-1943, 2099-2235) on the B200.

In the reference every one of them is a Python loop of QuTiP ``mesolve`` calls over operators the caller
hands in.  The CUDA path never materialises operators: it propagates the reduced Lindblad systems of a
*parametrised* model (csrc/rydcz_systems_d*.inc).  These functions therefore

1. project the caller's Hamiltonian(s) and collapse operators onto that model (a least-squares fit to the
   constant matrices of ``codegen.model`` -- exact for anything the reference's own builders produce; an
   operator outside the model raises ``ValueError`` instead of being silently approximated),
2. build the pulse schedule the reference's loop would walk (same boundaries, same skipped segments), and
3. run ONE kernel evaluation (``rydcz_evaluate_host`` with states) and hand back the final kets / density
   matrices of the requested inputs as NumPy arrays.

The kernel starts from the computational-basis inputs |00>, |01>, |10>, |11> (kets or their projectors); any
other initial state raises ``NotImplementedError`` -- there is no CPU integrator to fall back to.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .hamiltonians import HilbertSpace, build_full_hamiltonian, build_hilbert_space, build_phase_modulated_hamiltonian

LABELS = ("00", "01", "10", "11")
_FIT_TOL = 1e-9


def _model(dim: int):
    from ..codegen import model
    return model


def _dim_of(hs, *ops) -> int:
    if hs is not None:
        return int(hs.dim)
    for op in ops:
        if op is not None:
            n = int(round(np.sqrt(np.asarray(op).shape[0])))
            if n * n == np.asarray(op).shape[0]:
                return n
    raise ValueError("cannot infer the Hilbert-space dimension: pass hs")


def hamiltonian_parameters(H, dim: int) -> Dict[str, float]:
    """(om_re, om_im, det_a, det_b, V, delta1) of a two-atom Hamiltonian built by the reference's builders
    (RG/hamiltonians.py:584-1274); ValueError if H is not of that form."""
    model = _model(dim)
    H = np.asarray(H, dtype=complex)
    if dim not in (3, 4) or H.shape != (dim * dim, dim * dim):
        raise ValueError(f"expected a {dim * dim}x{dim * dim} two-atom Hamiltonian of the 3- or 4-level model")
    terms = model._hamiltonian_terms(dim)
    names = list(terms)
    A = np.stack([terms[k].reshape(-1) for k in names], axis=1)
    A = np.concatenate([A.real, A.imag])
    b = np.concatenate([H.reshape(-1).real, H.reshape(-1).imag])
    p, *_ = np.linalg.lstsq(A, b, rcond=None)
    scale = max(np.abs(H).max(), 1e-300)
    if np.abs(A @ p - b).max() > _FIT_TOL * scale:
        raise ValueError("Hamiltonian is outside the model the CUDA path propagates "
                         "(laser coupling |1><r| per atom, detuning on |r>, shift on |1>, V|rr><rr|)")
    return dict(zip(names, map(float, p)))


def collapse_rates(c_ops: Optional[Sequence], dim: int) -> Dict[str, float]:
    """The kernel's rates (g_r1, g_r0, g_phi, g_sc[, g_mj]) from a list of collapse operators: Lindblad terms
    with the same operator add in rate, so the dissipator is linear in them.  ValueError if the list contains
    an operator outside the reference's channel set (RG/noise_models.py:1199-1446)."""
    model = _model(dim)
    names, L = model.liouvillian_terms(dim)
    rate_names = [n for n in names if n.startswith("g_")]
    if not c_ops:
        return {n: 0.0 for n in rate_names}
    n = dim * dim
    I = np.eye(n, dtype=complex)
    D = np.zeros((n * n, n * n), dtype=complex)
    for Cq in c_ops:
        Cq = np.asarray(Cq, dtype=complex)
        if Cq.shape != (n, n):
            raise ValueError(f"collapse operator of shape {Cq.shape}, expected {(n, n)}")
        CdC = Cq.conj().T @ Cq
        D += np.kron(Cq, Cq.conj()) - 0.5 * np.kron(CdC, I) - 0.5 * np.kron(I, CdC.T)
    A = np.stack([L[names.index(k)].reshape(-1) for k in rate_names], axis=1)
    A = np.concatenate([A.real, A.imag])
    b = np.concatenate([D.reshape(-1).real, D.reshape(-1).imag])
    g, *_ = np.linalg.lstsq(A, b, rcond=None)
    scale = max(np.abs(D).max(), 1e-300)
    if np.abs(A @ g - b).max() > _FIT_TOL * scale or (g < -_FIT_TOL * scale).any():
        raise ValueError("collapse operators are outside the channel set the CUDA path propagates "
                         "(|r>->|1>, |r>->|0>, dephasing on |r><r|, scattering on |1><1|, r+<->r- mixing)")
    return dict(zip(rate_names, (max(float(x), 0.0) for x in g)))


def _basis_label(psi0, dim: int) -> Tuple[str, bool]:
    """('ab', is_density_matrix) for a computational-basis input; NotImplementedError otherwise."""
    a = np.asarray(psi0, dtype=complex)
    n = dim * dim
    for lab in LABELS:
        idx = dim * int(lab[0]) + int(lab[1])
        if a.size == n:
            e = np.zeros(n, dtype=complex)
            e[idx] = 1.0
            if np.allclose(a.reshape(-1), e, rtol=0, atol=1e-14):
                return lab, False
        elif a.shape == (n, n):
            P = np.zeros((n, n), dtype=complex)
            P[idx, idx] = 1.0
            if np.allclose(a, P, rtol=0, atol=1e-14):
                return lab, True
    raise NotImplementedError(
        "the CUDA path propagates the computational-basis inputs |00>, |01>, |10>, |11> (kets or projectors); "
        "an arbitrary initial state has no kernel and there is no CPU integrator to fall back to")


def _propagate(dim: int, segments: List[Tuple[float, np.ndarray]], c_ops, device: int = 0):
    """Run one kernel evaluation over [(duration, H), ...]; returns ({label: ket or rho}, noisy)."""
    from ..engine import get_evaluator
    from .simulation import _states_to_results

    rates = collapse_rates(c_ops, dim)
    noisy = bool(c_ops)
    if not segments:
        raise ValueError("empty pulse schedule")
    ps = [hamiltonian_parameters(H, dim) for _, H in segments]
    p0 = ps[0]
    tol = 1e-12
    for p in ps[1:]:
        for k in ("V", "delta1"):
            if abs(p[k] - p0[k]) > tol * max(abs(p0[k]), 1.0):
                raise ValueError(f"'{k}' must be the same in every segment of one evolution")
    lib_stride = 12
    pt = np.zeros((1, lib_stride))
    pt[0, 0], pt[0, 1] = p0["V"], p0["delta1"]
    pt[0, 4], pt[0, 5], pt[0, 6], pt[0, 7] = rates["g_r1"], rates["g_r0"], rates["g_phi"], rates["g_sc"]
    pt[0, 8] = rates.get("g_mj", 0.0)
    seg = np.zeros((1, len(segments), 4))
    off_a = off_b = None
    for s, ((dt, _), p) in enumerate(zip(segments, ps)):
        # per-atom detunings: a common part goes into the segment row, a constant difference into the offsets
        if off_a is None:
            off_a, off_b = 0.0, p["det_b"] - p["det_a"]
        if abs((p["det_b"] - p["det_a"]) - off_b) > tol * max(abs(p["det_a"]), abs(p["det_b"]), 1.0):
            raise ValueError("the detuning difference between the atoms must be the same in every segment")
        seg[0, s] = (dt, p["om_re"], p["om_im"], p["det_a"])
    pt[0, 2], pt[0, 3] = off_a, off_b
    ev = get_evaluator(device)
    out, st = ev.evaluate_host(dim, noisy, pt, seg, np.array([len(segments)], dtype=np.int32), want_states=True)
    results, _ = _states_to_results(dim, noisy, st[0])
    return results, noisy


def _shape_like(x, template, noisy):
    """Kets come back in the caller's ket shape (column or flat); density matrices as (n, n)."""
    t = np.asarray(template)
    if noisy or np.asarray(x).ndim == 2:
        return x
    return x.reshape(t.shape) if t.size == x.size else x


def _select(results, noisy, initial_states):
    final = {}
    for label, psi0 in initial_states.items():
        lab, _ = _basis_label(psi0, int(round(np.sqrt(np.asarray(results["00"]).shape[0]))))
        final[label] = _shape_like(results[lab], psi0, noisy)
    return final


def evolve_state(H, psi0, tlist, c_ops=None, options=None):
    """Final state of ``mesolve(H, psi0, tlist, c_ops)`` (RG/simulation.py:647-690) for a constant model
    Hamiltonian and a computational-basis input: one exact segment of length ``tlist[-1] - tlist[0]`` (the
    solver tolerances in ``options`` do not apply -- the propagation is exact to rounding)."""
    dim = _dim_of(None, H)
    lab, _ = _basis_label(psi0, dim)
    t = np.asarray(tlist, dtype=float).reshape(-1)
    results, noisy = _propagate(dim, [(float(t[-1] - t[0]), H)], c_ops or [])
    return _shape_like(results[lab], psi0, noisy)


def evolve_two_pulse(H1, H2, psi0, tau_single: float, c_ops=None, n_steps: int = 100):
    """RG/simulation.py:693-737: pulse 1 with H1 for tau, pulse 2 with H2 for tau."""
    dim = _dim_of(None, H1)
    lab, _ = _basis_label(psi0, dim)
    results, noisy = _propagate(dim, [(float(tau_single), H1), (float(tau_single), H2)], c_ops or [])
    return _shape_like(results[lab], psi0, noisy)


def evolve_two_pulse_lp(initial_states: Dict[str, np.ndarray], H1, H2, tau_single: float, c_ops=None):
    """RG/simulation.py:740-776: every input through both Levine-Pichler pulses -- one kernel evaluation."""
    dim = _dim_of(None, H1)
    results, noisy = _propagate(dim, [(float(tau_single), H1), (float(tau_single), H2)], c_ops or [])
    return _select(results, noisy, initial_states)


def evolve_bangbang_jp(initial_states, Omega: float, V: float, omega_tau: float, switching_times, phases,
                       hs: HilbertSpace, c_ops=None, delta_zeeman: float = 0.0, delta_stark: float = 0.0,
                       trap_laser_on: bool = True, n_steps_per_segment: int = 50, verbose: bool = False):
    """RG/simulation.py:1795-1943: boundaries [0] + switching_times + [omega_tau] divided by Omega, one
    constant Hamiltonian H(Omega e^{i phi_k}, Delta = 0) per segment, segments shorter than 1e-18 s skipped."""
    assert len(phases) == len(switching_times) + 1, (
        f"Need len(phases) = len(switching_times) + 1, got {len(phases)} phases and {len(switching_times)} switching times")
    bounds = [b / Omega for b in [0.0] + list(switching_times) + [omega_tau]]
    segments = []
    for k, phase in enumerate(phases):
        dt = bounds[k + 1] - bounds[k]
        if dt < 1e-18:
            continue
        segments.append((dt, build_phase_modulated_hamiltonian(Omega=Omega, V=V, phase=phase, hs=hs, Delta=0.0,
                                                               delta_zeeman=delta_zeeman, delta_stark=delta_stark,
                                                               trap_laser_on=trap_laser_on)))
    results, noisy = _propagate(hs.dim, segments, c_ops or [])
    info = {"protocol": "bangbang_jp", "omega_tau": omega_tau, "omega_tau_over_pi": omega_tau / np.pi,
            "tau_us": omega_tau / Omega * 1e6, "V_over_Omega": V / Omega, "n_segments": len(phases),
            "switching_times": list(switching_times), "phases": list(phases),
            "phases_over_pi": [p / np.pi for p in phases]}
    return _select(results, noisy, initial_states), info


def evolve_smooth_sinusoidal_jp(initial_states, Omega: float, V: float, tau_total: float, hs: HilbertSpace, c_ops=None,
                                A: float = None, omega_mod: float = None, phi_offset: float = 0.0,
                                delta_over_omega: float = None, n_steps: int = 200, delta_zeeman: float = 0.0,
                                delta_stark: float = 0.0, trap_laser_on: bool = True,
                                intermediate_detuning_sign: str = "positive", verbose: bool = False):
    """RG/simulation.py:1502-1760: ``n_steps`` equal steps, phase A cos(omega_mod t_mid - phi_offset) sampled at
    the midpoints, static detuning Delta = delta_over_omega * Omega (the repeated-segment path of the kernel)."""
    A = 0.311 * np.pi if A is None else A
    omega_mod = 1.242 * Omega if omega_mod is None else omega_mod
    delta_over_omega = 0.0205 if delta_over_omega is None else delta_over_omega
    Delta = delta_over_omega * Omega
    valid = delta_over_omega < 0 if intermediate_detuning_sign == "positive" else delta_over_omega > 0
    if not valid and delta_over_omega != 0:
        want = "negative" if intermediate_detuning_sign == "positive" else "positive"
        warnings.warn(f"Dark state condition violated! With {intermediate_detuning_sign} intermediate-state detuning, "
                      f"two-photon detuning should be {want} but got delta_over_omega = {delta_over_omega:+.4f}. "
                      f"This will increase scattering error.", UserWarning)
    if V / Omega < 5:
        warnings.warn(f"V/Ω = {V / Omega:.1f} may be too weak for reliable CZ operation. "
                      f"Recommend V/Ω > 10 for high-fidelity gates.", UserWarning)
    tl = np.linspace(0, tau_total, n_steps + 1)
    dt = tau_total / n_steps
    phase_evolution, segments = [], []
    for i in range(n_steps):
        t_mid = 0.5 * (tl[i] + tl[i + 1])
        phi = A * np.cos(omega_mod * t_mid - phi_offset)
        phase_evolution.append(phi)
        segments.append((dt, build_phase_modulated_hamiltonian(Omega=Omega, V=V, phase=phi, hs=hs, Delta=Delta,
                                                               delta_zeeman=delta_zeeman, delta_stark=delta_stark,
                                                               trap_laser_on=trap_laser_on)))
    results, noisy = _propagate(hs.dim, segments, c_ops or [])
    info = {"protocol": "smooth_sinusoidal_jp", "omega_tau": Omega * tau_total, "tau_us": tau_total * 1e6,
            "V_over_Omega": V / Omega, "A": A, "omega_mod": omega_mod, "omega_mod_ratio": omega_mod / Omega,
            "phi_offset": phi_offset, "delta_over_omega": delta_over_omega, "n_steps": n_steps,
            "dark_state_valid": bool(valid), "phase_evolution": np.array(phase_evolution), "tlist": tl}
    return _select(results, noisy, initial_states), info


def evolve_shaped_pulse(initial_states, pulse_shape: str, Omega: float, Delta: float, V: float, xi: complex,
                        tau_single: float, hs: HilbertSpace, c_ops=None, delta_zeeman: float = 0.0,
                        delta_stark: float = 0.0, trap_laser_on: bool = True, drag_lambda: float = None,
                        n_time_steps: int = 500, verbose: bool = False):
    """RG/simulation.py:2099-2235: two pulses of ``n_time_steps - 1`` midpoint-sampled steps each; the envelope is
    evaluated on a *scalar* time like the reference does (so max-normalised shapes degenerate to 1)."""
    from .pulse_shaping import area_correction_factor, get_pulse_envelope

    correction = area_correction_factor(pulse_shape, tau_single)
    peak = Omega * correction
    t_pulse = np.linspace(0, tau_single, n_time_steps)
    dt = tau_single / n_time_steps
    segments = []
    for factor in (1.0, xi):
        for i in range(n_time_steps - 1):
            t_mid = 0.5 * (t_pulse[i] + t_pulse[i + 1])
            env = get_pulse_envelope(pulse_shape, t_mid, tau_single)
            segments.append((dt, build_full_hamiltonian(peak * env * factor, Delta, V, hs, delta_stark=delta_stark,
                                                        delta_zeeman=delta_zeeman, trap_laser_on=trap_laser_on)))
    results, noisy = _propagate(hs.dim, segments, c_ops or [])
    info = {"pulse_shape": pulse_shape, "area_correction": correction, "peak_scaling": correction,
            "drag_lambda": drag_lambda}
    return _select(results, noisy, initial_states), info


__all__ = ["hamiltonian_parameters", "collapse_rates", "evolve_state", "evolve_two_pulse", "evolve_two_pulse_lp",
           "evolve_bangbang_jp", "evolve_smooth_sinusoidal_jp", "evolve_shaped_pulse"]
