"""CPU restatement (NumPy/SciPy) of the reference's Rydberg-CZ gate evaluation.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

What is restated
----------------
The reference evaluates a gate by building QuTiP operators and calling
``mesolve`` once per (input state, pulse segment).  QuTiP (``qutip>=5.0.0``,
un-pinned, ``/root/reference/pyproject.toml:38``) is not installable in this
environment, so this module restates the arithmetic QuTiP performs for the
reference's call sites.  ``RG/`` abbreviates
``src/qpu_simulator/micro_physics/neutral_atoms/rydberg_gates/`` of the
reference tree.

* Hamiltonian assembly        -> ``hamiltonian``          (RG/hamiltonians.py:584-989, 1057-1274)
* collapse operators          -> ``collapse_operators``   (RG/noise_models.py:1199-1620)
* ``mesolve`` of a constant H -> ``propagate``            (RG/simulation.py:647-690)
    - mode "exact": the mathematical limit of mesolve, a matrix exponential
      (``scipy.linalg.expm``) of -iH or of the vectorised Liouvillian;
    - mode "zvode": the integrator QuTiP's default method wraps
      (``scipy.integrate.ode('zvode', method='adams')`` with the reference's
      ``atol=1e-10, rtol=1e-8, nsteps``) stepped over the reference's tlist.
* segment schedules           -> ``segments_*``           (RG/simulation.py:693-776, 1502-1760, 1795-1943, 2099-2231)
* fidelity epilogue           -> ``cz_fidelity_pure`` / ``cz_fidelity_mixed``
                                                          (RG/simulation.py:186-222, 225-633)

Pinning: ``tests/test_oracle_golden.py`` checks this module against the twelve
saved QuTiP outputs of the reference's own notebook
(``RG/cz_gate_optimization_demo.ipynb`` cells 6, 10, 18; listed in
SURVEY.md section 8c / BASELINE.md section 2).  Those are all noiseless runs; the noisy
(``include_noise=True``) branch has no saved reference output anywhere, so for
that branch parity is UNPINNED (ρ entries and populations are compared against
this oracle only).

Conventions: two-atom basis index = d*a + b (QuTiP ``tensor(a, b)``), levels
0,1,r (d=3) or 0,1,r+,r- (d=4); vec(ρ) is row-major, so A ρ B <-> (A ⊗ B^T) vec(ρ).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as sla
from scipy.integrate import ode

LABELS = ("00", "01", "10", "11")


# --------------------------------------------------------------------------
# Operators
# --------------------------------------------------------------------------
def _ket(d: int, i: int) -> np.ndarray:
    v = np.zeros(d, dtype=complex)
    v[i] = 1.0
    return v


def _op(d: int, i: int, j: int) -> np.ndarray:
    """|i><j| on one atom."""
    m = np.zeros((d, d), dtype=complex)
    m[i, j] = 1.0
    return m


def _two(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.kron(a, b)


def hamiltonian(
    dim: int,
    Omega: complex,
    Delta: float,
    V: float,
    delta_1: float = 0.0,
    detuning_offsets: Tuple[float, float] = (0.0, 0.0),
) -> np.ndarray:
    """Two-atom rotating-frame Hamiltonian (rad/s).

    H = sum_atoms [ (Ω/2)|r><1| + (Ω*/2)|1><r| - Δ|r><r| + δ1 |1><1| ] + V |rr><rr|

    RG/hamiltonians.py:650-653 (laser, d=3), :661-663 and :678-679 (d=4, default
    ``sigma+`` couples only r+), :739 / :747-753 (detuning; r- gets the same Δ
    because ``zeeman_splitting=0`` at the call sites RG/simulation.py:3180-3205),
    :833 / :841-853 (V on every rr pair for d=4), :931-933 and :987-989
    (δ_zeeman + δ_stark on |1>).  ``detuning_offsets`` is the extension used for
    Monte-Carlo Doppler shots: atom k sees Δ + offset_k.
    """
    I = np.eye(dim, dtype=complex)
    ryd = [2] if dim == 3 else [2, 3]
    h_laser = 0.5 * (Omega * _op(dim, 2, 1) + np.conj(Omega) * _op(dim, 1, 2))
    H = _two(h_laser, I) + _two(I, h_laser)
    for atom, off in enumerate(detuning_offsets):
        for r in ryd:
            P = _op(dim, r, r)
            H = H - (Delta + off) * (_two(P, I) if atom == 0 else _two(I, P))
    for ra in ryd:
        for rb in ryd:
            H = H + V * _two(_op(dim, ra, ra), _op(dim, rb, rb))
    P1 = _op(dim, 1, 1)
    H = H + delta_1 * (_two(P1, I) + _two(I, P1))
    return H


@dataclass
class NoiseRates:
    """The ten rates ``build_all_noise_operators`` receives (RG/simulation.py:3322-3334)."""

    gamma_r: float = 0.0
    gamma_bbr: float = 0.0
    gamma_phi_laser: float = 0.0
    gamma_phi_thermal: float = 0.0
    gamma_phi_zeeman: float = 0.0
    gamma_loss_antitrap: float = 0.0
    gamma_loss_background: float = 0.0
    gamma_scatter_intermediate: float = 0.0
    gamma_leakage: float = 0.0
    mJ_leakage_rate: float = 0.0
    branching_1: float = 0.5

    def collapsed(self) -> Tuple[float, float, float, float, float]:
        """(Γ_r→1, Γ_r→0, γ_φ on |r><r|, γ_sc on |1><1|, γ_mJ): Lindblad terms with the
        same operator add in rate ('collapsed form')."""
        g_r1 = self.gamma_r * self.branching_1
        g_r0 = (self.gamma_r * (1.0 - self.branching_1) + self.gamma_bbr
                + self.gamma_loss_antitrap + self.gamma_loss_background + self.gamma_leakage)
        g_phi = self.gamma_phi_laser + self.gamma_phi_thermal + self.gamma_phi_zeeman
        return g_r1, g_r0, g_phi, self.gamma_scatter_intermediate, self.mJ_leakage_rate


def collapse_operators(dim: int, rates: NoiseRates) -> List[np.ndarray]:
    """Collapse operators in the reference's order and with its ``> 0`` guards.

    RG/noise_models.py:1576 decay (:1233-1251 d=3, :1255-1295 d=4), :1579-1580
    dephasing (:1324-1354), :1583/:1586/:1592 three 'rydberg' loss channels
    (:1380-1406), :1589 scatter (:1435-1444).
    """
    I = np.eye(dim, dtype=complex)
    ryd = [2] if dim == 3 else [2, 3]
    ops: List[np.ndarray] = []

    def both(a: np.ndarray, rate: float) -> None:
        s = math.sqrt(rate)
        ops.append(s * _two(a, I))
        ops.append(s * _two(I, a))

    g = rates
    if g.gamma_r > 0:
        for r in ryd:
            both(_op(dim, 1, r), g.gamma_r * g.branching_1)
            both(_op(dim, 0, r), g.gamma_r * (1.0 - g.branching_1))
    if g.gamma_bbr > 0:
        for r in ryd:
            both(_op(dim, 0, r), g.gamma_bbr)
    if dim == 4 and g.mJ_leakage_rate > 0:
        both(_op(dim, 3, 2), g.mJ_leakage_rate)
        both(_op(dim, 2, 3), g.mJ_leakage_rate)
    g_phi = g.gamma_phi_laser + g.gamma_phi_thermal + g.gamma_phi_zeeman
    if g_phi > 0:
        for r in ryd:
            both(_op(dim, r, r), g_phi)
    for rate in (g.gamma_loss_antitrap, g.gamma_loss_background):
        if rate > 0:
            for r in ryd:
                both(_op(dim, 0, r), rate)
    if g.gamma_scatter_intermediate > 0:
        both(_op(dim, 1, 1), g.gamma_scatter_intermediate)
    if g.gamma_leakage > 0:
        for r in ryd:
            both(_op(dim, 0, r), g.gamma_leakage)
    return ops


def liouvillian(H: np.ndarray, c_ops: Sequence[np.ndarray]) -> np.ndarray:
    """Row-major-vec Lindblad generator: what QuTiP's ``liouvillian(H, c_ops)`` builds
    for ``mesolve`` (RG/simulation.py:689), up to QuTiP's column-major convention."""
    n = H.shape[0]
    I = np.eye(n, dtype=complex)
    L = -1j * (np.kron(H, I) - np.kron(I, H.T))
    for C in c_ops:
        CdC = C.conj().T @ C
        L = L + np.kron(C, C.conj()) - 0.5 * np.kron(CdC, I) - 0.5 * np.kron(I, CdC.T)
    return L


# --------------------------------------------------------------------------
# Segment schedules: (duration, complex Ω, Δ, n_tlist) per constant-H piece
# --------------------------------------------------------------------------
@dataclass
class Segment:
    dt: float
    Omega: complex
    Delta: float
    n_tlist: int = 2  # number of tlist points the reference hands to mesolve


def compute_xi(Delta: float, Omega: float, tau: float) -> complex:
    """LP phase jump (RG/protocols.py:798-819)."""
    if abs(Omega) < 1e-10:
        return 1.0 + 0j
    y = Delta / abs(Omega)
    s = abs(Omega) * tau
    a = math.sqrt(y * y + 1.0)
    b = s * a / 2.0
    num = a * math.cos(b) + 1j * y * math.sin(b)
    den = -a * math.cos(b) + 1j * y * math.sin(b)
    if abs(den) < 1e-12:
        return 1.0 + 0j
    return num / den


def segments_lp(Omega: float, Delta: float, tau_single: float) -> List[Segment]:
    """Square-pulse LP: H1 for τ, H2 = H(Ω·ξ) for τ; tlist = linspace(0, τ, 100)
    (RG/simulation.py:728-735, 3178-3205)."""
    xi = compute_xi(Delta, Omega, tau_single)
    return [Segment(tau_single, complex(Omega), Delta, 100),
            Segment(tau_single, Omega * xi, Delta, 100)]


def segments_bangbang(Omega: float, omega_tau: float, switching_times: Sequence[float],
                      phases: Sequence[float]) -> List[Segment]:
    """JP bang-bang (RG/simulation.py:1859-1924): boundaries [0]+switching+[Ωτ] over Ω;
    segments with dt < 1e-18 skipped; Δ = 0; tlist has 51 points."""
    assert len(phases) == len(switching_times) + 1, (
        f"Need len(phases) = len(switching_times) + 1, got {len(phases)} phases "
        f"and {len(switching_times)} switching times")
    bounds = [b / Omega for b in [0.0] + list(switching_times) + [omega_tau]]
    segs = []
    for k, ph in enumerate(phases):
        dt = bounds[k + 1] - bounds[k]
        if dt < 1e-18:
            continue
        segs.append(Segment(dt, Omega * np.exp(1j * ph), 0.0, 51))
    return segs


def segments_smooth_jp(Omega: float, tau_total: float, A: float, omega_mod: float,
                       phi_offset: float, Delta: float, n_steps: int = 300) -> List[Segment]:
    """Smooth sinusoidal JP (RG/simulation.py:1693-1733): n_steps equal pieces, phase
    sampled at each midpoint, static Δ; tlist = [0, dt]."""
    tl = np.linspace(0, tau_total, n_steps + 1)
    dt = tau_total / n_steps
    segs = []
    for i in range(n_steps):
        t_mid = tl[i] + dt / 2
        ph = A * np.cos(omega_mod * t_mid - phi_offset)
        segs.append(Segment(dt, Omega * np.exp(1j * ph), Delta, 2))
    return segs


def segments_shaped_lp(Omega_peak_env, Delta: float, tau_single: float, xi: complex,
                       n_time_steps: int = 500) -> List[Segment]:
    """Shaped LP (RG/simulation.py:2179-2220).  ``Omega_peak_env(t_mid)`` must return
    Ω_peak·envelope(t_mid) exactly as the reference evaluates it on a *scalar*
    t_mid (which is what makes gaussian/blackman degenerate to 1.0)."""
    t_pulse = np.linspace(0, tau_single, n_time_steps)
    dt = tau_single / n_time_steps
    mids = [(t_pulse[i] + t_pulse[i + 1]) / 2 for i in range(n_time_steps - 1)]
    segs = [Segment(dt, complex(Omega_peak_env(t)), Delta, 2) for t in mids]
    segs += [Segment(dt, complex(Omega_peak_env(t)) * xi, Delta, 2) for t in mids]
    return segs


# --------------------------------------------------------------------------
# Propagation
# --------------------------------------------------------------------------
def _zvode_advance(gen: np.ndarray, y0: np.ndarray, dt: float, n_tlist: int,
                   nsteps: int) -> np.ndarray:
    """dy/dt = gen @ y over linspace(0, dt, n_tlist) with zvode/Adams at the
    reference's tolerances (RG/simulation.py:687)."""
    import scipy.sparse as sp

    G = sp.csr_matrix(gen)
    solver = ode(lambda t, y: G @ y)
    solver.set_integrator("zvode", method="adams", atol=1e-10, rtol=1e-8, nsteps=nsteps)
    solver.set_initial_value(y0.astype(complex), 0.0)
    for t in np.linspace(0.0, dt, n_tlist)[1:]:
        solver.integrate(t)
        if not solver.successful():
            raise RuntimeError("zvode failed")
    return solver.y


def initial_kets(dim: int) -> Dict[str, np.ndarray]:
    """|00>,|01>,|10>,|11> (RG/simulation.py:3364-3377)."""
    out = {}
    for lab in LABELS:
        a, b = int(lab[0]), int(lab[1])
        out[lab] = np.kron(_ket(dim, a), _ket(dim, b))
    return out


def propagate(
    dim: int,
    segments: Sequence[Segment],
    V: float,
    delta_1: float,
    rates: Optional[NoiseRates],
    mode: str = "exact",
    detuning_offsets: Tuple[float, float] = (0.0, 0.0),
    nsteps: int = 50000,
    initial: Optional[Dict[str, np.ndarray]] = None,
) -> Dict[str, np.ndarray]:
    """Evolve the four computational inputs through the segment list.

    ``rates is None`` or no collapse operator survives the ``> 0`` guards -> kets
    (QuTiP's Schrödinger path for ``c_ops=[]``); otherwise density matrices.
    ``initial`` may override the start operators (used for the coherence
    probes |ab><00| of the gauge-free epilogue, which QuTiP never evolves).
    """
    c_proto = collapse_operators(dim, rates) if rates is not None else []
    pure = len(c_proto) == 0 and initial is None
    kets = initial_kets(dim)
    if initial is not None:
        states = {k: np.asarray(v, dtype=complex).reshape(-1) for k, v in initial.items()}
    elif pure:
        states = {k: v.copy() for k, v in kets.items()}
    else:
        states = {k: np.outer(v, v.conj()).reshape(-1) for k, v in kets.items()}
    for seg in segments:
        H = hamiltonian(dim, seg.Omega, seg.Delta, V, delta_1, detuning_offsets)
        gen = -1j * H if pure else liouvillian(H, c_proto)
        if mode == "exact":
            U = sla.expm(gen * seg.dt)
            states = {k: U @ v for k, v in states.items()}
        elif mode == "zvode":
            states = {k: _zvode_advance(gen, v, seg.dt, seg.n_tlist, nsteps)
                      for k, v in states.items()}
        else:
            raise ValueError(mode)
    if pure:
        return states
    n = dim * dim
    return {k: v.reshape(n, n) for k, v in states.items()}


# --------------------------------------------------------------------------
# Fidelity epilogue
# --------------------------------------------------------------------------
def _wrap_pi(x: float) -> float:
    return (x + math.pi) % (2 * math.pi) - math.pi


def _finish(fid: Dict[str, float], phase_info: Dict) -> Tuple[Dict[str, float], float, Dict]:
    """RG/simulation.py:619-633: F11 is multiplied by the CZ-phase fidelity, then mean."""
    czf = phase_info.get("cz_phase_fidelity", 1.0)
    phase_info["F11_population"] = fid["11"]
    phase_info["F11_with_phase"] = fid["11"] * czf
    phase_info["cz_phase_condition_met"] = phase_info.get("phase_error_from_pi_rad", 0) < 0.2
    fid = dict(fid)
    fid["11"] = fid["11"] * czf
    return fid, float(np.mean([fid[k] for k in LABELS])), phase_info


def _phase_block(phis: Dict[str, float]) -> Dict:
    ctrl = _wrap_pi(phis["11"] - phis["01"] - phis["10"] + phis["00"])
    err = min(abs(ctrl - math.pi), abs(ctrl + math.pi))
    return {
        "controlled_phase_rad": ctrl,
        "controlled_phase_deg": math.degrees(ctrl),
        "phase_error_from_pi_rad": err,
        "phase_error_from_pi_deg": math.degrees(err),
        "cz_phase_fidelity": math.cos(err / 2) ** 2,
    }


def cz_fidelity_pure(dim: int, finals: Dict[str, np.ndarray]):
    """Pure-state branch of ``compute_CZ_fidelity`` (RG/simulation.py:482-561, 594-617)."""
    kets = initial_kets(dim)
    ov = {k: complex(np.vdot(kets[k], finals[k])) for k in LABELS}
    fid = {k: abs(ov[k]) ** 2 for k in LABELS}
    phis = {k: float(np.angle(ov[k])) for k in LABELS}
    info = {
        "phi_01_rad": phis["01"], "phi_01_deg": math.degrees(phis["01"]),
        "phi_11_rad": float(np.angle(-ov["11"])), "phi_11_plus_rad": phis["11"],
        "amp_01": abs(ov["01"]), "amp_11": abs(ov["11"]),
        "is_mixed_state": False,
    }
    info.update(_phase_block(phis))
    return _finish(fid, info)


def dominant_phase(rho: np.ndarray, ket: np.ndarray) -> float:
    """RG/simulation.py:424-434: phase of <ab|v_max>, v_max the eigenvector of the
    largest eigenvalue.  QuTiP's dense Hermitian ``eigenstates`` calls
    ``scipy.linalg.eigh``; the eigenvector's global phase is a LAPACK convention,
    so this number is gauge dependent (hard part 1)."""
    w, v = sla.eigh(rho)
    vec = v[:, int(np.argmax(w))]
    return float(np.angle(np.vdot(ket, vec)))


def cz_fidelity_mixed(dim: int, finals: Dict[str, np.ndarray], phase_mode: str = "compat",
                      coherences: Optional[Dict[str, complex]] = None):
    """Mixed-state branch (RG/simulation.py:384-481).  F_ab = <ab|ρ_ab|ab> (QuTiP
    ``fidelity(ρ, |t><t|)**2``, :186-222).  ``phase_mode``:
      "compat"    -- the reference's dominant-eigenvector phases (gauge dependent);
      "coherence" -- gauge-free: φ_ab = arg of the [ab,00] element of E(|ab><00|),
                     supplied in ``coherences`` (reduces to the pure branch as noise -> 0).
    """
    kets = initial_kets(dim)
    fid = {k: float(np.real(np.vdot(kets[k], finals[k] @ kets[k]))) for k in LABELS}
    if phase_mode == "compat":
        phis = {k: dominant_phase(finals[k], kets[k]) for k in LABELS}
    else:
        assert coherences is not None
        phis = {k: float(np.angle(coherences[k])) for k in LABELS}
    info = {
        "phi_01_rad": phis["01"], "phi_11_rad": phis["11"],
        "amp_01": math.sqrt(max(0.0, fid["01"])), "amp_11": math.sqrt(max(0.0, fid["11"])),
        "pop_00": fid["00"], "pop_01": fid["01"], "pop_11": fid["11"],
        "is_mixed_state": True,
    }
    info.update(_phase_block(phis))
    return _finish(fid, info)


def coherence_amplitudes(dim: int, segments, V, delta_1, rates, detuning_offsets=(0.0, 0.0)):
    """<ab| E(|ab><00|) |00> for the four inputs: the gauge-free phase carrier."""
    kets = initial_kets(dim)
    init = {k: np.outer(kets[k], kets["00"].conj()) for k in LABELS}
    fin = propagate(dim, segments, V, delta_1, rates if rates is not None else NoiseRates(),
                    "exact", detuning_offsets, initial=init)
    n = dim * dim
    return {k: complex(np.vdot(kets[k], fin[k].reshape(n, n) @ kets["00"])) for k in LABELS}


def evaluate(dim, segments, V, delta_1, rates, mode="exact", phase_mode="coherence",
             detuning_offsets=(0.0, 0.0)):
    """One gate evaluation -> (fidelities, avg_fidelity, phase_info, finals)."""
    finals = propagate(dim, segments, V, delta_1, rates, mode, detuning_offsets)
    if next(iter(finals.values())).ndim == 1:
        fid, avg, info = cz_fidelity_pure(dim, finals)
    else:
        coh = None
        if phase_mode == "coherence":
            coh = coherence_amplitudes(dim, segments, V, delta_1, rates, detuning_offsets)
        fid, avg, info = cz_fidelity_mixed(dim, finals, phase_mode, coh)
    return fid, avg, info, finals


def compute_cost(avg_fid: float, f11: float, cz_phase_fid: float, gate_time_us: float,
                 time_weight: float = 0.01) -> float:
    """RG/optimize_cz_gate.py:406-431."""
    if any(math.isnan(x) for x in (avg_fid, f11, cz_phase_fid)) or avg_fid < 0.5:
        return 1e6
    return (10.0 * ((1 - avg_fid) * 100) ** 2 + 5.0 * ((1 - f11) * 100) ** 2
            + 2.0 * ((1 - cz_phase_fid) * 100) ** 2 + time_weight * gate_time_us)


# --------------------------------------------------------------------------
# Five levels per atom (|0>, |1>, |e>, |r>, |l>): the intermediate state of the two-photon
# excitation kept explicitly.  EXTENSION -- the reference has no such model (SURVEY.md section 8d,
# config 5; BASELINE.json configs[4]); it uses the reference's own ingredients: the single-photon
# Rabi frequencies Ω1, Ω2 (RG/simulation.py:2896-2912), the intermediate detuning Δe and linewidth
# Γe (RG/atom_database.py:230, 402) and the Rydberg rates of RG/simulation.py:3230-3334.  Parity for
# this model is UNPINNED (no reference output can exist); it is checked against this restatement and
# against its own Δe -> infinity limit, the reference's 3-level model.
# --------------------------------------------------------------------------
LEVEL_E, LEVEL_R5, LEVEL_L = 2, 3, 4


@dataclass
class NoiseRates5:
    """Collapsed rates of the 5-level model (one jump operator per atom and rate)."""

    g_e1: float = 0.0    # |e> -> |1>   Γe · branching
    g_e0: float = 0.0    # |e> -> |0>   Γe · (1 - branching)
    g_r1: float = 0.0    # |r> -> |1>   γ_r · branching_1                 (RG/noise_models.py:1237)
    g_r0: float = 0.0    # |r> -> |0>   γ_r (1 - branching_1) + γ_bbr      (RG/noise_models.py:1242-1251)
    g_rl: float = 0.0    # |r> -> |l>   antitrap + background + leakage    (RG/noise_models.py:1583-1592)
    g_phi: float = 0.0   # dephasing on |r><r|                            (RG/noise_models.py:1579)

    def any(self) -> bool:
        return any(v > 0 for v in (self.g_e1, self.g_e0, self.g_r1, self.g_r0, self.g_rl, self.g_phi))


def hamiltonian5(Omega_eff: complex, Delta: float, V: float, delta_1: float, om1: float, det_e: float,
                 ls_comp: float = 1.0, detuning_offsets: Tuple[float, float] = (0.0, 0.0)) -> np.ndarray:
    """H = Σ_atoms [ Ω1/2 (|e><1| + h.c.) + (Ω2/2 |r><e| + h.c.) − Δe |e><e| − Δ' |r><r| + δ1' |1><1| ]
           + V |rr><rr|,
    with the second leg chosen so that the effective two-photon Rabi frequency is the pulse's:
    Ω2 = Ω_eff · 2Δe/Ω1 (Ω = Ω1Ω2/2Δe, RG/laser_physics.py:265-359), and -- ``ls_comp`` = 1 -- the
    light shifts Ω1²/4Δe on |1> and |Ω2|²/4Δe on |r> compensated (Δ' = Δ + |Ω2|²/4Δe,
    δ1' = δ1 − Ω1²/4Δe), so that Δe -> ∞ at fixed Ω_eff is the reference's 3-level Hamiltonian."""
    d = 5
    I = np.eye(d, dtype=complex)
    om2 = Omega_eff * 2.0 * det_e / om1
    ls = ls_comp / (4.0 * det_e)
    h = 0.5 * om1 * (_op(d, LEVEL_E, 1) + _op(d, 1, LEVEL_E))
    h = h + 0.5 * (om2 * _op(d, LEVEL_R5, LEVEL_E) + np.conj(om2) * _op(d, LEVEL_E, LEVEL_R5))
    h = h - det_e * _op(d, LEVEL_E, LEVEL_E) + (delta_1 - ls * om1 * om1) * _op(d, 1, 1)
    Pr = _op(d, LEVEL_R5, LEVEL_R5)
    H = _two(h, I) + _two(I, h)
    for atom, off in enumerate(detuning_offsets):
        det = Delta + off + ls * abs(om2) ** 2
        H = H - det * (_two(Pr, I) if atom == 0 else _two(I, Pr))
    return H + V * _two(Pr, Pr)


def collapse_operators5(rates: NoiseRates5) -> List[np.ndarray]:
    d = 5
    I = np.eye(d, dtype=complex)
    ops: List[np.ndarray] = []
    for rate, (to, frm) in ((rates.g_e1, (1, LEVEL_E)), (rates.g_e0, (0, LEVEL_E)), (rates.g_r1, (1, LEVEL_R5)),
                            (rates.g_r0, (0, LEVEL_R5)), (rates.g_rl, (LEVEL_L, LEVEL_R5)),
                            (rates.g_phi, (LEVEL_R5, LEVEL_R5))):
        if rate > 0:
            a = math.sqrt(rate) * _op(d, to, frm)
            ops += [_two(a, I), _two(I, a)]
    return ops


def propagate5(segments: Sequence[Segment], V: float, delta_1: float, om1: float, det_e: float,
               rates: Optional[NoiseRates5], ls_comp: float = 1.0, mode: str = "exact",
               detuning_offsets: Tuple[float, float] = (0.0, 0.0),
               initial: Optional[Dict[str, np.ndarray]] = None, nsteps: int = 200000) -> Dict[str, np.ndarray]:
    """Same contract as ``propagate`` for the 5-level model."""
    c_ops = collapse_operators5(rates) if rates is not None else []
    pure = len(c_ops) == 0 and initial is None
    kets = initial_kets(5)
    if initial is not None:
        states = {k: np.asarray(v, dtype=complex).reshape(-1) for k, v in initial.items()}
    elif pure:
        states = {k: v.copy() for k, v in kets.items()}
    else:
        states = {k: np.outer(v, v.conj()).reshape(-1) for k, v in kets.items()}
    for seg in segments:
        H = hamiltonian5(seg.Omega, seg.Delta, V, delta_1, om1, det_e, ls_comp, detuning_offsets)
        gen = -1j * H if pure else liouvillian(H, c_ops)
        if mode == "exact":
            U = sla.expm(gen * seg.dt)
            states = {k: U @ v for k, v in states.items()}
        elif mode == "zvode":
            states = {k: _zvode_advance(gen, v, seg.dt, seg.n_tlist, nsteps) for k, v in states.items()}
        else:
            raise ValueError(mode)
    if pure:
        return states
    return {k: v.reshape(25, 25) for k, v in states.items()}


def evaluate5(segments, V, delta_1, om1, det_e, rates: Optional[NoiseRates5], ls_comp: float = 1.0,
              mode: str = "exact", phase_mode: str = "coherence", detuning_offsets=(0.0, 0.0)):
    """One 5-level gate evaluation -> (fidelities, avg_fidelity, phase_info, finals)."""
    finals = propagate5(segments, V, delta_1, om1, det_e, rates, ls_comp, mode, detuning_offsets)
    if next(iter(finals.values())).ndim == 1:
        fid, avg, info = cz_fidelity_pure(5, finals)
    else:
        coh = None
        if phase_mode == "coherence":
            kets = initial_kets(5)
            init = {k: np.outer(kets[k], kets["00"].conj()) for k in LABELS}
            fin = propagate5(segments, V, delta_1, om1, det_e, rates, ls_comp, "exact", detuning_offsets, initial=init)
            coh = {k: complex(np.vdot(kets[k], fin[k].reshape(25, 25) @ kets["00"])) for k in LABELS}
        fid, avg, info = cz_fidelity_mixed(5, finals, phase_mode, coh)
    return fid, avg, info, finals
