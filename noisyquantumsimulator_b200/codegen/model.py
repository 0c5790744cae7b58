"""Linear-parametric description of the two-atom Rydberg Lindblad generator.

The generator the reference integrates (H from RG/hamiltonians.py:584-989, collapse
operators from RG/noise_models.py:1199-1620, assembled in RG/simulation.py:3178-3334)
is *linear* in a handful of scalars p_k:

    d/dt vec(X) = ( sum_k p_k L_k ) vec(X)

with p = (Re Ω, Im Ω, Δ_A, Δ_B, V, δ_1, Γ_r→1, Γ_r→0, γ_φ, γ_sc[, γ_mJ]).  Lindblad terms
that share an operator add in rate, so the reference's ~14 collapse operators
collapse to these rates.  This module builds the constant matrices
L_k, restricts them to the operator subspace reachable from one start operator
(|ab><ab| for a density-matrix input, |ab><00| for the phase-carrying coherence),
and turns the complex system into a real one.  ``emit.py`` prints the result as
straight-line CUDA.

Basis index of the two-atom space: d*a + b; vec(X) is row-major.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

PARAMS3 = ("om_re", "om_im", "det_a", "det_b", "V", "delta1", "g_r1", "g_r0", "g_phi", "g_sc")
PARAMS4 = PARAMS3 + ("g_mj",)
# d = 5: the intermediate state of the two-photon excitation is kept (levels 0, 1, e, r, l; DESIGN.md
# section 2.6).  om1 = first-leg Rabi frequency (real), om_re/om_im = second leg (carries the pulse phase),
# det_e = intermediate detuning, det_a/det_b = two-photon detuning per atom, g_e1/g_e0 = decay of |e>
# to |1> / |0>, g_rl = loss channels of |r> that leave the computational manifold (-> |l>).
PARAMS5 = ("om_re", "om_im", "det_a", "det_b", "V", "delta1", "g_r1", "g_r0", "g_phi", "om1", "det_e", "g_e1",
           "g_e0", "g_rl")
LEVEL_E, LEVEL_R5, LEVEL_L = 2, 3, 4


def params_of(d: int) -> Tuple[str, ...]:
    return {3: PARAMS3, 4: PARAMS4, 5: PARAMS5}[d]


def _op(d, i, j):
    m = np.zeros((d, d), dtype=complex)
    m[i, j] = 1.0
    return m


def _hamiltonian_terms(d: int) -> Dict[str, np.ndarray]:
    """H = sum_k p_k H_k for the coherent parameters."""
    I = np.eye(d, dtype=complex)
    if d == 5:
        return _hamiltonian_terms_5()
    ryd = [2] if d == 3 else [2, 3]
    two = np.kron
    up = _op(d, 2, 1)  # |r(+)><1|: the default sigma+ coupling reaches only r+ for d=4
    # (Ω/2)|r><1| + (Ω*/2)|1><r| with Ω = re + i im
    h_re = 0.5 * (up + up.conj().T)
    h_im = 0.5 * (1j * up - 1j * up.conj().T)
    Pr = sum(_op(d, r, r) for r in ryd)
    P1 = _op(d, 1, 1)
    return {
        "om_re": two(h_re, I) + two(I, h_re),
        "om_im": two(h_im, I) + two(I, h_im),
        "det_a": -two(Pr, I),
        "det_b": -two(I, Pr),
        "V": two(Pr, Pr),
        "delta1": two(P1, I) + two(I, P1),
    }


def _hamiltonian_terms_5() -> Dict[str, np.ndarray]:
    """Ladder 1 -(om1)- e -(om2)- r per atom:
    H = sum_atoms [ om1/2 (|e><1| + h.c.) + (om2/2 |r><e| + h.c.) - det_e |e><e| - det |r><r| + delta1 |1><1| ]
        + V |rr><rr|.  Eliminating |e> for |det_e| >> om1, |om2| gives the 3-level model with
    Omega = om1 om2 / (2 det_e) (the reference's effective Rabi frequency, RG/laser_physics.py:265-359)
    plus light shifts om1^2/(4 det_e) on |1> and |om2|^2/(4 det_e) on |r>."""
    d = 5
    I = np.eye(d, dtype=complex)
    two = np.kron
    e, r = LEVEL_E, LEVEL_R5
    up1 = _op(d, e, 1)
    up2 = _op(d, r, e)
    h1 = 0.5 * (up1 + up1.conj().T)
    h_re = 0.5 * (up2 + up2.conj().T)
    h_im = 0.5 * (1j * up2 - 1j * up2.conj().T)
    Pe, Pr, P1 = _op(d, e, e), _op(d, r, r), _op(d, 1, 1)
    return {
        "om1": two(h1, I) + two(I, h1),
        "om_re": two(h_re, I) + two(I, h_re),
        "om_im": two(h_im, I) + two(I, h_im),
        "det_e": -(two(Pe, I) + two(I, Pe)),
        "det_a": -two(Pr, I),
        "det_b": -two(I, Pr),
        "V": two(Pr, Pr),
        "delta1": two(P1, I) + two(I, P1),
    }


def _jump_terms(d: int) -> Dict[str, List[np.ndarray]]:
    """Unit-rate jump operators grouped by the rate that multiplies them."""
    I = np.eye(d, dtype=complex)
    if d == 5:
        e, r, l = LEVEL_E, LEVEL_R5, LEVEL_L
        both5 = lambda a: [np.kron(a, I), np.kron(I, a)]
        return {"g_r1": both5(_op(d, 1, r)), "g_r0": both5(_op(d, 0, r)), "g_rl": both5(_op(d, l, r)),
                "g_phi": both5(_op(d, r, r)), "g_e1": both5(_op(d, 1, e)), "g_e0": both5(_op(d, 0, e))}
    ryd = [2] if d == 3 else [2, 3]
    two = np.kron

    def both(a):
        return [two(a, I), two(I, a)]

    out = {
        "g_r1": sum((both(_op(d, 1, r)) for r in ryd), []),
        "g_r0": sum((both(_op(d, 0, r)) for r in ryd), []),
        "g_phi": sum((both(_op(d, r, r)) for r in ryd), []),
        "g_sc": both(_op(d, 1, 1)),
    }
    if d == 4:
        out["g_mj"] = both(_op(d, 3, 2)) + both(_op(d, 2, 3))
    return out


@functools.lru_cache(maxsize=None)
def liouvillian_terms(d: int) -> Tuple[Tuple[str, ...], List[np.ndarray]]:
    """(param names, [L_k]) with L_k of shape (d^4, d^4), row-major vec convention."""
    names = params_of(d)
    n = d * d
    I = np.eye(n, dtype=complex)
    H = _hamiltonian_terms(d)
    J = _jump_terms(d)
    terms = []
    for nm in names:
        if nm in H:
            h = H[nm]
            terms.append(-1j * (np.kron(h, I) - np.kron(I, h.T)))
        else:
            L = np.zeros((n * n, n * n), dtype=complex)
            for C in J[nm]:
                CdC = C.conj().T @ C
                L += np.kron(C, C.conj()) - 0.5 * np.kron(CdC, I) - 0.5 * np.kron(I, CdC.T)
            terms.append(L)
    return names, terms


@dataclass
class RealSystem:
    """x' = (sum_k p_k M_k) x, x real of length n."""

    name: str
    d: int
    hermitian: bool
    n: int
    params: Tuple[str, ...]
    M: List[np.ndarray]                 # one (n, n) real matrix per parameter
    x0: np.ndarray                      # start vector
    # variable bookkeeping: var index -> (row i, col j, part) with part in {"re", "im", "diag"}
    var_info: List[Tuple[int, int, str]] = field(default_factory=list)

    def generator(self, p: Sequence[float]) -> np.ndarray:
        return sum(pk * Mk for pk, Mk in zip(p, self.M))

    def to_operator(self, x: np.ndarray) -> np.ndarray:
        """Rebuild the (d^2 x d^2) complex operator from the real state."""
        N = self.d * self.d
        X = np.zeros((N, N), dtype=complex)
        for v, (i, j, part) in zip(x, self.var_info):
            if part == "diag":
                X[i, i] += v
            elif part == "re":
                X[i, j] += v
                if self.hermitian:
                    X[j, i] += v
            else:
                X[i, j] += 1j * v
                if self.hermitian:
                    X[j, i] -= 1j * v
        return X

    def index_of(self, i: int, j: int, part: str) -> int:
        return self.var_info.index((i, j, part))


def reduce_system(name: str, d: int, X0: np.ndarray, hermitian: bool, readout=None) -> RealSystem:
    """Restrict to the operator entries reachable from X0 and make the system real.

    ``readout`` (a list of operator entries (i, j)) additionally drops every entry that cannot
    influence any readout entry: sink sectors (an atom lost to |0> or |l>) are fed by the active
    block and never feed back, so the remaining system is still closed and exact."""
    names, L = liouvillian_terms(d)
    N = d * d
    pattern = sum(np.abs(Lk) for Lk in L) > 0
    reach = np.abs(X0.reshape(-1)) > 0
    while True:
        new = reach | (pattern @ reach.astype(float) > 0)
        if (new == reach).all():
            break
        reach = new
    if readout is not None:
        need = np.zeros(N * N, dtype=bool)
        for (i, j) in readout:
            need[i * N + j] = True
            if hermitian:
                need[j * N + i] = True
        while True:
            new = need | (pattern.T @ need.astype(float) > 0)
            if (new == need).all():
                break
            need = new
        reach = reach & need
    entries = [(idx // N, idx % N) for idx in np.nonzero(reach)[0]]
    eset = set(entries)
    var_info: List[Tuple[int, int, str]] = []
    if hermitian:
        for (i, j) in entries:
            assert (j, i) in eset
        for (i, j) in sorted(eset):
            if i == j:
                var_info.append((i, i, "diag"))
            elif i < j:
                var_info.append((i, j, "re"))
                var_info.append((i, j, "im"))
    else:
        for (i, j) in sorted(eset):
            var_info.append((i, j, "re"))
            var_info.append((i, j, "im"))
    n = len(var_info)
    vidx = {t: k for k, t in enumerate(var_info)}

    def expand(col_entry):
        """entry (k,l) of X as a list of (var index, complex weight)."""
        k, l = col_entry
        if hermitian:
            if k == l:
                return [(vidx[(k, k, "diag")], 1.0 + 0j)]
            if k < l:
                return [(vidx[(k, l, "re")], 1.0 + 0j), (vidx[(k, l, "im")], 1j)]
            return [(vidx[(l, k, "re")], 1.0 + 0j), (vidx[(l, k, "im")], -1j)]
        return [(vidx[(k, l, "re")], 1.0 + 0j), (vidx[(k, l, "im")], 1j)]

    M = []
    for Lk in L:
        Mk = np.zeros((n, n))
        for row_var, (i, j, part) in enumerate(var_info):
            row = Lk[i * N + j]
            for col in np.nonzero(row)[0]:
                kl = (col // N, col % N)
                if kl not in eset:
                    continue
                for (cv, w) in expand(kl):
                    c = row[col] * w
                    Mk[row_var, cv] += c.imag if part == "im" else c.real
        M.append(Mk)
    x0 = np.zeros(n)
    for k, (i, j, part) in enumerate(var_info):
        x0[k] = X0[i, j].imag if part == "im" else X0[i, j].real
    return RealSystem(name, d, hermitian, n, names, M, x0, var_info)


def standard_systems(d: int) -> Dict[str, RealSystem]:
    """The six systems one gate evaluation needs (|00> is inert: H and every collapse
    operator annihilate it, so neither its ρ nor its coherence moves)."""
    N = d * d
    out = {}

    def ket(a, b):
        v = np.zeros(N, dtype=complex)
        v[d * a + b] = 1.0
        return v

    k00 = ket(0, 0)
    for lab in ("01", "10", "11"):
        a, b = int(lab[0]), int(lab[1])
        k = ket(a, b)
        ro_rho = ro_coh = None
        if d == 5:
            # keep the block in which no atom has been lost: populations of every state an atom that
            # started in |1> can still coherently occupy (1, e, r); an atom that started in |0> stays there
            act = lambda s: (1, LEVEL_E, LEVEL_R5) if s == 1 else (0,)
            states = [d * x + y for x in act(a) for y in act(b)]
            ro_rho = [(i, i) for i in states]
            ro_coh = [(i, 0) for i in states]
        out["rho" + lab] = reduce_system("rho" + lab, d, np.outer(k, k.conj()), True, ro_rho)
        out["coh" + lab] = reduce_system("coh" + lab, d, np.outer(k, k00.conj()), False, ro_coh)
    return out


# --------------------------------------------------------------------------------------------
# Exchange-symmetric specialisation: identical atoms (det_a == det_b)
# --------------------------------------------------------------------------------------------
def sym_params(d: int) -> Tuple[str, ...]:
    """Parameter list with the two per-atom detunings merged into one."""
    return tuple("det" if n == "det_a" else n for n in params_of(d) if n != "det_b")


@dataclass
class FoldedSystem(RealSystem):
    """A RealSystem restricted to the atom-exchange-invariant subspace.  ``expand`` (n_full x n)
    maps its state back to the variables of the general system it was folded from."""

    expand: np.ndarray = None
    full_n: int = 0


def fold_exchange(full: RealSystem) -> FoldedSystem:
    """When both atoms see the same parameters the generator commutes with swapping the atoms, and
    the start operators |11><11|, |11><00| are swap invariant, so X(t) = P X(t) P for all t.  In
    the real variables of ``full`` the swap is a signed permutation S; the invariant subspace is
    spanned by one representative per orbit.  Reduced generator: rows of the representatives of
    M E, where E embeds orbit variables into the full space."""
    d, N = full.d, full.d * full.d

    def swap(i):
        a, b = divmod(i, d)
        return d * b + a

    idx = {t: k for k, t in enumerate(full.var_info)}
    S_perm = np.zeros(full.n, dtype=int)
    S_sign = np.ones(full.n)
    for k, (i, j, part) in enumerate(full.var_info):
        i2, j2 = swap(i), swap(j)
        sign = 1.0
        if full.hermitian and i2 > j2:          # entry (i2, j2) is the conjugate of (j2, i2)
            i2, j2 = j2, i2
            sign = -1.0 if part == "im" else 1.0
        S_perm[k], S_sign[k] = idx[(i2, j2, part)], sign
    reps, cols = [], []
    seen = set()
    for k in range(full.n):
        if k in seen:
            continue
        k2 = int(S_perm[k])
        seen.update((k, k2))
        col = np.zeros(full.n)
        if k2 == k:
            if S_sign[k] < 0:
                continue                         # x = -x: identically zero on the invariant subspace
            col[k] = 1.0
        else:
            col[k], col[k2] = 1.0, S_sign[k]
        reps.append(k)
        cols.append(col)
    E = np.stack(cols, axis=1)                   # (n_full, n_sym)
    names = list(full.params)
    ia, ib = names.index("det_a"), names.index("det_b")
    M_sym = []
    for nm in sym_params(d):
        Mk = full.M[ia] + full.M[ib] if nm == "det" else full.M[names.index(nm)]
        MkE = Mk @ E
        # invariance check: M E must lie in the range of E
        resid = MkE - E @ MkE[reps, :]
        assert np.abs(resid).max() < 1e-12, (full.name, nm)
        M_sym.append(MkE[reps, :])
    x0 = full.x0[reps]
    assert np.allclose(E @ x0, full.x0)
    return FoldedSystem(name=full.name + "s", d=d, hermitian=full.hermitian, n=len(reps),
                        params=sym_params(d), M=M_sym, x0=x0,
                        var_info=[full.var_info[k] for k in reps], expand=E, full_n=full.n)


def symmetric_systems(d: int) -> Dict[str, FoldedSystem]:
    """rho11 and coh11 folded by atom exchange.  (The |01> input is a single-atom problem already,
    and by the same symmetry the |10> input equals it.)"""
    full = standard_systems(d)
    return {"rho11s": fold_exchange(full["rho11"]), "coh11s": fold_exchange(full["coh11"])}


def direct_sum(a: RealSystem, b: RealSystem, name: str) -> RealSystem:
    """Two systems driven by the same parameters, advanced as one (block-diagonal generator, no
    spectral shift).  Used for the single-atom inputs: rho_01 (5 variables) and the |01><00|
    coherence (4) have similar, small spectral radii, so sharing one Chebyshev loop costs a few
    extra FMAs and saves one loop with all its per-segment set-up."""
    assert a.params == b.params and a.d == b.d
    n = a.n + b.n
    M = []
    for Ma, Mb in zip(a.M, b.M):
        Mk = np.zeros((n, n))
        Mk[:a.n, :a.n] = Ma
        Mk[a.n:, a.n:] = Mb
        M.append(Mk)
    return RealSystem(name=name, d=a.d, hermitian=True, n=n, params=a.params, M=M,
                      x0=np.concatenate([a.x0, b.x0]), var_info=list(a.var_info) + list(b.var_info))


def combined_input_systems(d: int) -> Dict[str, RealSystem]:
    full = standard_systems(d)
    return {"in01": direct_sum(full["rho01"], full["coh01"], "in01"),
            "in10": direct_sum(full["rho10"], full["coh10"], "in10")}
