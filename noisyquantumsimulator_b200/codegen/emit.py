"""Print the reduced real systems of ``model.py`` as straight-line CUDA/C++.

Each system becomes a struct with compile-time sizes and three inline functions:

    coefs(p, c)      c[q] = distinct linear forms  sum_k m_k p[k]  found among the
                     nonzeros of  sum_k p_k M_k  (p has one extra trailing entry, the
                     spectral shift used by coherence systems);
    bounds(c, ...)   Gershgorin information for the Chebyshev interval;
    apply(c, x, y)   y += G x, fully unrolled: one FMA per structural nonzero, every
                     index a literal so x, y, c stay in registers.

Run ``python -m noisyquantumsimulator_b200.codegen.emit`` to regenerate
``csrc/rydcz_systems_d3.inc`` / ``_d4.inc`` / ``_d5.inc`` (committed; the build does not need Python).
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from . import model


BIG_CHUNKS = 3        # pieces of one big-system matvec between scheduling fences
BIG_SYSTEM_N = 40    # more real variables than this: coefficients and start vector live in shared memory


def _fmt(v: float) -> str:
    r = repr(float(v))
    return r if ("." in r or "e" in r or "inf" in r) else r + ".0"


def _linear_forms(sysm: model.RealSystem, shift_matrix):
    """Group nonzeros by their parameter-coefficient vector (up to sign)."""
    mats = list(sysm.M) + ([shift_matrix] if shift_matrix is not None else [])
    K = len(mats)
    stack = np.stack(mats, axis=0)  # (K, n, n)
    forms: Dict[Tuple, int] = {}
    form_vecs: List[np.ndarray] = []
    entries: List[Tuple[int, int, int, int]] = []  # (row, col, q, sign)
    n = sysm.n
    for i in range(n):
        for j in range(n):
            v = stack[:, i, j]
            if not np.any(v != 0):
                continue
            first = v[np.nonzero(v)[0][0]]
            sgn = 1 if first > 0 else -1
            key = tuple(np.round(v * sgn, 12))
            if key not in forms:
                forms[key] = len(form_vecs)
                form_vecs.append(v * sgn)
            entries.append((i, j, forms[key], sgn))
    return K, form_vecs, entries


def emit_system(sysm: model.RealSystem, struct_name: str, cm: bool = False) -> str:
    n = sysm.n
    # the spectral shift: coherence systems are complex-linear, so multiplying the state by
    # exp(-i c t) commutes with the generator.  In real (re, im) layout "multiply by -i c"
    # is  re' += c*im, im' -= c*re ; we add p_shift * S with S = that matrix for c = -1,
    # i.e. the generator actually exponentiated is G - i*cen (centre moved to zero).
    shift = None
    if not sysm.hermitian:
        shift = np.zeros((n, n))
        for k in range(0, n, 2):
            assert sysm.var_info[k][2] == "re" and sysm.var_info[k + 1][2] == "im"
            shift[k, k + 1] = -1.0
            shift[k + 1, k] = +1.0
    K, forms, entries = _linear_forms(sysm, shift)
    pnames = list(sysm.params) + (["shift"] if shift is not None else [])
    L: List[str] = []
    A = L.append
    A(f"// ---- {struct_name}: start operator '{sysm.name}', d={sysm.d}, n={n} real variables,")
    A(f"//      {len(entries)} structural nonzeros, {len(forms)} distinct coefficients")
    for k, (i, j, part) in enumerate(sysm.var_info):
        A(f"//   x[{k}] = {part}( X[{i},{j}] )")
    A(f"struct {struct_name} {{")
    A(f"  static constexpr int N = {n};")
    A(f"  static constexpr int NC = {len(forms)};")
    A(f"  static constexpr int NNZ = {len(entries)};")
    A(f"  static constexpr int NP = {len(pnames)};   // {', '.join(pnames)}")
    A(f"  static constexpr bool HERMITIAN = {'true' if sysm.hermitian else 'false'};")
    nz0 = np.nonzero(sysm.x0)[0]
    assert all(sysm.x0[k] == 1.0 for k in nz0)
    A("  RYDCZ_HD static void init(double* __restrict__ x) {   // the start operator")
    A("    " + " ".join(f"x[{k}] = {'1.0' if k in nz0 else '0.0'};" for k in range(n)))
    A("  }")
    plist = list(sysm.params)

    def pidx(*names):
        for nm in names:
            if nm in plist:
                return plist.index(nm)
        return -1

    A(f"  static constexpr int IDX_OM_RE = {pidx('om_re')}, IDX_OM_IM = {pidx('om_im')}, "
      f"IDX_DET_A = {pidx('det_a', 'det')}, IDX_DET_B = {pidx('det_b', 'det')}, IDX_V = {pidx('V')}, "
      f"IDX_DELTA1 = {pidx('delta1')};")
    A(f"  static constexpr int IDX_G_R1 = {pidx('g_r1')}, IDX_G_R0 = {pidx('g_r0')}, IDX_G_PHI = {pidx('g_phi')}, "
      f"IDX_G_SC = {pidx('g_sc')}, IDX_G_MJ = {pidx('g_mj')};")
    A(f"  static constexpr int DIM = {sysm.d}, IDX_OM1 = {pidx('om1')}, IDX_DET_E = {pidx('det_e')}, "
      f"IDX_G_E1 = {pidx('g_e1')}, IDX_G_E0 = {pidx('g_e0')}, IDX_G_RL = {pidx('g_rl')};")
    if getattr(sysm, "expand", None) is not None:
        A(f"  static constexpr int FULL_N = {sysm.full_n};")
        A("  // state of the general (unfolded) system from the exchange-folded one")
        A("  RYDCZ_HD static void expand(const double* __restrict__ xs, double* __restrict__ xg) {")
        E = sysm.expand
        for r in range(E.shape[0]):
            nz = np.nonzero(E[r])[0]
            assert len(nz) <= 1
            if len(nz) == 0:
                A(f"    xg[{r}] = 0.0;")
                continue
            cidx = int(nz[0])
            A(f"    xg[{r}] = {'' if E[r, cidx] > 0 else '-'}xs[{cidx}];")
        A("  }")
    # Big systems keep their coefficients (and the segment's start vector) in per-thread shared-memory
    # columns instead of registers: coefs/bounds take any indexable container, and apply_cm walks the
    # nonzeros coefficient by coefficient so that each c[q] is loaded once per matvec.
    big = n > BIG_SYSTEM_N or cm     # cm: templated containers although the state fits in registers
    # coefs
    if big:
        A("  template <class CV> RYDCZ_HD static void coefs(const double* __restrict__ p, CV c) {")
    else:
        A("  RYDCZ_HD static void coefs(const double* __restrict__ p, double* __restrict__ c) {")
    for q, v in enumerate(forms):
        terms = []
        for k in np.nonzero(v)[0]:
            coef = v[k]
            if coef == 1.0:
                terms.append(f"+ p[{k}]")
            elif coef == -1.0:
                terms.append(f"- p[{k}]")
            else:
                terms.append(f"{'+' if coef > 0 else '-'} {_fmt(abs(coef))} * p[{k}]")
        expr = " ".join(terms).lstrip("+ ").strip()
        A(f"    c[{q}] = {expr};")
    A("  }")
    # bounds: per-row |diag| and off-diagonal absolute row sums
    A("  // Gershgorin: every eigenvalue lies in a disc centred at the (real, <=0) diagonal entry.")
    A("  // rho systems: returns radius = max_i(|G_ii| + sum_{j!=i}|G_ij|), lo = hi = 0.")
    A("  // coh systems: the pair frequency E_i = G[re_i][im_i]; returns the interval [lo, hi] that")
    A("  // contains every E_i -/+ (other off-diagonals of the row) and damp = max |G_ii|.")
    if big:
        A("  template <class CV> RYDCZ_HD static void bounds(const CV c, double& lo, double& hi, double& damp) {")
    else:
        A("  RYDCZ_HD static void bounds(const double* __restrict__ c, double& lo, double& hi, double& damp) {")
    rows: Dict[int, List[Tuple[int, int, int]]] = {}
    for (i, j, q, s) in entries:
        rows.setdefault(i, []).append((j, q, s))
    if sysm.hermitian:
        A("    double r = 0.0, t; damp = 0.0;")
        for i in range(n):
            parts = [f"fabs(c[{q}])" for (j, q, s) in rows.get(i, [])]
            if not parts:
                continue
            A(f"    t = {' + '.join(parts)}; r = fmax(r, t);")
            dq = [q for (j, q, s) in rows.get(i, []) if j == i]
            if dq:
                A(f"    damp = fmax(damp, fabs(c[{dq[0]}]));")
        A("    lo = -r; hi = r;")
    else:
        A("    double t, e; lo = 1e300; hi = -1e300; damp = 0.0;")
        for k in range(0, n, 2):
            # row of re_k: entry on im_k carries +E (d re/dt = E*im)
            E = [(q, s) for (j, q, s) in rows.get(k, []) if j == k + 1]
            others = [f"fabs(c[{q}])" for (j, q, s) in rows.get(k, []) if j not in (k, k + 1)]
            dq = [q for (j, q, s) in rows.get(k, []) if j == k]
            A(f"    e = {('%sc[%d]' % ('' if E[0][1] > 0 else '-', E[0][0])) if E else '0.0'};"
              f" t = {' + '.join(others) if others else '0.0'};")
            A("    lo = fmin(lo, e - t); hi = fmax(hi, e + t);")
            if dq:
                A(f"    damp = fmax(damp, fabs(c[{dq[0]}]));")
    A("  }")
    # apply
    A("  RYDCZ_HD static void apply(const double* __restrict__ c, const double* __restrict__ x, double* __restrict__ y) {")
    order = os.environ.get("RYDCZ_EMIT_ORDER", "col")   # measured on B200: col 19.53 ms, row 19.79, rr 19.80, cm 20.60
    stmts = [(i, j, q, s) for i in range(n) for (j, q, s) in rows.get(i, [])]
    if order == "col":        # walk the source vector: consecutive FMAs hit different accumulators
        stmts.sort(key=lambda t: (t[1], t[0]))
    elif order == "cm":       # coefficient-major: consecutive FMAs share the coefficient operand
        stmts.sort(key=lambda t: (t[2], t[0]))
    elif order == "rr":       # round-robin over rows: k-th nonzero of every row, then the (k+1)-th ...
        per = {}
        for t in stmts:
            per.setdefault(t[0], []).append(t)
        stmts = [per[i][k] for k in range(max(len(v) for v in per.values())) for i in sorted(per) if k < len(per[i])]
    for (i, j, q, s) in stmts:
        A(f"    y[{i}] {'+=' if s > 0 else '-='} c[{q}] * x[{j}];")
    A("  }")
    if big:
        A("  // coefficient-major: c lives in a shared-memory column, each c[q] is read once")
        A("  template <class CV> RYDCZ_HD static void apply_cm(const CV c, const double* __restrict__ x, double* __restrict__ y) {")
        A("    double cq;")
        by_q: Dict[int, List[Tuple[int, int, int]]] = {}
        for (i, j, q, s) in stmts:
            by_q.setdefault(q, []).append((i, j, s))
        for q in sorted(by_q):
            A(f"    cq = coef_load(c, {q});")
            for (i, j, s) in by_q[q]:
                A(f"    y[{i}] {'+=' if s > 0 else '-='} cq * x[{j}];")
        A("  }")
        A("  // One Clenshaw half-step  y <- ak * xs + G x + y  with xs (the segment's start vector) and c in")
        A("  // shared-memory columns.  The loads are spread over CHUNKS pieces separated by scheduling fences")
        A("  // so that only a few of them are in flight at a time (otherwise ptxas hoists all ~86 loads to the")
        A("  // top of the block and spills the Clenshaw vectors).")
        qs = sorted(by_q)
        n_chunks = BIG_CHUNKS if n > BIG_SYSTEM_N else 1
        A("  template <class CV, class XV> RYDCZ_HD static void apply_cm_axpy(const CV c, const XV xs, double ak, const double* __restrict__ x, double* __restrict__ y) {")
        A("    double cq;")
        for ch in range(n_chunks):
            for i in range(ch * n // n_chunks, (ch + 1) * n // n_chunks):
                A(f"    y[{i}] = fma(ak, coef_load(xs, {i}), y[{i}]);")
            for q in qs[ch * len(qs) // n_chunks:(ch + 1) * len(qs) // n_chunks]:
                A(f"    cq = coef_load(c, {q});")
                for (i, j, s) in by_q[q]:
                    A(f"    y[{i}] {'+=' if s > 0 else '-='} cq * x[{j}];")
            if ch != n_chunks - 1:
                A("    RYDCZ_SCHED_FENCE();")
        A("  }")
    if n > 2 * BIG_SYSTEM_N:
        # A system whose three Clenshaw vectors do not fit in registers at all (the general rho_11 of the 5-level
        # model: 81 variables): the vectors live in per-thread shared-memory columns and the matvec is emitted ROW
        # by row -- one accumulator, one load per nonzero, one store per row; the coefficients stay in registers.
        A("  // One Clenshaw half-step  bo <- ak * xs + G bi + bo  with every vector in a container (shared-memory columns)")
        A("  template <class XV> RYDCZ_HD static void apply_rows_axpy(const double* __restrict__ c, const XV xs, double ak, const XV bi, const XV bo) {")
        A("    double a;")
        for i in range(n):
            A(f"    a = fma(ak, xs[{i}], bo[{i}]);")
            for (j, q, sgn) in sorted(rows.get(i, [])):
                A(f"    a {'+=' if sgn > 0 else '-='} c[{q}] * bi[{j}];")
            A(f"    bo[{i}] = a;")
        A("  }")
    if sysm.d in (3, 4):
        # Laser-phase covariance: G(Omega e^{i phi}) = R(phi) G(Omega) R(phi)^T, R rotating the (re, im) pair of
        # entry X[i,j] by q phi, q = (Rydberg excitations of |i>) - (of |j>).  Used by the repeated-segment path
        # (csrc/rydcz_repeat.cuh): one propagator serves every segment of a phase-modulated pulse.
        d = sysm.d

        def n_ryd(i):
            return sum(1 for lev in divmod(int(i), d) if lev >= 2)

        A("  // x <- R(phi) x given (c1, s1) = (cos phi, sin phi)")
        A("  RYDCZ_HD static void rotate(double* __restrict__ x, double c1, double s1) {")
        A("    const double c2 = c1 * c1 - s1 * s1, s2 = 2.0 * s1 * c1;")
        A("    double re;")
        used_c2 = False
        for k, (i, j, part) in enumerate(sysm.var_info):
            if part != "re":
                continue
            q = n_ryd(i) - n_ryd(j)
            if (i, j, "im") not in sysm.var_info:
                assert q == 0, (sysm.name, i, j)
                continue
            k2 = sysm.var_info.index((i, j, "im"))
            if q == 0:
                continue
            assert abs(q) in (1, 2)
            c, sn = ("c1", "s1") if abs(q) == 1 else ("c2", "s2")
            sg = "" if q > 0 else "-"
            used_c2 |= abs(q) == 2
            A(f"    re = x[{k}]; x[{k}] = {c} * re - ({sg}{sn}) * x[{k2}]; x[{k2}] = ({sg}{sn}) * re + {c} * x[{k2}];")
        if not used_c2:
            A("    (void)c2; (void)s2;")
        A("  }")
    A("};")
    A("")
    return "\n".join(L)


def emit_dim(d: int) -> str:
    systems = model.standard_systems(d)
    out = [f"// GENERATED by noisyquantumsimulator_b200/codegen/emit.py -- do not edit.",
           f"// Two-atom {d}-level Rydberg model: reduced real Lindblad systems.",
           f"// Parameters p[]: {', '.join(model.params_of(d))} (, shift for coh systems)",
           ""]
    for key, sysm in systems.items():
        out.append(emit_system(sysm, f"Sys_{key}_d{d}"))
    folded = model.symmetric_systems(d)
    for key, sysm in folded.items():
        out.append(emit_system(sysm, f"Sys_{key}_d{d}", cm=(key == "rho11s" and d == 3)))
    for key, sysm in model.combined_input_systems(d).items():
        out.append(emit_system(sysm, f"Sys_{key}_d{d}"))
        out.append(f"static constexpr int {key.upper()}_COH_OFFSET_D{d} = {systems['rho' + key[2:]].n};")
    # readout indices
    N = d * d
    out.append(f"// readout indices (d={d})")
    idx11 = d * 1 + 1
    out.append(f"static constexpr int RHO11S_POP_D{d} = {folded['rho11s'].index_of(idx11, idx11, 'diag')};")
    out.append(f"static constexpr int COH11S_RE_D{d} = {folded['coh11s'].index_of(idx11, 0, 're')};")
    out.append(f"static constexpr int COH11S_IM_D{d} = {folded['coh11s'].index_of(idx11, 0, 'im')};")
    for lab in ("01", "10", "11"):
        a, b = int(lab[0]), int(lab[1])
        idx = d * a + b
        rho = systems["rho" + lab]
        coh = systems["coh" + lab]
        out.append(f"static constexpr int RHO{lab}_POP_D{d} = {rho.index_of(idx, idx, 'diag')};")
        out.append(f"static constexpr int COH{lab}_RE_D{d} = {coh.index_of(idx, 0, 're')};")
        out.append(f"static constexpr int COH{lab}_IM_D{d} = {coh.index_of(idx, 0, 'im')};")
    # diagonal (population) variables of rho11 by sector, for loss / Rydberg-residual readouts
    rho11 = systems["rho11"]
    diag = [(k, i) for k, (i, j, part) in enumerate(rho11.var_info) if part == "diag"]
    ryd = {3: {2}, 4: {2, 3}, 5: {model.LEVEL_R5}}[d]

    def cls(i):
        a, b = i // d, i % d
        return (a in ryd or b in ryd, a == 0 or b == 0)

    ryd_terms = " + ".join(f"x[{k}]" for k, i in diag if cls(i)[0])
    if d == 5:
        # only the block in which no atom has been lost is evolved: loss = 1 - its trace
        out.append("// |11> input: probability that at least one atom has left {1, e, r} (to |0> or |l>) / sits in |r>")
        out.append(f"RYDCZ_HD double rho11_loss_d{d}(const double* __restrict__ x) {{ return 1.0 - ("
                   + " + ".join(f"x[{k}]" for k, i in diag) + "); }")
    else:
        loss_terms = " + ".join(f"x[{k}]" for k, i in diag if cls(i)[1])
        out.append(f"// |11> input: probability that at least one atom sits in |0> / in a Rydberg level")
        out.append(f"RYDCZ_HD double rho11_loss_d{d}(const double* __restrict__ x) {{ return {loss_terms}; }}")
    out.append(f"RYDCZ_HD double rho11_ryd_d{d}(const double* __restrict__ x) {{ return {ryd_terms}; }}")
    if d != 5:
        # invariants used by the clustered-spectrum path to validate its result (trace of rho_11; norm of the ket)
        out.append(f"RYDCZ_HD double rho11_trace_d{d}(const double* __restrict__ x) {{ return "
                   + " + ".join(f"x[{k}]" for k, i in diag) + "; }")
        out.append(f"RYDCZ_HD double coh11_norm2_d{d}(const double* __restrict__ x) {{ return "
                   + " + ".join(f"x[{k}] * x[{k}]" for k in range(systems['coh11'].n)) + "; }")
    if d == 5:
        coh11 = systems["coh11"]
        terms = [f"x[{k}] * x[{k}]" for k, (i, j, part) in enumerate(coh11.var_info) if cls(i)[0]]
        out.append("// noiseless |11> input (x = the ket, coh11 layout): probability that an atom sits in |r>")
        out.append(f"RYDCZ_HD double coh11_ryd_d{d}(const double* __restrict__ x) {{ return {' + '.join(terms)}; }}")
    out.append("")
    return "\n".join(out)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(os.path.dirname(here), "csrc")
    for d in (3, 4, 5):
        path = os.path.join(csrc, f"rydcz_systems_d{d}.inc")
        with open(path, "w") as f:
            f.write(emit_dim(d))
        print("wrote", path)


if __name__ == "__main__":
    main()
