// rydcz_closed.cuh -- the |ab><00| coherences and the single-atom density operators of the 3-level model in closed form.
// (The single-atom rho_01 / rho_10 -- optical Bloch equations, a real 4 x 4 -- have their own header comment further
//  down; what follows is about the coherences.)
//
// The phase-carrying inputs X = |ab><00| (compute_CZ_fidelity's phases, RG/simulation.py:225-633; the whole state
// when the run is noiseless) obey a CLOSED linear equation of their own: <00| is inert under H and under every
// collapse operator (C|00> = 0), so the Lindblad feed terms C X C^+ vanish and
//
//     dX/dt = M X,     M = -i H - (1/2) sum_k rate_k C_k^+ C_k       (codegen/model.py: liouvillian_terms)
//
// acts on the ket side only.  For identical atoms the |11> input lives on {|11>, |1r> = |r1>, |rr>}: M is a
// complex TRIDIAGONAL 3 x 3 matrix; a single-atom input lives on {|1>, |r>}: 2 x 2.  exp(h M) v is then a polynomial
// of degree 2 (1) in M -- the one that interpolates exp(h mu) at the eigenvalues (Newton form, divided differences) --
// and the eigenvalues are the roots of a cubic (quadratic).  A pulse segment costs a few hundred instructions
// instead of K ~ 25-60 Clenshaw steps of three matvecs: the coherence loops were 28 % (|11>) and half of 12 %
// (single atom) of the clustered evaluate kernel (gpurun_out/ab_r.log, round 2).
//
//   * roots: Cardano on the depressed cubic (complex coefficients), then two Newton steps on the cubic in PRODUCT
//     form (a b c - q2 a - q1 c with a, b, c = mu - diagonal), which is accurate to ~eps |Omega| for every root --
//     also for the slow pair next to a root at ~V, where Cardano alone loses (V / Omega) digits;
//   * divided differences: (e^{h b} - e^{h a}) / (b - a), by a series in h (b - a) when that is small; the closest
//     pair of roots comes first, so no difference is ever divided by a gap smaller than the one it spans.  An error
//     in a divided difference is multiplied by the product of the gaps it spans (Newton form), an error delta in a
//     root costs h delta (second order for a double root): measured against 40-digit expm over V / Omega = 0.03 ...
//     1000, Omega = 0, degenerate diagonals, negative durations: <= 1e-13 + 4e-16 |h V| (tests/test_closed_coherence.py);
//   * roots and exponentials are kept across segments that repeat (|Omega|, Delta) (and the duration): bang-bang and
//     smooth-JP pulses change only the laser phase, which enters the off-diagonal elements of M alone.
//
// Plain C++ when RYDCZ_HD is empty (tests/hostsim), like rydcz_core.cuh.  3-level model only: with r+ / r- the |11>
// input has five folded components (quintic) and keeps the Chebyshev loop, as does everything that is not a
// coherence.  The |11> input of UNEQUAL atoms has four components (a quartic): its form is further down.
#pragma once
// (included by rydcz_core.cuh inside namespace rydcz, after run_system)

struct Cx { double re, im; };
RYDCZ_HD Cx cx(double re, double im) { Cx z; z.re = re; z.im = im; return z; }
RYDCZ_HD Cx operator+(Cx a, Cx b) { return cx(a.re + b.re, a.im + b.im); }
RYDCZ_HD Cx operator-(Cx a, Cx b) { return cx(a.re - b.re, a.im - b.im); }
RYDCZ_HD Cx operator-(Cx a) { return cx(-a.re, -a.im); }
RYDCZ_HD Cx operator*(Cx a, Cx b) { return cx(fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)); }
RYDCZ_HD Cx operator*(double s, Cx a) { return cx(s * a.re, s * a.im); }
RYDCZ_HD double cx_abs2(Cx a) { return fma(a.re, a.re, a.im * a.im); }
RYDCZ_HD Cx cx_conj(Cx a) { return cx(a.re, -a.im); }
RYDCZ_HD Cx operator/(Cx a, Cx b) {
  const double inv = 1.0 / cx_abs2(b);
  return cx(fma(a.re, b.re, a.im * b.im) * inv, fma(a.im, b.re, -a.re * b.im) * inv);
}
RYDCZ_HD Cx cx_sqrt(Cx z) {     // principal branch
  const double r = sqrt(cx_abs2(z));
  if (r == 0.0) return cx(0.0, 0.0);
  const double t = sqrt(0.5 * (r + fabs(z.re)));
  return z.re >= 0.0 ? cx(t, z.im / (2.0 * t)) : cx(fabs(z.im) / (2.0 * t), z.im < 0.0 ? -t : t);
}
RYDCZ_HD Cx cx_exp(Cx z) {
  double s, c;
  sincos(z.im, &s, &c);
  const double e = exp(z.re);
  return cx(e * c, e * s);
}
// (e^{h b} - e^{h a}) / (b - a) from Ea = e^{h a}, Eb = e^{h b}; by h e^{h (a + b) / 2} sinh(z) / z, z = h (b - a) / 2,
// when |z| < 0.1 (series to z^8: 3e-14 relative)
RYDCZ_HD Cx cx_exp_dd(Cx a, Cx b, double h, Cx Ea, Cx Eb) {
  const Cx d = b - a;
  const Cx z = 0.5 * h * d;
  if (cx_abs2(z) >= 0.01) return (Eb - Ea) / d;
  const Cx z2 = z * z;
  Cx s = cx(1.0, 0.0) + (1.0 / 72.0) * z2;
  s = cx(1.0, 0.0) + (1.0 / 42.0) * (z2 * s);
  s = cx(1.0, 0.0) + (1.0 / 20.0) * (z2 * s);
  s = cx(1.0, 0.0) + (1.0 / 6.0) * (z2 * s);
  return h * (cx_exp(0.5 * h * (a + b)) * s);
}

// Roots of (mu - d0)(mu - d1)(mu - d2) - q2 (mu - d0) - q1 (mu - d2): the characteristic polynomial of the tridiagonal
// matrix with diagonal d and products of opposite off-diagonal elements q1 (rows 0, 1) and q2 (rows 1, 2).
RYDCZ_HD void tridiagonal3_eigenvalues(Cx d0, Cx d1, Cx d2, double q1, double q2, Cx (&n)[3]) {
  const Cx s = (1.0 / 3.0) * (d0 + d1 + d2);
  const Cx e0 = d0 - s, e1 = d1 - s, e2 = d2 - s;
  const Cx P = e0 * e1 + e0 * e2 + e1 * e2 - cx(q1 + q2, 0.0);
  const Cx Q = q2 * e0 + q1 * e2 - e0 * e1 * e2;
  // nu^3 + P nu + Q = 0:  nu = u - P / (3 u),  u^3 = -Q/2 +- sqrt(Q^2/4 + P^3/27) (the larger of the two)
  const Cx D = cx_sqrt(0.25 * (Q * Q) + (1.0 / 27.0) * (P * P * P));
  Cx u3 = D - 0.5 * Q;
  const Cx u3b = -D - 0.5 * Q;
  if (cx_abs2(u3b) > cx_abs2(u3)) u3 = u3b;
  const double m2 = cx_abs2(u3);
  if (m2 == 0.0) {
    n[0] = s; n[1] = s; n[2] = s;
  } else {
    const double r = cbrt(sqrt(m2));
    double sn, cs;
    sincos(atan2(u3.im, u3.re) * (1.0 / 3.0), &sn, &cs);
    Cx u = cx(r * cs, r * sn);
    const Cx w = cx(-0.5, 0.86602540378443864676);      // e^{2 pi i / 3}
    const Cx P3 = (1.0 / 3.0) * P;
    RYDCZ_UNROLL
    for (int k = 0; k < 3; ++k) {
      n[k] = s + u - P3 / u;
      u = u * w;
    }
  }
  RYDCZ_UNROLL
  for (int k = 0; k < 3; ++k) {
    RYDCZ_UNROLL
    for (int it = 0; it < 2; ++it) {
      const Cx a = n[k] - d0, b = n[k] - d1, c = n[k] - d2;
      const Cx bc = b * c;
      const Cx f = a * bc - q2 * a - q1 * c;
      const Cx fp = bc + a * c + a * b - cx(q1 + q2, 0.0);
      if (cx_abs2(fp) > 0.0) n[k] = n[k] - f / fp;
    }
  }
  // the closest pair first (see the header)
  const double g01 = cx_abs2(n[1] - n[0]), g02 = cx_abs2(n[2] - n[0]), g12 = cx_abs2(n[2] - n[1]);
  if (g02 <= g01 && g02 <= g12) { const Cx t = n[1]; n[1] = n[2]; n[2] = t; }
  else if (g12 <= g01 && g12 <= g02) { const Cx t = n[0]; n[0] = n[2]; n[2] = t; }
}

// what the closed forms keep from one segment to the next
struct Closed3Cache {
  double om2, det, h;          // |Omega|^2, Delta and duration the entries below belong to (om2 < 0: none)
  Cx n[3];                     // eigenvalues
  Cx f0, f01, f012;            // divided differences of exp(h mu) over them
};
struct Closed2Cache {
  double om2, det, h;
  Cx E, ch, sh;                // e^{h mean}, cosh(h s), h sinh(h s) / (h s)
};

RYDCZ_HD bool cx_finite(Cx a) { return fabs(a.re) <= 1.7e308 && fabs(a.im) <= 1.7e308; }

// One segment of the exchange-folded |11><00| coherence (variables of Sys_coh11s_d3: re/im of the |11>, |1r> = |r1>
// and |rr> components):  v <- exp(h M) v,
//   M = [ d0      -i Om*    0     ]   d0 = -2 i delta1 - g_sc
//       [ -i Om/2  d1    -i Om*/2 ]   d1 = -i (delta1 - Delta) - (G + g_sc) / 2,   G = g_r1 + g_r0 + g_phi
//       [ 0       -i Om     d2    ]   d2 = -i (V - 2 Delta) - G
RYDCZ_HD bool coh11s_closed_segment_d3(double (&x)[6], Closed3Cache& cache, double h, double om_re, double om_im,
                                       double det, double V, double delta1, double G, double g_sc) {
  const Cx d0 = cx(-g_sc, -2.0 * delta1), d1 = cx(-0.5 * (G + g_sc), det - delta1), d2 = cx(-G, 2.0 * det - V);
  const double om2 = fma(om_re, om_re, om_im * om_im);
  const bool same_roots = cache.om2 >= 0.0 && fabs(om2 - cache.om2) <= 1e-12 * cache.om2 && det == cache.det;
  if (!same_roots) {
    tridiagonal3_eigenvalues(d0, d1, d2, -0.5 * om2, -0.5 * om2, cache.n);
    cache.om2 = om2; cache.det = det;
  }
  if (!same_roots || h != cache.h) {
    const Cx E0 = cx_exp(h * cache.n[0]), E1 = cx_exp(h * cache.n[1]), E2 = cx_exp(h * cache.n[2]);
    cache.f0 = E0;
    cache.f01 = cx_exp_dd(cache.n[0], cache.n[1], h, E0, E1);
    const Cx f12 = cx_exp_dd(cache.n[1], cache.n[2], h, E1, E2);
    const Cx g = cache.n[2] - cache.n[0];
    cache.f012 = cx_abs2(g) > 0.0 ? (f12 - cache.f01) / g : (0.5 * h * h) * E0;
    cache.h = h;
  }
  const Cx m01 = cx(-om_im, -om_re), m10 = cx(0.5 * om_im, -0.5 * om_re);     // -i Om*, -i Om / 2
  const Cx m12 = cx(-0.5 * om_im, -0.5 * om_re), m21 = cx(om_im, -om_re);     // -i Om* / 2, -i Om
  const Cx v0 = cx(x[0], x[1]), v1 = cx(x[2], x[3]), v2 = cx(x[4], x[5]);
  const Cx a0 = d0 - cache.n[0], a1 = d1 - cache.n[0], a2 = d2 - cache.n[0];
  const Cx u0 = a0 * v0 + m01 * v1, u1 = m10 * v0 + a1 * v1 + m12 * v2, u2 = m21 * v1 + a2 * v2;
  const Cx b0 = d0 - cache.n[1], b1 = d1 - cache.n[1], b2 = d2 - cache.n[1];
  const Cx t0 = b0 * u0 + m01 * u1, t1 = m10 * u0 + b1 * u1 + m12 * u2, t2 = m21 * u1 + b2 * u2;
  const Cx w0 = cache.f0 * v0 + cache.f01 * u0 + cache.f012 * t0;
  const Cx w1 = cache.f0 * v1 + cache.f01 * u1 + cache.f012 * t1;
  const Cx w2 = cache.f0 * v2 + cache.f01 * u2 + cache.f012 * t2;
  x[0] = w0.re; x[1] = w0.im; x[2] = w1.re; x[3] = w1.im; x[4] = w2.re; x[5] = w2.im;
  return cx_finite(w0) && cx_finite(w1) && cx_finite(w2);
}

// One segment of a single-atom coherence (variables of Sys_coh01_d3 / Sys_coh10_d3: re/im of the |1> and |r>
// components of the atom that is driven; det = Delta + that atom's offset):  v <- exp(h M) v,
//   M = [ -i delta1 - g_sc / 2    -i Om* / 2       ]
//       [ -i Om / 2                i Delta - G / 2 ]
// exp(h M) = e^{h mean} [ cosh(h s) + (sinh(h s) / s) (M - mean) ],  s^2 = ((d0 - d1) / 2)^2 - |Om|^2 / 4
// (cosh and sinh(z) / z are even: the branch of s does not matter).
RYDCZ_HD bool coh01_closed_segment_d3(double (&x)[4], Closed2Cache& cache, double h, double om_re, double om_im,
                                      double det, double delta1, double G, double g_sc) {
  const Cx d0 = cx(-0.5 * g_sc, -delta1), d1 = cx(-0.5 * G, det);
  const double om2 = fma(om_re, om_re, om_im * om_im);
  const Cx dl = 0.5 * (d0 - d1);
  if (!(cache.om2 >= 0.0 && fabs(om2 - cache.om2) <= 1e-12 * cache.om2 && det == cache.det && h == cache.h)) {
    const Cx z = h * cx_sqrt(dl * dl - cx(0.25 * om2, 0.0));
    const Cx ez = cx_exp(z), emz = cx(1.0, 0.0) / ez;
    cache.ch = 0.5 * (ez + emz);
    if (cx_abs2(z) >= 0.01) cache.sh = h * ((0.5 * (ez - emz)) / z);
    else {
      const Cx z2 = z * z;
      Cx s = cx(1.0, 0.0) + (1.0 / 72.0) * z2;
      s = cx(1.0, 0.0) + (1.0 / 42.0) * (z2 * s);
      s = cx(1.0, 0.0) + (1.0 / 20.0) * (z2 * s);
      s = cx(1.0, 0.0) + (1.0 / 6.0) * (z2 * s);
      cache.sh = h * s;
    }
    cache.E = cx_exp(0.5 * h * (d0 + d1));
    cache.om2 = om2; cache.det = det; cache.h = h;
  }
  const Cx m01 = cx(-0.5 * om_im, -0.5 * om_re), m10 = cx(0.5 * om_im, -0.5 * om_re);
  const Cx v0 = cx(x[0], x[1]), v1 = cx(x[2], x[3]);
  const Cx w0 = cache.E * (cache.ch * v0 + cache.sh * (dl * v0 + m01 * v1));
  const Cx w1 = cache.E * (cache.ch * v1 + cache.sh * (m10 * v0 - dl * v1));
  x[0] = w0.re; x[1] = w0.im; x[2] = w1.re; x[3] = w1.im;
  return cx_finite(w0) && cx_finite(w1);
}

// ---- the |11><00| coherence of UNEQUAL atoms (variables of Sys_coh11_d3: re/im of |11>, |1r>, |r1>, |rr>) -----------
//   M = [ d0   mc   mc   0  ]   d0 = -2 i delta1 - g_sc,  dA = -i (delta1 - Delta_B) - (G + g_sc) / 2  (|1r>: atom B driven),
//       [ m    dA   0    mc ]   dB = -i (delta1 - Delta_A) - (G + g_sc) / 2,  d3 = -i (V - Delta_A - Delta_B) - G,
//       [ m    0    dB   mc ]   m = -i Om / 2,  mc = -i Om* / 2.
//       [ 0    m    m    d3 ]
// In the basis (|11>, S, |rr>, A) of the symmetric / antisymmetric combinations of |1r>, |r1> this is the tridiagonal block
// of identical atoms (diagonal d0, dS = (dA + dB) / 2, d3) with A hanging on S by kappa = (dA - dB) / 2:
//     p4(mu) = (mu - dS) p3(mu) - kappa^2 (mu - d0)(mu - d3),      p3 = the cubic of tridiagonal3_eigenvalues.
// Roots: Newton on p4 (product form) from the three roots of p3 and from dS -- a Doppler shot has |kappa| << |Omega|.
// A degree-3 Newton form through all four roots would carry terms of size (h V)^2 that cancel (the products (M - n) v
// grow by V per factor); instead the FAST eigenvector (root n3 ~ -i V) is split off with its right / left eigenvectors,
//     exp(h M) v = e^{h n3} P3 v + p(M) (v - P3 v),        p: degree 2 through the three slow roots,
// and the rounding-level fast component that every matvec re-creates is projected out again.  Accepted only when the
// iteration converged to four different roots and the fast root is isolated (V >~ 16 |Omega|: what the clustered path
// needs anyway); else false, and the caller runs the Chebyshev loop.  Against 40-digit expm: <= 2e-13 + 4e-16 |h V|.
struct Closed4Cache {
  double om2, det_a, det_b, h;
  Cx n[4];                     // n[0..2]: slow roots, closest pair first; n[3]: the fast one
  Cx E3, f0, f01, f012;
  bool ok;
};

RYDCZ_HD bool coh11_closed_segment_d3(double (&x)[8], Closed4Cache& cache, double h, double om_re, double om_im,
                                      double det_a, double det_b, double V, double delta1, double G, double g_sc) {
  const Cx d0 = cx(-g_sc, -2.0 * delta1), dA = cx(-0.5 * (G + g_sc), det_b - delta1), dB = cx(-0.5 * (G + g_sc), det_a - delta1);
  const Cx d3 = cx(-G, det_a + det_b - V);
  const double om2 = fma(om_re, om_re, om_im * om_im);
  const bool same_roots = cache.om2 >= 0.0 && fabs(om2 - cache.om2) <= 1e-12 * cache.om2 && det_a == cache.det_a &&
                          det_b == cache.det_b;
  if (!same_roots) {
    const Cx dS = 0.5 * (dA + dB), kap = 0.5 * (dA - dB), k2 = kap * kap;
    const double q = -0.5 * om2;
    Cx c[3];
    tridiagonal3_eigenvalues(d0, dS, d3, q, q, c);        // c[2]: the one far from the closest pair
    Cx start[4];
    start[0] = c[0]; start[1] = c[1]; start[2] = dS; start[3] = c[2];
    const double sc = sqrt(cx_abs2(d0)) + sqrt(cx_abs2(dS)) + sqrt(cx_abs2(d3)) + sqrt(om2) + 1e-300;
    const double tiny2 = (1e-15 * sc) * (1e-15 * sc), lim = 1e-13 * sqrt(om2) + 1e-15 * sc;
    bool ok = true;
    RYDCZ_UNROLL
    for (int k = 0; k < 4; ++k) {
      Cx mu = start[k];
      double step2 = 0.0;
      for (int it = 0; it < 8; ++it) {
        const Cx a = mu - d0, s_ = mu - dS, cc = mu - d3;
        const Cx p3 = a * (s_ * cc) - q * cc - q * a;
        const Cx p3p = s_ * cc + a * cc + a * s_ - cx(2.0 * q, 0.0);
        const Cx f = s_ * p3 - k2 * (a * cc);
        const Cx fp = p3 + s_ * p3p - k2 * (a + cc);
        if (!(cx_abs2(fp) > 0.0)) { step2 = 0.0; break; }
        const Cx st = f / fp;
        mu = mu - st;
        step2 = cx_abs2(st);
        if (step2 <= tiny2) break;
      }
      ok = ok && step2 <= lim * lim;
      cache.n[k] = mu;
    }
    // four different roots (their sum is the trace), the fast one isolated
    const Cx tr = d0 + dA + dB + d3;
    const Cx sum = cache.n[0] + cache.n[1] + cache.n[2] + cache.n[3];
    const double mag = sqrt(cx_abs2(cache.n[0])) + sqrt(cx_abs2(cache.n[1])) + sqrt(cx_abs2(cache.n[2])) + sqrt(cx_abs2(cache.n[3])) + 1e-300;
    ok = ok && cx_abs2(sum - tr) <= (1e-9 * mag) * (1e-9 * mag);
    const double g01 = cx_abs2(cache.n[1] - cache.n[0]), g02 = cx_abs2(cache.n[2] - cache.n[0]), g12 = cx_abs2(cache.n[2] - cache.n[1]);
    const double gmax = fmax(g01, fmax(g02, g12));
    const double iso = fmin(cx_abs2(cache.n[3] - cache.n[0]), fmin(cx_abs2(cache.n[3] - cache.n[1]), cx_abs2(cache.n[3] - cache.n[2])));
    ok = ok && iso >= 64.0 * gmax && iso > 0.0;
    // closest slow pair first
    if (g02 <= g01 && g02 <= g12) { const Cx t = cache.n[1]; cache.n[1] = cache.n[2]; cache.n[2] = t; }
    else if (g12 <= g01 && g12 <= g02) { const Cx t = cache.n[0]; cache.n[0] = cache.n[2]; cache.n[2] = t; }
    cache.ok = ok;
    cache.om2 = om2; cache.det_a = det_a; cache.det_b = det_b;
  }
  if (!cache.ok) return false;
  if (!same_roots || h != cache.h) {
    const Cx E0 = cx_exp(h * cache.n[0]), E1 = cx_exp(h * cache.n[1]), E2 = cx_exp(h * cache.n[2]);
    cache.E3 = cx_exp(h * cache.n[3]);
    cache.f0 = E0;
    cache.f01 = cx_exp_dd(cache.n[0], cache.n[1], h, E0, E1);
    const Cx f12 = cx_exp_dd(cache.n[1], cache.n[2], h, E1, E2);
    const Cx g = cache.n[2] - cache.n[0];
    cache.f012 = cx_abs2(g) > 0.0 ? (f12 - cache.f01) / g : (0.5 * h * h) * E0;
    cache.h = h;
  }
  const Cx m = cx(0.5 * om_im, -0.5 * om_re), mc = cx(-0.5 * om_im, -0.5 * om_re);      // -i Om / 2, -i Om* / 2
  const Cx n3 = cache.n[3];
  // right / left eigenvectors of the fast root, last component 1 (they turn with the laser phase: not cached)
  const Cx one = cx(1.0, 0.0);
  const Cx gA = one / (dA - n3), gB = one / (dB - n3), g = gA + gB;
  const Cx den = (d0 - n3) - (mc * m) * g;
  const Cx r0 = ((mc * mc) * g) / den, l0 = ((m * m) * g) / den;
  const Cx tr_ = m * r0 + mc, tl_ = mc * l0 + m;
  const Cx r1 = -(tr_ * gA), r2 = -(tr_ * gB), l1 = -(tl_ * gA), l2 = -(tl_ * gB);
  const Cx lr = one / (l0 * r0 + l1 * r1 + l2 * r2 + one);
  Cx v0 = cx(x[0], x[1]), v1 = cx(x[2], x[3]), v2 = cx(x[4], x[5]), v3 = cx(x[6], x[7]);
  // fast part / slow part of v
  const Cx cf = (l0 * v0 + l1 * v1 + l2 * v2 + v3) * lr;
  const Cx w0f = cf * r0, w1f = cf * r1, w2f = cf * r2, w3f = cf;
  v0 = v0 - w0f; v1 = v1 - w1f; v2 = v2 - w2f; v3 = v3 - w3f;
  // u = (M - n0) v, re-projected; t = (M - n1) u, re-projected
  const Cx a0 = cache.n[0], a1 = cache.n[1];
  Cx u0 = (d0 - a0) * v0 + mc * (v1 + v2), u1 = m * v0 + (dA - a0) * v1 + mc * v3;
  Cx u2 = m * v0 + (dB - a0) * v2 + mc * v3, u3 = m * (v1 + v2) + (d3 - a0) * v3;
  Cx cu = (l0 * u0 + l1 * u1 + l2 * u2 + u3) * lr;
  u0 = u0 - cu * r0; u1 = u1 - cu * r1; u2 = u2 - cu * r2; u3 = u3 - cu;
  Cx t0 = (d0 - a1) * u0 + mc * (u1 + u2), t1 = m * u0 + (dA - a1) * u1 + mc * u3;
  Cx t2 = m * u0 + (dB - a1) * u2 + mc * u3, t3 = m * (u1 + u2) + (d3 - a1) * u3;
  cu = (l0 * t0 + l1 * t1 + l2 * t2 + t3) * lr;
  t0 = t0 - cu * r0; t1 = t1 - cu * r1; t2 = t2 - cu * r2; t3 = t3 - cu;
  const Cx y0 = cache.E3 * w0f + cache.f0 * v0 + cache.f01 * u0 + cache.f012 * t0;
  const Cx y1 = cache.E3 * w1f + cache.f0 * v1 + cache.f01 * u1 + cache.f012 * t1;
  const Cx y2 = cache.E3 * w2f + cache.f0 * v2 + cache.f01 * u2 + cache.f012 * t2;
  const Cx y3 = cache.E3 * w3f + cache.f0 * v3 + cache.f01 * u3 + cache.f012 * t3;
  x[0] = y0.re; x[1] = y0.im; x[2] = y1.re; x[3] = y1.im; x[4] = y2.re; x[5] = y2.im; x[6] = y3.re; x[7] = y3.im;
  return cx_finite(y0) && cx_finite(y1) && cx_finite(y2) && cx_finite(y3);
}

// FP64 operations of one segment, for the kernels' own flop count (products 6, sums 2, quotients 11; the
// transcendental functions are not counted): roots, exponentials + divided differences, the two matvecs + combination
constexpr double kClosed3FlopsRoots = 400.0, kClosed3FlopsExp = 90.0, kClosed3FlopsApply = 200.0;
constexpr double kClosed2FlopsExp = 70.0, kClosed2FlopsApply = 70.0;

// The whole pulse: every active segment row (next_segment's rule), start operator |11><00| (|01><00|).  false: a
// non-finite value appeared (the caller then runs the Chebyshev loop of the same system).
template <bool NOISY>
RYDCZ_HD bool coh11s_closed_d3(double (&x)[6], const PointView pt, const double* __restrict__ seg, int n_seg,
                               double& flops) {
  Sys_coh11s_d3::init(x);
  Closed3Cache cache;
  cache.om2 = -1.0; cache.det = 0.0; cache.h = 0.0;
  const double V = pt[PT_V], delta1 = pt[PT_DELTA1], off = pt[PT_OFF_A];
  const double G = NOISY ? pt[PT_G_R1] + pt[PT_G_R0] + pt[PT_G_PHI] : 0.0, g_sc = NOISY ? pt[PT_G_SC] : 0.0;
  bool ok = true;
  for (int s = 0; s < n_seg; ++s) {
    const double* sg = seg + (long long)s * SG_STRIDE;
    const double dt = sg[SG_DT], re = sg[SG_RE], im = sg[SG_IM], det = sg[SG_DET] + off;
    if (!segment_active(dt)) continue;
    const double om2_before = cache.om2, det_before = cache.det, h_before = cache.h;
    ok = coh11s_closed_segment_d3(x, cache, dt, re, im, det, V, delta1, G, g_sc) && ok;
    flops += kClosed3FlopsApply + ((cache.om2 != om2_before || cache.det != det_before) ? kClosed3FlopsRoots + kClosed3FlopsExp
                                   : (cache.h != h_before ? kClosed3FlopsExp : 0.0));
  }
  return ok;
}
constexpr double kClosed4FlopsRoots = 1200.0, kClosed4FlopsExp = 110.0, kClosed4FlopsApply = 600.0;
// unequal atoms: the general |11><00| coherence (Sys_coh11_d3).  false: not accepted (see above) or non-finite.
template <bool NOISY>
RYDCZ_HD bool coh11_closed_d3(double (&x)[8], const PointView pt, const double* __restrict__ seg, int n_seg,
                              double& flops) {
  Sys_coh11_d3::init(x);
  Closed4Cache cache;
  cache.om2 = -1.0; cache.det_a = 0.0; cache.det_b = 0.0; cache.h = 0.0; cache.ok = false;
  const double V = pt[PT_V], delta1 = pt[PT_DELTA1], off_a = pt[PT_OFF_A], off_b = pt[PT_OFF_B];
  const double G = NOISY ? pt[PT_G_R1] + pt[PT_G_R0] + pt[PT_G_PHI] : 0.0, g_sc = NOISY ? pt[PT_G_SC] : 0.0;
  for (int s = 0; s < n_seg; ++s) {
    const double* sg = seg + (long long)s * SG_STRIDE;
    const double dt = sg[SG_DT], re = sg[SG_RE], im = sg[SG_IM], det = sg[SG_DET];
    if (!segment_active(dt)) continue;
    const double om2_before = cache.om2, da_before = cache.det_a, db_before = cache.det_b, h_before = cache.h;
    if (!coh11_closed_segment_d3(x, cache, dt, re, im, det + off_a, det + off_b, V, delta1, G, g_sc)) return false;
    flops += kClosed4FlopsApply + ((cache.om2 != om2_before || cache.det_a != da_before || cache.det_b != db_before)
                                       ? kClosed4FlopsRoots + kClosed4FlopsExp
                                       : (cache.h != h_before ? kClosed4FlopsExp : 0.0));
  }
  return true;
}
// ---- the single-atom density operator rho_01 / rho_10 (variables of Sys_rho01_d3 / Sys_rho10_d3) -------------------
// x = (p0, p1, re c, im c, pr) of the driven atom: the optical Bloch equations with decay r -> 1 (g_r1) and r -> 0
// (g_r0).  The trace is conserved (p0' = g_r0 pr = -(p1 + pr)'), so p0 follows from the other four, which obey a REAL
// 4 x 4 system whose characteristic polynomial has coefficients that are sums of positive terms
//     a3 = 2 G2 + G,   a2 = G2^2 + 2 G2 G + d^2 + |Om|^2,   a1 = G (G2^2 + d^2) + |Om|^2 (G2 + g_r0 / 2),
//     a0 = G2 g_r0 |Om|^2 / 2          (G = g_r1 + g_r0, G2 = (G + g_phi + g_sc) / 2, d = Delta + delta1):
// a FAST pair lam, lam* ~ -G' +- i sqrt(|Om|^2 + d^2) (Newton on the quartic) and a SLOW pair of size ~ G, the roots of
// mu^2 + b' mu + c' with c' = a0 / |lam|^2, b' = (a1 - b c') / |lam|^2 (no cancellation: the slow roots come out with
// their full relative accuracy although they are 1e-4 of the fast ones).  With q(mu) = (mu - lam)(mu - lam*),
//     exp(h M) = al0 + al1 M + q(M) (be0 + be1 M),
// al: the line through exp(h mu) at the fast pair; be: the line through (exp(h mu) - al0 - al1 mu) / q(mu) at the slow
// pair (divided differences by Leibniz' rule: nothing is divided by the small gap of the slow pair).  Every term is
// O(|v|): q(M) removes the fast components, so nothing of size (h Omega)^k has to cancel.  Declines (false: the caller
// runs the Chebyshev loop) when Newton did not converge or the pairs are not separated (|q(slow)| < a2 / 256: drive
// and detuning both ~ 0, or damping ~ drive).  Against 40-digit expm, |Om| 1e2..1e8, rates up to 3e6: <= 2e-14 + 4e-16 |h Om_eff| (tests/test_closed_coherence.py).
struct ClosedRhoCache {
  double om2, det, h;          // |Omega|^2, Delta (+ offset) and duration the entries below belong to (om2 < 0: none)
  double b, c;                 // q(mu) = mu^2 + b mu + c
  Cx lam, r1, r2;              // fast root (upper half plane), slow roots (both real, or a conjugate pair)
  double al0, al1, be0, be1;
  bool ok;
};

RYDCZ_HD bool rho01_closed_segment_d3(double (&x)[5], ClosedRhoCache& cache, double h, double om_re, double om_im,
                                      double det, double delta1, double g_r1, double g_r0, double g_phi, double g_sc) {
  const double om2 = fma(om_re, om_re, om_im * om_im);
  const double G = g_r1 + g_r0, G2 = 0.5 * (G + g_phi + g_sc), d = det + delta1;
  const bool same_roots = cache.om2 >= 0.0 && fabs(om2 - cache.om2) <= 1e-12 * cache.om2 && det == cache.det;
  if (!same_roots) {
    const double a3 = 2.0 * G2 + G, a2 = G2 * G2 + 2.0 * G2 * G + d * d + om2;
    const double a1 = G * (G2 * G2 + d * d) + om2 * (G2 + 0.5 * g_r0), a0 = 0.5 * G2 * g_r0 * om2;
    cache.om2 = om2; cache.det = det; cache.h = 0.0;
    cache.ok = false;
    if (!(a2 > 0.0) || !(a2 <= 1.7e308)) return false;
    Cx lam = cx(-0.5 * (a3 - a1 / a2), sqrt(a2));
    double step2 = 0.0;
    bool ok = true;
    RYDCZ_UNROLL
    for (int it = 0; it < 5; ++it) {
      const Cx p = (((lam + cx(a3, 0.0)) * lam + cx(a2, 0.0)) * lam + cx(a1, 0.0)) * lam + cx(a0, 0.0);
      const Cx dp = ((4.0 * lam + cx(3.0 * a3, 0.0)) * lam + cx(2.0 * a2, 0.0)) * lam + cx(a1, 0.0);
      if (!(cx_abs2(dp) > 0.0)) { ok = false; break; }
      const Cx st = p / dp;
      lam = lam - st;
      step2 = cx_abs2(st);
    }
    const double c = cx_abs2(lam), b = -2.0 * lam.re;
    ok = ok && c > 0.0 && step2 <= 1e-24 * c;
    if (!ok) return false;
    const double cp = a0 / c, bp = (a1 - b * cp) / c;
    const double disc = bp * bp - 4.0 * cp;
    Cx r1, r2;
    if (disc >= 0.0) {
      const double sq = sqrt(disc);
      const double q = -0.5 * (bp + (bp >= 0.0 ? sq : -sq));
      r1 = cx(q, 0.0);
      r2 = cx(q != 0.0 ? cp / q : 0.0, 0.0);
    } else {
      r1 = cx(-0.5 * bp, 0.5 * sqrt(-disc));
      r2 = cx_conj(r1);
    }
    // two Newton steps on the quartic itself (Horner is stable next to 0: every term is of the size of a0)
    RYDCZ_UNROLL
    for (int k = 0; k < 2; ++k) {
      Cx r = k == 0 ? r1 : r2;
      RYDCZ_UNROLL
      for (int it = 0; it < 2; ++it) {
        const Cx p = (((r + cx(a3, 0.0)) * r + cx(a2, 0.0)) * r + cx(a1, 0.0)) * r + cx(a0, 0.0);
        const Cx dp = ((4.0 * r + cx(3.0 * a3, 0.0)) * r + cx(2.0 * a2, 0.0)) * r + cx(a1, 0.0);
        if (cx_abs2(dp) > 0.0) r = r - p / dp;
      }
      if (k == 0) r1 = r; else r2 = r;
    }
    if (disc < 0.0) r2 = cx_conj(r1);
    const Cx q1 = (r1 + cx(b, 0.0)) * r1 + cx(c, 0.0), q2 = (r2 + cx(b, 0.0)) * r2 + cx(c, 0.0);
    const double sep = a2 * (1.0 / 256.0);
    if (!(cx_abs2(q1) >= sep * sep) || !(cx_abs2(q2) >= sep * sep)) return false;
    cache.lam = lam; cache.r1 = r1; cache.r2 = r2; cache.b = b; cache.c = c;
    cache.ok = true;
  }
  if (!cache.ok) return false;
  if (!same_roots || h != cache.h) {
    const Cx lam = cache.lam, r1 = cache.r1, r2 = cache.r2;
    const Cx E = cx_exp(h * lam);
    cache.al1 = cx_exp_dd(lam, cx_conj(lam), h, E, cx_conj(E)).re;
    cache.al0 = (E - cache.al1 * lam).re;
    const Cx q1 = (r1 + cx(cache.b, 0.0)) * r1 + cx(cache.c, 0.0), q2 = (r2 + cx(cache.b, 0.0)) * r2 + cx(cache.c, 0.0);
    Cx E1, E2, F12;
    if (r1.im == 0.0 && r2.im == 0.0) {
      // the usual case, two real slow roots: real exponentials, and exp(h (r1 + r2) / 2) = sqrt(E1 E2)
      E1 = cx(exp(h * r1.re), 0.0); E2 = cx(exp(h * r2.re), 0.0);
      const double z = 0.5 * h * (r2.re - r1.re), z2 = z * z;
      double dd;
      if (z2 >= 0.01) dd = (E2.re - E1.re) / (r2.re - r1.re);
      else {
        double sr = fma(z2, 1.0 / 72.0, 1.0);
        sr = fma(z2 * sr, 1.0 / 42.0, 1.0);
        sr = fma(z2 * sr, 1.0 / 20.0, 1.0);
        sr = fma(z2 * sr, 1.0 / 6.0, 1.0);
        dd = h * sqrt(E1.re * E2.re) * sr;
      }
      F12 = cx(dd - cache.al1, 0.0);
    } else {
      E1 = cx_exp(h * r1); E2 = cx_exp(h * r2);
      F12 = cx_exp_dd(r1, r2, h, E1, E2) - cx(cache.al1, 0.0);
    }
    const Cx g1 = (E1 - cx(cache.al0, 0.0) - cache.al1 * r1) / q1;
    cache.be1 = ((F12 - g1 * (r1 + r2 + cx(cache.b, 0.0))) / q2).re;
    cache.be0 = (g1 - cache.be1 * r1).re;
    cache.h = h;
  }
  // M on (p1, re c, im c, pr): the rows of Sys_rho01_d3::apply
  const double hr = 0.5 * om_re, hi = 0.5 * om_im, Gr = g_r1 + g_r0;
#define RYDCZ_RHO01_M(o, v)                                                  \
  const double o##0 = -om_im * v##1 - om_re * v##2 + g_r1 * v##3;           \
  const double o##1 = hi * (v##0 - v##3) - G2 * v##1 + d * v##2;            \
  const double o##2 = hr * (v##0 - v##3) - d * v##1 - G2 * v##2;            \
  const double o##3 = om_im * v##1 + om_re * v##2 - Gr * v##3;
  const double v0 = x[1], v1 = x[2], v2 = x[3], v3 = x[4];
  RYDCZ_RHO01_M(u, v)
  RYDCZ_RHO01_M(t, u)
  const double w0 = t0 + cache.b * u0 + cache.c * v0, w1 = t1 + cache.b * u1 + cache.c * v1;
  const double w2 = t2 + cache.b * u2 + cache.c * v2, w3 = t3 + cache.b * u3 + cache.c * v3;
  RYDCZ_RHO01_M(z, w)
#undef RYDCZ_RHO01_M
  const double y0 = cache.al0 * v0 + cache.al1 * u0 + cache.be0 * w0 + cache.be1 * z0;
  const double y1 = cache.al0 * v1 + cache.al1 * u1 + cache.be0 * w1 + cache.be1 * z1;
  const double y2 = cache.al0 * v2 + cache.al1 * u2 + cache.be0 * w2 + cache.be1 * z2;
  const double y3 = cache.al0 * v3 + cache.al1 * u3 + cache.be0 * w3 + cache.be1 * z3;
  x[0] += (v0 + v3) - (y0 + y3);      // the trace is conserved
  x[1] = y0; x[2] = y1; x[3] = y2; x[4] = y3;
  return fabs(y0) <= 1.7e308 && fabs(y1) <= 1.7e308 && fabs(y2) <= 1.7e308 && fabs(y3) <= 1.7e308;
}
constexpr double kClosedRhoFlopsRoots = 900.0, kClosedRhoFlopsExp = 160.0, kClosedRhoFlopsApply = 130.0;
// The whole pulse.  atom_b: rho_01 (atom B driven, offset PT_OFF_B), else rho_10 (atom A, PT_OFF_A).  false: declined
// (x is then unspecified and the caller runs the Chebyshev loop from the start operator).
RYDCZ_HD bool rho01_closed_d3(double (&x)[5], const PointView pt, const double* __restrict__ seg, int n_seg,
                              bool atom_b, double& flops) {
  Sys_rho01_d3::init(x);
  ClosedRhoCache cache;
  cache.om2 = -1.0; cache.det = 0.0; cache.h = 0.0; cache.ok = false;
  const double delta1 = pt[PT_DELTA1], off = atom_b ? pt[PT_OFF_B] : pt[PT_OFF_A];
  const double g_r1 = pt[PT_G_R1], g_r0 = pt[PT_G_R0], g_phi = pt[PT_G_PHI], g_sc = pt[PT_G_SC];
  for (int s = 0; s < n_seg; ++s) {
    const double* sg = seg + (long long)s * SG_STRIDE;
    const double dt = sg[SG_DT], re = sg[SG_RE], im = sg[SG_IM], det = sg[SG_DET] + off;
    if (!segment_active(dt)) continue;
    const double h_before = cache.h, om2_before = cache.om2, det_before = cache.det;
    if (!rho01_closed_segment_d3(x, cache, dt, re, im, det, delta1, g_r1, g_r0, g_phi, g_sc)) return false;
    flops += kClosedRhoFlopsApply + ((cache.om2 != om2_before || cache.det != det_before) ? kClosedRhoFlopsRoots + kClosedRhoFlopsExp
                                     : (cache.h != h_before ? kClosedRhoFlopsExp : 0.0));
  }
  return true;
}

// atom_b: the |01> input (atom B driven, offset PT_OFF_B), else the |10> input (atom A, PT_OFF_A)
template <bool NOISY>
RYDCZ_HD bool coh01_closed_d3(double (&x)[4], const PointView pt, const double* __restrict__ seg, int n_seg,
                              bool atom_b, double& flops) {
  Sys_coh01_d3::init(x);
  Closed2Cache cache;
  cache.om2 = -1.0; cache.det = 0.0; cache.h = 0.0;
  const double delta1 = pt[PT_DELTA1], off = atom_b ? pt[PT_OFF_B] : pt[PT_OFF_A];
  const double G = NOISY ? pt[PT_G_R1] + pt[PT_G_R0] + pt[PT_G_PHI] : 0.0, g_sc = NOISY ? pt[PT_G_SC] : 0.0;
  bool ok = true;
  for (int s = 0; s < n_seg; ++s) {
    const double* sg = seg + (long long)s * SG_STRIDE;
    const double dt = sg[SG_DT], re = sg[SG_RE], im = sg[SG_IM], det = sg[SG_DET] + off;
    if (!segment_active(dt)) continue;
    const double h_before = cache.h, om2_before = cache.om2, det_before = cache.det;
    ok = coh01_closed_segment_d3(x, cache, dt, re, im, det, delta1, G, g_sc) && ok;
    flops += kClosed2FlopsApply +
             ((cache.om2 != om2_before || cache.det != det_before || cache.h != h_before) ? kClosed2FlopsExp : 0.0);
  }
  return ok;
}

