// rydcz_closed.cuh -- the |ab><00| coherences of the 3-level model in closed form.
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
// input has five folded components (quintic), and the general (unequal-atom) |11> input four (quartic); those keep
// the Chebyshev loop, as does everything that is not a coherence.
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
    const Cx ez = cx_exp(z), emz = cx_exp(-z);
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

