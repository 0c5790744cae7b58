// rydcz_repeat.cuh -- one propagator for every segment of a phase-modulated pulse.
//
// The smooth Jandura-Pupillo protocol is 300 equal steps that differ only in the laser phase
// (RG/simulation.py:1693-1733); any schedule whose active segments share (duration, |Omega|, Delta) is the
// same case.  The generator is covariant under the phase:
//
//     G(Omega e^{i phi}) = R(phi) G(Omega) R(phi)^T,     R(phi): the (re, im) pair of X[i,j] turned by q phi,
//                                                        q = Rydberg excitations of |i> minus those of |j>
//
// (U = exp(-i phi sum_atoms |r><r|) conjugates the coupling, commutes with every diagonal term, and the
// collapse operators only pick up a phase; Sys::rotate is generated from the variable table and checked
// against the generated systems in tests/test_repeat_hostsim.py).  Hence exp(h G_s) = R(phi_s) E0 R(-phi_s)
// with ONE dense E0 = exp(h G(|Omega|)) per system and point:
//
//   build:  column c of E0 = the plain Chebyshev propagation (run_system) of the unit vector e_c through one
//           segment; one thread per (point, column), 30 columns for the noisy 3-level evaluation;
//   apply:  per segment  x <- E0 R(phi_prev - phi_s) x  for the four systems at once (the phasors come from
//           the segment rows, no trigonometry), one thread per point, E0 in shared memory ([entry][lane]).
//
// 300 x (N^2 + a rotation) FMAs instead of 300 x ~40 sparse matvecs: ~9x fewer flops for the noisy smooth-JP
// evaluation, whatever V / Omega.  Exact to rounding like the other paths.
//
// Plain C++ when RYDCZ_HD is empty (tests/hostsim).
#pragma once

namespace rydcz {

constexpr int RP_MIN_COUNT = 32;    // fewer equal segments than this: not worth the N columns of E0

// offsets (doubles per point, times the batch stride) of the four propagators and of the per-column step counts
template <int D> struct RepeatLayout;
template <> struct RepeatLayout<3> {
  static constexpr int NA = Sys_rho11s_d3::N, NB = Sys_coh11s_d3::N, NC_ = Sys_rho01_d3::N, ND = Sys_coh01_d3::N;
  static constexpr int RHO11 = 0, COH11 = RHO11 + NA * NA, RHO01 = COH11 + NB * NB, COH01 = RHO01 + NC_ * NC_;
  static constexpr int STEPS = COH01 + ND * ND;       // one entry per column: Chebyshev terms its build took
  static constexpr int COLS = NA + NB + NC_ + ND;
  static constexpr int STRIDE = STEPS + COLS;
};

// All active segments equal in (duration, |Omega|^2 to 1e-9, Delta), |Omega| > 0, at least RP_MIN_COUNT of them.
// first = index of the first active segment row.
RYDCZ_HD bool repeat_eligible(const double* __restrict__ seg, int n_seg, int& first, int& count) {
  first = -1; count = 0;
  double dt = 0.0, det = 0.0, om2 = 0.0;
  for (int s = 0; s < n_seg; ++s) {
    const double* sg = seg + (long long)s * SG_STRIDE;
    const double d = sg[SG_DT];
    if (!(d != 0.0 && d == d)) continue;
    const double o2 = sg[SG_RE] * sg[SG_RE] + sg[SG_IM] * sg[SG_IM];
    if (first < 0) {
      first = s; dt = d; det = sg[SG_DET]; om2 = o2;
      if (!(om2 > 0.0)) return false;
    } else if (!(d == dt && sg[SG_DET] == det && fabs(o2 - om2) <= 1e-9 * om2)) {
      return false;
    }
    ++count;
  }
  return count >= RP_MIN_COUNT;
}

// Column c of E0 for one system: the unit vector e_c through one segment at phase 0.  E[(r N + c) stride].
// Returns the Chebyshev terms spent.
template <class Sys>
RYDCZ_HD double repeat_build_column(int c, const PointView pt, const double* __restrict__ seg_row, bool noisy,
                                    double* __restrict__ E, long long stride) {
  constexpr int N = Sys::N;
  double row[SG_STRIDE];
  row[SG_DT] = seg_row[SG_DT];
  row[SG_RE] = sqrt(seg_row[SG_RE] * seg_row[SG_RE] + seg_row[SG_IM] * seg_row[SG_IM]);
  row[SG_IM] = 0.0;
  row[SG_DET] = seg_row[SG_DET];
  double x[N];
  RYDCZ_UNROLL
  for (int i = 0; i < N; ++i) x[i] = i == c ? 1.0 : 0.0;
  double st = 0.0;
  // the coherence systems have a closed form (rydcz_closed.cuh): their columns cost one segment formula, not a
  // Chebyshev propagation -- 10 of the 30 columns of the noisy 3-level evaluation
  bool done = false;
  if constexpr (RYDCZ_CLOSED_COHERENCE != 0 && Sys::DIM == 3 && !Sys::HERMITIAN && (N == 6 || N == 4)) {
    const double G = noisy ? pt[PT_G_R1] + pt[PT_G_R0] + pt[PT_G_PHI] : 0.0, g_sc = noisy ? pt[PT_G_SC] : 0.0;
    const double det = row[SG_DET] + pt[PT_OFF_A];        // (identical atoms on this path: OFF_A = OFF_B)
    if constexpr (N == 6) {
      Closed3Cache cache;
      cache.om2 = -1.0; cache.det = 0.0; cache.h = 0.0;
      done = coh11s_closed_segment_d3(x, cache, row[SG_DT], row[SG_RE], 0.0, det, pt[PT_V], pt[PT_DELTA1], G, g_sc);
    } else {
      Closed2Cache cache;
      cache.om2 = -1.0; cache.det = 0.0; cache.h = 0.0;
      done = coh01_closed_segment_d3(x, cache, row[SG_DT], row[SG_RE], 0.0, det, pt[PT_DELTA1], G, g_sc);
    }
    if (!done) { RYDCZ_UNROLL for (int i = 0; i < N; ++i) x[i] = i == c ? 1.0 : 0.0; }
  }
  // ... and so has the single-atom density operator (optical Bloch equations): 5 more columns
  if constexpr (RYDCZ_CLOSED_COHERENCE != 0 && RYDCZ_CLOSED_RHO01 != 0 && Sys::DIM == 3 && Sys::HERMITIAN && N == 5) {
    if (noisy) {
      ClosedRhoCache cache;
      cache.om2 = -1.0; cache.det = 0.0; cache.h = 0.0; cache.ok = false;
      done = rho01_closed_segment_d3(x, cache, row[SG_DT], row[SG_RE], 0.0, row[SG_DET] + pt[PT_OFF_A], pt[PT_DELTA1],
                                     pt[PT_G_R1], pt[PT_G_R0], pt[PT_G_PHI], pt[PT_G_SC]);
      if (!done) { RYDCZ_UNROLL for (int i = 0; i < N; ++i) x[i] = i == c ? 1.0 : 0.0; }
    }
  }
  if (!done) run_system<Sys>(x, pt, row, 1, noisy, st, false);
  RYDCZ_UNROLL
  for (int r = 0; r < N; ++r) E[(long long)(r * N + c) * stride] = x[r];
  return st;
}

// One dense step of one system: x <- E0 R(phi_prev - phi_s) x, (c1, s1) = cos / sin of the turn.
template <class Sys>
RYDCZ_HD void repeat_step(double (&x)[Sys::N], const double* __restrict__ E, long long stride, double c1, double s1) {
  constexpr int N = Sys::N;
  Sys::rotate(x, c1, s1);
  // column by column: N independent accumulators, so consecutive FMAs never wait for one another (one warp
  // per scheduler here: there is no other warp to fill a dependent chain's latency)
  double y[N];
  RYDCZ_UNROLL
  for (int r = 0; r < N; ++r) y[r] = E[(long long)(r * N) * stride] * x[0];
  RYDCZ_UNROLL
  for (int c = 1; c < N; ++c) {
    RYDCZ_UNROLL
    for (int r = 0; r < N; ++r) y[r] = fma(E[(long long)(r * N + c) * stride], x[c], y[r]);
  }
  RYDCZ_UNROLL
  for (int r = 0; r < N; ++r) x[r] = y[r];
}

template <bool NOISY>
RYDCZ_HD void repeat_finish_d3(const double (&xa)[Sys_rho11s_d3::N], const double (&xb)[Sys_coh11s_d3::N],
                               const double (&xc)[Sys_rho01_d3::N], const double (&xd)[Sys_coh01_d3::N], int count,
                               const double* __restrict__ E, long long stride, double* __restrict__ out,
                               double* __restrict__ states);

// One full gate evaluation from the four propagators (identical atoms, 3-level model).  All systems advance
// in ONE pass over the segment rows: x <- prod_s R(phi_s) E0 R(-phi_s) x_start, the state kept in the frame of
// the previous segment; the next row is loaded while the current step computes.
template <bool NOISY>
RYDCZ_HD void evaluate_point_repeat_d3(const double* __restrict__ E, long long stride, const double* __restrict__ seg,
                                       int n_seg, double* __restrict__ out, double* __restrict__ states) {
  typedef StateLayout<3> SL;
  typedef RepeatLayout<3> RL;
  double xa[Sys_rho11s_d3::N], xb[Sys_coh11s_d3::N], xc[Sys_rho01_d3::N], xd[Sys_coh01_d3::N];
  Sys_rho11s_d3::init(xa); Sys_coh11s_d3::init(xb); Sys_rho01_d3::init(xc); Sys_coh01_d3::init(xd);
  double pr = 1.0, pi = 0.0;      // phasor of the previous segment
  int count = 0;                  // active segments
  double nd = 0.0, nre = 0.0, nim = 0.0;
  if (n_seg > 0) { nd = seg[SG_DT]; nre = seg[SG_RE]; nim = seg[SG_IM]; }
  for (int s = 0; s < n_seg; ++s) {
    const double d = nd, re = nre, im = nim;
    if (s + 1 < n_seg) {
      const double* nx = seg + (long long)(s + 1) * SG_STRIDE;
      nd = nx[SG_DT]; nre = nx[SG_RE]; nim = nx[SG_IM];
    }
    if (!(d != 0.0 && d == d)) continue;
    ++count;
    const double inv = 1.0 / sqrt(re * re + im * im);
    const double ur = re * inv, ui = im * inv;
    const double c1 = pr * ur + pi * ui, s1 = pi * ur - pr * ui;     // turn by phi_prev - phi_s
    if (NOISY) {
      repeat_step<Sys_rho11s_d3>(xa, E + (long long)RL::RHO11 * stride, stride, c1, s1);
      repeat_step<Sys_rho01_d3>(xc, E + (long long)RL::RHO01 * stride, stride, c1, s1);
    }
    repeat_step<Sys_coh11s_d3>(xb, E + (long long)RL::COH11 * stride, stride, c1, s1);
    repeat_step<Sys_coh01_d3>(xd, E + (long long)RL::COH01 * stride, stride, c1, s1);
    pr = ur; pi = ui;
  }
  // back to the laboratory frame
  if (NOISY) { Sys_rho11s_d3::rotate(xa, pr, pi); Sys_rho01_d3::rotate(xc, pr, pi); }
  Sys_coh11s_d3::rotate(xb, pr, pi); Sys_coh01_d3::rotate(xd, pr, pi);
  repeat_finish_d3<NOISY>(xa, xb, xc, xd, count, E, stride, out, states);
}

// Readouts, epilogue and work counters from the four final states (laboratory frame).
template <bool NOISY>
RYDCZ_HD void repeat_finish_d3(const double (&xa)[Sys_rho11s_d3::N], const double (&xb)[Sys_coh11s_d3::N],
                               const double (&xc)[Sys_rho01_d3::N], const double (&xd)[Sys_coh01_d3::N], int count,
                               const double* __restrict__ E, long long stride, double* __restrict__ out,
                               double* __restrict__ states) {
  typedef StateLayout<3> SL;
  typedef RepeatLayout<3> RL;
  double p01 = 1.0, p11 = 1.0, loss = 0.0, ryd = 0.0, flops = 0.0, build_steps = 0.0;
  int n_sys = 2;
  if (NOISY) {
    p11 = xa[RHO11S_POP_D3];
    double xf[Sys_rho11s_d3::FULL_N];
    Sys_rho11s_d3::expand(xa, xf);
    loss = rho11_loss_d3(xf); ryd = rho11_ryd_d3(xf);
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_rho11s_d3::FULL_N; ++i) states[SL::RHO11 + i] = xf[i]; }
    p01 = xc[RHO01_POP_D3];
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_rho01_d3::N; ++i) { states[SL::RHO01 + i] = xc[i]; states[SL::RHO10 + i] = xc[i]; } }
    n_sys = 4;
    flops += 2.0 * count * (RL::NA * RL::NA + RL::NC_ * RL::NC_);
    for (int c = 0; c < RL::NA; ++c) {
      const double st = E[(long long)(RL::STEPS + c) * stride];
      build_steps += st; flops += st * 2.0 * (Sys_rho11s_d3::NNZ + Sys_rho11s_d3::N);
    }
    for (int c = 0; c < RL::NC_; ++c) {
      const double st = E[(long long)(RL::STEPS + RL::NA + RL::NB + c) * stride];
      build_steps += st; flops += st * 2.0 * (Sys_rho01_d3::NNZ + Sys_rho01_d3::N);
    }
  }
  const double c11r = xb[COH11S_RE_D3], c11i = xb[COH11S_IM_D3];
  const double c01r = xd[COH01_RE_D3], c01i = xd[COH01_IM_D3];
  if (!NOISY) ryd = 1.0 - (c11r * c11r + c11i * c11i);
  if (states) {
    double xf[Sys_coh11s_d3::FULL_N];
    Sys_coh11s_d3::expand(xb, xf);
    RYDCZ_UNROLL for (int i = 0; i < Sys_coh11s_d3::FULL_N; ++i) states[SL::COH11 + i] = xf[i];
    RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d3::N; ++i) { states[SL::COH01 + i] = xd[i]; states[SL::COH10 + i] = xd[i]; }
  }
  flops += 2.0 * count * (RL::NB * RL::NB + RL::ND * RL::ND);
  for (int c = 0; c < RL::NB; ++c) {
    const double st = E[(long long)(RL::STEPS + RL::NA + c) * stride];
    build_steps += st; flops += st * 2.0 * (Sys_coh11s_d3::NNZ + Sys_coh11s_d3::N);
  }
  for (int c = 0; c < RL::ND; ++c) {
    const double st = E[(long long)(RL::STEPS + RL::NA + RL::NB + RL::NC_ + c) * stride];
    build_steps += st; flops += st * 2.0 * (Sys_coh01_d3::NNZ + Sys_coh01_d3::N);
  }
  cz_epilogue(NOISY, p01, p01, p11, c01r, c01i, c01r, c01i, c11r, c11i, out);
  out[OUT_LOSS] = loss; out[OUT_RYD] = ryd;
  out[OUT_STEPS] = build_steps + (double)count * n_sys;    // operator applications: sparse (build) + dense (apply)
  out[OUT_FLOPS] = flops;
}

// The build of one column, dispatched on the column index (the host harness loops over all of them; on the
// GPU a warp holds one column index, so the branch is uniform).
template <bool NOISY>
RYDCZ_HD void repeat_build_d3(int col, const PointView pt, const double* __restrict__ seg_row,
                              double* __restrict__ E, long long stride) {
  typedef RepeatLayout<3> RL;
  double st = 0.0;
  if (col < RL::NA) {
    if (!NOISY) return;
    st = repeat_build_column<Sys_rho11s_d3>(col, pt, seg_row, true, E + (long long)RL::RHO11 * stride, stride);
  } else if (col < RL::NA + RL::NB) {
    st = repeat_build_column<Sys_coh11s_d3>(col - RL::NA, pt, seg_row, NOISY, E + (long long)RL::COH11 * stride, stride);
  } else if (col < RL::NA + RL::NB + RL::NC_) {
    if (!NOISY) return;
    st = repeat_build_column<Sys_rho01_d3>(col - RL::NA - RL::NB, pt, seg_row, true, E + (long long)RL::RHO01 * stride, stride);
  } else {
    st = repeat_build_column<Sys_coh01_d3>(col - RL::NA - RL::NB - RL::NC_, pt, seg_row, NOISY,
                                           E + (long long)RL::COH01 * stride, stride);
  }
  E[(long long)(RL::STEPS + col) * stride] = st;
}

}  // namespace rydcz
