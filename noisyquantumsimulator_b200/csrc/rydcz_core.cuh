// rydcz_core.cuh -- per-evaluation arithmetic of the Rydberg-CZ hot path.
//
// One gate evaluation = the reference's simulate_CZ_gate steps 11-13
// (RG/simulation.py:3364-3572): evolve |00>,|01>,|10>,|11> through S constant-H
// segments (one QuTiP mesolve per input and segment, RG/simulation.py:689) and reduce to
// fidelities (compute_CZ_fidelity, RG/simulation.py:225-633).
//
// Method.  The Lindblad generator is linear in a few scalars and closes on a small operator
// subspace for each start operator (codegen/model.py).  Each such real system x' = G x
// (n = 25 for rho_11, 8 for the |11><00| coherence, 5/4 for the single-atom inputs) lives
// entirely in registers.  A constant segment is advanced exactly (to rounding) by the
// Chebyshev expansion of exp(tau G) on the imaginary-axis interval that Gershgorin bounds,
//
//     exp(tau G) x = J_0(z) x + 2 sum_{k>=1} J_k(z) phi_k(G/R) x,   z = R tau,
//     phi_0 = I, phi_1 = A, phi_{k+1} = 2 A phi_k + phi_{k-1}      (phi_k(A) = i^k T_k(A/i)),
//
// summed with Clenshaw's backward recurrence while Miller's backward recurrence produces the
// (unnormalised) Bessel coefficients on the fly, so no coefficient table is stored and the
// work is K ~ z + 10.3 z^(1/3) + 4 sparse matvecs per segment.  Cost is proportional to
// ||H|| tau, with no step-size control, no rejected steps and no dependence on the number of
// tlist points the reference requests.
//
// This header is plain C++ when RYDCZ_HD is empty, so tests/hostsim can compile the very same
// arithmetic with g++ (a test harness; the product path is the CUDA kernel only).
#pragma once
#include <math.h>

#ifndef RYDCZ_HD
#ifdef __CUDACC__
#define RYDCZ_HD __host__ __device__ __forceinline__
#else
#define RYDCZ_HD inline
#endif
#endif

#ifndef RYDCZ_UNROLL
#ifdef __CUDACC__
#define RYDCZ_UNROLL _Pragma("unroll")
#else
#define RYDCZ_UNROLL
#endif
#endif

// Scheduling fence for the big-system matvec: shared-memory loads may not be moved across a warp
// barrier, which bounds how many of them ptxas keeps in flight.  Lanes that have left the loop are
// not in the active mask.  No-op on the host.
// Reconverge the lanes of a warp that evaluate points (between two systems: lanes leave a system's
// loop at different iterations and must not start the next loop in separate groups).
#ifndef RYDCZ_RECONVERGE
#ifdef __CUDA_ARCH__
#define RYDCZ_RECONVERGE(mask) __syncwarp(mask)
#else
#define RYDCZ_RECONVERGE(mask) ((void)0)
#endif
#endif
#ifndef RYDCZ_SCHED_FENCE
#ifdef __CUDA_ARCH__
#define RYDCZ_SCHED_FENCE() __syncwarp(__activemask())
#else
#define RYDCZ_SCHED_FENCE() ((void)0)
#endif
#endif

namespace rydcz {

// a writable strided column (per-thread vector in shared memory; stride 1 on the host)
struct Col {
  double* base;
  int stride;
  RYDCZ_HD double& operator[](int k) const { return base[(long long)k * stride]; }
};
// One generator coefficient for the coefficient-major matvec of a big system.  The load is
// volatile so that the compiler issues it where it is written (41 short-lived values per matvec)
// instead of hoisting all coefficients of both Clenshaw half-steps into registers and spilling.
RYDCZ_HD double coef_load(const Col c, int q) { return *(const volatile double*)(c.base + (long long)q * c.stride); }
RYDCZ_HD double coef_load(const double* c, int q) { return c[q]; }

#include "rydcz_systems_d3.inc"
#include "rydcz_systems_d4.inc"
#include "rydcz_systems_d5.inc"

// ---- layout of one point row (doubles) -------------------------------------------------
// Everything the generator needs that is constant over the pulse.
enum PointSlot {
  PT_V = 0,        // blockade shift V (rad/s)                      RG/simulation.py:2958-2959
  PT_DELTA1 = 1,   // delta_zeeman + delta_stark on |1> (rad/s)      RG/simulation.py:3129-3159
  PT_OFF_A = 2,    // extra detuning seen by atom A (Doppler shot), rad/s
  PT_OFF_B = 3,    // extra detuning seen by atom B
  PT_G_R1 = 4,     // Gamma r->1 = gamma_r * b                       RG/noise_models.py:1237-1239
  PT_G_R0 = 5,     // Gamma r->0 = gamma_r(1-b)+bbr+antitrap+bg+leak  RG/noise_models.py:1242-1251,1583-1592
  PT_G_PHI = 6,    // dephasing on |r><r| (laser+thermal+zeeman)      RG/noise_models.py:1579-1580
  PT_G_SC = 7,     // scattering on |1><1|                            RG/noise_models.py:1589
  PT_G_MJ = 8,     // r+ <-> r- mixing (4-level only)                 RG/noise_models.py:1287-1295
  PT_SIG_DET = 9,  // Monte-Carlo: std of the per-atom Doppler detuning k_eff*v_rms (rad/s)
  PT_SIG_RELR = 10,// Monte-Carlo: std of dR/R0 = sqrt(2) sigma_r / R0
  PT_RESERVED = 11,
  PT_STRIDE = 12,
  // dim = 5 only (rows of PT5_STRIDE doubles): the intermediate state of the two-photon excitation
  PT5_OM1 = 12,     // first-leg Rabi frequency Omega_1 (rad/s)                RG/simulation.py:2910
  PT5_DET_E = 13,   // intermediate detuning Delta_e (rad/s)
  PT5_G_E1 = 14,    // decay |e> -> |1>
  PT5_G_E0 = 15,    // decay |e> -> |0>
  PT5_G_RL = 16,    // |r> -> |l>: the loss channels that leave the atom untrapped / outside the manifold
  PT5_LS_COMP = 17, // 1 = light shifts of both legs compensated (the reference's effective model), 0 = bare
  PT5_USED = 18,    // slots of a dim-5 row that carry data
  PT5_STRIDE = 20,
  PT_MAX_STRIDE = 18, // slots of the per-thread copy of a row, any dim (the scratch slots follow)
  PT_SCRATCH = 8    // extra per-thread slots behind a PointView (not part of the C-ABI row)
};
template <int D> struct PointStride { static constexpr int value = D == 5 ? PT5_STRIDE : PT_STRIDE; };
// one segment row: dt (s), Re Omega, Im Omega, Delta (rad/s)
enum SegSlot { SG_DT = 0, SG_RE = 1, SG_IM = 2, SG_DET = 3, SG_STRIDE = 4 };

// per-point output row
enum OutSlot {
  OUT_AVG_F = 0,      // avg_fidelity                         RG/simulation.py:631
  OUT_F00 = 1, OUT_F01 = 2, OUT_F10 = 3,
  OUT_F11 = 4,        // F11 population * cz_phase_fidelity   RG/simulation.py:619-629
  OUT_F11_POP = 5,
  OUT_CZ_PHASE_FID = 6,
  OUT_CTRL_PHASE = 7, // controlled phase in (-pi, pi]        RG/simulation.py:607-608
  OUT_PHASE_ERR = 8,  // min(|c-pi|, |c+pi|)                  RG/simulation.py:610
  OUT_PHI01 = 9, OUT_PHI10 = 10, OUT_PHI11 = 11,
  OUT_LOSS = 12,      // |11> input: probability that an atom ended in |0> (all 'loss' channels land there)
  OUT_RYD = 13,       // |11> input: probability that an atom is left in a Rydberg level
  OUT_STEPS = 14,     // Chebyshev matvecs performed (work counter for the FLOP accounting)
  OUT_FLOPS = 15,     // useful FP64 flops performed = sum 2*(NNZ+N)*steps over the systems
  OUT_STRIDE = 16
};

// raw reduced states, written when the caller asks for rho (compat epilogue, rho parity)
template <int D> struct StateLayout;
template <> struct StateLayout<3> {
  static constexpr int RHO11 = 0, RHO01 = 25, RHO10 = 30, COH11 = 35, COH01 = 43, COH10 = 47, STRIDE = 51;
};
template <> struct StateLayout<4> {
  static constexpr int RHO11 = 0, RHO01 = 36, RHO10 = 42, COH11 = 48, COH01 = 56, COH10 = 60, STRIDE = 64;
};
template <> struct StateLayout<5> {   // only the block in which no atom has been lost
  static constexpr int RHO11 = 0, RHO01 = 81, RHO10 = 90, COH11 = 99, COH01 = 117, COH10 = 123, STRIDE = 129;
};

// Strided read-only view of one point row: stride 1 for a plain row, blockDim for the
// transposed shared-memory copy the kernels keep (conflict-free, and the row does not occupy
// registers across the Chebyshev loop).
struct PointView {
  double* base;
  int stride;
  unsigned live = 0xffffffffu;   // lanes of this warp that evaluate a point (device: reconvergence mask)
  int sbase = PT_MAX_STRIDE;     // first scratch slot (kernels that only see 12-slot rows put it at PT_STRIDE)
  RYDCZ_HD double operator[](int k) const { return base[(long long)k * stride]; }
  // slots sbase .. sbase+PT_SCRATCH-1 of the same column: per-thread scratch that
  // must survive the Chebyshev loop without holding registers (the cached Chebyshev interval)
  RYDCZ_HD double& scratch(int k) const { return base[(long long)(sbase + k) * stride]; }
};

RYDCZ_HD int cheb_terms(double z) {
  // smallest even K with |J_k(z)| < ~1e-15 for all k > K (fit checked in tests/test_oracle_golden.py);
  // the single-precision cube root is padded so it never undershoots
  int K = (int)ceil(z + 10.3 * (double)(cbrtf((float)z) * 1.000001f) + 4.0);
  return K + (K & 1);
}
// The plain loop starts its backward recurrences at a multiple of 8 (more terms are always valid): the
// lanes of a warp then reach their segment transitions at common iterations, which matters for the short
// segments of the single-atom systems once the clustered path has made everything else cheap.
#ifndef RYDCZ_PLAIN_ALIGN
#define RYDCZ_PLAIN_ALIGN 8
#endif
RYDCZ_HD int cheb_terms_aligned(double z) { return (cheb_terms(z) + RYDCZ_PLAIN_ALIGN - 1) & ~(RYDCZ_PLAIN_ALIGN - 1); }

// Work estimate used to order points (same formula the evolution uses).
RYDCZ_HD double cheb_work(double radius, double dt) {
  double z = radius * fabs(dt);
  if (!(z > 0.0)) return 0.0;
  return z + 10.3 * (double)cbrtf((float)z) + 5.0;      // (an ordering key and a cost estimate: single precision will do)
}

// Interval-cache test for a detuning.  dim 5 adds the light-shift compensation |Omega_2|^2/4Delta_e to the
// detuning, which differs by rounding between equal-amplitude segments: compare to 1e-9 like |Omega|^2.
template <class Sys>
RYDCZ_HD bool same_detuning(double a, double b) {
  return Sys::DIM == 5 ? fabs(a - b) <= 1e-9 * fabs(b) || a == b : a == b;
}

// The generator's parameter vector p for one active segment row (re, im = Omega; det = Delta).
template <class Sys>
RYDCZ_HD void segment_params(double* __restrict__ p, const PointView pt, double re, double im, double det, bool noisy) {
  p[Sys::IDX_OM_RE] = re;
  p[Sys::IDX_OM_IM] = im;
  p[Sys::IDX_DET_A] = det + pt[PT_OFF_A];
  p[Sys::IDX_DET_B] = det + pt[PT_OFF_B];   // same slot as DET_A for exchange-folded systems
  p[Sys::IDX_V] = pt[PT_V];
  p[Sys::IDX_DELTA1] = pt[PT_DELTA1];
  p[Sys::IDX_G_R1] = noisy ? pt[PT_G_R1] : 0.0;
  p[Sys::IDX_G_R0] = noisy ? pt[PT_G_R0] : 0.0;
  p[Sys::IDX_G_PHI] = noisy ? pt[PT_G_PHI] : 0.0;
  constexpr int kSc = Sys::IDX_G_SC >= 0 ? Sys::IDX_G_SC : 0;
  if (Sys::IDX_G_SC >= 0) p[kSc] = noisy ? pt[PT_G_SC] : 0.0;
  constexpr int kMj = Sys::IDX_G_MJ >= 0 ? Sys::IDX_G_MJ : 0;
  if (Sys::IDX_G_MJ >= 0) p[kMj] = noisy ? pt[PT_G_MJ] : 0.0;
  if (Sys::DIM == 5) {
    // The segment row carries the effective two-photon Rabi frequency (so every pulse schedule of
    // the 3-level path applies unchanged); the second leg is Omega_2 = Omega_eff * 2 Delta_e / Omega_1
    // and, with PT5_LS_COMP = 1, the light shifts of both legs are compensated.
    constexpr int kOm1 = Sys::IDX_OM1 >= 0 ? Sys::IDX_OM1 : 0, kDe = Sys::IDX_DET_E >= 0 ? Sys::IDX_DET_E : 0,
                  kE1 = Sys::IDX_G_E1 >= 0 ? Sys::IDX_G_E1 : 0, kE0 = Sys::IDX_G_E0 >= 0 ? Sys::IDX_G_E0 : 0,
                  kRl = Sys::IDX_G_RL >= 0 ? Sys::IDX_G_RL : 0;
    const double om1 = pt[PT5_OM1], dete = pt[PT5_DET_E];
    const double k2 = 2.0 * dete / om1, ls = pt[PT5_LS_COMP] / (4.0 * dete);
    const double re2 = p[Sys::IDX_OM_RE] * k2, im2 = p[Sys::IDX_OM_IM] * k2;
    const double shift_r = ls * (re2 * re2 + im2 * im2);
    p[Sys::IDX_OM_RE] = re2; p[Sys::IDX_OM_IM] = im2;
    const double da = p[Sys::IDX_DET_A], db = p[Sys::IDX_DET_B];
    p[Sys::IDX_DET_A] = da + shift_r;
    if (Sys::IDX_DET_B != Sys::IDX_DET_A) p[Sys::IDX_DET_B] = db + shift_r;
    p[Sys::IDX_DELTA1] = pt[PT_DELTA1] - ls * (om1 * om1);
    p[kOm1] = om1; p[kDe] = dete;
    p[kE1] = noisy ? pt[PT5_G_E1] : 0.0;
    p[kE0] = noisy ? pt[PT5_G_E0] : 0.0;
    p[kRl] = noisy ? pt[PT5_G_RL] : 0.0;
  }
}

// A segment row is active when its duration is neither zero nor NaN.
RYDCZ_HD bool segment_active(double dt) { return dt != 0.0 && dt == dt; }

// Advance `s` to the next segment with a non-zero duration and fill the generator's parameter
// vector p for it; false when the schedule is exhausted.  The four fields of a row are loaded together
// (one memory latency per row, not one for the duration and a second one for the rest).
template <class Sys>
RYDCZ_HD bool next_segment(double* __restrict__ p, const PointView pt, const double* __restrict__ seg, int n_seg,
                           int& s, double& dt, bool noisy) {
  while (s < n_seg) {
    const double* sg = seg + (long long)s * SG_STRIDE;
    const double d = sg[SG_DT], re = sg[SG_RE], im = sg[SG_IM], det = sg[SG_DET];
    ++s;
    if (segment_active(d)) {
      dt = d;
      segment_params<Sys>(p, pt, re, im, det, noisy);
      return true;
    }
  }
  return false;
}

// Run one system through all segments of a point.
//
// The loop is FLAT: one iteration = two Clenshaw steps of whatever segment this lane is in.
// Segment (and sub-step) transitions happen inside the same loop under `if (k < 2)`, so the 32
// lanes of a warp never wait for one another at segment boundaries -- a lane whose first pulse
// segment is long and second short costs the same as one with the opposite split.  (With one
// loop per segment the warp pays sum_s max_lanes K_s instead of max_lanes sum_s K_s, which for
// bang-bang candidates was ~3x.)
//
// Per segment:  coefs -> Gershgorin interval [cen-R, cen+R] -> z = R h, K = cheb_terms(z) ->
// c <- (2/R) c.  dt < 0 (the reference reaches it when Delta_e < 0 makes Omega, hence tau,
// negative; mesolve then integrates backwards) applies exp(-|dt| G).  Strong damping pushes the
// spectrum off the imaginary axis, so a segment is cut into nsub pieces with damp*h <= 2, which
// keeps the Chebyshev polynomials O(e^2) ("Chebyshev interval").
template <class Sys>
RYDCZ_HD void run_system(double (&x)[Sys::N], const PointView pt,
                         const double* __restrict__ seg, int n_seg, bool noisy, double& steps,
                         bool from_start_operator = true) {
  constexpr int N = Sys::N;
  if (from_start_operator) Sys::init(x);   // false: x holds the vector to propagate (rydcz_repeat.cuh)
  double c[Sys::NC], b1[N], b2[N];
  double jp = 0.0, jk = 0.0, norm = 0.0, kd = 0.0, tz = 0.0, rot = 0.0;
  double &rot_total = pt.scratch(6), &rot_comp = pt.scratch(7);   // Kahan-summed shift phase
  rot_total = 0.0; rot_comp = 0.0;
  // cached Chebyshev interval, kept in the point's scratch slots (shared memory on the GPU)
  double &iv_om2 = pt.scratch(0), &iv_da = pt.scratch(1), &iv_db = pt.scratch(2), &iv_R = pt.scratch(3),
         &iv_cen = pt.scratch(4), &iv_damp = pt.scratch(5);
  iv_om2 = -1.0;
  int k = 0;         // Clenshaw index still to process in the active sub-step
  int K = 0;         // its starting value
  int sub_left = 0;  // further sub-steps of the current segment
  int s = 0;         // next segment row to load
  bool active = false;
  // Small systems (the single-atom inputs, the coherences) keep the NEXT segment row in registers, loaded one
  // transition ahead: their segments are short, and the two dependent global loads at every transition showed
  // up as long-scoreboard stalls once the clustered path had made everything else cheap (ncu, round 2).
  constexpr bool PREFETCH = Sys::N <= 10;
  double nrow[SG_STRIDE] = {0.0, 0.0, 0.0, 0.0};
  if (PREFETCH && n_seg > 0) {
    nrow[0] = seg[SG_DT]; nrow[1] = seg[SG_RE]; nrow[2] = seg[SG_IM]; nrow[3] = seg[SG_DET];
  }
  for (;;) {
    if (k < 2) {
      if (active) {
        // k = 0:  exp(hG) x = (J_0 x + A b1 + b2) / (J_0 + 2 sum_even J_k),  A b1 = (2A)(b1/2)
        RYDCZ_UNROLL
        for (int i = 0; i < N; ++i) { b1[i] *= 0.5; b2[i] = fma(jk, x[i], b2[i]); }
        Sys::apply(c, b1, b2);
        const double inv = 1.0 / (norm + jk);
        RYDCZ_UNROLL
        for (int i = 0; i < N; ++i) x[i] = b2[i] * inv;
        steps += (double)(K + 1);
        if (!Sys::HERMITIAN) {
          // the shift's phase factor commutes with the (complex-linear) evolution: accumulate the
          // angle (compensated sum: ~1000 terms for shaped pulses) and rotate once after the last segment
          const double yk = rot - rot_comp, tk = rot_total + yk;
          rot_comp = (tk - rot_total) - yk;
          rot_total = tk;
        }
        active = false;
      }
      if (sub_left > 0) {
        --sub_left;
      } else {
        double dt = 0.0;
        double p[Sys::NP];          // rebuilt per segment so it is not live across the hot loop
        if (PREFETCH) {
          // the row was loaded one transition ago; the loads for the row after it are issued now and are
          // consumed at the next transition (small systems only: eight registers live across the hot loop)
          bool found = false;
          double re = 0.0, im = 0.0, det = 0.0;
          while (s < n_seg && !found) {
            dt = nrow[0]; re = nrow[1]; im = nrow[2]; det = nrow[3];
            ++s;
            if (s < n_seg) {
              const double* sg = seg + (long long)s * SG_STRIDE;
              nrow[0] = sg[SG_DT]; nrow[1] = sg[SG_RE]; nrow[2] = sg[SG_IM]; nrow[3] = sg[SG_DET];
            }
            found = segment_active(dt);
          }
          if (!found) break;
          segment_params<Sys>(p, pt, re, im, det, noisy);
        } else if (!next_segment<Sys>(p, pt, seg, n_seg, s, dt, noisy)) {
          break;
        }
        const double sgn = dt < 0.0 ? -1.0 : 1.0;
        const double adt = fabs(dt);
        // Chebyshev interval.  The Gershgorin bound is taken with |Re Ω| = |Im Ω| = |Ω| (valid for
        // any laser phase, at most ~|Ω| wider), so it depends only on (|Ω|, Δ_A, Δ_B) and the point
        // constants: bang-bang, smooth-JP and LP segments differ only in the phase of Ω and reuse it.
        const double om2 = p[Sys::IDX_OM_RE] * p[Sys::IDX_OM_RE] + p[Sys::IDX_OM_IM] * p[Sys::IDX_OM_IM];
        // (|Ω|² of equal-amplitude segments differs by rounding: compare to 1e-9, far inside the
        //  slack of the bound)
        if (!(fabs(om2 - iv_om2) <= 1e-9 * iv_om2 && same_detuning<Sys>(p[Sys::IDX_DET_A], iv_da) &&
              same_detuning<Sys>(p[Sys::IDX_DET_B], iv_db))) {
          const double re = p[Sys::IDX_OM_RE], im = p[Sys::IDX_OM_IM];
          const double da = p[Sys::IDX_DET_A], db = p[Sys::IDX_DET_B];
          const double mag = sqrt(om2);
          p[Sys::IDX_OM_RE] = mag; p[Sys::IDX_OM_IM] = mag;
          if (!Sys::HERMITIAN) p[Sys::NP - 1] = 0.0;
          double lo, hi;
          Sys::coefs(p, c);
          Sys::bounds(c, lo, hi, iv_damp);
          if (Sys::HERMITIAN) { iv_cen = 0.0; iv_R = hi; }
          else { iv_cen = 0.5 * (lo + hi); iv_R = 0.5 * (hi - lo) + iv_damp; }
          p[Sys::IDX_OM_RE] = re; p[Sys::IDX_OM_IM] = im;
          iv_om2 = om2; iv_da = da; iv_db = db;
        }
        const double damp = iv_damp, cen = iv_cen;
        double R = iv_R;
        if (!Sys::HERMITIAN) p[Sys::NP - 1] = cen;
        Sys::coefs(p, c);
        int nsub = 1;
        if (damp * adt > 2.0) nsub = (int)ceil(0.5 * damp * adt);
        const double h = adt / (double)nsub;
        // a wider interval is always valid; never let z underflow the Miller recurrence
        if (!(R * h >= 1e-2)) R = 1e-2 / h;
        const double z = R * h;
        const double sc = sgn * 2.0 / R;
        RYDCZ_UNROLL
        for (int q = 0; q < Sys::NC; ++q) c[q] *= sc;
        K = cheb_terms_aligned(z);
        tz = 2.0 / z;
        rot = cen * h * sgn;
        sub_left = nsub - 1;
      }
      RYDCZ_UNROLL
      for (int i = 0; i < N; ++i) { b1[i] = 0.0; b2[i] = 0.0; }
      jp = 0.0; jk = 1e-30;   // J_{K+1}, J_K up to a common factor (Miller's backward recurrence)
      norm = 0.0;             // J_0 + 2 sum_{even k>0} J_k = 1 fixes the factor
      kd = (double)K;
      k = K;
      active = true;
    }
    // k even:   b2 <- 2 J_k x + (2A) b1 + b2
    double ak = jk + jk;
    RYDCZ_UNROLL
    for (int i = 0; i < N; ++i) b2[i] = fma(ak, x[i], b2[i]);
    Sys::apply(c, b1, b2);
    norm += ak;
    double jm = fma(kd * tz, jk, -jp); jp = jk; jk = jm; kd -= 1.0;
    // k-1 odd:  b1 <- 2 J_{k-1} x + (2A) b2 + b1
    ak = jk + jk;
    RYDCZ_UNROLL
    for (int i = 0; i < N; ++i) b1[i] = fma(ak, x[i], b1[i]);
    Sys::apply(c, b2, b1);
    jm = fma(kd * tz, jk, -jp); jp = jk; jk = jm; kd -= 1.0;
    k -= 2;
  }
  if (!Sys::HERMITIAN) {
    // undo the spectral shifts: every complex component times exp(-i sum_s cen_s h_s)
    double sn, cs;
    sincos(rot_total, &sn, &cs);
    RYDCZ_UNROLL
    for (int i = 0; i < N; i += 2) {
      const double re = x[i], im = x[i + 1];
      x[i] = fma(re, cs, im * sn);
      x[i + 1] = fma(im, cs, -re * sn);
    }
  }
}

// run_system for a system too large for registers (rho_11 of the 5-level model: 45 real variables).
// The two Clenshaw vectors b1, b2 stay in registers; the segment's start vector x and the
// generator's coefficients c live in per-thread shared-memory columns.  x is read once per Clenshaw
// step (the 2 J_k x term) and every c[q] once per matvec (Sys::apply_cm walks the nonzeros
// coefficient by coefficient), so shared-memory traffic stays ~4x below the FMA count.
// Hermitian systems only (no spectral shift).
template <class Sys, class XV, class CV>
RYDCZ_HD void run_system_big(const XV x, const CV c, const PointView pt, const double* __restrict__ seg,
                             int n_seg, bool noisy, double& steps) {
  constexpr int N = Sys::N;
  static_assert(Sys::HERMITIAN, "run_system_big: density-matrix systems only");
  {
    double x0[N];
    Sys::init(x0);
    RYDCZ_UNROLL
    for (int i = 0; i < N; ++i) x[i] = x0[i];
  }
  double b1[N], b2[N];
  double jp = 0.0, jk = 0.0, norm = 0.0, kd = 0.0, tz = 0.0;
  double &iv_om2 = pt.scratch(0), &iv_da = pt.scratch(1), &iv_db = pt.scratch(2), &iv_R = pt.scratch(3),
         &iv_damp = pt.scratch(5);
  iv_om2 = -1.0;
  int k = 0, K = 0, sub_left = 0, s = 0;
  bool active = false;
  for (;;) {
    if (k < 2) {
      if (active) {
        RYDCZ_UNROLL
        for (int i = 0; i < N; ++i) { b1[i] *= 0.5; b2[i] = fma(jk, x[i], b2[i]); }
        Sys::apply_cm(c, b1, b2);
        const double inv = 1.0 / (norm + jk);
        RYDCZ_UNROLL
        for (int i = 0; i < N; ++i) x[i] = b2[i] * inv;
        steps += (double)(K + 1);
        active = false;
      }
      if (sub_left > 0) {
        --sub_left;
      } else {
        double dt = 0.0;
        double p[Sys::NP];
        if (!next_segment<Sys>(p, pt, seg, n_seg, s, dt, noisy)) break;
        const double sgn = dt < 0.0 ? -1.0 : 1.0;
        const double adt = fabs(dt);
        const double om2 = p[Sys::IDX_OM_RE] * p[Sys::IDX_OM_RE] + p[Sys::IDX_OM_IM] * p[Sys::IDX_OM_IM];
        if (!(fabs(om2 - iv_om2) <= 1e-9 * iv_om2 && same_detuning<Sys>(p[Sys::IDX_DET_A], iv_da) &&
              same_detuning<Sys>(p[Sys::IDX_DET_B], iv_db))) {
          const double re = p[Sys::IDX_OM_RE], im = p[Sys::IDX_OM_IM];
          const double mag = sqrt(om2);
          p[Sys::IDX_OM_RE] = mag; p[Sys::IDX_OM_IM] = mag;
          double lo, hi;
          Sys::coefs(p, c);
          Sys::bounds(c, lo, hi, iv_damp);
          iv_R = hi;
          p[Sys::IDX_OM_RE] = re; p[Sys::IDX_OM_IM] = im;
          iv_om2 = om2; iv_da = p[Sys::IDX_DET_A]; iv_db = p[Sys::IDX_DET_B];
        }
        const double damp = iv_damp;
        double R = iv_R;
        Sys::coefs(p, c);
        int nsub = 1;
        if (damp * adt > 2.0) nsub = (int)ceil(0.5 * damp * adt);
        const double h = adt / (double)nsub;
        if (!(R * h >= 1e-2)) R = 1e-2 / h;
        const double z = R * h;
        const double sc = sgn * 2.0 / R;
        RYDCZ_UNROLL
        for (int q = 0; q < Sys::NC; ++q) c[q] *= sc;
        K = cheb_terms_aligned(z);
        tz = 2.0 / z;
        sub_left = nsub - 1;
      }
      RYDCZ_UNROLL
      for (int i = 0; i < N; ++i) { b1[i] = 0.0; b2[i] = 0.0; }
      jp = 0.0; jk = 1e-30;
      norm = 0.0;
      kd = (double)K;
      k = K;
      active = true;
    }
    double ak = jk + jk;
    Sys::apply_cm_axpy(c, x, ak, b1, b2);
    norm += ak;
    double jm = fma(kd * tz, jk, -jp); jp = jk; jk = jm; kd -= 1.0;
    ak = jk + jk;
    Sys::apply_cm_axpy(c, x, ak, b2, b1);
    jm = fma(kd * tz, jk, -jp); jp = jk; jk = jm; kd -= 1.0;
    k -= 2;
  }
}

// run_system for a system whose Clenshaw vectors do not fit in registers at all (the general rho_11 of the 5-level
// model: 81 real variables, unequal atoms / Monte-Carlo shots).  The start vector x and both Clenshaw vectors live in
// per-thread shared-memory columns; the 45 coefficients stay in registers and the matvec walks the rows
// (Sys::apply_rows_axpy: one accumulator, one shared-memory load per nonzero, one store per row).  Same arithmetic as
// run_system_big; an order of magnitude slower (two warps per SM, a load per FMA) -- it exists so that the 5-level
// model is complete, the identical-atom kernel is the fast path.
template <class Sys, class XV>
RYDCZ_HD void run_system_huge(const XV x, const XV b1, const XV b2, const PointView pt, const double* __restrict__ seg,
                              int n_seg, bool noisy, double& steps) {
  constexpr int N = Sys::N;
  static_assert(Sys::HERMITIAN, "run_system_huge: density-matrix systems only");
  {
    double x0[N];
    Sys::init(x0);
    RYDCZ_UNROLL
    for (int i = 0; i < N; ++i) x[i] = x0[i];
  }
  double c[Sys::NC];
  double jp = 0.0, jk = 0.0, norm = 0.0, kd = 0.0, tz = 0.0;
  double &iv_om2 = pt.scratch(0), &iv_da = pt.scratch(1), &iv_db = pt.scratch(2), &iv_R = pt.scratch(3),
         &iv_damp = pt.scratch(5);
  iv_om2 = -1.0;
  int k = 0, K = 0, sub_left = 0, s = 0;
  bool active = false;
  for (;;) {
    if (k < 2) {
      if (active) {
        // k = 0:  exp(hG) x = (J_0 x + A b1 + b2) / (J_0 + 2 sum_even J_k),  A b1 = (2A)(b1/2)
        for (int i = 0; i < N; ++i) b1[i] *= 0.5;
        Sys::apply_rows_axpy(c, x, jk, b1, b2);
        const double inv = 1.0 / (norm + jk);
        for (int i = 0; i < N; ++i) x[i] = b2[i] * inv;
        steps += (double)(K + 1);
        active = false;
      }
      if (sub_left > 0) {
        --sub_left;
      } else {
        double dt = 0.0;
        double p[Sys::NP];
        if (!next_segment<Sys>(p, pt, seg, n_seg, s, dt, noisy)) break;
        const double sgn = dt < 0.0 ? -1.0 : 1.0;
        const double adt = fabs(dt);
        const double om2 = p[Sys::IDX_OM_RE] * p[Sys::IDX_OM_RE] + p[Sys::IDX_OM_IM] * p[Sys::IDX_OM_IM];
        if (!(fabs(om2 - iv_om2) <= 1e-9 * iv_om2 && same_detuning<Sys>(p[Sys::IDX_DET_A], iv_da) &&
              same_detuning<Sys>(p[Sys::IDX_DET_B], iv_db))) {
          const double re = p[Sys::IDX_OM_RE], im = p[Sys::IDX_OM_IM];
          const double mag = sqrt(om2);
          p[Sys::IDX_OM_RE] = mag; p[Sys::IDX_OM_IM] = mag;
          double lo, hi;
          Sys::coefs(p, c);
          Sys::bounds(c, lo, hi, iv_damp);
          iv_R = hi;
          p[Sys::IDX_OM_RE] = re; p[Sys::IDX_OM_IM] = im;
          iv_om2 = om2; iv_da = p[Sys::IDX_DET_A]; iv_db = p[Sys::IDX_DET_B];
        }
        const double damp = iv_damp;
        double R = iv_R;
        Sys::coefs(p, c);
        int nsub = 1;
        if (damp * adt > 2.0) nsub = (int)ceil(0.5 * damp * adt);
        const double h = adt / (double)nsub;
        if (!(R * h >= 1e-2)) R = 1e-2 / h;
        const double z = R * h;
        const double sc = sgn * 2.0 / R;
        RYDCZ_UNROLL
        for (int q = 0; q < Sys::NC; ++q) c[q] *= sc;
        K = cheb_terms_aligned(z);
        tz = 2.0 / z;
        sub_left = nsub - 1;
      }
      for (int i = 0; i < N; ++i) { b1[i] = 0.0; b2[i] = 0.0; }
      jp = 0.0; jk = 1e-30;
      norm = 0.0;
      kd = (double)K;
      k = K;
      active = true;
    }
    double ak = jk + jk;
    Sys::apply_rows_axpy(c, x, ak, b1, b2);
    norm += ak;
    double jm = fma(kd * tz, jk, -jp); jp = jk; jk = jm; kd -= 1.0;
    ak = jk + jk;
    Sys::apply_rows_axpy(c, x, ak, b2, b1);
    jm = fma(kd * tz, jk, -jp); jp = jk; jk = jm; kd -= 1.0;
    k -= 2;
  }
}

// ---- closed forms for the |ab><00| coherences of the 3-level model (rydcz_closed.cuh) --------------------------
#ifndef RYDCZ_CLOSED_COHERENCE
#define RYDCZ_CLOSED_COHERENCE 1     // 0: every system on the Chebyshev loops (A/B builds, round-1 behaviour)
#endif
#ifndef RYDCZ_CLOSED_RHO01
#define RYDCZ_CLOSED_RHO01 1         // 0: single-atom rho on its Chebyshev loop (A/B builds)
#endif
#include "rydcz_closed.cuh"

// The exchange-folded |11><00| coherence in closed form, where there is one: false = not available for this model
// or a non-finite value appeared -- the caller runs the system's Chebyshev loop instead.
template <int D> struct ClosedCoh11s {
  static constexpr bool available = false;
  template <bool NOISY, int N>
  RYDCZ_HD static bool run(double (&)[N], const PointView, const double* __restrict__, int, double&) { return false; }
};
template <> struct ClosedCoh11s<3> {
  static constexpr bool available = RYDCZ_CLOSED_COHERENCE != 0;
  template <bool NOISY>
  RYDCZ_HD static bool run(double (&x)[Sys_coh11s_d3::N], const PointView pt, const double* __restrict__ seg, int n_seg,
                           double& flops) {
    return available && coh11s_closed_d3<NOISY>(x, pt, seg, n_seg, flops);
  }
};

// ... and the |11><00| coherence of unequal atoms (4 x 4; declines outside the strong-blockade, small-offset regime)
template <int D> struct ClosedCoh11 {
  template <bool NOISY, int N>
  RYDCZ_HD static bool run(double (&)[N], const PointView, const double* __restrict__, int, double&) { return false; }
};
template <> struct ClosedCoh11<3> {
  template <bool NOISY>
  RYDCZ_HD static bool run(double (&x)[Sys_coh11_d3::N], const PointView pt, const double* __restrict__ seg, int n_seg,
                           double& flops) {
    return RYDCZ_CLOSED_COHERENCE != 0 && coh11_closed_d3<NOISY>(x, pt, seg, n_seg, flops);
  }
};

// A single-atom coherence |01><00| (atom_b) or |10><00|: closed form for the 3-level model, else (or when that
// fails) the Chebyshev loop of the system.  Every live lane of the warp reaches the reconvergence point.
template <class SysCoh, bool NOISY>
RYDCZ_HD void single_atom_coherence(double (&y)[SysCoh::N], const PointView pt, const double* __restrict__ seg,
                                    int n_seg, bool atom_b, double& steps, double& flops) {
  bool done = false;
  if constexpr (SysCoh::DIM == 3 && RYDCZ_CLOSED_COHERENCE != 0) {
    double cf = 0.0;
    done = coh01_closed_d3<NOISY>(y, pt, seg, n_seg, atom_b, cf);
    if (done) flops += cf;
  }
  if (!done) {
    double st = 0.0;
    run_system<SysCoh>(y, pt, seg, n_seg, NOISY, st);
    steps += st; flops += st * 2.0 * (SysCoh::NNZ + SysCoh::N);
  }
  RYDCZ_RECONVERGE(pt.live);
}


// A single-atom density operator rho_01 (atom_b) or rho_10: closed form for the 3-level model (optical Bloch
// equations, rydcz_closed.cuh), else (or when that declines) the Chebyshev loop of the system.
template <class SysRho>
RYDCZ_HD void single_atom_rho(double (&x)[SysRho::N], const PointView pt, const double* __restrict__ seg,
                              int n_seg, bool atom_b, double& steps, double& flops) {
  bool done = false;
  if constexpr (SysRho::DIM == 3 && RYDCZ_CLOSED_COHERENCE != 0 && RYDCZ_CLOSED_RHO01 != 0) {
    double cf = 0.0;
    done = rho01_closed_d3(x, pt, seg, n_seg, atom_b, cf);
    if (done) flops += cf;
  }
  if (!done) {
    double st = 0.0;
    run_system<SysRho>(x, pt, seg, n_seg, true, st);
    steps += st; flops += st * 2.0 * (SysRho::NNZ + SysRho::N);
  }
  RYDCZ_RECONVERGE(pt.live);
}

RYDCZ_HD double wrap_pi(double x) {
  // Python: (x + pi) % (2 pi) - pi, with % taking the sign of the divisor  (RG/simulation.py:608)
  const double twopi = 6.283185307179586476925286766559;
  const double pi = 3.141592653589793238462643383279;
  double y = x + pi;
  y = y - twopi * floor(y / twopi);
  return y - pi;
}

// Fidelity epilogue (compute_CZ_fidelity, RG/simulation.py:225-633).
//   noiseless: pure-state branch -- F_ab = |<ab|psi_ab>|^2, phases from the overlaps.
//   noisy    : populations F_ab = <ab|rho_ab|ab>; phases from the coherences <ab|E(|ab><00|)|00>
//              (gauge-free replacement for the reference's dominant-eigenvector phases; the
//              reference's own number is recomputed on the host from rho in 'compat' mode).
RYDCZ_HD void cz_epilogue(bool noisy, double p01, double p10, double p11,
                          double c01r, double c01i, double c10r, double c10i, double c11r, double c11i,
                          double* __restrict__ out) {
  const double pi = 3.141592653589793238462643383279;
  double f00 = 1.0, f01, f10, f11;
  if (noisy) { f01 = p01; f10 = p10; f11 = p11; }
  else {
    f01 = c01r * c01r + c01i * c01i;
    f10 = c10r * c10r + c10i * c10i;
    f11 = c11r * c11r + c11i * c11i;
  }
  const double phi01 = atan2(c01i, c01r), phi10 = atan2(c10i, c10r), phi11 = atan2(c11i, c11r);
  const double ctrl = wrap_pi(phi11 - phi01 - phi10 + 0.0);
  const double err = fmin(fabs(ctrl - pi), fabs(ctrl + pi));
  const double ch = cos(0.5 * err);
  const double czf = ch * ch;
  out[OUT_F00] = f00; out[OUT_F01] = f01; out[OUT_F10] = f10;
  out[OUT_F11_POP] = f11;
  out[OUT_F11] = f11 * czf;
  out[OUT_CZ_PHASE_FID] = czf;
  out[OUT_CTRL_PHASE] = ctrl;
  out[OUT_PHASE_ERR] = err;
  out[OUT_PHI01] = phi01; out[OUT_PHI10] = phi10; out[OUT_PHI11] = phi11;
  out[OUT_AVG_F] = 0.25 * (f00 + f01 + f10 + f11 * czf);
}

// One full gate evaluation.  `states` (may be null) receives the raw reduced states.
#define RYDCZ_DEFINE_EVALUATE(D)                                                                    \
  template <bool NOISY>                                                                             \
  RYDCZ_HD void evaluate_point_d##D(const PointView pt, const double* __restrict__ seg,            \
                                    int n_seg, double* __restrict__ out,                            \
                                    double* __restrict__ states) {                                  \
    typedef StateLayout<D> SL;                                                                      \
    double steps = 0.0, flops = 0.0, st;                                                            \
    double p01 = 1.0, p10 = 1.0, p11 = 1.0, loss = 0.0, ryd = 0.0;                                  \
    double c01r, c01i, c10r, c10i, c11r, c11i;                                                      \
    if (NOISY) {                                                                                    \
      double x[Sys_rho11_d##D::N]; st = 0.0;                                                        \
      run_system<Sys_rho11_d##D>(x, pt, seg, n_seg, true, st); RYDCZ_RECONVERGE(pt.live);           \
      steps += st; flops += st * 2.0 * (Sys_rho11_d##D::NNZ + Sys_rho11_d##D::N);                   \
      p11 = x[RHO11_POP_D##D];                                                                      \
      loss = rho11_loss_d##D(x); ryd = rho11_ryd_d##D(x);                                           \
      if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_rho11_d##D::N; ++i) states[SL::RHO11 + i] = x[i]; } \
    }                                                                                               \
    {                                                                                               \
      double x[Sys_coh11_d##D::N], cf = 0.0; st = 0.0;                                              \
      if (ClosedCoh11<D>::template run<NOISY>(x, pt, seg, n_seg, cf)) flops += cf;                  \
      else {                                                                                        \
        run_system<Sys_coh11_d##D>(x, pt, seg, n_seg, NOISY, st);                                   \
        steps += st; flops += st * 2.0 * (Sys_coh11_d##D::NNZ + Sys_coh11_d##D::N);                 \
      }                                                                                             \
      RYDCZ_RECONVERGE(pt.live);                                                                    \
      c11r = x[COH11_RE_D##D]; c11i = x[COH11_IM_D##D];                                             \
      if (!NOISY) ryd = 1.0 - (c11r * c11r + c11i * c11i);                                          \
      if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh11_d##D::N; ++i) states[SL::COH11 + i] = x[i]; } \
    }                                                                                               \
    if (NOISY && ClosedCoh11s<D>::available) {                                                      \
      /* single-atom inputs: rho_ab on its Chebyshev loop, its |ab><00| coherence in closed form */   \
      double x[Sys_rho01_d##D::N];                                                                  \
      single_atom_rho<Sys_rho01_d##D>(x, pt, seg, n_seg, true, steps, flops);                       \
      p01 = x[RHO01_POP_D##D];                                                                      \
      double xc[Sys_coh01_d##D::N];                                                                 \
      single_atom_coherence<Sys_coh01_d##D, true>(xc, pt, seg, n_seg, true, steps, flops);          \
      c01r = xc[COH01_RE_D##D]; c01i = xc[COH01_IM_D##D];                                           \
      double y[Sys_rho10_d##D::N];                                                                  \
      single_atom_rho<Sys_rho10_d##D>(y, pt, seg, n_seg, false, steps, flops);                      \
      p10 = y[RHO10_POP_D##D];                                                                      \
      double yc[Sys_coh10_d##D::N];                                                                 \
      single_atom_coherence<Sys_coh10_d##D, true>(yc, pt, seg, n_seg, false, steps, flops);         \
      c10r = yc[COH10_RE_D##D]; c10i = yc[COH10_IM_D##D];                                           \
      if (states) {                                                                                 \
        RYDCZ_UNROLL for (int i = 0; i < Sys_rho01_d##D::N; ++i) states[SL::RHO01 + i] = x[i];      \
        RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d##D::N; ++i) states[SL::COH01 + i] = xc[i];     \
        RYDCZ_UNROLL for (int i = 0; i < Sys_rho10_d##D::N; ++i) states[SL::RHO10 + i] = y[i];      \
        RYDCZ_UNROLL for (int i = 0; i < Sys_coh10_d##D::N; ++i) states[SL::COH10 + i] = yc[i];     \
      }                                                                                             \
    } else if (NOISY) {                                                                             \
      /* single-atom inputs: rho_ab and its |ab><00| coherence share one Chebyshev loop */           \
      double x[Sys_in01_d##D::N]; st = 0.0;                                                         \
      run_system<Sys_in01_d##D>(x, pt, seg, n_seg, true, st); RYDCZ_RECONVERGE(pt.live);            \
      steps += st; flops += st * 2.0 * (Sys_in01_d##D::NNZ + Sys_in01_d##D::N);                     \
      p01 = x[RHO01_POP_D##D];                                                                      \
      c01r = x[IN01_COH_OFFSET_D##D + COH01_RE_D##D]; c01i = x[IN01_COH_OFFSET_D##D + COH01_IM_D##D]; \
      if (states) {                                                                                 \
        RYDCZ_UNROLL for (int i = 0; i < Sys_rho01_d##D::N; ++i) states[SL::RHO01 + i] = x[i];      \
        RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d##D::N; ++i) states[SL::COH01 + i] = x[IN01_COH_OFFSET_D##D + i]; \
      }                                                                                             \
      double y[Sys_in10_d##D::N]; st = 0.0;                                                         \
      run_system<Sys_in10_d##D>(y, pt, seg, n_seg, true, st); RYDCZ_RECONVERGE(pt.live);            \
      steps += st; flops += st * 2.0 * (Sys_in10_d##D::NNZ + Sys_in10_d##D::N);                     \
      p10 = y[RHO10_POP_D##D];                                                                      \
      c10r = y[IN10_COH_OFFSET_D##D + COH10_RE_D##D]; c10i = y[IN10_COH_OFFSET_D##D + COH10_IM_D##D]; \
      if (states) {                                                                                 \
        RYDCZ_UNROLL for (int i = 0; i < Sys_rho10_d##D::N; ++i) states[SL::RHO10 + i] = y[i];      \
        RYDCZ_UNROLL for (int i = 0; i < Sys_coh10_d##D::N; ++i) states[SL::COH10 + i] = y[IN10_COH_OFFSET_D##D + i]; \
      }                                                                                             \
    } else {                                                                                        \
      double x[Sys_coh01_d##D::N];                                                                  \
      single_atom_coherence<Sys_coh01_d##D, false>(x, pt, seg, n_seg, true, steps, flops);          \
      c01r = x[COH01_RE_D##D]; c01i = x[COH01_IM_D##D];                                             \
      if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d##D::N; ++i) states[SL::COH01 + i] = x[i]; } \
      double y[Sys_coh10_d##D::N];                                                                  \
      single_atom_coherence<Sys_coh10_d##D, false>(y, pt, seg, n_seg, false, steps, flops);         \
      c10r = y[COH10_RE_D##D]; c10i = y[COH10_IM_D##D];                                             \
      if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh10_d##D::N; ++i) states[SL::COH10 + i] = y[i]; } \
    }                                                                                               \
    cz_epilogue(NOISY, p01, p10, p11, c01r, c01i, c10r, c10i, c11r, c11i, out);                     \
    out[OUT_LOSS] = loss; out[OUT_RYD] = ryd; out[OUT_STEPS] = steps; out[OUT_FLOPS] = flops;       \
  }

RYDCZ_DEFINE_EVALUATE(3)
RYDCZ_DEFINE_EVALUATE(4)

// Identical atoms (off_a == off_b): rho_11 and the |11><00| coherence are invariant under swapping
// the atoms, so they are evolved in the folded variables (15 instead of 25, 6 instead of 8; fewer
// than half the FMAs per step) and expanded at the end; the |10> input equals the |01> input.
// Same outputs, same `states` layout as the general path.
#define RYDCZ_DEFINE_EVALUATE_SYM(D)                                                                \
  template <bool NOISY>                                                                             \
  RYDCZ_HD void evaluate_point_sym_d##D(const PointView pt, const Col xcol,                         \
                                        const double* __restrict__ seg, int n_seg,                  \
                                        double* __restrict__ out, double* __restrict__ states) {    \
    typedef StateLayout<D> SL;                                                                      \
    double steps = 0.0, flops = 0.0, st;                                                            \
    double p01 = 1.0, p11 = 1.0, loss = 0.0, ryd = 0.0;                                             \
    double c01r, c01i, c11r, c11i;                                                                  \
    if (NOISY) {                                                                                    \
      double x[Sys_rho11s_d##D::N]; st = 0.0;                                                       \
      run_rho11s_d##D(x, xcol, pt, seg, n_seg, st);                                                 \
      steps += st; flops += st * 2.0 * (Sys_rho11s_d##D::NNZ + Sys_rho11s_d##D::N);                 \
      p11 = x[RHO11S_POP_D##D];                                                                     \
      double xf[Sys_rho11s_d##D::FULL_N];                                                           \
      Sys_rho11s_d##D::expand(x, xf);                                                               \
      loss = rho11_loss_d##D(xf); ryd = rho11_ryd_d##D(xf);                                         \
      if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_rho11s_d##D::FULL_N; ++i) states[SL::RHO11 + i] = xf[i]; } \
    }                                                                                               \
    {                                                                                               \
      double x[Sys_coh11s_d##D::N], cf = 0.0; st = 0.0;                                             \
      if (ClosedCoh11s<D>::template run<NOISY>(x, pt, seg, n_seg, cf)) flops += cf;                 \
      else {                                                                                        \
        run_system<Sys_coh11s_d##D>(x, pt, seg, n_seg, NOISY, st);                                  \
        steps += st; flops += st * 2.0 * (Sys_coh11s_d##D::NNZ + Sys_coh11s_d##D::N);               \
      }                                                                                             \
      RYDCZ_RECONVERGE(pt.live);                                                                    \
      c11r = x[COH11S_RE_D##D]; c11i = x[COH11S_IM_D##D];                                           \
      if (!NOISY) ryd = 1.0 - (c11r * c11r + c11i * c11i);                                          \
      if (states) {                                                                                 \
        double xf[Sys_coh11s_d##D::FULL_N];                                                         \
        Sys_coh11s_d##D::expand(x, xf);                                                             \
        RYDCZ_UNROLL for (int i = 0; i < Sys_coh11s_d##D::FULL_N; ++i) states[SL::COH11 + i] = xf[i]; \
      }                                                                                             \
    }                                                                                               \
    if (NOISY && ClosedCoh11s<D>::available) {                                                      \
      double x[Sys_rho01_d##D::N];                                                                  \
      single_atom_rho<Sys_rho01_d##D>(x, pt, seg, n_seg, true, steps, flops);                       \
      p01 = x[RHO01_POP_D##D];                                                                      \
      double xc[Sys_coh01_d##D::N];                                                                 \
      single_atom_coherence<Sys_coh01_d##D, true>(xc, pt, seg, n_seg, true, steps, flops);          \
      c01r = xc[COH01_RE_D##D]; c01i = xc[COH01_IM_D##D];                                           \
      if (states) {                                                                                 \
        RYDCZ_UNROLL for (int i = 0; i < Sys_rho01_d##D::N; ++i) {                                  \
          states[SL::RHO01 + i] = x[i]; states[SL::RHO10 + i] = x[i]; }                             \
        RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d##D::N; ++i) {                                  \
          states[SL::COH01 + i] = xc[i]; states[SL::COH10 + i] = xc[i]; }                           \
      }                                                                                             \
    } else if (NOISY) {                                                                             \
      double x[Sys_in01_d##D::N]; st = 0.0;                                                         \
      run_system<Sys_in01_d##D>(x, pt, seg, n_seg, true, st); RYDCZ_RECONVERGE(pt.live);            \
      steps += st; flops += st * 2.0 * (Sys_in01_d##D::NNZ + Sys_in01_d##D::N);                     \
      p01 = x[RHO01_POP_D##D];                                                                      \
      c01r = x[IN01_COH_OFFSET_D##D + COH01_RE_D##D]; c01i = x[IN01_COH_OFFSET_D##D + COH01_IM_D##D]; \
      if (states) {                                                                                 \
        RYDCZ_UNROLL for (int i = 0; i < Sys_rho01_d##D::N; ++i) {                                  \
          states[SL::RHO01 + i] = x[i]; states[SL::RHO10 + i] = x[i]; }                             \
        RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d##D::N; ++i) {                                  \
          states[SL::COH01 + i] = x[IN01_COH_OFFSET_D##D + i]; states[SL::COH10 + i] = x[IN01_COH_OFFSET_D##D + i]; } \
      }                                                                                             \
    } else {                                                                                        \
      double x[Sys_coh01_d##D::N];                                                                  \
      single_atom_coherence<Sys_coh01_d##D, false>(x, pt, seg, n_seg, true, steps, flops);          \
      c01r = x[COH01_RE_D##D]; c01i = x[COH01_IM_D##D];                                             \
      if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d##D::N; ++i) {                      \
        states[SL::COH01 + i] = x[i]; states[SL::COH10 + i] = x[i]; } }                             \
    }                                                                                               \
    cz_epilogue(NOISY, p01, p01, p11, c01r, c01i, c01r, c01i, c11r, c11i, out);                     \
    out[OUT_LOSS] = loss; out[OUT_RYD] = ryd; out[OUT_STEPS] = steps; out[OUT_FLOPS] = flops;       \
  }

// rho_11 of the folded 3-level model.  RYDCZ_D3_VARIANT: 0 = everything in registers (run_system),
// 1 = coefficients in the shared-memory column, 2 = start vector in the shared-memory column.
#ifndef RYDCZ_D3_VARIANT
#define RYDCZ_D3_VARIANT 0
#endif
RYDCZ_HD void run_rho11s_d3(double (&x)[Sys_rho11s_d3::N], const Col col, const PointView pt,
                            const double* __restrict__ seg, int n_seg, double& st) {
#if RYDCZ_D3_VARIANT == 1
  run_system_big<Sys_rho11s_d3, double*, Col>(x, col, pt, seg, n_seg, true, st);
#elif RYDCZ_D3_VARIANT == 2
  double c[Sys_rho11s_d3::NC];
  run_system_big<Sys_rho11s_d3, Col, double*>(col, c, pt, seg, n_seg, true, st);
  RYDCZ_UNROLL for (int i = 0; i < Sys_rho11s_d3::N; ++i) x[i] = col[i];
#else
  run_system<Sys_rho11s_d3>(x, pt, seg, n_seg, true, st);
#endif
  RYDCZ_RECONVERGE(pt.live);
}
RYDCZ_HD void run_rho11s_d4(double (&x)[Sys_rho11s_d4::N], const Col, const PointView pt,
                            const double* __restrict__ seg, int n_seg, double& st) {
  run_system<Sys_rho11s_d4>(x, pt, seg, n_seg, true, st);
  RYDCZ_RECONVERGE(pt.live);
}
RYDCZ_DEFINE_EVALUATE_SYM(3)
RYDCZ_DEFINE_EVALUATE_SYM(4)

// dim = 5 (identical atoms only): rho_11 through run_system_big with its start vector and
// coefficients in the shared-memory columns xcol / ccol, everything else in registers as above.
template <bool NOISY>
RYDCZ_HD void evaluate_point_sym_d5(const PointView pt, const Col xcol, const Col ccol,
                                    const double* __restrict__ seg, int n_seg, double* __restrict__ out,
                                    double* __restrict__ states) {
  typedef StateLayout<5> SL;
  double steps = 0.0, flops = 0.0, st;
  double p01 = 1.0, p11 = 1.0, loss = 0.0, ryd = 0.0;
  double c01r, c01i, c11r, c11i;
  if (NOISY) {
    st = 0.0;
    run_system_big<Sys_rho11s_d5, Col, Col>(xcol, ccol, pt, seg, n_seg, true, st);
    RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_rho11s_d5::NNZ + Sys_rho11s_d5::N);
    double x[Sys_rho11s_d5::N];
    RYDCZ_UNROLL for (int i = 0; i < Sys_rho11s_d5::N; ++i) x[i] = xcol[i];
    p11 = x[RHO11S_POP_D5];
    double xf[Sys_rho11s_d5::FULL_N];
    Sys_rho11s_d5::expand(x, xf);
    loss = rho11_loss_d5(xf); ryd = rho11_ryd_d5(xf);
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_rho11s_d5::FULL_N; ++i) states[SL::RHO11 + i] = xf[i]; }
  }
  {
    double x[Sys_coh11s_d5::N]; st = 0.0;
    run_system<Sys_coh11s_d5>(x, pt, seg, n_seg, NOISY, st);
    RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_coh11s_d5::NNZ + Sys_coh11s_d5::N);
    c11r = x[COH11S_RE_D5]; c11i = x[COH11S_IM_D5];
    if (!NOISY) {
      double xf[Sys_coh11s_d5::FULL_N];
      Sys_coh11s_d5::expand(x, xf);
      ryd = coh11_ryd_d5(xf);
    }
    if (states) {
      double xf[Sys_coh11s_d5::FULL_N];
      Sys_coh11s_d5::expand(x, xf);
      RYDCZ_UNROLL for (int i = 0; i < Sys_coh11s_d5::FULL_N; ++i) states[SL::COH11 + i] = xf[i];
    }
  }
  if (NOISY) {
    double x[Sys_in01_d5::N]; st = 0.0;
    run_system<Sys_in01_d5>(x, pt, seg, n_seg, true, st);
    RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_in01_d5::NNZ + Sys_in01_d5::N);
    p01 = x[RHO01_POP_D5];
    c01r = x[IN01_COH_OFFSET_D5 + COH01_RE_D5]; c01i = x[IN01_COH_OFFSET_D5 + COH01_IM_D5];
    if (states) {
      RYDCZ_UNROLL for (int i = 0; i < Sys_rho01_d5::N; ++i) {
        states[SL::RHO01 + i] = x[i]; states[SL::RHO10 + i] = x[i]; }
      RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d5::N; ++i) {
        states[SL::COH01 + i] = x[IN01_COH_OFFSET_D5 + i]; states[SL::COH10 + i] = x[IN01_COH_OFFSET_D5 + i]; }
    }
  } else {
    double x[Sys_coh01_d5::N]; st = 0.0;
    run_system<Sys_coh01_d5>(x, pt, seg, n_seg, false, st);
    RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_coh01_d5::NNZ + Sys_coh01_d5::N);
    c01r = x[COH01_RE_D5]; c01i = x[COH01_IM_D5];
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d5::N; ++i) {
      states[SL::COH01 + i] = x[i]; states[SL::COH10 + i] = x[i]; } }
  }
  cz_epilogue(NOISY, p01, p01, p11, c01r, c01i, c01r, c01i, c11r, c11i, out);
  out[OUT_LOSS] = loss; out[OUT_RYD] = ryd; out[OUT_STEPS] = steps; out[OUT_FLOPS] = flops;
}

// dim = 5, atoms that see different detunings (Monte-Carlo shots with Doppler offsets): the general systems.  rho_11
// (81 variables) through run_system_huge with x, b1, b2 in the shared-memory columns xcol / b1col / b2col; the
// coherence (18 variables) and the two single-atom inputs (15 each) in registers.
template <bool NOISY>
RYDCZ_HD void evaluate_point_d5(const PointView pt, const Col xcol, const Col b1col, const Col b2col,
                                const double* __restrict__ seg, int n_seg, double* __restrict__ out,
                                double* __restrict__ states) {
  typedef StateLayout<5> SL;
  double steps = 0.0, flops = 0.0, st;
  double p01 = 1.0, p10 = 1.0, p11 = 1.0, loss = 0.0, ryd = 0.0;
  double c01r, c01i, c10r, c10i, c11r, c11i;
  if (NOISY) {
    st = 0.0;
    run_system_huge<Sys_rho11_d5, Col>(xcol, b1col, b2col, pt, seg, n_seg, true, st);
    RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_rho11_d5::NNZ + Sys_rho11_d5::N);
    double x[Sys_rho11_d5::N];
    for (int i = 0; i < Sys_rho11_d5::N; ++i) x[i] = xcol[i];
    p11 = x[RHO11_POP_D5];
    loss = rho11_loss_d5(x); ryd = rho11_ryd_d5(x);
    if (states) { for (int i = 0; i < Sys_rho11_d5::N; ++i) states[SL::RHO11 + i] = x[i]; }
  }
  {
    double x[Sys_coh11_d5::N]; st = 0.0;
    run_system<Sys_coh11_d5>(x, pt, seg, n_seg, NOISY, st);
    RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_coh11_d5::NNZ + Sys_coh11_d5::N);
    c11r = x[COH11_RE_D5]; c11i = x[COH11_IM_D5];
    if (!NOISY) ryd = coh11_ryd_d5(x);
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh11_d5::N; ++i) states[SL::COH11 + i] = x[i]; }
  }
  if (NOISY) {
    double x[Sys_in01_d5::N]; st = 0.0;
    run_system<Sys_in01_d5>(x, pt, seg, n_seg, true, st); RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_in01_d5::NNZ + Sys_in01_d5::N);
    p01 = x[RHO01_POP_D5];
    c01r = x[IN01_COH_OFFSET_D5 + COH01_RE_D5]; c01i = x[IN01_COH_OFFSET_D5 + COH01_IM_D5];
    if (states) {
      RYDCZ_UNROLL for (int i = 0; i < Sys_rho01_d5::N; ++i) states[SL::RHO01 + i] = x[i];
      RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d5::N; ++i) states[SL::COH01 + i] = x[IN01_COH_OFFSET_D5 + i];
    }
    double y[Sys_in10_d5::N]; st = 0.0;
    run_system<Sys_in10_d5>(y, pt, seg, n_seg, true, st); RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_in10_d5::NNZ + Sys_in10_d5::N);
    p10 = y[RHO10_POP_D5];
    c10r = y[IN10_COH_OFFSET_D5 + COH10_RE_D5]; c10i = y[IN10_COH_OFFSET_D5 + COH10_IM_D5];
    if (states) {
      RYDCZ_UNROLL for (int i = 0; i < Sys_rho10_d5::N; ++i) states[SL::RHO10 + i] = y[i];
      RYDCZ_UNROLL for (int i = 0; i < Sys_coh10_d5::N; ++i) states[SL::COH10 + i] = y[IN10_COH_OFFSET_D5 + i];
    }
  } else {
    double x[Sys_coh01_d5::N]; st = 0.0;
    run_system<Sys_coh01_d5>(x, pt, seg, n_seg, false, st); RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_coh01_d5::NNZ + Sys_coh01_d5::N);
    c01r = x[COH01_RE_D5]; c01i = x[COH01_IM_D5];
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d5::N; ++i) states[SL::COH01 + i] = x[i]; }
    double y[Sys_coh10_d5::N]; st = 0.0;
    run_system<Sys_coh10_d5>(y, pt, seg, n_seg, false, st); RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (Sys_coh10_d5::NNZ + Sys_coh10_d5::N);
    c10r = y[COH10_RE_D5]; c10i = y[COH10_IM_D5];
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh10_d5::N; ++i) states[SL::COH10 + i] = y[i]; }
  }
  cz_epilogue(NOISY, p01, p10, p11, c01r, c01i, c10r, c10i, c11r, c11i, out);
  out[OUT_LOSS] = loss; out[OUT_RYD] = ryd; out[OUT_STEPS] = steps; out[OUT_FLOPS] = flops;
}

// Gershgorin radius of the stiffest system (rho_11, or coh_11 when noiseless) for one segment:
// the work estimate used to sort points so that the threads of a warp run equally long.
template <int D>
RYDCZ_HD double point_work(const double* __restrict__ pt, const double* __restrict__ seg, int n_seg) {
  double w = 0.0;
  const double V = fabs(pt[PT_V]);
  double damp = pt[PT_G_R1] + pt[PT_G_R0] + pt[PT_G_PHI] + pt[PT_G_SC];
  double extra = 0.0, k2 = 1.0;
  if (D == 5) {
    // the two-atom spectrum spans V + 2 Delta_e plus both legs' Rabi frequencies
    damp += pt[PT5_G_E1] + pt[PT5_G_E0] + pt[PT5_G_RL];
    extra = 2.0 * fabs(pt[PT5_DET_E]) + 2.0 * fabs(pt[PT5_OM1]);
    k2 = fabs(2.0 * pt[PT5_DET_E] / pt[PT5_OM1]);
  }
  for (int s = 0; s < n_seg; ++s) {
    const double* sg = seg + (long long)s * SG_STRIDE;
    const double R = V + extra + 2.0 * k2 * (fabs(sg[SG_RE]) + fabs(sg[SG_IM])) + 2.0 * fabs(sg[SG_DET]) +
                     fabs(pt[PT_OFF_A]) + fabs(pt[PT_OFF_B]) + 2.0 * fabs(pt[PT_DELTA1]) + 2.0 * damp;
    w += cheb_work(R, sg[SG_DT]);
  }
  return w;
}

}  // namespace rydcz

#include "rydcz_cluster.cuh"
#include "rydcz_repeat.cuh"
