// rydcz_cluster.cuh -- exact segment propagation when the blockade splits the spectrum into clusters.
//
// The plain path (run_system, rydcz_core.cuh) expands exp(tau G) in Chebyshev polynomials on ONE
// interval that spans the whole spectrum of the generator: K ~ ||G|| tau ~ V tau matvecs per segment,
// almost all of them spent resolving the fast phase exp(+-i V t) of the doubly excited state.  For
// V >> Omega the spectrum of G is not an interval but three short ones: with H = H_small + V |rr><rr|,
//
//   * all eigenvalues of H but one stay inside [m, M], the spectral range of H_small (Cauchy
//     interlacing for the rank-one term V |rr><rr|), and the last one lies in
//     [E_rr, E_rr + eta], eta = |Omega|^2 / (2 (E_rr - M)), E_rr = V - Delta_A - Delta_B; for identical
//     atoms the tighter range of the exchange-symmetric block without |rr> is used, widened by eta below;
//   * the eigenvalues of the Liouvillian -i[H, .] are differences of those: clusters at 0 and +-i V_c of
//     half-widths w0 = M - m and w1 = (M - m + eta) / 2; the coherence systems (real form of
//     v' = -i H_eff v, no spectral shift) have the same three clusters;
//   * the dissipator moves every eigenvalue by at most ||L_D|| <= 4 sum(rates) (Bauer-Fike, -i[H, .] is normal).
//
// The cubic q(x) = x (x^2 + V_c^2) maps all three clusters onto ONE short interval of the imaginary
// axis, |q| <= rho ~ 2 V_c^2 w.  Every entire f can be written f(x) = a0(q) + x a1(q) + x^2 a2(q) with a_j
// entire (a_j(s) = coefficients of the quadratic that interpolates f at the three pre-images of s), so
//
//     exp(h G) x = sum_k [ d0_k + d1_k G~ + d2_k G~^2 ] phi_k(S~) x,     S~ = q(G) / rho,
//
// with phi_k the same Chebyshev recurrence the plain path uses and d_jk the Chebyshev coefficients of
// a_j(i rho t) on [-1, 1] (computed per segment by a cosine transform over K + 2 Gauss nodes; only
// half of them by conjugate symmetry).  Their number is K ~ h (M - m) + O((h w)^(1/3)): set by Omega tau
// and no longer by V tau.  One Clenshaw step costs three sparse matvecs, so the path pays when
// V >~ 15 Omega; for the reference's medium apparatus (V/Omega = 342) a bang-bang candidate needs ~16x
// fewer matvecs.  The result is exact to rounding like the plain path (tests: < 1e-12 on rho entries).
//
// Plain C++ when RYDCZ_HD is empty (tests/hostsim), like rydcz_core.cuh.
#pragma once

namespace rydcz {

constexpr int CL_REC = 4;          // doubles per coefficient record (d0, d1, d2, pad): one 32-byte sector
constexpr int CL_MAX_SLOTS = 8;    // distinct (duration, |Omega|, Delta) groups of consecutive segments per point
constexpr int CL_MAX_K = 62;       // terms per (sub)step; longer segments are cut into equal sub-steps
constexpr int CL_HALF_MAX = (CL_MAX_K + 2) / 2;
constexpr int CL_HDR = 8;          // doubles of a slot header (two records)
constexpr int CL_OWN0 = CL_MAX_SLOTS * CL_HDR / CL_REC;     // a point's region: its slot headers first, then its own records
constexpr int CL_RECORDS = CL_MAX_SLOTS * (CL_MAX_K + 3);   // records per point: header (2) + K + 1 terms per slot
constexpr int CL_STRIDE = CL_RECORDS * CL_REC;              // doubles per point in the coefficient scratch
#ifndef RYDCZ_CL_ALIGN
#define RYDCZ_CL_ALIGN 8
#endif
constexpr int CL_ALIGN = RYDCZ_CL_ALIGN;        // Clenshaw steps per (sub)step are padded to a multiple of this, so the
                                   // lanes of a warp reach their segment transitions at common iterations
                                   // (4: less padding but more scattered transitions, 2.36 ms against 2.24 ms)
constexpr double CL_Z_MAX = 26.0;  // largest h * freq of one sub-step (keeps K <= CL_MAX_K: 26 + 10.3 * 26^(1/3) + 4 - 2 < 62)
constexpr double CL_STEP_COST = 3.4;    // one clustered Clenshaw step in units of one plain Chebyshev term
constexpr double CL_SLOT_COST = 60.0;   // coefficient generation of one slot, same units
constexpr double CL_SLOT_COST_SHARED = 2.0;   // ... when the slot takes its records from the shared table
constexpr double CL_SUBSTEP_COST = 3.0; // start vectors of one (sub)step, same units

struct ClusterGeom {
  double Vc;    // centre of the outer clusters (rad/s)
  double rho;   // |q(x)| <= rho on all three clusters
  double freq;  // largest d|x|/d(rho t) * rho on the interval: h * freq bounds the Chebyshev degree
  double damp;  // bound on the real parts (4 x sum of rates)
  double w0, w1; // half-widths of the central and of the outer clusters (rad/s, slack included)
  double iVc, w0h, w1h;   // 1 / Vc and the half-widths in units of Vc (the shared table works in those units)
};

// D only selects nothing here (the bounds hold for the 3- and 4-level models); kept for symmetry with the callers.
RYDCZ_HD bool cluster_geometry(double V, double d1, double rates, double om2, double detA, double detB, bool sym,
                               ClusterGeom& g) {
  if (!(V > 0.0)) return false;
  // single-atom block {|1>, |r>}: [[d1, Omega/2], [Omega*/2, -Delta]], eigenvalues a -+ b (|0> sits at 0)
  const double sA = d1 + detA, aA = 0.5 * (d1 - detA), bA = 0.5 * sqrt(sA * sA + om2);
  double m, M;
  if (sym) {
    // identical atoms, exchange-symmetric sector: {|11>, |S>} block [[2 d1, Omega/sqrt2], [., d1 - Delta]]
    const double a2 = 0.5 * (3.0 * d1 - detA), b2 = 0.5 * sqrt(sA * sA + 2.0 * om2);
    m = fmin(0.0, fmin(aA - bA, a2 - b2));
    M = fmax(0.0, fmax(aA + bA, a2 + b2));
  } else {
    const double sB = d1 + detB, aB = 0.5 * (d1 - detB), bB = 0.5 * sqrt(sB * sB + om2);
    m = fmin(0.0, aA - bA) + fmin(0.0, aB - bB);
    M = fmax(0.0, aA + bA) + fmax(0.0, aB + bB);
  }
  const double Err = V - detA - detB;
  const double gap = Err - M;
  if (!(gap > 0.0)) return false;
  const double eta = 0.5 * om2 / gap, d = 4.0 * rates;
  // identical atoms: [m, M] is the range of the block WITHOUT |rr>; coupling to |rr> (|F|^2 = |Omega|^2 / 2)
  // can push its lowest level down by eta as well (Cauchy interlacing for the bordered matrix)
  if (sym) m -= eta;
  const double W = M - m;
  const double Vc = Err + 0.5 * (eta - M - m);
  const double w0 = 1.02 * (W + d), w1 = 1.02 * (0.5 * (W + eta) + d);
  const double Vc2 = Vc * Vc;
  const double rho = fmax(w0 * Vc2, (Vc + w1) * w1 * (2.0 * Vc + w1));
  // the branch points of q^{-1} sit at |q| = 0.385 Vc^3: stay well inside
  if (!(rho <= 0.25 * Vc2 * Vc) || !(rho > 0.0)) return false;
  double y = rho / Vc2;   // largest |y| of the central branch: y (Vc^2 - y^2) = rho
  for (int i = 0; i < 6; ++i) {          // (two or three steps unless rho is close to its limit)
    const double dy = (y * (Vc2 - y * y) - rho) / (Vc2 - 3.0 * y * y);
    y -= dy;
    if (fabs(dy) <= 1e-15 * y) break;
  }
  g.Vc = Vc; g.rho = rho; g.freq = rho / (Vc2 - 3.0 * y * y); g.damp = d; g.w0 = w0; g.w1 = w1;
  g.iVc = 1.0 / Vc; g.w0h = w0 * g.iVc; g.w1h = w1 * g.iVc;
  return true;
}

// cos(pi x): the CUDA intrinsic on the device, libm on the host harness
RYDCZ_HD double cluster_cospi(double x) {
#ifdef __CUDA_ARCH__
  return cospi(x);
#else
  return cos(3.141592653589793238462643383279 * x);
#endif
}

// prefetch of a coefficient record into L2 (L1 is too small for two lines per resident thread); no-op on the host
#ifndef RYDCZ_PREFETCH
#ifdef __CUDA_ARCH__
#define RYDCZ_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define RYDCZ_PREFETCH(p) ((void)0)
#endif
#endif

// Asynchronous 8-byte copies global -> shared (cp.async), committed in per-thread groups: the ring of
// coefficient records that run_system_cluster consumes is filled RING - 1 steps ahead of its use without
// holding registers.  Plain copies on the host.
#ifndef RYDCZ_ASYNC_COPY8
#ifdef __CUDA_ARCH__
#define RYDCZ_ASYNC_COPY8(dst, src)                                                             \
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define RYDCZ_ASYNC_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define RYDCZ_ASYNC_WAIT(n) asm volatile("cp.async.wait_group %0;" ::"n"(n) : "memory")
#else
#define RYDCZ_ASYNC_COPY8(dst, src) (*(dst) = *(src))
#define RYDCZ_ASYNC_COMMIT() ((void)0)
#define RYDCZ_ASYNC_WAIT(n) ((void)0)
#endif
#endif

RYDCZ_HD int cluster_roundup(int steps) { return (steps + CL_ALIGN - 1) / CL_ALIGN * CL_ALIGN; }

// ---- the shared coefficient table ------------------------------------------------------------------------------
// The records of a (sub)step depend on three numbers only: K, theta = h Vc and rho^ = rho / Vc^3 (everything in
// units of Vc).  And the expansion they encode is a polynomial approximation of exp(theta x) that is accurate
// wherever x lies in the three clusters it was built for -- so a (sub)step does not need records of its own:
//
//     exp(h G) = exp(theta_n (lambda G / Vc)),   theta_n = n CT_DTHETA the grid value next to theta = |h| Vc,
//                                                lambda = +-theta / theta_n,
//
// is evaluated EXACTLY (to the same truncation error) with the records of (theta_n, rho^_c) applied to the operator
// lambda G / Vc, as long as the spectrum of THAT operator lies inside the table entry's clusters: centre i lambda
// instead of i (|lambda - 1| <= CT_DTHETA / 2 theta_n) and widths scaled by lambda.  The planner picks the
// narrowest class c = floor(8 log2(0.25 / rho^_req)) that still contains them; the price is z = theta_n freq^_c up
// by <= CT_DTHETA plus <= 9 % (one class), about half a term of the ~20-60 a (sub)step needs.
// In an optimiser population and in the bench every candidate shares (V, |Omega|, Delta): 1.3 M slots of a
// 262,144-candidate bang-bang batch map onto ~2,000 table entries, which are generated once per call (or kept
// across calls, rydcz_set_table) and stay in L2 -- the per-slot coefficient kernel (29 % of the step in round 1)
// and its 4 GB of records per step are gone.  Slots that do not fit the table (theta < 0.25, rho^ < 1e-3,
// ...) keep records of their own, as before.
constexpr int CT_CLASSES = 64;               // rho^_c = 0.25 * 2^(-c / 8), c = 0 .. 63 (down to 1.07e-3: V/Omega ~ 2000)
constexpr double CT_PER_OCTAVE = 8.0;
constexpr double CT_DTHETA = 0.5;            // grid step in theta = h Vc
constexpr double CT_Z_MAX = 27.0;            // largest theta_n * freq^_c of an entry (K <= CL_MAX_K)
constexpr int CT_ENTRY_REC = CL_MAX_K + 1;   // records per entry
struct ClusterTableDir {
  double rho[CT_CLASSES], freq[CT_CLASSES], kappa[CT_CLASSES], inv_gs[CT_CLASSES];
  int base[CT_CLASSES + 1];                  // entry (c, n), n = 1 .. base[c + 1] - base[c], sits at base[c] + n - 1
};
inline void cluster_table_dir_init(ClusterTableDir& d) {
  d.base[0] = 0;
  for (int c = 0; c < CT_CLASSES; ++c) {
    const double rho = 0.25 * exp2(-(double)c / CT_PER_OCTAVE);
    double y = rho;                          // central branch: y (1 - y^2) = rho^
    for (int i = 0; i < 40; ++i) y -= (y * (1.0 - y * y) - rho) / (1.0 - 3.0 * y * y);
    const double gs = cbrt(0.5 * rho);
    d.rho[c] = rho; d.freq[c] = rho / (1.0 - 3.0 * y * y); d.kappa[c] = 1.0 / (gs * gs); d.inv_gs[c] = 1.0 / gs;
    d.base[c + 1] = d.base[c] + (int)floor(CT_Z_MAX / (d.freq[c] * CT_DTHETA));
  }
}
// The planner's view: the directory and one request stamp per entry -- dir == nullptr: no table.  A planner only
// STORES the call's epoch into the stamp of every entry its point needs (no atomic, no read: the planner is a chain of
// dependent memory operations and an atomic with a result cost it 0.35 ms per 262,144 points); a pass over the
// requested range (cluster_table_kernel; cluster_table_generate on the host) then generates the entries whose stamp
// is the current epoch and whose generation mark is not the current generation, and marks them.  The range of
// requested entries comes back through lo / hi.
struct ClusterTable {
  const ClusterTableDir* dir;
  unsigned* stamp;                         // nullptr: a dry run (cost and eligibility only)
  unsigned epoch;
  RYDCZ_HD bool on() const { return dir != nullptr; }
  RYDCZ_HD const ClusterTableDir& d() const { return *dir; }
  RYDCZ_HD void request(long long e) const {
    if (stamp) stamp[e] = epoch;
  }
};
// (The planner kernel passes a view of the same shape whose directory is a __shared__ copy: the walker, the lookup
//  and cluster_plan take the view's type as a template parameter so that those reads compile to shared-memory loads.)
// (K, theta_n, rho^_c) of a table entry from its index: what the lookup below computed for the slots that use it
RYDCZ_HD void cluster_table_entry(const ClusterTableDir& d, const int* __restrict__ base, long long e, int& K, double& theta_n,
                                  double& rho) {
  int c = 0;
  RYDCZ_UNROLL
  for (int step = CT_CLASSES / 2; step > 0; step >>= 1)
    if (c + step < CT_CLASSES && (long long)base[c + step] <= e) c += step;
  const int n = (int)(e - base[c]) + 1;
  theta_n = (double)n * CT_DTHETA;
  rho = d.rho[c];
  double zs = theta_n * d.freq[c];
  if (!(zs >= 1e-2)) zs = 1e-2;
  K = cheb_terms(zs) - 2;
  if (K < 6) K = 6;
}

// One slot: `count` consecutive active segments with the same duration, |Omega| and Delta, each advanced
// in `nsub` equal sub-steps of length h with K + 1 coefficient records -- its own (entry < 0) or a table entry's.
struct ClusterSlot {
  ClusterGeom g;
  double h;
  int K, nsub, count;
  long long entry;        // table entry, -1 = records of its own
  double kappa, inv_gs;   // what the evaluator needs: (Vc / gs)^2 and the factor that turns G into G~
};

// The table entry for a (sub)step of length h under geometry g, if there is one.
template <class TB>
RYDCZ_HD bool cluster_table_lookup(const TB& tb, const ClusterGeom& g, double h, ClusterSlot& sl) {
  if (!tb.on()) return false;
  const ClusterTableDir& dir = tb.d();
  const double theta = fabs(h) * g.Vc;
  const double nn = rint(theta / CT_DTHETA);
  if (!(nn >= 1.0) || !(nn < 1e8)) return false;
  const double theta_n = nn * CT_DTHETA, lam = theta / theta_n;
  const double w1 = fabs(lam - 1.0) + lam * g.w1h;
  const double need = fmax(lam * g.w0h, (1.0 + w1) * w1 * (2.0 + w1));
  if (!(need <= 0.25) || !(need > 0.0)) return false;
  // class = floor(8 log2(0.25 / need)): a single-precision logarithm, then made exact against the directory
  int c = (int)floorf((float)CT_PER_OCTAVE * log2f(0.25f / (float)need));
  if (c >= CT_CLASSES) c = CT_CLASSES - 1;
  if (c < 0) c = 0;
  while (c + 1 < CT_CLASSES && dir.rho[c + 1] >= need) ++c;
  while (c >= 0 && dir.rho[c] < need) --c;
  if (c < 0) return false;
  const int n = (int)nn;
  if (n > dir.base[c + 1] - dir.base[c]) return false;     // theta_n freq^_c > CT_Z_MAX: more sub-steps needed
  double zs = theta_n * dir.freq[c];
  if (!(zs >= 1e-2)) zs = 1e-2;
  int K = cheb_terms(zs) - 2;
  if (K < 6) K = 6;
  if (K > CL_MAX_K) return false;
  sl.K = K;
  sl.entry = (long long)dir.base[c] + (n - 1);
  sl.kappa = dir.kappa[c];
  sl.inv_gs = (h < 0.0 ? -lam : lam) * g.iVc * dir.inv_gs[c];
  return true;
}

// Walks the active segments of a point (the ones next_segment visits) and groups them into slots.
template <int D, class TB = ClusterTable>
struct ClusterWalker {
  const double* pt; const double* seg;
  int n_seg, s;
  bool sym;
  double rates;
  // the geometry of the previous slot: consecutive slots of a bang-bang pulse share (|Omega|, Delta) and differ
  // only in their duration, so the bounds (square roots, a cube root's worth of Newton steps) are computed once
  double last_om2 = -1.0, last_det = 0.0;
  bool last_ok = false;
  ClusterGeom last_g;
  const TB& tb;
  RYDCZ_HD ClusterWalker(const double* pt_, const double* seg_, int n_seg_, bool noisy, bool sym_, const TB& tb_)
      : pt(pt_), seg(seg_), n_seg(n_seg_), s(0), sym(sym_), rates(0.0), tb(tb_) {
    if (noisy) {
      rates = pt[PT_G_R1] + pt[PT_G_R0] + pt[PT_G_PHI] + pt[PT_G_SC];
      if (D == 4) rates += pt[PT_G_MJ];
    }
  }
  RYDCZ_HD static bool is_active(double dt) { return dt != 0.0 && dt == dt; }
  // 1 = slot filled, 0 = no segment left, -1 = this point cannot take the clustered path
  RYDCZ_HD int next(ClusterSlot& sl) {
    while (s < n_seg && !is_active(seg[(long long)s * SG_STRIDE + SG_DT])) ++s;
    if (s >= n_seg) return 0;
    const double* sg = seg + (long long)s * SG_STRIDE;
    const double dt = sg[SG_DT], det = sg[SG_DET];
    const double om2 = sg[SG_RE] * sg[SG_RE] + sg[SG_IM] * sg[SG_IM];
    // (|Omega|^2 of equal-amplitude segments differs by rounding: the bounds carry 2 % of slack)
    if (!(det == last_det && fabs(om2 - last_om2) <= 1e-12 * last_om2)) {
      last_ok = cluster_geometry(pt[PT_V], pt[PT_DELTA1], rates, om2, det + pt[PT_OFF_A], det + pt[PT_OFF_B], sym, last_g);
      last_om2 = om2; last_det = det;
    }
    if (!last_ok) return -1;
    sl.g = last_g;
    const double adt = fabs(dt);
    const double z = adt * sl.g.freq, dz = sl.g.damp * adt;
    if (!(z < 1e6) || !(dz < 1e3)) return -1;
    // sub-steps: keep K <= CL_MAX_K, and damp * h <= 1 so that the Chebyshev polynomials stay O(e^2) on the
    // strip the damped spectrum occupies (same reason as the plain path's nsub); nearly free, K ~ z + O(z^(1/3))
    sl.nsub = z > CL_Z_MAX ? (int)ceil(z / CL_Z_MAX) : 1;
    if (dz > (double)sl.nsub) sl.nsub = (int)ceil(dz);
    sl.h = dt / (double)sl.nsub;
    double zs = z / (double)sl.nsub;
    if (!(zs >= 1e-2)) zs = 1e-2;
    // records from the shared table where an entry contains this (sub)step; the grid value next to h Vc widens the
    // clusters a little, so a (sub)step at the very edge of the table's range is tried once more, one piece shorter
    sl.entry = -1;
    bool shared = tb.on() && cluster_table_lookup(tb, sl.g, sl.h, sl);
    if (!shared && tb.on() && zs > 0.5 * CL_Z_MAX) {
      const double h2 = dt / (double)(sl.nsub + 1);
      shared = cluster_table_lookup(tb, sl.g, h2, sl);
      if (shared) { sl.h = h2; sl.nsub += 1; }
    }
    if (!shared) {
      // records of its own.  Two terms fewer than the plain path's rule (measured: the fixtures' worst error against
      // the expm oracle stays at 6.1e-13; it reaches 7e-12 at four fewer and 1.2e-11 at six)
      sl.K = cheb_terms(zs) - 2;
      if (sl.K < 6) sl.K = 6;
      if (sl.K > CL_MAX_K) return -1;
      const double gs = cbrt(0.5 * sl.g.rho);
      sl.kappa = (sl.g.Vc / gs) * (sl.g.Vc / gs);
      sl.inv_gs = 1.0 / gs;
    }
    sl.count = 1;
    ++s;
    // absorb the following active segments that repeat (dt, |Omega|, Delta): smooth-JP steps, the second LP pulse
    for (;;) {
      int t = s;
      while (t < n_seg && !is_active(seg[(long long)t * SG_STRIDE + SG_DT])) ++t;
      if (t >= n_seg) break;
      const double* nx = seg + (long long)t * SG_STRIDE;
      const double o2 = nx[SG_RE] * nx[SG_RE] + nx[SG_IM] * nx[SG_IM];
      if (!(nx[SG_DT] == dt && nx[SG_DET] == det && fabs(o2 - om2) <= 1e-9 * om2)) break;
      ++sl.count;
      s = t + 1;
    }
    return 1;
  }
};

// Slot header (CL_HDR doubles; slot q of a point sits at q * CL_HDR of the point's region):
//   K + 1024 nsub, kappa = (Vc / gs)^2, the factor that turns G into G~ (1 / gs; table: +-lambda / (Vc gs^)),
//   count + 65536 K_next | where this slot's records are, where the NEXT slot's records are, 0, 0
// (gs = (rho/2)^(1/3).  K_next = term count of the slot that follows, 0 if there is none: the evaluator's record
//  ring runs across slot boundaries and must know it, and where those records are, before that slot's own header
//  has been read.)  A record position >= 0 counts records from the start of the point's region (records of its
//  own, from CL_OWN0 on); -(1 + r) is record r of the shared table.
RYDCZ_HD double cluster_terms_position(const ClusterSlot& sl, int own_offset) {
  return sl.entry >= 0 ? -(double)(1 + sl.entry * CT_ENTRY_REC) : (double)own_offset;
}
RYDCZ_HD void cluster_write_header(const ClusterSlot& sl, double position, int K_next, double next_position,
                                    double* __restrict__ rec) {
  rec[0] = (double)(sl.K + 1024 * sl.nsub);
  rec[1] = sl.kappa;
  rec[2] = sl.inv_gs;
  rec[3] = (double)sl.count + 65536.0 * (double)K_next;
  rec[4] = position;
  rec[5] = next_position;
  rec[6] = 0.0;
  rec[7] = 0.0;
}
RYDCZ_HD const double* cluster_terms_at(double position, const double* __restrict__ rec, const double* __restrict__ table) {
  const long long v = (long long)position;
  return v >= 0 ? rec + v * CL_REC : table + (-(v + 1)) * CL_REC;
}

// Records the clustered path needs in the point's region (0 = not eligible) and its cost in plain Chebyshev terms.
// With `rec` the slot headers are written on the way.  Per slot q: slot_K[q] = term count of a coefficient task to
// run for records of the point's own (0: no such slot, or the slot reads the shared table), slot_val[q] = record
// index in the point's region where that task writes, `dense` (4 doubles per slot) = what cluster_coefficients_of
// needs for it: K (+ 1024 nsub), h, Vc, rho.  Table entries are requested through tb (ClusterTable::request);
// n_shared counts the slots that read the table, [e_lo, e_hi] is widened to the entries requested.
// (Every header is written ONCE, when the slot after it is known -- it carries that slot's term count and record
//  position; reading a header back to patch it was 16 % of the planner's stall samples.)
template <int D, class TB>
RYDCZ_HD int cluster_plan(const double* pt, const double* seg, int n_seg, bool noisy, bool sym, double& cost,
                          const TB& tb, int& n_shared, long long& e_lo, long long& e_hi,
                          int* __restrict__ slot_K = nullptr, unsigned* __restrict__ slot_val = nullptr,
                          double* __restrict__ rec = nullptr, double* __restrict__ dense = nullptr) {
  ClusterWalker<D, TB> w(pt, seg, n_seg, noisy, sym, tb);
  ClusterSlot sl, prev;
  double prev_position = 0.0;
  int records = CL_OWN0, slots = 0;
  cost = 0.0;
  n_shared = 0;
  for (;;) {
    const int r = w.next(sl);
    if (r < 0) return 0;
    if (r > 0 && ++slots > CL_MAX_SLOTS) return 0;
    const double position = r > 0 ? cluster_terms_position(sl, records) : 0.0;
    if (rec && slots > (r > 0 ? 1 : 0))       // the header of the previous slot, now that this one is known
      cluster_write_header(prev, prev_position, r > 0 ? sl.K : 0, position, rec + (slots - (r > 0 ? 2 : 1)) * CL_HDR);
    if (r == 0) break;
    const int q = slots - 1;
    const bool shared = sl.entry >= 0;
    if (shared) {
      tb.request(sl.entry);
      ++n_shared;
      if (sl.entry < e_lo) e_lo = sl.entry;
      if (sl.entry > e_hi) e_hi = sl.entry;
    }
    if (slot_K) slot_K[q] = shared ? 0 : sl.K;
    if (slot_val) slot_val[q] = (unsigned)records;
    if (dense && !shared) {
      double* d = dense + q * 4;
      d[0] = (double)(sl.K + 1024 * sl.nsub); d[1] = sl.h; d[2] = sl.g.Vc; d[3] = sl.g.rho;
    }
    if (!shared) records += sl.K + 1;
    cost += (double)sl.count * (double)sl.nsub * (CL_STEP_COST * (double)cluster_roundup(sl.K + 2) + CL_SUBSTEP_COST) +
            (shared ? CL_SLOT_COST_SHARED : CL_SLOT_COST);
    prev = sl; prev_position = position;
  }
  return slots == 0 ? 0 : records;
}

// Coefficient records of one slot, from its header:
//   record k = 0..K: d0_k, d1_k, d2_k for  C_k = d0_k + d1_k G~ + d2_k G~^2
// Nodes t_n = cos(pi (n + 1/2) / N), N = K + 2; the three pre-images of s = i rho t are x = i y with
// y (Vc^2 - y^2) = rho t: y_0 near 0 and y_+- = +-Vc (1 - e_+-), in units of Vc roots of
//     y - y^3 = c,   phi(e) = 2 e - 3 e^2 + e^3 = +-c,   c = rho t / Vc^3 in [-1/4, 1/4],
// found by two Newton steps from a polynomial fit (start errors 5e-8 / 5e-6).  The fast phases
// exp(+-i h Vc) are factored out and evaluated once, so the per-node sincos arguments stay O(h w) and the
// rounding of h Vc acts as one consistent perturbation of V.
RYDCZ_HD double cluster_root_e(double c) {
  double e = c * (0.49996334654598612 + c * (0.37665288212020709 + c * (0.50740503889100941 + c * (0.69820872531829381 +
             c * (1.1095386292091427 + c * (5.605127239220935 + c * 12.998643520117961)))))) - 3.3907227034220971e-06;
  // chord iteration with the slope at the start point: the error shrinks by ~2.5e-5 per step
  const double r = 1.0 / (2.0 + e * (3.0 * e - 6.0));
  RYDCZ_UNROLL
  for (int it = 0; it < 3; ++it) e = fma(-(e * (2.0 + e * (e - 3.0)) - c), r, e);
  return e;
}
RYDCZ_HD double cluster_root_y(double c) {
  const double q = c * c;
  double y = c * (0.99999984362350458 + q * (1.0001738811032581 + q * (2.9690741824804951 + q * (13.967400857958946 +
             q * (0.85149574480882961 + q * 886.98095006735718)))));
  const double r = 1.0 / (1.0 - 3.0 * y * y);
  RYDCZ_UNROLL
  for (int it = 0; it < 2; ++it) y = fma(-(fma(-y * y, y, y) - c), r, y);
  return y;
}

// G lanes work on one slot (G = 4 consecutive lanes of a warp on the GPU, 1 on the host); lane m takes the
// nodes n = m, m + G, ... and then the blocks of eight k starting at 8 m, 8 (m + G), ...  The node values
// a_j(t_n) go through `A` (element (v, n) at A[(v * CL_HALF_MAX + n) * a_stride]; shared memory on the
// GPU), with a group barrier between the two phases.
#ifndef RYDCZ_GROUP_SYNC
#ifdef __CUDA_ARCH__
#define RYDCZ_GROUP_SYNC() __syncwarp()
#else
#define RYDCZ_GROUP_SYNC() ((void)0)
#endif
#endif
RYDCZ_HD void cluster_coefficients_of(int K, double h, double Vc, double rho, double* __restrict__ out,
                                      double* __restrict__ A, int a_stride, int m, int G);
// ... from (K, h, Vc, rho) handed over directly; `out` = the K + 1 records behind the slot header
RYDCZ_HD void cluster_coefficients_of(int K, double h, double Vc, double rho, double* __restrict__ out,
                                      double* __restrict__ A, int a_stride, int m, int G) {
  const int N = K + 2, half = N / 2;
  const double gs = cbrt(0.5 * rho);
#define RYDCZ_A(v, n) A[((v) * CL_HALF_MAX + (n)) * a_stride]
  const double rhat = rho / (Vc * Vc * Vc);
  const double hv = h * Vc, iVc = 1.0 / Vc;
  const double g1 = gs * iVc, g2 = g1 * g1;      // a1 and a2 are built in units of 1/Vc, 1/Vc^2
  double sV, cV;
  sincos(hv, &sV, &cV);
  // node angles theta_n = pi (n + 1/2) / N, advanced by rotation
  const double pi = 3.141592653589793238462643383279;
  const double dth = pi / (double)N;
  {
    double st, ct, sd, cd;
    sincos(((double)m + 0.5) * dth, &st, &ct);
    sincos((double)G * dth, &sd, &cd);
    for (int n = m; n < half; n += G) {
      const double c = rhat * ct;
      const double y0 = cluster_root_y(c), ep = cluster_root_e(c), em = cluster_root_e(-c);
      // f_m = exp(i h y_m)
      double s0, c0, sp, cp, sm, cm;
      sincos(hv * y0, &s0, &c0);
      sincos(hv * ep, &sp, &cp);
      sincos(hv * em, &sm, &cm);
      const double fpr = cV * cp + sV * sp, fpi = sV * cp - cV * sp;      // exp(i hv) exp(-i hv ep)
      const double fmr = cV * cm + sV * sm, fmi = -(sV * cm - cV * sm);   // exp(-i hv) exp(+i hv em)
      const double Yp = 1.0 - ep, Ym = -(1.0 - em);                       // roots in units of Vc
      // divided differences on x = i y:  f[a, b] = -i (f_b - f_a) / (y_b - y_a); one division for the three gaps
      const double g01 = Yp - y0, g12 = Ym - Yp, g02 = Ym - y0;
      const double inv = 1.0 / (g01 * g12 * g02);
      const double i01 = inv * g12 * g02, i12 = inv * g01 * g02, i02 = inv * g01 * g12;
      const double d01r = (fpi - s0) * i01, d01i = -(fpr - c0) * i01;
      const double d12r = (fmi - fpi) * i12, d12i = -(fmr - fpr) * i12;
      const double d012r = (d12i - d01i) * i02, d012i = -(d12r - d01r) * i02;
      // p(x) = f_0 + f[0,+] (x - x_0) + f[0,+,-] (x - x_0)(x - x_+) = a0 + a1 x + a2 x^2
      const double ys = y0 + Yp, yp = y0 * Yp;
      const double a1r = d01r + ys * d012i, a1i = d01i - ys * d012r;            // d01 - i ys d012
      RYDCZ_A(0, n) = c0 + y0 * d01i - yp * d012r;                              // f0 - i y0 d01 - y0 yp d012
      RYDCZ_A(1, n) = s0 - y0 * d01r - yp * d012i;
      RYDCZ_A(2, n) = a1r * g1; RYDCZ_A(3, n) = a1i * g1;
      RYDCZ_A(4, n) = d012r * g2; RYDCZ_A(5, n) = d012i * g2;
      const double cn = ct * cd - st * sd;
      st = st * cd + ct * sd; ct = cn;
    }
  }
  RYDCZ_GROUP_SYNC();
  // cosine transform, eight k at a time for all three j (24 accumulators in registers).  Within a block
  // the Chebyshev recurrence runs from T_kb, T_kb+1 = cos(kb theta), cos((kb + 1) theta), which are carried
  // from node to node by rotation, so every record is written once, whole.  Even k pick up Re a_j, odd k
  // Im a_j (conjugate symmetry of the +-t node pairs).
  const double scale = 4.0 / (double)N;
  double st0, ct0, sd, cd;
  sincos(0.5 * dth, &st0, &ct0);
  sincos(dth, &sd, &cd);
  for (int kb = 8 * m; kb <= K; kb += 8 * G) {
    double acc[8][3];
    RYDCZ_UNROLL
    for (int q = 0; q < 8; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; acc[q][2] = 0.0; }
    // cos / sin of kb theta_0 and of the step kb pi / N
    double sk, ck, sdk, cdk;
    sincos((double)kb * (0.5 * dth), &sk, &ck);
    sincos((double)kb * dth, &sdk, &cdk);
    double ct = ct0, st = st0;
    for (int n = 0; n < half; ++n) {
      const double t2 = ct + ct;
      double Tc = ck, Tn = ck * ct - sk * st;                        // T_kb, T_kb+1
      const double r0 = RYDCZ_A(0, n), i0 = RYDCZ_A(1, n), r1 = RYDCZ_A(2, n), i1 = RYDCZ_A(3, n),
                   r2 = RYDCZ_A(4, n), i2 = RYDCZ_A(5, n);
      RYDCZ_UNROLL
      for (int q = 0; q < 8; q += 2) {
        acc[q][0] = fma(r0, Tc, acc[q][0]); acc[q][1] = fma(r1, Tc, acc[q][1]); acc[q][2] = fma(r2, Tc, acc[q][2]);
        acc[q + 1][0] = fma(i0, Tn, acc[q + 1][0]); acc[q + 1][1] = fma(i1, Tn, acc[q + 1][1]);
        acc[q + 1][2] = fma(i2, Tn, acc[q + 1][2]);
        Tc = fma(t2, Tn, -Tc);          // T_{k+2}
        Tn = fma(t2, Tc, -Tn);          // T_{k+3}
      }
      double cn = ct * cd - st * sd;
      st = st * cd + ct * sd; ct = cn;
      cn = ck * cdk - sk * sdk;
      sk = sk * cdk + ck * sdk; ck = cn;
    }
    RYDCZ_UNROLL
    for (int q = 0; q < 8; ++q) {
      const int k = kb + q;
      if (k <= K) {
        const double sc = ((k & 2) ? -scale : scale) * (k == 0 ? 0.5 : 1.0);   // i^{-k} on the surviving part
        out[k * CL_REC + 0] = acc[q][0] * sc;
        out[k * CL_REC + 1] = acc[q][1] * sc;
        out[k * CL_REC + 2] = acc[q][2] * sc;
        out[k * CL_REC + 3] = 0.0;
      }
    }
  }
#undef RYDCZ_A
}

// The table pass on the host: every entry of [e_lo, e_hi] that was requested in this epoch and does not carry the
// current generation mark.  Returns how many it generated.  (cluster_table_kernel on the GPU.)
inline int cluster_table_generate(const ClusterTableDir& dir, const unsigned* stamp, unsigned epoch, unsigned* mark,
                                  unsigned generation, long long e_lo, long long e_hi, double* table) {
  int generated = 0;
  double A[6 * CL_HALF_MAX];
  for (long long e = e_lo; e <= e_hi; ++e) {
    if (stamp[e] != epoch || mark[e] == generation) continue;
    int K; double theta_n, rho;
    cluster_table_entry(dir, dir.base, e, K, theta_n, rho);
    cluster_coefficients_of(K, theta_n, 1.0, rho, table + e * (CT_ENTRY_REC * CL_REC), A, 1, 0, 1);
    mark[e] = generation;
    ++generated;
  }
  return generated;
}

// All coefficient tasks of a point's own records, one after the other, and the table entries it requests (the host
// harness; the GPU runs four lanes per task / entry).  `mark` = generation marks of the table entries, `table` = its
// records.  Returns the number of table entries it generated.
template <int D>
RYDCZ_HD int cluster_prepare(const double* pt, const double* seg, int n_seg, bool noisy, bool sym,
                             double* __restrict__ rec, const ClusterTable& tb, unsigned* mark, unsigned generation,
                             double* __restrict__ table) {
  int slot_K[CL_MAX_SLOTS];
  unsigned slot_val[CL_MAX_SLOTS];
  double dense[CL_MAX_SLOTS * 4];
  for (int q = 0; q < CL_MAX_SLOTS; ++q) slot_K[q] = 0;
  double cost;
  int n_shared = 0;
  long long e_lo = 0x7fffffffffffffffLL, e_hi = -1;
  const int records = cluster_plan<D>(pt, seg, n_seg, noisy, sym, cost, tb, n_shared, e_lo, e_hi, slot_K, slot_val, rec, dense);
  if (records <= 0) return 0;
  double A[6 * CL_HALF_MAX];
  for (int q = 0; q < CL_MAX_SLOTS; ++q) {
    if (slot_K[q] <= 0) continue;
    cluster_coefficients_of(slot_K[q], dense[q * 4 + 1], dense[q * 4 + 2], dense[q * 4 + 3],
                            rec + (long long)slot_val[q] * CL_REC, A, 1, 0, 1);
  }
  return (tb.on() && tb.stamp && e_hi >= e_lo)
             ? cluster_table_generate(*tb.dir, tb.stamp, tb.epoch, mark, generation, e_lo, e_hi, table) : 0;
}

// One Clenshaw step of the clustered expansion:  dst <- C_k x + (2 S~) src + dst,
// (2 S~) v = G~ (G~^2 + kappa) v,  C_k x = d0 X0 + d1 X1 + d2 X2  (X_j = G~^j x in the columns xc).
// A PRIMING lane runs the very same instructions with src = x and zero coefficients and keeps the two
// intermediate products: X1 = G~ x and X2 = G~^2 x, the start vectors of the (sub)step it is about to begin.
template <class Sys, class XC>
RYDCZ_HD void cluster_step(const double* __restrict__ c, double kappa, double d0, double d1, double d2, const XC xc,
                           const double* __restrict__ src, double* __restrict__ dst, bool priming) {
  constexpr int N = Sys::N;
  double t[N], u[N];
  RYDCZ_UNROLL
  for (int i = 0; i < N; ++i) t[i] = 0.0;
  Sys::apply(c, src, t);
  RYDCZ_UNROLL
  for (int i = 0; i < N; ++i) u[i] = kappa * src[i];
  Sys::apply(c, t, u);
  if (priming) {
    RYDCZ_UNROLL
    for (int i = 0; i < N; ++i) { xc[i] = src[i]; xc[N + i] = t[i]; xc[2 * N + i] = fma(-kappa, src[i], u[i]); }
  }
  RYDCZ_UNROLL
  for (int i = 0; i < N; ++i) dst[i] = fma(d0, xc[i], fma(d1, xc[N + i], fma(d2, xc[2 * N + i], dst[i])));
  Sys::apply(c, u, dst);
}

// run_system for the clustered path.  Same flat loop: one iteration = two Clenshaw steps of whatever
// (sub)step the lane is in.  A (sub)step takes roundup(K + 2, CL_ALIGN) steps -- the priming step, K + 1
// recurrence steps, padding with zero coefficients in front -- so the lanes of a warp reach their
// transitions at common iterations, and a transition itself is only the header / segment bookkeeping.
// `rec` = the point's coefficient records, `xc` = 3 N per-thread columns for the start vectors, `rc` =
// 3 RING per-thread columns: the ring through which the records arrive (record k sits in slot k % RING,
// copied asynchronously RING - 1 steps before step k reads it; every step commits one copy group), followed
// by CL_LOOKAHEAD columns: the NEXT segment row and the NEXT slot header, copied asynchronously one transition
// ahead, so that a transition reads shared memory instead of waiting for two dependent global loads (ncu,
// round 2: those two loads were 12 % of the kernel's stall samples).
constexpr int CL_LOOKAHEAD = 10;    // per-thread columns behind the record ring: segment row (4) + slot header (6)
template <class Sys, int RING, class XC>
RYDCZ_HD void run_system_cluster(double (&x)[Sys::N], const XC xc, const XC rc, const PointView pt,
                                 const double* __restrict__ seg, int n_seg, bool noisy,
                                 const double* __restrict__ rec, const double* __restrict__ table, double& steps) {
  constexpr int N = Sys::N;
  constexpr int LA = 3 * RING;      // first lookahead column
  static_assert((RING & (RING - 1)) == 0 && RING >= 2, "ring size");
  double c[Sys::NC], b1[N], b2[N];
  Sys::init(b1);
  int k = -1, K = -1, nsub = 1, sub_left = 0, seg_left = 0, s = 0;
  bool priming = false;
  double kappa = 0.0, inv_gs = 0.0;
  const double* terms = rec;
  const double* hdr_next = rec;       // header of the slot after the current one
  double next_position = 0.0;         // where that slot's records are (cluster_terms_at)
  // the (sub)step after the current one, for the ring: its records and term count (nK < 0: there is none)
  const double* nterms = rec;
  int nK = -1, K_next = 0;
  // Record kk of the current (sub)step into its ring slot.  kk < 0 continues the countdown ACROSS the boundary: the
  // first records of the next (sub)step (index roundup(nK + 2) + kk), so that a (sub)step never starts by waiting
  // for records it has only just asked for (ncu, round 2: that wait was 7 % of the kernel's stall samples).
#define RYDCZ_RING_FETCH(kk)                                                                    \
  do {                                                                                          \
    int kf_ = (kk);                                                                             \
    const double* tf_ = terms;                                                                  \
    int Kf_ = K;                                                                                \
    if (kf_ < 0) { kf_ += cluster_roundup(nK + 2); tf_ = nterms; Kf_ = nK; }                    \
    if (kf_ >= 0 && kf_ <= Kf_) {                                                               \
      const int sl_ = (kf_ & (RING - 1)) * 3;                                                   \
      RYDCZ_ASYNC_COPY8(&rc[sl_], tf_ + kf_ * CL_REC);                                          \
      RYDCZ_ASYNC_COPY8(&rc[sl_ + 1], tf_ + kf_ * CL_REC + 1);                                  \
      RYDCZ_ASYNC_COPY8(&rc[sl_ + 2], tf_ + kf_ * CL_REC + 2);                                  \
    }                                                                                           \
    RYDCZ_ASYNC_COMMIT();                                                                       \
  } while (0)
#define RYDCZ_LOOKAHEAD_ROW(ss)                                                                 \
  do {                                                                                          \
    const double* sg_ = seg + (long long)(ss) * SG_STRIDE;                                      \
    RYDCZ_UNROLL                                                                                \
    for (int q_ = 0; q_ < SG_STRIDE; ++q_) RYDCZ_ASYNC_COPY8(&rc[LA + q_], sg_ + q_);           \
  } while (0)
#define RYDCZ_LOOKAHEAD_HEADER(hh)                                                              \
  do {                                                                                          \
    RYDCZ_UNROLL                                                                                \
    for (int q_ = 0; q_ < 6; ++q_) RYDCZ_ASYNC_COPY8(&rc[LA + 4 + q_], (hh) + q_);              \
  } while (0)
  // the first row and the first header: their latency is paid once per system, before the first transition
  if (n_seg > 0) RYDCZ_LOOKAHEAD_ROW(0);
  RYDCZ_LOOKAHEAD_HEADER(rec);
  RYDCZ_ASYNC_COMMIT();
  RYDCZ_ASYNC_WAIT(0);
  bool first_substep = true;
  for (;;) {
    if (k < 0) {
      // b1 holds the state.  Next sub-step of this segment, or the next segment (and maybe the next slot).
      if (sub_left > 0) {
        --sub_left;
      } else {
        // Row s sits in the lookahead columns (copied one transition ago: its copy group was waited for by the
        // first step of the (sub)step that just ended).  Inactive rows (zero or NaN duration: the reference
        // skips them) are rare and walked with direct loads.
        double dt = 0.0, re = 0.0, im = 0.0, det = 0.0;
        bool found = false;
        if (s < n_seg) {
          dt = rc[LA]; re = rc[LA + 1]; im = rc[LA + 2]; det = rc[LA + 3];
          ++s;
          found = segment_active(dt);
        }
        while (!found && s < n_seg) {
          const double* sg = seg + (long long)s * SG_STRIDE;
          dt = sg[SG_DT]; re = sg[SG_RE]; im = sg[SG_IM]; det = sg[SG_DET];
          ++s;
          found = segment_active(dt);
        }
        if (!found) break;
        double p[Sys::NP];
        segment_params<Sys>(p, pt, re, im, det, noisy);
        if (seg_left == 0) {
          const int kn = (int)rc[LA + 4];
          K = kn & 1023; nsub = kn >> 10;
          kappa = rc[LA + 5]; inv_gs = rc[LA + 6];
          const int cn = (int)rc[LA + 7];
          seg_left = cn & 0xffff; K_next = cn >> 16;
          terms = cluster_terms_at(rc[LA + 8], rec, table);
          next_position = rc[LA + 9];
          hdr_next += CL_HDR;
          RYDCZ_LOOKAHEAD_HEADER(hdr_next);          // (past the last slot: the point's own records, harmless)
        }
        --seg_left;
        if (s < n_seg) RYDCZ_LOOKAHEAD_ROW(s);       // these copies join the first ring group committed below
        if (!Sys::HERMITIAN) p[Sys::NP - 1] = 0.0;   // no spectral shift on this path
        Sys::coefs(p, c);
        RYDCZ_UNROLL
        for (int q = 0; q < Sys::NC; ++q) c[q] *= inv_gs;
        sub_left = nsub - 1;
      }
      steps += (double)(K + 1);
      RYDCZ_UNROLL
      for (int i = 0; i < N; ++i) b2[i] = 0.0;
      k = cluster_roundup(K + 2) - 1;    // the priming step (k > K: zero coefficients), then k - 1 ... 0
      priming = true;
      nK = -1;                           // the (sub)step after this one is looked up when the ring first needs it
      if (first_substep) {
        // fill the ring: records k ... k - (RING - 2); step j then fetches record j - (RING - 1).  Later (sub)steps
        // find their first records already in the ring (fetched across the boundary, below).
        RYDCZ_UNROLL
        for (int q = 0; q < RING - 1; ++q) RYDCZ_RING_FETCH(k - q);
        first_substep = false;
      }
    }
    if (k - (RING - 1) < 0 && nK < 0) {
      // The ring is about to run past this (sub)step: what comes next?  The same records again (further sub-steps,
      // or further segments of this slot), or the next slot, whose term count this slot's header carries.
      if (sub_left > 0 || seg_left > 0) { nterms = terms; nK = K; }
      else if (K_next > 0) { nterms = cluster_terms_at(next_position, rec, table); nK = K_next; }
      else { nterms = terms; nK = 0; }
    }
    // step k: b2 <- C_k x + 2S~ b1 + b2      (priming: b1 = x, and both vectors restart from zero afterwards)
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    RYDCZ_ASYNC_WAIT(RING - 2);
    if (k <= K) { const int sl = (k & (RING - 1)) * 3; d0 = rc[sl]; d1 = rc[sl + 1]; d2 = rc[sl + 2]; }
    RYDCZ_RING_FETCH(k - (RING - 1));
    cluster_step<Sys>(c, kappa, d0, d1, d2, xc, b1, b2, priming);
    if (priming) {
      RYDCZ_UNROLL
      for (int i = 0; i < N; ++i) { b1[i] = 0.0; b2[i] = 0.0; }
      priming = false;
    }
    // step k-1: b1 <- C_{k-1} x + 2S~ b2 + b1; the last one (k-1 = 0) takes S~ b2 instead: halve b2
    d0 = 0.0; d1 = 0.0; d2 = 0.0;
    RYDCZ_ASYNC_WAIT(RING - 2);
    if (k - 1 <= K) { const int sl = ((k - 1) & (RING - 1)) * 3; d0 = rc[sl]; d1 = rc[sl + 1]; d2 = rc[sl + 2]; }
    RYDCZ_RING_FETCH(k - 1 - (RING - 1));
    if (k == 1) {
      RYDCZ_UNROLL
      for (int i = 0; i < N; ++i) b2[i] *= 0.5;
    }
    cluster_step<Sys>(c, kappa, d0, d1, d2, xc, b2, b1, false);
    k -= 2;
  }
#undef RYDCZ_RING_FETCH
#undef RYDCZ_LOOKAHEAD_ROW
#undef RYDCZ_LOOKAHEAD_HEADER
  RYDCZ_ASYNC_WAIT(0);
  RYDCZ_UNROLL
  for (int i = 0; i < N; ++i) x[i] = b1[i];
}

// One full gate evaluation on the clustered path (identical atoms; the |01> input has no blockade shift
// and stays on the plain path).  Returns false when the invariant check fails (trace of rho_11, norm of
// the noiseless ket): the caller then leaves the point to the plain kernel.
// slots of the record ring: 8 for the 3-level model; 4 for the 4-level one, whose 63 start-vector columns
// leave less shared memory at two CTAs per SM
#ifndef RYDCZ_CL_RING3
#define RYDCZ_CL_RING3 8
#endif
template <int D> struct ClusterRing { static constexpr int value = D == 3 ? RYDCZ_CL_RING3 : 4; };
// (the ring runs across (sub)step boundaries: a (sub)step's length must be a multiple of the ring size, or record k of
//  the next (sub)step would not land in slot k % RING -- measured the hard way with CL_ALIGN = 4 and a ring of 8: every
//  point failed its invariant check and was handed back to the plain kernel)
static_assert(CL_ALIGN % RYDCZ_CL_RING3 == 0 && CL_ALIGN % 4 == 0, "CL_ALIGN must be a multiple of the record ring sizes");
// (rho_11 and the coherence advanced in ONE loop -- records read once, transitions shared, two independent
// chains -- was measured on B200 in round 1: slower, 2.44 ms against 2.22 ms per 262,144 bang-bang candidates,
// 255 registers with 45 doubles spilled; gpurun_out/cluster_ab_r01p.log.  The code was removed in round 2.)

#define RYDCZ_DEFINE_EVALUATE_CLUSTER(D)                                                            \
  template <bool NOISY>                                                                             \
  RYDCZ_HD bool evaluate_point_cluster_d##D(const PointView pt, const Col xcol, const Col rcol,     \
                                            const double* __restrict__ seg, int n_seg,              \
                                            const double* __restrict__ rec,                         \
                                            const double* __restrict__ table,                       \
                                            double* __restrict__ out, double* __restrict__ states) {\
    typedef StateLayout<D> SL;                                                                      \
    double steps = 0.0, flops = 0.0, st;                                                            \
    double p01 = 1.0, p11 = 1.0, loss = 0.0, ryd = 0.0;                                             \
    double c01r, c01i, c11r, c11i;                                                                  \
    bool ok = true;                                                                                 \
    double xcoh[Sys_coh11s_d##D::N];                                                                \
    if (NOISY) {                                                                                    \
      double x[Sys_rho11s_d##D::N]; st = 0.0;                                                       \
      run_system_cluster<Sys_rho11s_d##D, ClusterRing<D>::value>(x, xcol, rcol, pt, seg, n_seg, true, rec, table, st); \
      RYDCZ_RECONVERGE(pt.live);                                                                    \
      steps += st; flops += st * 2.0 * (3 * Sys_rho11s_d##D::NNZ + 4 * Sys_rho11s_d##D::N);         \
      p11 = x[RHO11S_POP_D##D];                                                                     \
      double xf[Sys_rho11s_d##D::FULL_N];                                                           \
      Sys_rho11s_d##D::expand(x, xf);                                                               \
      loss = rho11_loss_d##D(xf); ryd = rho11_ryd_d##D(xf);                                         \
      ok = fabs(rho11_trace_d##D(xf) - 1.0) <= 1e-9;                                                \
      if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_rho11s_d##D::FULL_N; ++i) states[SL::RHO11 + i] = xf[i]; } \
    }                                                                                               \
    {                                                                                               \
      double (&x)[Sys_coh11s_d##D::N] = xcoh;                                                       \
      double cf = 0.0;                                                                              \
      st = 0.0;                                                                                     \
      if (ClosedCoh11s<D>::template run<NOISY>(x, pt, seg, n_seg, cf)) flops += cf;                 \
      else {                                                                                        \
        run_system_cluster<Sys_coh11s_d##D, ClusterRing<D>::value>(x, xcol, rcol, pt, seg, n_seg, NOISY, rec, table, st); \
        steps += st; flops += st * 2.0 * (3 * Sys_coh11s_d##D::NNZ + 4 * Sys_coh11s_d##D::N);       \
      }                                                                                             \
      RYDCZ_RECONVERGE(pt.live);                                                                    \
      c11r = x[COH11S_RE_D##D]; c11i = x[COH11S_IM_D##D];                                           \
      if (!NOISY) ryd = 1.0 - (c11r * c11r + c11i * c11i);                                          \
      if (!NOISY || states) {                                                                       \
        double xf[Sys_coh11s_d##D::FULL_N];                                                         \
        Sys_coh11s_d##D::expand(x, xf);                                                             \
        if (!NOISY) ok = fabs(coh11_norm2_d##D(xf) - 1.0) <= 1e-9;                                  \
        if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh11s_d##D::FULL_N; ++i) states[SL::COH11 + i] = xf[i]; } \
      }                                                                                             \
    }                                                                                               \
    if (!ok) return false;                                                                          \
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
    return true;                                                                                    \
  }

RYDCZ_DEFINE_EVALUATE_CLUSTER(3)
RYDCZ_DEFINE_EVALUATE_CLUSTER(4)

// The same for atoms that see different detunings (Monte-Carlo shots with Doppler offsets): the general
// 25-variable rho_11 and 8-variable coherence on the clustered expansion with the two-atom bounds of
// cluster_geometry(sym = false); the |01> and |10> inputs are different single-atom problems and stay on the
// plain loop.  3-level model only (the general 4-level rho_11 has 36 variables: four vectors of it do not fit).
constexpr int CL_RING_GENERAL = 4;   // 75 start-vector columns + a four-slot record ring + the lookahead: 107 columns, two CTAs per SM
template <bool NOISY>
RYDCZ_HD bool evaluate_point_cluster_general_d3(const PointView pt, const Col xcol, const Col rcol,
                                                const double* __restrict__ seg, int n_seg,
                                                const double* __restrict__ rec, const double* __restrict__ table,
                                                double* __restrict__ out, double* __restrict__ states) {
  typedef StateLayout<3> SL;
  double steps = 0.0, flops = 0.0, st;
  double p01 = 1.0, p10 = 1.0, p11 = 1.0, loss = 0.0, ryd = 0.0;
  double c01r, c01i, c10r, c10i, c11r, c11i;
  bool ok = true;
  if (NOISY) {
    double x[Sys_rho11_d3::N]; st = 0.0;
    run_system_cluster<Sys_rho11_d3, CL_RING_GENERAL>(x, xcol, rcol, pt, seg, n_seg, true, rec, table, st);
    RYDCZ_RECONVERGE(pt.live);
    steps += st; flops += st * 2.0 * (3 * Sys_rho11_d3::NNZ + 4 * Sys_rho11_d3::N);
    p11 = x[RHO11_POP_D3];
    loss = rho11_loss_d3(x); ryd = rho11_ryd_d3(x);
    ok = fabs(rho11_trace_d3(x) - 1.0) <= 1e-9;
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_rho11_d3::N; ++i) states[SL::RHO11 + i] = x[i]; }
  }
  {
    double x[Sys_coh11_d3::N], cf = 0.0; st = 0.0;
    // closed form (rydcz_closed.cuh) where it is accepted: strong blockade and Doppler-sized offsets, i.e. every shot
    if (ClosedCoh11<3>::run<NOISY>(x, pt, seg, n_seg, cf)) flops += cf;
    else {
      run_system_cluster<Sys_coh11_d3, CL_RING_GENERAL>(x, xcol, rcol, pt, seg, n_seg, NOISY, rec, table, st);
      steps += st; flops += st * 2.0 * (3 * Sys_coh11_d3::NNZ + 4 * Sys_coh11_d3::N);
    }
    RYDCZ_RECONVERGE(pt.live);
    c11r = x[COH11_RE_D3]; c11i = x[COH11_IM_D3];
    if (!NOISY) { ryd = 1.0 - (c11r * c11r + c11i * c11i); ok = fabs(coh11_norm2_d3(x) - 1.0) <= 1e-9; }
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh11_d3::N; ++i) states[SL::COH11 + i] = x[i]; }
  }
  if (!ok) return false;
  if (NOISY) {
    // single-atom inputs: rho_ab on its Chebyshev loop, the |ab><00| coherence in closed form (rydcz_closed.cuh)
    double x[Sys_rho01_d3::N];
    single_atom_rho<Sys_rho01_d3>(x, pt, seg, n_seg, true, steps, flops);
    p01 = x[RHO01_POP_D3];
    double xc[Sys_coh01_d3::N];
    single_atom_coherence<Sys_coh01_d3, true>(xc, pt, seg, n_seg, true, steps, flops);
    c01r = xc[COH01_RE_D3]; c01i = xc[COH01_IM_D3];
    double y[Sys_rho10_d3::N];
    single_atom_rho<Sys_rho10_d3>(y, pt, seg, n_seg, false, steps, flops);
    p10 = y[RHO10_POP_D3];
    double yc[Sys_coh10_d3::N];
    single_atom_coherence<Sys_coh10_d3, true>(yc, pt, seg, n_seg, false, steps, flops);
    c10r = yc[COH10_RE_D3]; c10i = yc[COH10_IM_D3];
    if (states) {
      RYDCZ_UNROLL for (int i = 0; i < Sys_rho01_d3::N; ++i) states[SL::RHO01 + i] = x[i];
      RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d3::N; ++i) states[SL::COH01 + i] = xc[i];
      RYDCZ_UNROLL for (int i = 0; i < Sys_rho10_d3::N; ++i) states[SL::RHO10 + i] = y[i];
      RYDCZ_UNROLL for (int i = 0; i < Sys_coh10_d3::N; ++i) states[SL::COH10 + i] = yc[i];
    }
  } else {
    double x[Sys_coh01_d3::N];
    single_atom_coherence<Sys_coh01_d3, false>(x, pt, seg, n_seg, true, steps, flops);
    c01r = x[COH01_RE_D3]; c01i = x[COH01_IM_D3];
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh01_d3::N; ++i) states[SL::COH01 + i] = x[i]; }
    double y[Sys_coh10_d3::N];
    single_atom_coherence<Sys_coh10_d3, false>(y, pt, seg, n_seg, false, steps, flops);
    c10r = y[COH10_RE_D3]; c10i = y[COH10_IM_D3];
    if (states) { RYDCZ_UNROLL for (int i = 0; i < Sys_coh10_d3::N; ++i) states[SL::COH10 + i] = y[i]; }
  }
  cz_epilogue(NOISY, p01, p10, p11, c01r, c01i, c10r, c10i, c11r, c11i, out);
  out[OUT_LOSS] = loss; out[OUT_RYD] = ryd; out[OUT_STEPS] = steps; out[OUT_FLOPS] = flops;
  return true;
}

}  // namespace rydcz
