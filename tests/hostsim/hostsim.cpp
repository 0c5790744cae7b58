// Host build of the per-evaluation arithmetic in csrc/rydcz_core.cuh -- TEST HARNESS ONLY.
// Lets the CPU test-suite check the Chebyshev/Clenshaw evolution and the generated systems
// against the oracle without a GPU.  Never loaded by the product package.
#include <cstddef>
#define RYDCZ_HD inline
#define RYDCZ_UNROLL
#include "../../noisyquantumsimulator_b200/csrc/rydcz_core.cuh"
#include "../../noisyquantumsimulator_b200/csrc/rydcz_prologue.cuh"

#include <cstdlib>
#include <vector>

static int g_allow_sym = 1;
extern "C" void hostsim_allow_sym(int on) { g_allow_sym = on; }
// clustered-spectrum path (csrc/rydcz_cluster.cuh): 0 = off, 1 = where the planner prefers it (the CUDA
// library's rule), 2 = wherever it is eligible
static int g_cluster = 0;
static long long g_cluster_used = 0, g_cluster_retry = 0, g_repeat_used = 0;
extern "C" long long hostsim_repeat_used() { const long long n = g_repeat_used; g_repeat_used = 0; return n; }
extern "C" void hostsim_set_cluster(int mode) { g_cluster = mode; }
extern "C" void hostsim_cluster_stats(long long* used, long long* retried) {
  *used = g_cluster_used; *retried = g_cluster_retry;
  g_cluster_used = 0; g_cluster_retry = 0;
}

// shared coefficient table (csrc/rydcz_cluster.cuh): 0 = every slot generates records of its own, 1 = table entries
// generated once per hostsim_evaluate call, 2 = kept across calls.  The records live in one lazily committed
// calloc block (1.2 GB of address space, only the entries that are used get touched).
static int g_table_mode = 1;
static rydcz::ClusterTableDir g_dir;
static std::vector<unsigned> g_stamp, g_mark;
static double* g_table = nullptr;
static unsigned g_epoch = 1, g_generation = 1;
static long long g_table_generated = 0;
extern "C" void hostsim_set_table(int mode) { g_table_mode = mode; }
extern "C" long long hostsim_table_generated() { const long long n = g_table_generated; g_table_generated = 0; return n; }
static rydcz::ClusterTable host_table() {
  using namespace rydcz;
  if (!g_table_mode) return ClusterTable{nullptr, nullptr, 0u};
  if (!g_table) {
    cluster_table_dir_init(g_dir);
    g_stamp.assign((size_t)g_dir.base[CT_CLASSES], 0u);
    g_mark.assign((size_t)g_dir.base[CT_CLASSES], 0u);
    g_table = (double*)calloc((size_t)g_dir.base[CT_CLASSES] * CT_ENTRY_REC * CL_REC + 512, sizeof(double));
  }
  return ClusterTable{&g_dir, g_stamp.data(), g_epoch};
}
// the same directory without the stamps: cluster_plan then claims nothing (cost and eligibility only)
static rydcz::ClusterTable dry(const rydcz::ClusterTable& tb) { return rydcz::ClusterTable{tb.dir, nullptr, 0u}; }

template <int D>
static bool try_cluster(const rydcz::PointView pt, const double* row, const double* sg, int ns, bool noisy,
                        double* o, double* st) {
  using namespace rydcz;
  double cost = 0.0;
  const ClusterTable tb = host_table();
  // (like plan_kernel: the headers are written, and table entries claimed and generated, before the path is chosen)
  std::vector<double> rec((size_t)CL_STRIDE + 512, 0.0);    // (+ pad: header lookahead and ring prefetch past the last slot)
  int n_shared = 0;
  long long e_lo = 0x7fffffffffffffffLL, e_hi = -1;
  const int records = cluster_plan<D>(row, sg, ns, noisy, true, cost, dry(tb), n_shared, e_lo, e_hi);
  if (records <= 0 || records > CL_RECORDS) return false;
  if (g_cluster == 1 && !(cost < 0.8 * point_work<D>(row, sg, ns))) return false;
  g_table_generated += cluster_prepare<D>(row, sg, ns, noisy, true, rec.data(), tb, g_mark.data(), g_generation, g_table);
  double xc[3 * (21 + 6)], ring[3 * 8 + CL_LOOKAHEAD];
  bool ok;
  if (D == 3) ok = noisy ? evaluate_point_cluster_d3<true>(pt, Col{xc, 1}, Col{ring, 1}, sg, ns, rec.data(), g_table, o, st)
                         : evaluate_point_cluster_d3<false>(pt, Col{xc, 1}, Col{ring, 1}, sg, ns, rec.data(), g_table, o, st);
  else ok = noisy ? evaluate_point_cluster_d4<true>(pt, Col{xc, 1}, Col{ring, 1}, sg, ns, rec.data(), g_table, o, st)
                  : evaluate_point_cluster_d4<false>(pt, Col{xc, 1}, Col{ring, 1}, sg, ns, rec.data(), g_table, o, st);
  ++g_cluster_used;
  if (!ok) ++g_cluster_retry;
  return ok;
}

// unequal atoms, 3-level model: the general systems on the clustered path
static bool try_cluster_general(const rydcz::PointView pt, const double* row, const double* sg, int ns, bool noisy,
                                double* o, double* st) {
  using namespace rydcz;
  double cost = 0.0;
  const ClusterTable tb = host_table();
  int n_shared = 0;
  long long e_lo = 0x7fffffffffffffffLL, e_hi = -1;
  const int records = cluster_plan<3>(row, sg, ns, noisy, false, cost, dry(tb), n_shared, e_lo, e_hi);
  if (records <= 0 || records > CL_RECORDS) return false;
  if (g_cluster == 1 && !(cost < 0.8 * point_work<3>(row, sg, ns))) return false;
  std::vector<double> rec((size_t)CL_STRIDE + 512, 0.0);    // (+ pad: header lookahead and ring prefetch past the last slot)
  g_table_generated += cluster_prepare<3>(row, sg, ns, noisy, false, rec.data(), tb, g_mark.data(), g_generation, g_table);
  double xc[3 * 25], ring[3 * 8 + CL_LOOKAHEAD];
  const bool ok = noisy ? evaluate_point_cluster_general_d3<true>(pt, Col{xc, 1}, Col{ring, 1}, sg, ns, rec.data(), g_table, o, st)
                        : evaluate_point_cluster_general_d3<false>(pt, Col{xc, 1}, Col{ring, 1}, sg, ns, rec.data(), g_table, o, st);
  ++g_cluster_used;
  if (!ok) ++g_cluster_retry;
  return ok;
}

// coefficient records of one point, for tests of the cosine transform (returns the number of records)
extern "C" int hostsim_cluster_records(int dim, int noisy, const double* point, const double* segs, int n_seg,
                                       double* rec, double* cost) {
  using namespace rydcz;
  // (records of the point's own: the table is left out, so the test sees every slot's coefficients in `rec`)
  const ClusterTable none{nullptr, nullptr, 0u};
  int n_shared = 0;
  long long e_lo = 0x7fffffffffffffffLL, e_hi = -1;
  const int records = dim == 3 ? cluster_plan<3>(point, segs, n_seg, noisy != 0, true, *cost, none, n_shared, e_lo, e_hi)
                               : cluster_plan<4>(point, segs, n_seg, noisy != 0, true, *cost, none, n_shared, e_lo, e_hi);
  if (records > CL_RECORDS) return -records;
  if (records > 0 && rec) {
    if (dim == 3) cluster_prepare<3>(point, segs, n_seg, noisy != 0, true, rec, none, nullptr, 0u, nullptr);
    else cluster_prepare<4>(point, segs, n_seg, noisy != 0, true, rec, none, nullptr, 0u, nullptr);
  }
  return records;
}

extern "C" int hostsim_evaluate(int dim, int noisy, long long n_points, const double* points,
                                const double* segs, const int* n_seg, int max_seg, double* out,
                                double* states) {
  using namespace rydcz;
  const int stride = dim == 5 ? PT5_STRIDE : PT_STRIDE;
  ++g_epoch;                             // requests are per call ...
  if (g_table_mode == 1) ++g_generation; // ... and so are the entries, unless they are kept (mode 2)
  for (long long i = 0; i < n_points; ++i) {
    double row[PT_MAX_STRIDE + PT_SCRATCH];
    for (int k = 0; k < (dim == 5 ? (int)PT5_USED : (int)PT_STRIDE); ++k) row[k] = points[i * stride + k];
    const PointView pt{row, 1};
    const double* sg = segs + i * (long long)max_seg * SG_STRIDE;
    double* o = out + i * OUT_STRIDE;
    // same dispatch rule as the CUDA kernels: identical atoms take the exchange-folded path
    const bool sym = g_allow_sym && pt[PT_OFF_A] == pt[PT_OFF_B];
    // (noiseless, identical atoms, 3-level: closed forms only -- the planner keeps such points on the plain folded path)
    const bool closed_all = dim == 3 && RYDCZ_CLOSED_COHERENCE != 0 && !noisy;
    if (g_cluster && sym && dim == 3 && !closed_all) {
      // repeated-segment path (csrc/rydcz_repeat.cuh): same rule as the CUDA planner
      int first = 0, count = 0;
      if (repeat_eligible(sg, n_seg[i], first, count)) {
        typedef RepeatLayout<3> RL;
        std::vector<double> E((size_t)RL::STRIDE, 0.0);
        for (int col = 0; col < RL::COLS; ++col) {
          if (noisy) repeat_build_d3<true>(col, pt, sg + (long long)first * SG_STRIDE, E.data(), 1);
          else repeat_build_d3<false>(col, pt, sg + (long long)first * SG_STRIDE, E.data(), 1);
        }
        double* st = states ? states + i * StateLayout<3>::STRIDE : nullptr;
        if (noisy) evaluate_point_repeat_d3<true>(E.data(), 1, sg, n_seg[i], o, st);
        else evaluate_point_repeat_d3<false>(E.data(), 1, sg, n_seg[i], o, st);
        ++g_repeat_used;
        continue;
      }
    }
    if (g_cluster && !sym && dim == 3 && g_allow_sym && !closed_all) {
      double* st = states ? states + i * StateLayout<3>::STRIDE : nullptr;
      if (try_cluster_general(pt, row, sg, n_seg[i], noisy != 0, o, st)) continue;
    }
    if (g_cluster && sym && (dim == 3 || dim == 4) && !closed_all) {
      double* st = states ? states + i * (dim == 3 ? StateLayout<3>::STRIDE : StateLayout<4>::STRIDE) : nullptr;
      if (dim == 3 ? try_cluster<3>(pt, row, sg, n_seg[i], noisy != 0, o, st)
                   : try_cluster<4>(pt, row, sg, n_seg[i], noisy != 0, o, st)) continue;
    }
    if (dim == 3) {
      double* st = states ? states + i * StateLayout<3>::STRIDE : nullptr;
      double cc[Sys_rho11s_d3::NC];
      if (sym) { if (noisy) evaluate_point_sym_d3<true>(pt, Col{cc, 1}, sg, n_seg[i], o, st);
                 else evaluate_point_sym_d3<false>(pt, Col{cc, 1}, sg, n_seg[i], o, st); }
      else if (noisy) evaluate_point_d3<true>(pt, sg, n_seg[i], o, st);
      else evaluate_point_d3<false>(pt, sg, n_seg[i], o, st);
    } else if (dim == 4) {
      double* st = states ? states + i * StateLayout<4>::STRIDE : nullptr;
      double cc[Sys_rho11s_d4::NC];
      if (sym) { if (noisy) evaluate_point_sym_d4<true>(pt, Col{cc, 1}, sg, n_seg[i], o, st);
                 else evaluate_point_sym_d4<false>(pt, Col{cc, 1}, sg, n_seg[i], o, st); }
      else if (noisy) evaluate_point_d4<true>(pt, sg, n_seg[i], o, st);
      else evaluate_point_d4<false>(pt, sg, n_seg[i], o, st);
    } else if (dim == 5) {
      double* st = states ? states + i * StateLayout<5>::STRIDE : nullptr;
      if (!sym) {
        // unequal atoms: the general systems, rho_11's three vectors in "shared-memory" columns
        double xg[Sys_rho11_d5::N], b1g[Sys_rho11_d5::N], b2g[Sys_rho11_d5::N];
        if (noisy) evaluate_point_d5<true>(pt, Col{xg, 1}, Col{b1g, 1}, Col{b2g, 1}, sg, n_seg[i], o, st);
        else evaluate_point_d5<false>(pt, Col{xg, 1}, Col{b1g, 1}, Col{b2g, 1}, sg, n_seg[i], o, st);
        continue;
      }
      double xcol[Sys_rho11s_d5::N], ccol[Sys_rho11s_d5::NC];
      if (noisy) evaluate_point_sym_d5<true>(pt, Col{xcol, 1}, Col{ccol, 1}, sg, n_seg[i], o, st);
      else evaluate_point_sym_d5<false>(pt, Col{xcol, 1}, Col{ccol, 1}, sg, n_seg[i], o, st);
    } else return 1;
  }
  return 0;
}

// host build of the device prologue (csrc/rydcz_prologue.cuh)
extern "C" int hostsim_prologue(const double* model, long long n_points, const double* app, double* points,
                                double* segs, int* n_seg, double* scalars) {
  using namespace rydcz;
  const int S = (int)model[MD_NSEG];
  const int stride = (int)model[MD_DIM] == 5 ? (int)PT5_STRIDE : (int)PT_STRIDE;   // rows as the C ABI lays them out
  for (long long i = 0; i < n_points; ++i) {
    double pt[PT5_STRIDE];
    n_seg[i] = prologue_point(model, AppView{app + i, n_points}, pt, segs + i * (long long)S * SG_STRIDE,
                              scalars + i * SC_STRIDE);
    for (int k = 0; k < stride; ++k) points[i * stride + k] = pt[k];
  }
  return 0;
}

// closed-form coherences (csrc/rydcz_closed.cuh) of one point, for tests/test_closed_coherence.py:
// which = 0: exchange-folded |11><00| (6 doubles), 1: |01><00| (atom B driven), 2: |10><00| (4 doubles), 3: |11><00| of
// unequal atoms (8 doubles), 4 / 5: the single-atom density operators rho_01 / rho_10 (5 doubles).  Returns 1 when the closed form was accepted and finite.
extern "C" int hostsim_closed_coherence(int which, int noisy, const double* point, const double* segs, int n_seg,
                                        double* x) {
  using namespace rydcz;
  double row[PT_MAX_STRIDE + PT_SCRATCH];
  for (int k = 0; k < (int)PT_STRIDE; ++k) row[k] = point[k];
  const PointView pt{row, 1};
  double flops = 0.0;
  if (which == 0) {
    double v[6];
    const bool ok = noisy ? coh11s_closed_d3<true>(v, pt, segs, n_seg, flops) : coh11s_closed_d3<false>(v, pt, segs, n_seg, flops);
    for (int k = 0; k < 6; ++k) x[k] = v[k];
    return ok ? 1 : 0;
  }
  if (which == 3) {      // unequal atoms: the general |11><00| coherence (8 doubles)
    double v[8];
    const bool ok = noisy ? coh11_closed_d3<true>(v, pt, segs, n_seg, flops) : coh11_closed_d3<false>(v, pt, segs, n_seg, flops);
    for (int k = 0; k < 8; ++k) x[k] = v[k];
    return ok ? 1 : 0;
  }
  if (which == 4 || which == 5) {      // single-atom rho_01 (atom B driven) / rho_10 (5 doubles); noisy rates always on
    double v[5];
    const bool ok = rho01_closed_d3(v, pt, segs, n_seg, which == 4, flops);
    for (int k = 0; k < 5; ++k) x[k] = v[k];
    return ok ? 1 : 0;
  }
  double v[4];
  const bool ok = noisy ? coh01_closed_d3<true>(v, pt, segs, n_seg, which == 1, flops)
                        : coh01_closed_d3<false>(v, pt, segs, n_seg, which == 1, flops);
  for (int k = 0; k < 4; ++k) x[k] = v[k];
  return ok ? 1 : 0;
}
