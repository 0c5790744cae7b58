// Host build of the per-evaluation arithmetic in csrc/rydcz_core.cuh -- TEST HARNESS ONLY.
// Lets the CPU test-suite check the Chebyshev/Clenshaw evolution and the generated systems
// against the oracle without a GPU.  Never loaded by the product package.
#include <cstddef>
#define RYDCZ_HD inline
#define RYDCZ_UNROLL
#include "../../noisyquantumsimulator_b200/csrc/rydcz_core.cuh"

extern "C" int hostsim_evaluate(int dim, int noisy, long long n_points, const double* points,
                                const double* segs, const int* n_seg, int max_seg, double* out,
                                double* states) {
  using namespace rydcz;
  for (long long i = 0; i < n_points; ++i) {
    const double* pt = points + i * PT_STRIDE;
    const double* sg = segs + i * (long long)max_seg * SG_STRIDE;
    double* o = out + i * OUT_STRIDE;
    if (dim == 3) {
      double* st = states ? states + i * StateLayout<3>::STRIDE : nullptr;
      if (noisy) evaluate_point_d3<true>(pt, sg, n_seg[i], o, st);
      else evaluate_point_d3<false>(pt, sg, n_seg[i], o, st);
    } else if (dim == 4) {
      double* st = states ? states + i * StateLayout<4>::STRIDE : nullptr;
      if (noisy) evaluate_point_d4<true>(pt, sg, n_seg[i], o, st);
      else evaluate_point_d4<false>(pt, sg, n_seg[i], o, st);
    } else return 1;
  }
  return 0;
}
