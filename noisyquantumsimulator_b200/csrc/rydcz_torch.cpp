// rydcz_torch.cpp -- the PyTorch face of the C ABI (include/rydcz.h): torch.ops.rydcz.*
//
// A thin dispatcher registration, nothing else: CUDA float64 tensors in, CUDA tensors out, work enqueued on the
// current torch stream, errors as TORCH_CHECK failures, per-point trouble as NaN rows + the status column
// (rydcz_point_status), never an exception in the middle of a batch.  All arithmetic is
// in librydcz.so; this file only owns one evaluator context per device.
//
//   rydcz::evaluate(points[B,12|20], segs[B,S,4], n_seg[B] int32 or None, dim, noisy, want_states) -> (out[B,16], states)
//   rydcz::evaluate_shots(points, segs, n_seg, dim, noisy, n_shots, seed, want_shots) -> (mean[B,16], shots[B,n,16])
//   rydcz::simulate(model[32] (CPU), app[40,B], n_shots, seed) -> (out[B,16], scalars[B,8])
//   rydcz::point_status(n) -> int32[n] (CPU)      status column of the last call on the current device
//
// Replaces, for a batch, RG/simulation.py:3364-3572 (evolve_* + compute_CZ_fidelity) and, for simulate,
// RG/simulation.py:2820-3572.
#include <ATen/cuda/CUDAContext.h>
#include <cuda_runtime_api.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/torch.h>

#include <map>
#include <mutex>

#include "../../include/rydcz.h"

namespace {

std::mutex g_mutex;
std::map<int, rydcz_ctx*> g_ctx;

rydcz_ctx* context_for(int device) {
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g_ctx.find(device);
  if (it != g_ctx.end()) return it->second;
  rydcz_ctx* ctx = nullptr;
  const int rc = rydcz_create(device, &ctx);
  TORCH_CHECK(rc == 0 && ctx, "rydcz_create(device=", device, ") failed: ", rydcz_last_error(nullptr));
  g_ctx[device] = ctx;
  return ctx;
}

// torch's default stream has the handle 0, which the C ABI reads as "the context's own stream": name it explicitly
// (cudaStreamLegacy) so that the work is ordered with whatever the caller enqueues on the default stream next.
void* current_stream(int device) {
  cudaStream_t s = at::cuda::getCurrentCUDAStream(device).stream();
  return s ? (void*)s : (void*)cudaStreamLegacy;
}

void check_rc(rydcz_ctx* ctx, int rc, const char* what) {
  TORCH_CHECK(rc == 0, what, " failed (", rc, "): ", rydcz_last_error(ctx));
}

void check_batch(const at::Tensor& points, const at::Tensor& segs, const c10::optional<at::Tensor>& n_seg, int64_t dim) {
  TORCH_CHECK(points.is_cuda() && segs.is_cuda(), "rydcz: points and segs must be CUDA tensors");
  TORCH_CHECK(points.scalar_type() == at::kDouble && segs.scalar_type() == at::kDouble, "rydcz: float64 tensors expected");
  TORCH_CHECK(points.device() == segs.device(), "rydcz: points and segs on different devices");
  TORCH_CHECK(dim == 3 || dim == 4 || dim == 5, "rydcz: dim must be 3, 4 or 5");
  TORCH_CHECK(points.dim() == 2 && points.size(1) == rydcz_point_stride_dim((int)dim), "rydcz: points must be [B, ",
              rydcz_point_stride_dim((int)dim), "]");
  TORCH_CHECK(segs.dim() == 3 && segs.size(0) == points.size(0) && segs.size(2) == RYDCZ_SEG_STRIDE,
              "rydcz: segs must be [B, S, 4]");
  if (n_seg.has_value()) {
    TORCH_CHECK(n_seg->is_cuda() && n_seg->scalar_type() == at::kInt && n_seg->dim() == 1 &&
                    n_seg->size(0) == points.size(0), "rydcz: n_seg must be a CUDA int32 tensor of length B");
  }
}

std::tuple<at::Tensor, at::Tensor> evaluate(const at::Tensor& points_, const at::Tensor& segs_,
                                            const c10::optional<at::Tensor>& n_seg_, int64_t dim, bool noisy,
                                            bool want_states) {
  check_batch(points_, segs_, n_seg_, dim);
  const c10::cuda::CUDAGuard guard(points_.device());
  const at::Tensor points = points_.contiguous(), segs = segs_.contiguous();
  const int64_t B = points.size(0);
  at::Tensor out = at::empty({B, RYDCZ_OUT_STRIDE}, points.options());
  at::Tensor states = at::empty({want_states ? B : 0, rydcz_state_stride((int)dim)}, points.options());
  at::Tensor ns;
  if (n_seg_.has_value()) ns = n_seg_->contiguous();
  rydcz_ctx* ctx = context_for(points.get_device());
  check_rc(ctx, rydcz_evaluate_device(ctx, (int)dim, noisy ? 1 : 0, B, points.data_ptr<double>(), segs.data_ptr<double>(),
                                      n_seg_.has_value() ? ns.data_ptr<int32_t>() : nullptr, (int)segs.size(1),
                                      out.data_ptr<double>(), want_states ? states.data_ptr<double>() : nullptr,
                                      current_stream(points.get_device())),
           "rydcz_evaluate_device");
  return std::make_tuple(out, states);
}

std::tuple<at::Tensor, at::Tensor> evaluate_shots(const at::Tensor& points_, const at::Tensor& segs_,
                                                  const c10::optional<at::Tensor>& n_seg_, int64_t dim, bool noisy,
                                                  int64_t n_shots, int64_t seed, bool want_shots) {
  check_batch(points_, segs_, n_seg_, dim);
  TORCH_CHECK(n_shots >= 1, "rydcz: n_shots must be >= 1");
  const c10::cuda::CUDAGuard guard(points_.device());
  const at::Tensor points = points_.contiguous(), segs = segs_.contiguous();
  const int64_t B = points.size(0);
  at::Tensor mean = at::empty({B, RYDCZ_OUT_STRIDE}, points.options());
  at::Tensor shots = at::empty({want_shots ? B : 0, n_shots, RYDCZ_OUT_STRIDE}, points.options());
  at::Tensor ns;
  if (n_seg_.has_value()) ns = n_seg_->contiguous();
  rydcz_ctx* ctx = context_for(points.get_device());
  check_rc(ctx, rydcz_evaluate_shots_device(ctx, (int)dim, noisy ? 1 : 0, B, (int)n_shots, (uint64_t)seed,
                                            points.data_ptr<double>(), segs.data_ptr<double>(),
                                            n_seg_.has_value() ? ns.data_ptr<int32_t>() : nullptr, (int)segs.size(1),
                                            mean.data_ptr<double>(), want_shots ? shots.data_ptr<double>() : nullptr,
                                            current_stream(points.get_device())),
           "rydcz_evaluate_shots_device");
  return std::make_tuple(mean, shots);
}

std::tuple<at::Tensor, at::Tensor> simulate(const at::Tensor& model_, const at::Tensor& app_, int64_t n_shots,
                                            int64_t seed) {
  TORCH_CHECK(app_.is_cuda() && app_.scalar_type() == at::kDouble && app_.dim() == 2 && app_.size(0) == RYDCZ_APP_STRIDE,
              "rydcz: app must be a CUDA float64 tensor [", RYDCZ_APP_STRIDE, ", B] (slot-major)");
  TORCH_CHECK(model_.numel() == RYDCZ_MODEL_STRIDE && model_.scalar_type() == at::kDouble,
              "rydcz: model must hold ", RYDCZ_MODEL_STRIDE, " float64 values");
  const c10::cuda::CUDAGuard guard(app_.device());
  const at::Tensor model = model_.cpu().contiguous();
  const at::Tensor app = app_.contiguous();
  const int64_t B = app.size(1);
  at::Tensor out = at::empty({B, RYDCZ_OUT_STRIDE}, app.options());
  at::Tensor scalars = at::empty({B, RYDCZ_SCALAR_STRIDE}, app.options());
  rydcz_ctx* ctx = context_for(app.get_device());
  check_rc(ctx, rydcz_simulate_device(ctx, model.data_ptr<double>(), B, app.data_ptr<double>(), (int)n_shots, (uint64_t)seed,
                                      out.data_ptr<double>(), scalars.data_ptr<double>(), nullptr,
                                      current_stream(app.get_device())),
           "rydcz_simulate_device");
  return std::make_tuple(out, scalars);
}

at::Tensor point_status(int64_t n) {
  const int device = (int)c10::cuda::current_device();
  rydcz_ctx* ctx = context_for(device);
  at::Tensor st = at::empty({n}, at::TensorOptions().dtype(at::kInt).device(at::kCPU));
  check_rc(ctx, rydcz_point_status(ctx, n, st.data_ptr<int32_t>()), "rydcz_point_status");
  return st;
}

// shape functions for tracing / FakeTensor propagation
std::tuple<at::Tensor, at::Tensor> evaluate_meta(const at::Tensor& points, const at::Tensor& segs,
                                                 const c10::optional<at::Tensor>&, int64_t dim, bool, bool want_states) {
  return std::make_tuple(at::empty({points.size(0), RYDCZ_OUT_STRIDE}, points.options()),
                         at::empty({want_states ? points.size(0) : 0, rydcz_state_stride((int)dim)}, points.options()));
}
std::tuple<at::Tensor, at::Tensor> evaluate_shots_meta(const at::Tensor& points, const at::Tensor&,
                                                       const c10::optional<at::Tensor>&, int64_t, bool, int64_t n_shots,
                                                       int64_t, bool want_shots) {
  return std::make_tuple(at::empty({points.size(0), RYDCZ_OUT_STRIDE}, points.options()),
                         at::empty({want_shots ? points.size(0) : 0, n_shots, RYDCZ_OUT_STRIDE}, points.options()));
}
std::tuple<at::Tensor, at::Tensor> simulate_meta(const at::Tensor&, const at::Tensor& app, int64_t, int64_t) {
  return std::make_tuple(at::empty({app.size(1), RYDCZ_OUT_STRIDE}, app.options()),
                         at::empty({app.size(1), RYDCZ_SCALAR_STRIDE}, app.options()));
}

}  // namespace

TORCH_LIBRARY(rydcz, m) {
  m.def("evaluate(Tensor points, Tensor segs, Tensor? n_seg, int dim, bool noisy, bool want_states=False) -> (Tensor, Tensor)");
  m.def("evaluate_shots(Tensor points, Tensor segs, Tensor? n_seg, int dim, bool noisy, int n_shots, int seed=1234, "
        "bool want_shots=False) -> (Tensor, Tensor)");
  m.def("simulate(Tensor model, Tensor app, int n_shots=0, int seed=1234) -> (Tensor, Tensor)");
  m.def("point_status(int n) -> Tensor");
}
TORCH_LIBRARY_IMPL(rydcz, CUDA, m) {
  m.impl("evaluate", &evaluate);
  m.impl("evaluate_shots", &evaluate_shots);
  m.impl("simulate", &simulate);
}
TORCH_LIBRARY_IMPL(rydcz, Meta, m) {
  m.impl("evaluate", &evaluate_meta);
  m.impl("evaluate_shots", &evaluate_shots_meta);
  m.impl("simulate", &simulate_meta);
}
TORCH_LIBRARY_IMPL(rydcz, CompositeExplicitAutograd, m) { m.impl("point_status", &point_status); }
