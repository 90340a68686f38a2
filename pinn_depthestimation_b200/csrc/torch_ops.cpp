// torch custom-op registration over the C ABI ("C-ABI torch custom-op extension"): `torch.ops.pinn_b200.*`.
//
// Each op is a thin shim: it checks tensor types, takes the CURRENT CUDA stream from torch and calls the extern "C" entry
// point of libpinn_b200.so declared in include/pinn_b200.h -- no arithmetic happens here.  The descriptor travels as a
// CPU uint8 tensor holding the bytes of pinn_desc_t (PassSpec.to_desc() on the Python side), so the op schemas contain
// only tensors and scalars.  Reference interfaces replaced: see the header, entry point by entry point.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <cstring>

#include "../../include/pinn_b200.h"

namespace {

void check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "pinn_b200 ", what, " failed (code ", rc, "): ", pinn_last_error());
}
const pinn_desc_t* desc_of(const at::Tensor& d) {
  TORCH_CHECK(d.device().is_cpu() && d.scalar_type() == at::kByte && d.is_contiguous() &&
                  (size_t)d.numel() == sizeof(pinn_desc_t),
              "desc must be a contiguous CPU uint8 tensor of sizeof(pinn_desc_t) = ", sizeof(pinn_desc_t), " bytes");
  return reinterpret_cast<const pinn_desc_t*>(d.data_ptr());
}
const float* f32(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous(), name,
              " must be a contiguous float32 CUDA tensor (this path has no CPU fallback)");
  return t.data_ptr<float>();
}
const float* f32_opt(const c10::optional<at::Tensor>& t, const char* name) { return t.has_value() ? f32(*t, name) : nullptr; }
void* stream_of(const at::Tensor& t) { return (void*)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

int64_t workspace_bytes(const at::Tensor& desc, int64_t n_points, bool want_grad) {
  size_t b = 0;
  check(pinn_workspace_bytes_ex(desc_of(desc), n_points, want_grad ? 1 : 0, &b), "pinn_workspace_bytes_ex");
  return (int64_t)b;
}

// one pass: returns the raw sums double[16]; fills grad / out in place when given
at::Tensor jet_loss(const at::Tensor& desc, const at::Tensor& params, const at::Tensor& inputs,
                    const c10::optional<at::Tensor>& targets, const c10::optional<at::Tensor>& mask_count,
                    int64_t n_res_global, int64_t n_fid_global, at::Tensor workspace, c10::optional<at::Tensor> grad,
                    c10::optional<at::Tensor> out, int64_t flags) {
  const c10::cuda::CUDAGuard guard(params.device());
  pinn_eval_args_t a;
  std::memset(&a, 0, sizeof(a));
  a.params = f32(params, "params");
  a.inputs = inputs.numel() ? f32(inputs, "inputs") : nullptr;
  a.targets = f32_opt(targets, "targets");
  a.n_points = inputs.size(0);
  a.n_res_global = n_res_global;
  a.n_fid_global = n_fid_global;
  a.mask_count = f32_opt(mask_count, "mask_count");
  a.grad = grad.has_value() ? const_cast<float*>(f32(*grad, "grad")) : nullptr;
  a.out = out.has_value() ? const_cast<float*>(f32(*out, "out")) : nullptr;
  TORCH_CHECK(workspace.is_cuda() && workspace.scalar_type() == at::kByte, "workspace must be a CUDA uint8 tensor");
  const uintptr_t base = (uintptr_t)workspace.data_ptr();
  const uintptr_t al = (base + 255) & ~uintptr_t(255);
  a.workspace = (void*)al;
  a.workspace_bytes = (size_t)workspace.numel() - (size_t)(al - base);
  a.flags = (int32_t)flags;
  at::Tensor sums = at::zeros({PINN_NSUMS}, params.options().dtype(at::kDouble));
  a.sums = sums.data_ptr<double>();
  if (grad.has_value()) check(pinn_jet_loss_fwdbwd(desc_of(desc), &a, stream_of(params)), "pinn_jet_loss_fwdbwd");
  else check(pinn_jet_loss_fwd(desc_of(desc), &a, stream_of(params)), "pinn_jet_loss_fwd");
  return sums;
}

at::Tensor mask_count(const at::Tensor& desc, const at::Tensor& inputs) {
  const c10::cuda::CUDAGuard guard(inputs.device());
  at::Tensor c = at::zeros({1}, inputs.options());
  check(pinn_mask_count(desc_of(desc), f32(inputs, "inputs"), inputs.size(0), c.data_ptr<float>(), stream_of(inputs)),
        "pinn_mask_count");
  return c;
}

at::Tensor loss_finalize(const at::Tensor& desc, const at::Tensor& sums, const c10::optional<at::Tensor>& sums_b,
                         int64_t n_fid_global, int64_t n_res_global, const c10::optional<at::Tensor>& mask_count_t) {
  const c10::cuda::CUDAGuard guard(sums.device());
  TORCH_CHECK(sums.is_cuda() && sums.scalar_type() == at::kDouble && sums.numel() == PINN_NSUMS, "sums: CUDA double[16]");
  at::Tensor parts = at::zeros({4}, sums.options().dtype(at::kFloat));
  check(pinn_loss_finalize(desc_of(desc), sums.data_ptr<double>(), sums_b.has_value() ? sums_b->data_ptr<double>() : nullptr,
                           n_fid_global, n_res_global, f32_opt(mask_count_t, "mask_count"), parts.data_ptr<float>(),
                           stream_of(sums)),
        "pinn_loss_finalize");
  return parts;
}

at::Tensor lbfgs_direction(const at::Tensor& S, const at::Tensor& Y, const at::Tensor& rho, const at::Tensor& h_diag,
                           const at::Tensor& g, int64_t m_used, int64_t head) {
  const c10::cuda::CUDAGuard guard(g.device());
  at::Tensor d = at::empty_like(g);
  at::Tensor scratch = at::empty({2 * S.size(0) + 64}, g.options());
  check(pinn_lbfgs_direction(f32(S, "S"), f32(Y, "Y"), f32(rho, "rho"), f32(h_diag, "h_diag"), f32(g, "g"),
                             d.data_ptr<float>(), (int32_t)S.size(0), (int32_t)m_used, (int32_t)head, g.numel(),
                             scratch.data_ptr<float>(), stream_of(g)),
        "pinn_lbfgs_direction");
  return d;
}

at::Tensor vec_stats(const at::Tensor& a, const c10::optional<at::Tensor>& b) {
  const c10::cuda::CUDAGuard guard(a.device());
  at::Tensor out = at::zeros({6}, a.options());
  check(pinn_vec_stats(f32(a, "a"), f32_opt(b, "b"), a.numel(), out.data_ptr<float>(), stream_of(a)), "pinn_vec_stats");
  return out;
}

void axpy_(double alpha, const at::Tensor& x, at::Tensor y) {
  const c10::cuda::CUDAGuard guard(y.device());
  check(pinn_axpy((float)alpha, f32(x, "x"), const_cast<float*>(f32(y, "y")), y.numel(), stream_of(y)), "pinn_axpy");
}

void adam_step_(at::Tensor params, const at::Tensor& grad, at::Tensor exp_avg, at::Tensor exp_avg_sq, double lr, double beta1,
                double beta2, double eps, double weight_decay, int64_t step) {
  const c10::cuda::CUDAGuard guard(params.device());
  check(pinn_adam_step(const_cast<float*>(f32(params, "params")), f32(grad, "grad"), const_cast<float*>(f32(exp_avg, "exp_avg")),
                       const_cast<float*>(f32(exp_avg_sq, "exp_avg_sq")), params.numel(), (float)lr, (float)beta1, (float)beta2,
                       (float)eps, (float)weight_decay, step, stream_of(params)),
        "pinn_adam_step");
}

}  // namespace

TORCH_LIBRARY(pinn_b200, m) {
  m.def("workspace_bytes(Tensor desc, int n_points, bool want_grad) -> int", &workspace_bytes);
  m.def("jet_loss(Tensor desc, Tensor params, Tensor inputs, Tensor? targets, Tensor? mask_count, int n_res_global, "
        "int n_fid_global, Tensor(a!) workspace, Tensor(b!)? grad, Tensor(c!)? out, int flags) -> Tensor", &jet_loss);
  m.def("mask_count(Tensor desc, Tensor inputs) -> Tensor", &mask_count);
  m.def("loss_finalize(Tensor desc, Tensor sums, Tensor? sums_b, int n_fid_global, int n_res_global, Tensor? mask_count) -> Tensor",
        &loss_finalize);
  m.def("lbfgs_direction(Tensor S, Tensor Y, Tensor rho, Tensor h_diag, Tensor g, int m_used, int head) -> Tensor", &lbfgs_direction);
  m.def("vec_stats(Tensor a, Tensor? b) -> Tensor", &vec_stats);
  m.def("axpy_(float alpha, Tensor x, Tensor(a!) y) -> ()", &axpy_);
  m.def("adam_step_(Tensor(a!) params, Tensor grad, Tensor(b!) exp_avg, Tensor(c!) exp_avg_sq, float lr, float beta1, float beta2, "
        "float eps, float weight_decay, int step) -> ()", &adam_step_);
}
