// C-ABI entry points declared in include/pinn_b200.h: argument checks, error text, dispatch.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace pinn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  (void)cudaGetLastError();
  return PINN_E_CUDA;
}

// jet_fp32.cu
int validate_desc(const pinn_desc_t* D);
int workspace_bytes(const pinn_desc_t* D, long long n_points, bool bwd, size_t* bytes);
int run_pass(const pinn_desc_t* D, const pinn_eval_args_t* a, bool bwd, cudaStream_t st);
int run_mask_count(const pinn_desc_t* D, const float* inputs, long long n, float* out, cudaStream_t st);
int run_finalize(const pinn_desc_t* D, const double* s, const double* sb, long long n_fid, long long n_res,
                 const float* mask_count, float* parts, cudaStream_t st);
// optim.cu
int run_lbfgs_direction(const float* S, const float* Y, const float* rho, const float* h_diag,
                        const float* g, float* d, int hist, int m, int head, long long P,
                        float* scratch, cudaStream_t st);
int run_vec_stats(const float* a, const float* b, long long n, float* out, cudaStream_t st);
int run_axpy(float alpha, const float* x, float* y, long long n, cudaStream_t st);
int run_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
             float eps, float wd, long long step, cudaStream_t st);

}  // namespace pinn

using namespace pinn;

extern "C" {

const char* pinn_version(void) { return "pinn_b200 0.2 (sm_100a; jet_tc r2e, jet_fp32 r2a)"; }
const char* pinn_last_error(void) { return g_err; }

int pinn_param_count(const pinn_desc_t* desc, int64_t* n_params) {
  int rc = validate_desc(desc);
  if (rc) return rc;
  if (!n_params) return set_error("n_params is NULL"), PINN_E_ARG;
  int64_t p = 0;
  for (int i = 0; i < desc->n_linear; ++i)
    p += (int64_t)desc->widths[i] * desc->widths[i + 1] + desc->widths[i + 1];
  *n_params = p;
  return PINN_OK;
}

int pinn_workspace_bytes(const pinn_desc_t* desc, int64_t n_points, size_t* bytes) {
  if (!bytes) return set_error("bytes is NULL"), PINN_E_ARG;
  return workspace_bytes(desc, n_points, true, bytes);
}

int pinn_workspace_bytes_ex(const pinn_desc_t* desc, int64_t n_points, int32_t want_grad, size_t* bytes) {
  if (!bytes) return set_error("bytes is NULL"), PINN_E_ARG;
  return workspace_bytes(desc, n_points, want_grad != 0, bytes);
}

int pinn_jet_loss_fwd(const pinn_desc_t* desc, const pinn_eval_args_t* args, void* stream) {
  return run_pass(desc, args, false, (cudaStream_t)stream);
}

int pinn_jet_loss_fwdbwd(const pinn_desc_t* desc, const pinn_eval_args_t* args, void* stream) {
  return run_pass(desc, args, true, (cudaStream_t)stream);
}

int pinn_mask_count(const pinn_desc_t* desc, const float* inputs, int64_t n_points, float* count_out,
                    void* stream) {
  int rc = validate_desc(desc);
  if (rc) return rc;
  if (!count_out || (!inputs && n_points > 0)) return set_error("NULL pointer"), PINN_E_ARG;
  if (desc->mask_col < 0 || desc->mask_col >= desc->widths[0])
    return set_error("mask_col %d is not an input column", desc->mask_col), PINN_E_ARG;
  return run_mask_count(desc, inputs, n_points, count_out, (cudaStream_t)stream);
}

int pinn_loss_finalize(const pinn_desc_t* desc, const double* sums, const double* sums_b,
                       int64_t n_fid_global, int64_t n_res_global, const float* mask_count,
                       float* loss_parts, void* stream) {
  int rc = validate_desc(desc);
  if (rc) return rc;
  if (!sums || !loss_parts) return set_error("NULL pointer"), PINN_E_ARG;
  return run_finalize(desc, sums, sums_b, n_fid_global, n_res_global, mask_count, loss_parts,
                      (cudaStream_t)stream);
}

int pinn_lbfgs_direction(const float* hist_s, const float* hist_y, const float* rho, const float* h_diag,
                         const float* g, float* d, int32_t history_size, int32_t m_used, int32_t head,
                         int64_t n_params, float* scratch, void* stream) {
  if (!g || !d || !h_diag || (m_used > 0 && (!hist_s || !hist_y || !rho)))
    return set_error("NULL pointer"), PINN_E_ARG;
  return run_lbfgs_direction(hist_s, hist_y, rho, h_diag, g, d, history_size, m_used, head, n_params,
                             scratch, (cudaStream_t)stream);
}

int pinn_vec_stats(const float* a, const float* b, int64_t n, float* out6, void* stream) {
  if (!a || !out6) return set_error("NULL pointer"), PINN_E_ARG;
  return run_vec_stats(a, b, n, out6, (cudaStream_t)stream);
}

int pinn_axpy(float alpha_host, const float* x, float* y, int64_t n, void* stream) {
  if (n > 0 && (!x || !y)) return set_error("NULL pointer"), PINN_E_ARG;
  return run_axpy(alpha_host, x, y, n, (cudaStream_t)stream);
}

int pinn_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr_host, float beta1, float beta2, float eps, float weight_decay,
                   int64_t step_count, void* stream) {
  if (n > 0 && (!params || !grad || !exp_avg || !exp_avg_sq)) return set_error("NULL pointer"), PINN_E_ARG;
  return run_adam(params, grad, exp_avg, exp_avg_sq, n, lr_host, beta1, beta2, eps, weight_decay,
                  step_count, (cudaStream_t)stream);
}

}  // extern "C"
