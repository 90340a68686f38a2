/*
 * pinn_b200.h -- C ABI of the B200-native PINN training hot path.
 *
 * The reference (rezasalatin/PINN_depthEstimation) is pure Python and has no FFI; the boundary it
 * offers for this path is the set of Python call signatures listed in SURVEY.md section 8(b).  Each entry
 * point below names the reference interface (file:line under the reference tree) whose work it
 * replaces.  The Python facades in pinn_depthestimation_b200/ (dnn.py, physics.py, lbfgs.py,
 * l_bfgs_b_optimizer.py) bind these symbols with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless its name ends in _host
 *   - the caller owns every buffer; nothing is allocated, freed or synchronised inside
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*) of the current device
 *   - return value: 0 = ok, otherwise a PINN_E_* code; pinn_last_error() gives the text
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns
 *     PINN_E_CUDA
 *   - parameters are ONE flat fp32 vector in nn.Module.parameters() order of the reference DNN
 *     (dnn.py:31-34): layer_0.weight [out,in] row-major, layer_0.bias [out], layer_1.weight, ...
 *   - inputs are [n_points, widths[0]] row-major fp32, targets [n_points, n_targets] row-major
 */
#ifndef PINN_B200_H
#define PINN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_MAX_LINEAR 128   /* Linear layers per net (config_CMB_h.json has 101)            */
#define PINN_MAX_WIDTH 256    /* widest layer the round-1 kernels accept                       */
#define PINN_MAX_IN 8         /* input features                                                */
#define PINN_MAX_OUT 8        /* output features                                               */
#define PINN_MAX_DIRS 3       /* differentiated input directions (t,x,y)                       */
#define PINN_NSUMS 16         /* length of the raw-sum vector, see below                       */

/* activation: dnn.py:18-21 ('xavier' -> Tanh, 'kaiming' -> LeakyReLU(0.01)) */
enum { PINN_ACT_TANH = 0, PINN_ACT_LEAKY_RELU = 1 };

/* residual kind = which physics.py function the fused epilogue reproduces */
enum {
  PINN_RES_NONE = 0,       /* data misfit only (train.py:131-141 fidelity forward)                  */
  PINN_RES_CONT_ONLY = 1,  /* physics.py:18-33  continuity_only(x,y,h,U,V)      dirs (x,y)          */
  PINN_RES_CONT_FTEMP = 2, /* physics.py:37-47  continuity_ftemp(x,y,h,U,V)     dirs (x,y)          */
  PINN_RES_NSWE = 3,       /* physics.py:50-88  Navier_Stokes(t,x,y,h,z,u,v)    dirs (t,x,y)        */
  PINN_RES_WAVE_AVG = 4,   /* physics.py:91-120 physics_equation(x,y,h,U,V,eta_mean,Hrms,k) (x,y)   */
  PINN_RES_EXTERNAL = 5,   /* caller supplies d loss/d(out), d loss/d(out_j) (autograd facade)      */
  /* the historical physics_functions module (__pycache__/physics_functions.cpython-38.pyc, decompiled):     */
  PINN_RES_BOUSSINESQ = 6, /* :55-130 Boussinesq(output,t,x,y,device): fully nonlinear, input derivatives up  */
                           /*         to THIRD order -- order-3 Taylor jets (csrc/jet3.cu); dirs (t,x,y)      */
  PINN_RES_BOUSS_SIMPLE = 7 /* :18-52 Boussinesq_simple(output,t,x,y,device): first order; dirs (t,x,y)       */
};

/* arithmetic of the per-layer contractions */
enum {
  PINN_PREC_FP32 = 0,      /* FP32 FMA everywhere (parity mode)                                     */
  PINN_PREC_TF32 = 1,      /* tcgen05 kind::tf32, FP32 accumulate, for layers of width 256          */
  PINN_PREC_TF32X3 = 2     /* error-compensated 3xTF32 on the same path                             */
};

/* flags */
enum {
  PINN_FLAG_ACCUMULATE = 1, /* do not zero grad / sums first (second pass of train.py:128-157)       */
  PINN_FLAG_SKIP_PACK = 2   /* packed weights in the workspace are still valid for `params`          */
};

/* error codes */
enum {
  PINN_OK = 0, PINN_E_ARG = 1, PINN_E_UNSUPPORTED = 2, PINN_E_WORKSPACE = 3, PINN_E_CUDA = 4
};

/*
 * Description of one loss evaluation pass.  Plain data; mirrors what the reference scatters over
 * config*.json (layers.*, loss.*, data.*: train_newmethod.py:52-62,76-89; train.py:52-95) and
 * the argument order of the physics.py functions.
 *
 * dir_cols[j]   input column differentiated for direction j; direction order is the physics
 *               function's argument order: (x,y) or (t,x,y)
 * field_cols[f] output column of field f; field order is the physics function's argument order:
 *               CONT_*: (h,U,V)   NSWE, BOUSS*: (h,z,u,v)   WAVE_AVG: (h,U,V,eta_mean,Hrms,k)
 * mask_col      input column that continuity_only compares with cond_threshold (physics.py:27)
 * target_cols[i] output column supervised by targets[:,i], weight target_w[i]
 *               (train.py:136-141; weight 1 in train_newmethod.py:129-133)
 */
typedef struct pinn_desc {
  int32_t n_linear;
  int32_t widths[PINN_MAX_LINEAR + 1];
  int32_t activation;
  int32_t residual_kind;
  int32_t n_dirs;
  int32_t dir_cols[PINN_MAX_DIRS];
  int32_t field_cols[PINN_MAX_OUT];
  int32_t mask_col;
  float cond_threshold;  /* 25.5 in physics.py:27 */
  float cond_value;      /* 0.75 in physics.py:28 */
  int32_t n_targets;
  int32_t target_cols[PINN_MAX_OUT];
  float target_w[PINN_MAX_OUT];
  float w_fid;           /* loss.weight_fid_loss */
  float w_res;           /* loss.weight_res_loss */
  int32_t precision;
} pinn_desc_t;

/*
 * Raw sums produced by a pass (double[PINN_NSUMS], device).  They are SUMS over the points this
 * call saw, so shards on different GPUs add; pinn_loss_finalize turns them into the
 * reference's means.
 *   [0] sum fc^2   [1] sum fm_x^2   [2] sum fm_y^2   [3] sum (h-cond_value)^2 over masked points
 *   [4] number of masked points seen   [5+i] sum (pred-true)^2 of target i   [13] points seen
 */
enum { PINN_SUM_FC = 0, PINN_SUM_FX = 1, PINN_SUM_FY = 2, PINN_SUM_COND = 3, PINN_SUM_MASKCNT = 4,
       PINN_SUM_TARGET0 = 5, PINN_SUM_NPOINTS = 13 };

typedef struct pinn_eval_args {
  const float* params;        /* [P] flat parameters                                              */
  const float* inputs;        /* [n_points, widths[0]]                                            */
  const float* targets;       /* [n_points, n_targets] or NULL                                    */
  int64_t n_points;           /* points in THIS call (this GPU's shard)                           */
  int64_t n_res_global;       /* divisor of the residual means  (global point count)              */
  int64_t n_fid_global;       /* divisor of the data-misfit means                                 */
  const float* mask_count;    /* device scalar: global number of masked points (CONT_ONLY), or NULL */
  const float* seed_out;      /* EXTERNAL: d loss/d out   [n_points, o] or NULL                   */
  const float* seed_dout[PINN_MAX_DIRS]; /* EXTERNAL: d loss/d(d out/d dir j) [n_points,o] or NULL */
  float* grad;                /* [P] flat gradient (fwdbwd only)                                  */
  double* sums;               /* [PINN_NSUMS]                                                     */
  float* out;                 /* optional [n_points, o] network output                            */
  float* dout[PINN_MAX_DIRS]; /* optional [n_points, o] d out / d dir j                           */
  void* workspace;            /* pinn_workspace_bytes() bytes, 256-byte aligned                   */
  size_t workspace_bytes;
  int32_t flags;
} pinn_eval_args_t;

/* library / device probes */
const char* pinn_version(void);
const char* pinn_last_error(void);
int pinn_param_count(const pinn_desc_t* desc, int64_t* n_params);

/* Bytes of workspace pinn_jet_loss_* needs for this desc on the current device (packed weights +
 * per-CTA activation slabs); independent of n_points beyond a grid-size cap. */
int pinn_workspace_bytes(const pinn_desc_t* desc, int64_t n_points, size_t* bytes);

/* Same, for a workspace that will only ever be passed to pinn_jet_loss_fwd (want_grad == 0): no activation slabs.
 * (DNN.forward in dnn.py:54-55 and physics.compute_gradient in physics.py:6-15 are forward-only passes.) */
int pinn_workspace_bytes_ex(const pinn_desc_t* desc, int64_t n_points, int32_t want_grad, size_t* bytes);

/*
 * Loss evaluation without gradient: jet forward + fused residual / misfit sums.
 * Replaces DNN.forward (dnn.py:54-55) + compute_gradient (physics.py:6-15) + the physics.py
 * residual + the MSE terms of pinn.loss_func (train_newmethod.py:123-159, train.py:131-157).
 */
int pinn_jet_loss_fwd(const pinn_desc_t* desc, const pinn_eval_args_t* args, void* stream);

/*
 * Loss evaluation WITH flat weight gradient: the above plus the reverse sweep that replaces
 * loss.backward() (train_newmethod.py:200,207; train.py:191,198).  One fused kernel per call:
 * forward jets, epilogue, reverse, per tile of collocation points; activations never leave the
 * CTA's L2-resident slab.  grad receives d(total loss)/d params for the points of this call,
 * already divided by the global counts, so shards add.
 */
int pinn_jet_loss_fwdbwd(const pinn_desc_t* desc, const pinn_eval_args_t* args, void* stream);

/* Number of points with inputs[:,mask_col] < cond_threshold (physics.py:27) -> device float. */
int pinn_mask_count(const pinn_desc_t* desc, const float* inputs, int64_t n_points,
                    float* count_out, void* stream);

/*
 * sums (after any cross-GPU reduction) -> loss_parts[4] = {fidelity, residual, total, 0} exactly as
 * pinn.loss_func logs them (train_newmethod.py:133,156,159; train.py:141,154,157).
 * sums_b may be NULL; when given, its entries are added (two-pass form of train.py).
 */
int pinn_loss_finalize(const pinn_desc_t* desc, const double* sums, const double* sums_b,
                       int64_t n_fid_global, int64_t n_res_global, const float* mask_count,
                       float* loss_parts, void* stream);

/* ---- optimiser kernels on the flat vectors (replace torch.optim.LBFGS / Adam internals;
 *      call sites train_newmethod.py:95-117,197-209; torch/optim/lbfgs.py:386-457) ---- */

/*
 * L-BFGS two-loop recursion (torch/optim/lbfgs.py:423-447): d = -H g from the last m_used
 * (s,y) pairs.  hist_s / hist_y are [history_size, P] ring buffers; `head` is the slot of the OLDEST pair;
 * rho[i] = 1/(y_i . s_i) per slot; h_diag is a device scalar.  One thread-block-cluster kernel, no host sync.
 * scratch: >= (2*history_size + 64) floats.
 */
int pinn_lbfgs_direction(const float* hist_s, const float* hist_y, const float* rho,
                         const float* h_diag, const float* g, float* d, int32_t history_size,
                         int32_t m_used, int32_t head, int64_t n_params, float* scratch,
                         void* stream);

/*
 * Device-resident torch.optim.LBFGS.step (torch/optim/lbfgs.py:333-537 with _strong_wolfe :40-209 and
 * _cubic_interpolate :12-37; reference call sites train_newmethod.py:108-117,204-209).  The host only launches the
 * closure's evaluation and then pinn_lbfgs_advance, which performs everything up to the next evaluation in ONE
 * cluster kernel (line-search transition, gradient clones, termination tests, curvature-pair update, two-loop
 * recursion, first step length, next trial point written into flat_params) and copies a small status block to
 * status_host.  One stream synchronisation per evaluation, no other host round trips.
 */
typedef struct pinn_lbfgs_cfg {
  double lr, tolerance_grad, tolerance_change;
  int32_t max_iter, max_eval, history_size;
} pinn_lbfgs_cfg_t;

typedef struct pinn_lbfgs_status {
  int32_t code;             /* 1 = evaluate the closure at flat_params again, 2 = step() is over */
  int32_t n_iter;           /* iterations of this step() call                                          */
  int32_t current_evals;    /* evaluations of this step() call                                         */
  int32_t n_iter_total;     /* state['n_iter']                                                         */
  int32_t func_evals_total; /* state['func_evals']                                                     */
  int32_t phase, history_used, pad;
  double t, loss, first_loss, gtd, d_norm;
} pinn_lbfgs_status_t;

int pinn_lbfgs_workspace_bytes(int64_t n_params, int32_t history_size /* <= 512 */, size_t* bytes);
/* Start of a step() call.  reset != 0 also clears the history and the counters (a new optimiser). */
int pinn_lbfgs_begin(void* workspace, int64_t n_params, const pinn_lbfgs_cfg_t* cfg, int32_t reset, void* stream);
/* grad / *loss: gradient and loss (device) of the evaluation at the current flat_params.  status_host: HOST memory
 * (pinned), filled by an async copy on `stream`: synchronise the stream before reading it. */
int pinn_lbfgs_advance(void* workspace, int64_t n_params, int32_t history_size, float* flat_params, float* grad,
                       const float* loss, void* status_host, void* stream);

/* Measurement hook (bench.py): the two whole-chip passes of one direction computation over a FULL history of
 * history_size pairs in `workspace` (a pinn_lbfgs_workspace_bytes buffer the caller filled with anything finite);
 * *bytes_out (host) = the algorithmic bytes they move.  Time it with events on `stream`. */
int pinn_lbfgs_direction_probe(void* workspace, int64_t n_params, int32_t history_size, const float* grad,
                               double* bytes_out, void* stream);

/* out6 = [a.b, sum|a|, max|a|, max|b|, a.a, b.b]  (b may be NULL; one cluster launch, deterministic) */
int pinn_vec_stats(const float* a, const float* b, int64_t n, float* out6, void* stream);

/* y = y + alpha*x, alpha a host scalar (params += t*d, torch/optim/lbfgs.py:312-320) */
int pinn_axpy(float alpha_host, const float* x, float* y, int64_t n, void* stream);

/* Adam step with torch.optim.Adam defaults semantics (train_newmethod.py:95-98,198-202);
 * step_count is the 1-based step index; lr already includes the StepLR factor. */
int pinn_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int64_t n, float lr_host, float beta1, float beta2, float eps,
                   float weight_decay, int64_t step_count, void* stream);

/* ---- one-off data path on the device (train_newmethod.py:226-255, train.py:203-276, operations.py:4-30) ---- */

/* out2 = [nanmin(x), nanmax(x)] (operations.py:26-29: ranges of every input variable other than x / y). */
int pinn_nan_minmax(const float* x, int64_t n, float* out2, void* stream);

/*
 * Normalise, hstack and NaN-filter in one pass: cols_host is a HOST array of n_in + n_true DEVICE column pointers
 * (each [n]); input column c is mapped to 2 (x - lo[c]) / (hi[c] - lo[c]) - 1 (all zeros when hi == lo,
 * operations.py:4-7) and written to inputs_out [n_kept, n_in]; the true columns go to trues_out [n_kept, n_true]
 * unchanged.  Rows are dropped, order preserved, when they hold a NaN in a true column (nan_policy & 1,
 * train_newmethod.py:252-255) and / or in an input column (nan_policy & 2, train.py:274-276).  n_kept: device int64.
 * scratch: >= ceil(n / 256) int32.  lo_host / hi_host: host arrays [n_in].
 */
int pinn_assemble_points(const float* const* cols_host, int32_t n_in, int32_t n_true, const float* lo_host,
                         const float* hi_host, int32_t nan_policy, int64_t n, float* inputs_out, float* trues_out,
                         int64_t* n_kept, int32_t* scratch, void* stream);

/* Measurement helper (bench.py): launches an FP32-FMA-bound kernel of `ctas` x 256 threads and
 * reports the FLOPs it executes in *flops_out (host); time it with events on `stream`. */
int pinn_fma_probe(float* out, int32_t iters, int32_t ctas, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PINN_B200_H */
