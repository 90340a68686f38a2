// Device-resident L-BFGS (north_star (4): "two-loop recursion and line search run on device").
//
// Everything torch.optim.LBFGS.step does between two closure evaluations -- the strong-Wolfe bracket / zoom transition
// (torch/optim/lbfgs.py:40-209, cubic interpolation :12-37), the clones of the gradient it keeps for the bracket ends,
// the end-of-iteration termination tests (:511-526), the curvature-pair update (:404-421), the two-loop recursion
// (:432-447), the first step length (:454-457) and the next trial point x0 + t d -- runs as a fixed sequence of four
// launches per evaluation with all decisions taken on the device (`pinn_lbfgs_advance`).  The host launches the closure's
// kernels, then these, then reads ONE small status block (one sync per evaluation) that only says "evaluate again" or
// "finished".  Call sites in the reference: train_newmethod.py:108-117 (construction), :204-209 (the single step(closure)
// call with max_iter = 50000).
//
//   1. lbfgs_advance_kernel   8-CTA cluster: line-search transition, gradient clones, termination tests; either writes the
//                             next trial point of the running line search or raises `begin` (a new outer iteration starts)
//   2. vl_dots_kernel         whole chip, only if `begin`: writes the candidate pair (s, y) into its ring slot and takes, in
//                             ONE streaming pass over the history, the dot products of every stored vector with s, y and g;
//                             the last CTA to finish (fixed-order reduction of the per-CTA partials) accepts the pair
//                             (y.s > 1e-10), updates the Gram matrices S^T S, S^T Y, Y^T Y and runs the TWO-LOOP RECURSION IN
//                             COEFFICIENT SPACE: d = sum_j delta_j b_j over the basis {s_i, y_i, g}, every s_i.q / y_i.r of
//                             torch's loops being a (2m+1)-term sum over Gram entries in double precision -- the 2m dependent
//                             vector passes of the textbook recursion (2m cluster barriers, 0.6 ms at m = 100 however small
//                             the vectors are) collapse into one block-local scalar loop
//   3. vl_combine_kernel      whole chip, only if `begin`: d = sum_j delta_j b_j (the second and last pass over the history),
//                             prev_g = g, partial g.d / |g|_1 / max|d|; the last CTA forms the step length, tests
//                             g.d > -tolerance_change and sets up the line search
//   4. lbfgs_trial_kernel     whole chip, only if a line search starts: x0 = x, g_prev = g, x = x0 + t d
// Traffic per outer iteration: two passes over the 2 m P history (HBM / L2 bound, no dependent chain) instead of two
// passes broken into 2m barrier-separated steps on 8 SMs.
//
// All scalar decisions are taken in double precision from reductions with a fixed summation order, so every GPU of a
// sharded run (which all-reduce loss and gradient first) takes the same branch.  pinn_depthestimation_b200/lbfgs.py holds
// the same logic as host code (used for line_search_fn=None and as the CPU-tested restatement of torch's functions);
// tests/test_gpu_optim.py holds both to torch's iteration / evaluation counts.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pinn {

constexpr int kLbCtas = 8;
constexpr int kLbThreads = 1024;
constexpr int kLbMaxHistory = 512;

enum { LB_PH_START = 0, LB_PH_BRACKET = 1, LB_PH_ZOOM = 2, LB_PH_DONE = 3 };
enum { LB_STATUS_EVAL = 1, LB_STATUS_DONE = 2 };

struct LbState {
  // configuration (torch.optim.LBFGS defaults dict)
  double lr, tol_grad, tol_change;
  int max_iter, max_eval, hist;
  // persistent across step() calls (torch keeps these in self.state)
  int n_iter_total, func_evals_total;
  int head, used;
  double t, loss, prev_loss;
  // this step() call
  int ph, n_iter, current_evals;
  double first_loss;
  double gtd, d_norm;
  // line search (names as in torch's _strong_wolfe)
  double f0, gtd0, ls_t, t_prev, f_prev, gtd_prev;
  double br[2], br_f[2], br_gtd[2];
  int br_n, low, high, done, insuf, ls_iter, max_ls, ls_evals;
  int status;
  // hand-over between the kernels of one advance
  int begin;        // a new outer iteration starts: kernels 2-3 run
  int trial;        // kernel 4 writes the first trial point of a new line search
  int slot_new;     // ring slot of the candidate pair (or -1)
  double trial_t;
};

struct LbStatus {   // copied to the host after every advance
  int code, n_iter, current_evals, n_iter_total, func_evals_total, ph, used, pad;
  double t, loss, first_loss, gtd, d_norm;
};

struct LbVectors {  // all [P] unless noted; carved from the workspace
  float *d, *prev_g, *x0, *gp, *b0, *b1, *S, *Y;   // S, Y: [(hist+1), P] ring of curvature pairs
  double *SS, *SY, *YY;     // Gram matrices over ring slots, [cap][cap]: s_p.s_q, s_p.y_q, y_p.y_q
  double *delta;            // [2 cap + 2]: coefficients of d over {s_p}, {y_p}, g; last entry: H_diag
  double *dots_partial;     // [grid][2 cap][3]: per-CTA partial dots of every stored vector with s_new, y_new, g
  double *stats_partial;    // [grid][4]: per-CTA partial g.d, |g|_1, max|d|
  unsigned* counters;       // [2] arrival counters of the "last CTA finishes the job" reductions
  int grid;                 // CTAs of the whole-chip kernels
};

// ---- cluster-wide deterministic reductions: every thread of every CTA gets the same value ----
struct LbShared {
  float warp_buf[32];
  float slot[2][8];     // per-call partials of this CTA (up to 8 quantities), double-buffered
  float bcast[8];
  // decisions of thread 0 broadcast to the CTA
  int copy_src[4], copy_dst[4], n_copy;
  int do_finish, do_begin, do_trial, status;
  double trial_t;
};

__device__ __forceinline__ float lb_block_sum(float v, float* warp_buf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) warp_buf[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += warp_buf[i];  // fixed order
  return t;
}
__device__ __forceinline__ float lb_nan_max(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float lb_block_max(float v, float* warp_buf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = lb_nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) warp_buf[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t = lb_nan_max(t, warp_buf[i]);
  return t;
}
// out6 = [a.b, sum|a|, max|a|, max|b|, a.a, b.b] over [lo,hi) of this CTA, combined across the cluster -> sh.bcast
__device__ __forceinline__ void lb_stats(cg::cluster_group& cl, LbShared& sh, const float* a, const float* b, long long lo,
                                         long long hi, int& phase) {
  float ab = 0.f, l1 = 0.f, ma = 0.f, mb = 0.f, aa = 0.f, bb = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) {
    const float x = a[i], y = b ? b[i] : 0.f;
    ab = fmaf(x, y, ab);
    l1 += fabsf(x);
    ma = lb_nan_max(ma, fabsf(x));
    mb = lb_nan_max(mb, fabsf(y));
    aa = fmaf(x, x, aa);
    bb = fmaf(y, y, bb);
  }
  float mine[6];
  mine[0] = lb_block_sum(ab, sh.warp_buf);
  mine[1] = lb_block_sum(l1, sh.warp_buf);
  mine[2] = lb_block_max(ma, sh.warp_buf);
  mine[3] = lb_block_max(mb, sh.warp_buf);
  mine[4] = lb_block_sum(aa, sh.warp_buf);
  mine[5] = lb_block_sum(bb, sh.warp_buf);
  __syncthreads();
  if (threadIdx.x == 0)
    for (int q = 0; q < 6; ++q) sh.slot[phase][q] = mine[q];
  cl.sync();
  if (threadIdx.x < 6) {
    const int q = threadIdx.x;
    float t = 0.f;
    for (unsigned r = 0; r < cl.num_blocks(); ++r) {
      const float v = *cl.map_shared_rank(&sh.slot[phase][q], r);
      t = (q == 2 || q == 3) ? lb_nan_max(t, v) : t + v;
    }
    sh.bcast[q] = t;
  }
  __syncthreads();
  phase ^= 1;
}

// torch/optim/lbfgs.py:12-37 (IEEE division: a collapsed bracket yields inf / nan and falls through like torch's tensors)
__device__ double lb_cubic(double x1, double f1, double g1, double x2, double f2, double g2, bool has_bounds, double bmin,
                           double bmax) {
  double xmin_bound, xmax_bound;
  if (has_bounds) xmin_bound = bmin, xmax_bound = bmax;
  else if (x1 <= x2) xmin_bound = x1, xmax_bound = x2;
  else xmin_bound = x2, xmax_bound = x1;
  const double d1 = g1 + g2 - 3.0 * ((f1 - f2) / (x1 - x2));
  const double d2_square = d1 * d1 - g1 * g2;
  if (d2_square >= 0) {
    const double d2 = sqrt(d2_square);
    double min_pos;
    if (x1 <= x2) min_pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.0 * d2));
    else min_pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.0 * d2));
    // Python's min(max(min_pos, lo), hi): comparisons with NaN are false
    double r = (xmin_bound > min_pos) ? xmin_bound : min_pos;   // max(min_pos, xmin_bound) as Python evaluates it
    r = (xmax_bound < r) ? xmax_bound : r;                      // min(r, xmax_bound)
    return r;
  }
  return (xmin_bound + xmax_bound) / 2.0;
}

enum { LB_BUF_G = 0, LB_BUF_GP = 1, LB_BUF_B0 = 2, LB_BUF_B1 = 3, LB_BUF_PREVG = 4 };

__device__ __forceinline__ void lb_write_status(const LbState& st, LbStatus* out) {
  LbStatus o;
  o.code = st.status, o.n_iter = st.n_iter, o.current_evals = st.current_evals, o.n_iter_total = st.n_iter_total;
  o.func_evals_total = st.func_evals_total, o.ph = st.ph, o.used = st.used, o.pad = 0;
  o.t = st.t, o.loss = st.loss, o.first_loss = st.first_loss, o.gtd = st.gtd, o.d_norm = st.d_norm;
  *out = o;
}

// One transition of the optimiser between two closure evaluations.  `g` holds the gradient and *loss_dev the loss at
// `flat`; on return either `flat` is the next trial point (status EVAL) or the step() call is over (status DONE).
__global__ void __cluster_dims__(kLbCtas, 1, 1) __launch_bounds__(kLbThreads)
    lbfgs_advance_kernel(LbState* __restrict__ st_g, LbVectors V, float* __restrict__ flat, float* __restrict__ g,
                         const float* __restrict__ loss_dev, LbStatus* __restrict__ status_out, long long P) {
  cg::cluster_group cl = cg::this_cluster();
  __shared__ LbShared sh;
  __shared__ LbState st;      // thread 0's working copy
  const long long per = (P + kLbCtas - 1) / kLbCtas;
  const long long lo = per * cl.block_rank();
  const long long hi = lo + per < P ? lo + per : P;
  int phase = 0;
  if (threadIdx.x == 0) {
    st = *st_g;
    st.begin = 0, st.trial = 0, st.slot_new = -1;
  }
  __syncthreads();
  const int cap = st.hist + 1;
  float* bufs[5] = {g, V.gp, V.b0, V.b1, V.prev_g};

  // every evaluation: g.d, |g|_1, max|g|, max|d| (d is stale garbage in PH_START: only max|g| is used then)
  lb_stats(cl, sh, g, V.d, lo, hi, phase);
  const double gtd_new = (double)sh.bcast[0];
  const double gmax = (double)sh.bcast[2];
  const double f_new = (double)*loss_dev;

  if (threadIdx.x == 0) {
    sh.n_copy = 0, sh.do_finish = 0, sh.do_begin = 0, sh.do_trial = 0, sh.status = 0;
    auto copy = [&](int src, int dst) { sh.copy_src[sh.n_copy] = src, sh.copy_dst[sh.n_copy] = dst, ++sh.n_copy; };
    const double c1 = 1e-4, c2 = 0.9;
    if (st.ph == LB_PH_START) {
      st.first_loss = f_new;
      st.loss = f_new;
      st.current_evals = 1;
      st.func_evals_total += 1;
      st.n_iter = 0;
      if (gmax <= st.tol_grad) sh.status = LB_STATUS_DONE;     // lbfgs.py:386-388
      else sh.do_begin = 1;
    } else {
      bool to_zoom = false;
      st.ls_evals += 1;
      if (st.ph == LB_PH_BRACKET) {
        // body of the bracket loop for the evaluation that just came back (lbfgs.py:58-108)
        const double t = st.ls_t;
        bool bracketed = false;
        if (!(st.ls_iter < st.max_ls)) {
          bracketed = true;      // `while ls_iter < max_ls` is over without a bracket: the fallback below takes [0, t]
        } else if (f_new > (st.f0 + c1 * t * st.gtd0) || (st.ls_iter > 1 && f_new >= st.f_prev)) {
          st.br[0] = st.t_prev, st.br[1] = t, st.br_f[0] = st.f_prev, st.br_f[1] = f_new;
          st.br_gtd[0] = st.gtd_prev, st.br_gtd[1] = gtd_new, st.br_n = 2;
          copy(LB_BUF_GP, LB_BUF_B0), copy(LB_BUF_G, LB_BUF_B1);
          bracketed = true;
        } else if (fabs(gtd_new) <= -c2 * st.gtd0) {
          st.br[0] = t, st.br_f[0] = f_new, st.br_n = 1, st.done = 1;
          copy(LB_BUF_G, LB_BUF_B0);
          bracketed = true;
        } else if (gtd_new >= 0) {
          st.br[0] = st.t_prev, st.br[1] = t, st.br_f[0] = st.f_prev, st.br_f[1] = f_new;
          st.br_gtd[0] = st.gtd_prev, st.br_gtd[1] = gtd_new, st.br_n = 2;
          copy(LB_BUF_GP, LB_BUF_B0), copy(LB_BUF_G, LB_BUF_B1);
          bracketed = true;
        }
        if (!bracketed) {
          // extrapolate (lbfgs.py:96-108), then the loop condition `ls_iter < max_ls` decides whether it is evaluated
          const double min_step = t + 0.01 * (t - st.t_prev), max_step = t * 10.0;
          const double tn = lb_cubic(st.t_prev, st.f_prev, st.gtd_prev, t, f_new, gtd_new, true, min_step, max_step);
          st.t_prev = t, st.f_prev = f_new, st.gtd_prev = gtd_new;
          copy(LB_BUF_G, LB_BUF_GP);
          // torch evaluates at the new t inside the same loop iteration and then does ls_iter += 1; the evaluation is handed
          // to the host here, and the `while ls_iter < max_ls` test is made when its result comes back
          st.ls_t = tn;
          st.ls_iter += 1;
          sh.do_trial = 1, sh.trial_t = tn;
        } else {
          to_zoom = true;
        }
      } else {   // LB_PH_ZOOM: the evaluation requested by the zoom loop came back (lbfgs.py:170-203)
        const double t = st.ls_t;
        st.ls_iter += 1;
        if (f_new > (st.f0 + c1 * t * st.gtd0) || f_new >= st.br_f[st.low]) {
          st.br[st.high] = t, st.br_f[st.high] = f_new, st.br_gtd[st.high] = gtd_new;
          copy(LB_BUF_G, st.high ? LB_BUF_B1 : LB_BUF_B0);
          if (st.br_f[0] <= st.br_f[1]) st.low = 0, st.high = 1;
          else st.low = 1, st.high = 0;
        } else {
          if (fabs(gtd_new) <= -c2 * st.gtd0) {
            st.done = 1;
          } else if (gtd_new * (st.br[st.high] - st.br[st.low]) >= 0) {
            st.br[st.high] = st.br[st.low], st.br_f[st.high] = st.br_f[st.low], st.br_gtd[st.high] = st.br_gtd[st.low];
            copy(st.low ? LB_BUF_B1 : LB_BUF_B0, st.high ? LB_BUF_B1 : LB_BUF_B0);
          }
          st.br[st.low] = t, st.br_f[st.low] = f_new, st.br_gtd[st.low] = gtd_new;
          copy(LB_BUF_G, st.low ? LB_BUF_B1 : LB_BUF_B0);
        }
        to_zoom = true;
      }
      if (to_zoom) {
        // (first entry from the bracket phase) lbfgs.py:110-122
        if (st.ph == LB_PH_BRACKET) {
          if (st.ls_iter == st.max_ls || st.br_n == 0) {
            st.br[0] = 0.0, st.br[1] = st.ls_t, st.br_f[0] = st.f0, st.br_f[1] = f_new;
            st.br_gtd[0] = st.gtd0, st.br_gtd[1] = gtd_new, st.br_n = 2;
            sh.n_copy = 0;
            copy(LB_BUF_PREVG, LB_BUF_B0), copy(LB_BUF_G, LB_BUF_B1);
          }
          st.insuf = 0;
          if (st.br_n == 2 && !(st.br_f[0] <= st.br_f[1])) st.low = 1, st.high = 0;
          else st.low = 0, st.high = 1;
          st.ph = LB_PH_ZOOM;
        }
        // top of the zoom loop (lbfgs.py:130-168)
        bool stop = st.done || !(st.ls_iter < st.max_ls);
        if (!stop && fabs(st.br[1] - st.br[0]) * st.d_norm < 1e-9) stop = true;   // torch's _strong_wolfe default tolerance_change
        if (!stop) {
          double t = lb_cubic(st.br[0], st.br_f[0], st.br_gtd[0], st.br[1], st.br_f[1], st.br_gtd[1], false, 0, 0);
          const double bmax = st.br[0] > st.br[1] ? st.br[0] : st.br[1], bmin = st.br[0] < st.br[1] ? st.br[0] : st.br[1];
          const double eps = 0.1 * (bmax - bmin);
          const double m1 = bmax - t, m2 = t - bmin;
          // Python's min(a, b) with NaN: min(a, b) returns b if b < a else a
          const double mn = (m2 < m1) ? m2 : m1;
          if (mn < eps) {
            if (st.insuf || t >= bmax || t <= bmin) {
              if (fabs(t - bmax) < fabs(t - bmin)) t = bmax - eps;
              else t = bmin + eps;
              st.insuf = 0;
            } else {
              st.insuf = 1;
            }
          } else {
            st.insuf = 0;
          }
          st.ls_t = t;
          sh.do_trial = 1, sh.trial_t = t;
        } else {
          sh.do_finish = 1;
        }
      }
    }
  }
  __syncthreads();

  // ---- gradient clones decided above (torch's .clone() calls), in order ----
  for (int c = 0; c < sh.n_copy; ++c) {
    const float* src = bufs[sh.copy_src[c]];
    float* dst = bufs[sh.copy_dst[c]];
    for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) dst[i] = src[i];
    __syncthreads();
  }

  if (sh.do_finish) {
    // line search over: t, loss, flat_grad = bracket[low_pos] (lbfgs.py:205-209, 490-494)
    if (threadIdx.x == 0 && st.br_n == 1) st.low = 0;
    __syncthreads();
    const int low = st.low;
    const double t_fin = st.br[low];
    const float* gbest = low ? V.b1 : V.b0;
    for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) {
      g[i] = gbest[i];
      flat[i] = fmaf((float)t_fin, V.d[i], V.x0[i]);
    }
    cl.sync();
    lb_stats(cl, sh, g, V.d, lo, hi, phase);
    const double gmax_fin = (double)sh.bcast[2];
    if (threadIdx.x == 0) {
      st.t = t_fin;
      st.loss = st.br_f[low];
      st.current_evals += st.ls_evals;
      st.func_evals_total += st.ls_evals;
      // lbfgs.py:511-526
      bool brk = false;
      if (st.n_iter == st.max_iter) brk = true;
      else if (st.current_evals >= st.max_eval) brk = true;
      else if (gmax_fin <= st.tol_grad) brk = true;
      else if (st.d_norm * fabs(t_fin) <= st.tol_change) brk = true;
      else if (fabs(st.loss - st.prev_loss) < st.tol_change) brk = true;
      if (brk) sh.status = LB_STATUS_DONE;
      else sh.do_begin = 1;
    }
    __syncthreads();
  }

  if (sh.do_begin && threadIdx.x == 0) {
    // a new outer iteration starts (lbfgs.py:394-402): kernels 2-4 of this advance do the work
    st.n_iter += 1;
    st.n_iter_total += 1;
    st.begin = 1;
    st.slot_new = st.n_iter_total > 1 ? (st.head + st.used) % cap : -1;
  }
  __syncthreads();

  if (sh.do_trial) {   // next trial point of the running line search
    const float tt = (float)sh.trial_t;
    for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) flat[i] = fmaf(tt, V.d[i], V.x0[i]);
    if (threadIdx.x == 0) sh.status = LB_STATUS_EVAL;
  }
  __syncthreads();
  cl.sync();   // every CTA has taken its copy of the old state long ago; remote shared-memory reads are over
  if (cl.block_rank() == 0 && threadIdx.x == 0) {
    st.status = sh.status;
    if (sh.status == LB_STATUS_DONE) st.ph = LB_PH_DONE;
    *st_g = st;
    lb_write_status(st, status_out);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel 2: candidate pair + every dot product the iteration needs in one pass; last CTA: Gram update + coefficient two-loop
// ------------------------------------------------------------------------------------------------------------------
constexpr int kVlThreads = 256;

__device__ __forceinline__ double vl_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the CTA, result to every thread (fixed order)
__device__ __forceinline__ double vl_block_sum(double v, double* warp_buf) {
  v = vl_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) warp_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < kVlThreads / 32; ++i) t += warp_buf[i];
  return t;
}
__device__ __forceinline__ bool vl_slot_active(int p, int head, int used, int cap, int slot_new) {
  const int rel = (p - head + cap) % cap;
  return rel < used || p == slot_new;
}
// true in every thread of exactly one CTA: the last one to arrive (its view of the other CTAs' global writes is complete)
__device__ __forceinline__ bool vl_last_cta(unsigned* counter) {
  __shared__ unsigned ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(counter, 1u);
  __syncthreads();
  const bool last = ticket == gridDim.x - 1;
  if (last) {
    __threadfence();
    if (threadIdx.x == 0) *counter = 0;   // re-armed for the next launch
  }
  return last;
}

__global__ void __launch_bounds__(kVlThreads)
    vl_dots_kernel(LbState* __restrict__ st_g, LbVectors V, const float* __restrict__ g, long long P, int stage_bytes) {
  if (!st_g->begin) return;
  __shared__ double warp_buf[kVlThreads / 32];
  __shared__ double al[kLbMaxHistory + 1];
  const int cap = st_g->hist + 1;
  const int head = st_g->head, used = st_g->used, slot_new = st_g->slot_new;
  const long long per = ((P + gridDim.x - 1) / gridDim.x + 3) & ~3LL;
  const long long lo = per * blockIdx.x;
  const long long hi = lo + per < P ? lo + per : P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // candidate pair into the free ring slot: y = g - prev_g, s = t d (lbfgs.py:404-406)
  if (slot_new >= 0) {
    float* Ys = V.Y + (long long)slot_new * P;
    float* Ss = V.S + (long long)slot_new * P;
    const float tf = (float)st_g->t;
    for (long long i = lo + threadIdx.x; i < hi; i += kVlThreads) {
      Ys[i] = g[i] - V.prev_g[i];
      Ss[i] = V.d[i] * tf;
    }
  }
  __syncthreads();
  const float* s_new = slot_new >= 0 ? V.S + (long long)slot_new * P : nullptr;
  const float* y_new = slot_new >= 0 ? V.Y + (long long)slot_new * P : nullptr;
  double* part = V.dots_partial + (size_t)blockIdx.x * (2 * cap) * 3;
  constexpr int EL = 32;                        // elements of the slice per lane in the register-cached path
  if (per <= 32 * EL) {
    // the CTA's slices of s_new, y_new, g stay in registers (one coalesced load each); every stored vector then costs one
    // load per element -- the pass is a pure stream over the 2 m P history
    float sN[EL], yN[EL], gV[EL];
#pragma unroll
    for (int u = 0; u < EL; ++u) {
      const long long i = lo + lane + 32 * u;
      const bool in = i < hi;
      sN[u] = (in && s_new) ? s_new[i] : 0.f;
      yN[u] = (in && y_new) ? y_new[i] : 0.f;
      gV[u] = in ? g[i] : 0.f;
    }
    for (int v = warp; v < 2 * cap; v += kVlThreads / 32) {
      const int pslot = v < cap ? v : v - cap;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      if (vl_slot_active(pslot, head, used, cap, slot_new)) {
        const float* vec = (v < cap ? V.S : V.Y) + (long long)pslot * P;
        float x[EL];
#pragma unroll
        for (int u = 0; u < EL; ++u) {
          const long long i = lo + lane + 32 * u;
          x[u] = i < hi ? vec[i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < EL; ++u) {
          const double xd = (double)x[u];
          a0 = fma(xd, (double)sN[u], a0), a1 = fma(xd, (double)yN[u], a1), a2 = fma(xd, (double)gV[u], a2);
        }
        a0 = vl_warp_sum(a0), a1 = vl_warp_sum(a1), a2 = vl_warp_sum(a2);
      }
      if (lane == 0) part[v * 3] = a0, part[v * 3 + 1] = a1, part[v * 3 + 2] = a2;
    }
  } else {
    for (int v = warp; v < 2 * cap; v += kVlThreads / 32) {
      const int pslot = v < cap ? v : v - cap;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      if (vl_slot_active(pslot, head, used, cap, slot_new)) {
        const float* vec = (v < cap ? V.S : V.Y) + (long long)pslot * P;
        for (long long i = lo + lane; i < hi; i += 32) {
          const double x = (double)vec[i];
          if (s_new) a0 = fma(x, (double)s_new[i], a0), a1 = fma(x, (double)y_new[i], a1);
          a2 = fma(x, (double)g[i], a2);
        }
        a0 = vl_warp_sum(a0), a1 = vl_warp_sum(a1), a2 = vl_warp_sum(a2);
      }
      if (lane == 0) part[v * 3] = a0, part[v * 3 + 1] = a1, part[v * 3 + 2] = a2;
    }
  }
  if (!vl_last_cta(V.counters)) return;

  // ---------------- last CTA: reduce, accept the pair, update the Gram matrices, coefficient two-loop ----------------
  double* dS = V.delta;            // [cap] coefficient of s_p
  double* dY = V.delta + cap;      // [cap] coefficient of y_p
  __shared__ double dot_g_s[kLbMaxHistory + 1], dot_g_y[kLbMaxHistory + 1];
  __shared__ int sh_head, sh_used;
  // per stored vector: its dots with s_new, y_new, g (fixed summation order over the CTAs)
  for (int v = threadIdx.x; v < 2 * cap; v += kVlThreads) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 8
    for (unsigned c = 0; c < gridDim.x; ++c) {
      const double* pc = V.dots_partial + ((size_t)c * (2 * cap) + v) * 3;
      a0 += pc[0], a1 += pc[1], a2 += pc[2];
    }
    const int pslot = v < cap ? v : v - cap;
    if (v < cap) dot_g_s[pslot] = a2; else dot_g_y[pslot] = a2;
    if (slot_new >= 0) {
      // Gram rows / columns of the candidate slot (used only once the pair is accepted; harmless otherwise because a
      // rejected candidate's slot is rewritten before it can ever become active)
      if (v < cap) {
        V.SS[(size_t)slot_new * cap + pslot] = a0, V.SS[(size_t)pslot * cap + slot_new] = a0;   // s_new . s_p
        if (pslot != slot_new) V.SY[(size_t)pslot * cap + slot_new] = a1;                        // s_p . y_new (the diagonal
                                                                                                  // comes from the y row below)
      } else {
        V.SY[(size_t)slot_new * cap + pslot] = a0;                                                // s_new . y_p
        V.YY[(size_t)slot_new * cap + pslot] = a1, V.YY[(size_t)pslot * cap + slot_new] = a1;   // y_new . y_p
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    int h = head, u = used;
    double hd = V.delta[2 * cap + 1];
    if (st_g->n_iter_total == 1) {
      h = 0, u = 0, hd = 1.0;                                   // lbfgs.py:396-402
    } else if (slot_new >= 0) {
      const double ys = V.SY[(size_t)slot_new * cap + slot_new], yy = V.YY[(size_t)slot_new * cap + slot_new];
      if (ys > 1e-10) {                                          // lbfgs.py:408-421
        if (u == st_g->hist) h = (h + 1) % cap;
        else u += 1;
        hd = ys / yy;
      }
    }
    V.delta[2 * cap + 1] = hd;
    sh_head = h, sh_used = u;
    st_g->head = h, st_g->used = u;
  }
  __syncthreads();
  const int h2 = sh_head, m = sh_used;
  const double hdiag = V.delta[2 * cap + 1];
  // two-loop recursion on coefficients (lbfgs.py:432-447): q = -g; logical pair i sits in slot (h2 + i) % cap.
  //   backward  al_i = ro_i (s_i.q),  q -= al_i y_i      q has no s components yet, so s_i.q = sum_j cY_j (s_i.y_j) - s_i.g:
  //             a triangular recurrence over S^T Y -- sequential, done by ONE warp from a shared-memory copy of the m x m block
  //   r = H q;  w_i = sum_j cY_j (y_i.y_j) needs only the finished cY: a dense m x m product, all threads in parallel
  //   forward   be_i = ro_i (y_i.r) = ro_i (sum_j cS_j (s_j.y_i) + w_i + cG y_i.g),  cS_i += al_i - be_i: triangular again
  extern __shared__ double sy_stage[];                                // [m][m] logical block of S^T Y when it fits
  __shared__ double cSl[kLbMaxHistory + 1], cYl[kLbMaxHistory + 1];   // coefficients of s_i, y_i (logical order)
  __shared__ double rho_l[kLbMaxHistory + 1], wv[kLbMaxHistory + 1];
  auto slot = [&](int j) { return (h2 + j) % cap; };
  const bool staged = (size_t)m * m * sizeof(double) <= (size_t)stage_bytes;
  if (staged)
    for (int e = threadIdx.x; e < m * m; e += kVlThreads) sy_stage[e] = V.SY[(size_t)slot(e / m) * cap + slot(e % m)];
  for (int i = threadIdx.x; i < m; i += kVlThreads) {
    rho_l[i] = 1.0 / V.SY[(size_t)slot(i) * cap + slot(i)];           // ro_i = 1 / (y_i . s_i)
    cSl[i] = 0.0, cYl[i] = 0.0;
  }
  __syncthreads();
  auto sy = [&](int i, int j) { return staged ? sy_stage[i * m + j] : V.SY[(size_t)slot(i) * cap + slot(j)]; };   // s_i . y_j
  double cG = -1.0;
  if (warp == 0) {
    for (int i = m - 1; i >= 0; --i) {
      double acc = 0.0;
      for (int j = i + 1 + lane; j < m; j += 32) acc = fma(cYl[j], sy(i, j), acc);   // (cY_j is still 0 for j <= i)
      const double a_i = (vl_warp_sum(acc) + cG * dot_g_s[slot(i)]) * rho_l[i];
      if (lane == 0) al[i] = a_i, cYl[i] = -a_i;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += kVlThreads) cYl[i] *= hdiag;
  cG *= hdiag;
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += kVlThreads) {
    double acc = 0.0;
    const double* yyr = V.YY + (size_t)slot(i) * cap;
#pragma unroll 4
    for (int j = 0; j < m; ++j) acc = fma(cYl[j], yyr[slot(j)], acc);
    wv[i] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    for (int i = 0; i < m; ++i) {
      double acc = 0.0;
      for (int j = lane; j < i; j += 32) acc = fma(cSl[j], sy(j, i), acc);           // (cS_j is still 0 for j >= i)
      const double be = (vl_warp_sum(acc) + wv[i] + cG * dot_g_y[slot(i)]) * rho_l[i];
      if (lane == 0) cSl[i] = al[i] - be;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int p_ = threadIdx.x; p_ < cap; p_ += kVlThreads) dS[p_] = 0.0, dY[p_] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += kVlThreads) dS[slot(i)] = cSl[i], dY[slot(i)] = cYl[i];
  if (threadIdx.x == 0) V.delta[2 * cap] = cG;
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel 3: d = sum_j delta_j b_j (second pass over the history), prev_g = g, statistics; last CTA: step length, line-search set-up
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVlThreads)
    vl_combine_kernel(LbState* __restrict__ st_g, LbVectors V, const float* __restrict__ g, LbStatus* __restrict__ status_out,
                      long long P) {
  if (!st_g->begin) return;
  __shared__ double cS[kLbMaxHistory + 1], cY[kLbMaxHistory + 1];
  __shared__ double warp_buf[kVlThreads / 32];
  __shared__ float fmax_buf[kVlThreads / 32];
  const int cap = st_g->hist + 1;
  const int head = st_g->head, m = st_g->used;
  for (int p_ = threadIdx.x; p_ < cap; p_ += kVlThreads) cS[p_] = V.delta[p_], cY[p_] = V.delta[cap + p_];
  __syncthreads();
  const double cG = V.delta[2 * cap];
  double gd = 0.0, l1 = 0.0;
  float md = 0.f;
  for (long long i = (long long)blockIdx.x * kVlThreads + threadIdx.x; i < P; i += (long long)gridDim.x * kVlThreads) {
    const float gi = g[i];
    double acc = cG * (double)gi;
#pragma unroll 8
    for (int j = 0; j < m; ++j) {
      const int pj = (head + j) % cap;
      acc = fma(cS[pj], (double)V.S[(long long)pj * P + i], acc);
      acc = fma(cY[pj], (double)V.Y[(long long)pj * P + i], acc);
    }
    const float di = (float)acc;
    V.d[i] = di;
    V.prev_g[i] = gi;                                           // prev_flat_grad.copy_(flat_grad), lbfgs.py:449-452
    gd = fma((double)gi, (double)di, gd);
    l1 += (double)fabsf(gi);
    md = lb_nan_max(md, fabsf(di));
  }
  gd = vl_block_sum(gd, warp_buf);
  l1 = vl_block_sum(l1, warp_buf);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) md = lb_nan_max(md, __shfl_xor_sync(0xffffffffu, md, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) fmax_buf[threadIdx.x >> 5] = md;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = 0.f;
    for (int i = 0; i < kVlThreads / 32; ++i) mm = lb_nan_max(mm, fmax_buf[i]);
    double* sp = V.stats_partial + (size_t)blockIdx.x * 4;
    sp[0] = gd, sp[1] = l1, sp[2] = (double)mm;
  }
  if (!vl_last_cta(V.counters + 1)) return;
  // ---------------- last CTA: fixed-order reduction of the partials, then one thread: lbfgs.py:449-487 ----------------
  double gtd = 0.0, g_l1 = 0.0;
  float dmax = 0.f;
  for (unsigned c = threadIdx.x; c < gridDim.x; c += kVlThreads) {
    const double* sp = V.stats_partial + (size_t)c * 4;
    gtd += sp[0], g_l1 += sp[1];
    dmax = lb_nan_max(dmax, (float)sp[2]);
  }
  gtd = vl_block_sum(gtd, warp_buf);
  g_l1 = vl_block_sum(g_l1, warp_buf);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = lb_nan_max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) fmax_buf[threadIdx.x >> 5] = dmax;
  __syncthreads();
  if (threadIdx.x != 0) return;
  dmax = 0.f;
  for (int i = 0; i < kVlThreads / 32; ++i) dmax = lb_nan_max(dmax, fmax_buf[i]);
  LbState st = *st_g;
  st.prev_loss = st.loss;
  st.gtd = gtd;
  st.d_norm = (double)dmax;
  double t;
  if (st.n_iter_total == 1) {
    const double inv = 1.0 / g_l1;
    t = (inv < 1.0 ? inv : 1.0) * st.lr;          // min(1., 1. / |g|_1) * lr
  } else {
    t = st.lr;
  }
  st.t = t;
  if (st.gtd > -st.tol_change) {                  // lbfgs.py:463
    st.status = LB_STATUS_DONE;
    st.ph = LB_PH_DONE;
  } else {
    // line search set-up (lbfgs.py:40-56 with max_ls = max_eval - current_evals, :486)
    st.f0 = st.loss, st.gtd0 = st.gtd, st.ls_t = t;
    st.t_prev = 0.0, st.f_prev = st.loss, st.gtd_prev = st.gtd;
    st.br_n = 0, st.done = 0, st.insuf = 0, st.ls_iter = 0, st.ls_evals = 0;
    st.max_ls = st.max_eval - st.current_evals;
    st.low = 0, st.high = 1;
    st.ph = LB_PH_BRACKET;
    st.trial = 1, st.trial_t = t;
    st.status = LB_STATUS_EVAL;
  }
  *st_g = st;
  lb_write_status(st, status_out);
}

// Kernel 4: first trial point of a new line search: x0 = x, g_prev = g (torch clones both), x = x0 + t d
__global__ void __launch_bounds__(kVlThreads)
    lbfgs_trial_kernel(const LbState* __restrict__ st_g, LbVectors V, float* __restrict__ flat, const float* __restrict__ g, long long P) {
  if (!st_g->begin || !st_g->trial) return;
  const float tt = (float)st_g->trial_t;
  for (long long i = (long long)blockIdx.x * kVlThreads + threadIdx.x; i < P; i += (long long)gridDim.x * kVlThreads) {
    const float x = flat[i];
    V.x0[i] = x;
    V.gp[i] = g[i];
    flat[i] = fmaf(tt, V.d[i], x);
  }
}

static size_t lb_align(size_t x) { return (x + 255) & ~size_t(255); }

constexpr int kVlMaxGrid = 592;   // 148 SMs x 4: the whole-chip kernels never launch more CTAs than this

struct LbLayout {
  size_t state, status, counters, vec, hist, gram, delta, dots, stats, total;
};
static LbLayout lb_layout(long long P, int hist) {
  const size_t cap = (size_t)hist + 1;
  LbLayout L;
  L.state = 0;
  L.status = lb_align(sizeof(LbState));
  L.counters = L.status + lb_align(sizeof(LbStatus));
  size_t o = L.counters + 256;
  L.gram = o;
  o += 3 * lb_align(cap * cap * 8);
  L.delta = o;
  o += lb_align((2 * cap + 2) * 8);
  L.dots = o;
  o += lb_align((size_t)kVlMaxGrid * 2 * cap * 3 * 8);
  L.stats = o;
  o += lb_align((size_t)kVlMaxGrid * 4 * 8);
  L.vec = o;                       // everything before this offset is zeroed by a reset
  o += 6 * lb_align((size_t)P * 4);
  L.hist = o;
  o += 2 * lb_align(cap * (size_t)P * 4);
  L.total = o;
  return L;
}
// dynamic shared memory of vl_dots_kernel: the m x m block of S^T Y for the coefficient recurrences (when it fits)
static int vl_stage_bytes(int hist) {
  const long long want = (long long)hist * hist * 8;
  return (int)(want <= 160 * 1024 ? want : 0);
}
static int vl_grid(long long P) {
  static int sms = 0;          // (queried once: this sits on the per-evaluation host path)
  if (sms == 0) {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms = v;
  }
  long long g = (P + 1023) / 1024;           // at least 1024 elements per CTA
  const long long cap_g = (long long)sms * 4 < kVlMaxGrid ? (long long)sms * 4 : kVlMaxGrid;
  if (g > cap_g) g = cap_g;
  return (int)(g < 1 ? 1 : g);
}
static LbVectors lb_vectors(void* ws, long long P, int hist) {
  const LbLayout L = lb_layout(P, hist);
  const size_t cap = (size_t)hist + 1;
  char* b = reinterpret_cast<char*>(ws);
  const size_t vs = lb_align((size_t)P * 4);
  LbVectors V;
  V.d = reinterpret_cast<float*>(b + L.vec);
  V.prev_g = reinterpret_cast<float*>(b + L.vec + vs);
  V.x0 = reinterpret_cast<float*>(b + L.vec + 2 * vs);
  V.gp = reinterpret_cast<float*>(b + L.vec + 3 * vs);
  V.b0 = reinterpret_cast<float*>(b + L.vec + 4 * vs);
  V.b1 = reinterpret_cast<float*>(b + L.vec + 5 * vs);
  V.S = reinterpret_cast<float*>(b + L.hist);
  V.Y = reinterpret_cast<float*>(b + L.hist + lb_align(cap * (size_t)P * 4));
  V.SS = reinterpret_cast<double*>(b + L.gram);
  V.SY = reinterpret_cast<double*>(b + L.gram + lb_align(cap * cap * 8));
  V.YY = reinterpret_cast<double*>(b + L.gram + 2 * lb_align(cap * cap * 8));
  V.delta = reinterpret_cast<double*>(b + L.delta);
  V.dots_partial = reinterpret_cast<double*>(b + L.dots);
  V.stats_partial = reinterpret_cast<double*>(b + L.stats);
  V.counters = reinterpret_cast<unsigned*>(b + L.counters);
  V.grid = vl_grid(P);
  return V;
}

__global__ void lbfgs_begin_kernel(LbState* st, double lr, double tol_grad, double tol_change, int max_iter, int max_eval,
                                   int hist, int reset) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  LbState s = *st;
  if (reset) {
    s.n_iter_total = 0, s.func_evals_total = 0, s.head = 0, s.used = 0;
    s.t = 0, s.loss = 0, s.prev_loss = 0;
  }
  s.lr = lr, s.tol_grad = tol_grad, s.tol_change = tol_change;
  s.max_iter = max_iter, s.max_eval = max_eval, s.hist = hist;
  s.ph = LB_PH_START, s.n_iter = 0, s.current_evals = 0, s.status = 0;
  *st = s;
}

// bench hook: pretend the history is full and an iteration begins (the vectors hold whatever the caller put there)
__global__ void lbfgs_probe_setup_kernel(LbState* st, int hist) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->hist = hist, st->head = 0, st->used = hist - 1, st->slot_new = hist - 1;
  st->n_iter_total = 5, st->begin = 1, st->trial = 0, st->t = 1.0, st->lr = 1.0, st->tol_change = 0.0;
  st->max_eval = 100, st->current_evals = 1, st->loss = 1.0;
}

// opt in to the dynamic shared memory of vl_dots_kernel once per device (not on every evaluation)
static cudaError_t vl_allow_stage(int stage) {
  static int allowed[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || allowed[dev] < stage + 1) {
    e = cudaFuncSetAttribute(vl_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess && dev >= 0 && dev < 64) allowed[dev] = 160 * 1024 + 1;
  }
  return e;
}

}  // namespace pinn

using namespace pinn;

extern "C" int pinn_lbfgs_workspace_bytes(int64_t n_params, int32_t history_size, size_t* bytes) {
  if (!bytes || n_params <= 0 || history_size < 1) return set_error("lbfgs_workspace_bytes: bad arguments"), PINN_E_ARG;
  if (history_size > kLbMaxHistory) return set_error("history_size > %d", kLbMaxHistory), PINN_E_UNSUPPORTED;
  *bytes = lb_layout(n_params, history_size).total;
  return PINN_OK;
}

extern "C" int pinn_lbfgs_begin(void* workspace, int64_t n_params, const pinn_lbfgs_cfg_t* cfg, int32_t reset, void* stream) {
  if (!workspace || !cfg || n_params <= 0) return set_error("lbfgs_begin: bad arguments"), PINN_E_ARG;
  if (cfg->history_size < 1 || cfg->history_size > kLbMaxHistory) return set_error("lbfgs_begin: history_size out of range"), PINN_E_UNSUPPORTED;
  if (((uintptr_t)workspace & 255) != 0) return set_error("lbfgs workspace must be 256-byte aligned"), PINN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (reset) PINN_CUDA(cudaMemsetAsync(workspace, 0, lb_layout(n_params, cfg->history_size).vec, st));
  lbfgs_begin_kernel<<<1, 32, 0, st>>>(reinterpret_cast<LbState*>(workspace), cfg->lr, cfg->tolerance_grad, cfg->tolerance_change,
                                       cfg->max_iter, cfg->max_eval, cfg->history_size, reset);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

extern "C" int pinn_lbfgs_advance(void* workspace, int64_t n_params, int32_t history_size, float* flat_params, float* grad,
                                  const float* loss, void* status_host, void* stream) {
  if (!workspace || !flat_params || !grad || !loss || !status_host || n_params <= 0)
    return set_error("lbfgs_advance: bad arguments"), PINN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const LbLayout L = lb_layout(n_params, history_size);
  char* b = reinterpret_cast<char*>(workspace);
  LbState* state = reinterpret_cast<LbState*>(b);
  LbStatus* dev_status = reinterpret_cast<LbStatus*>(b + L.status);
  const LbVectors V = lb_vectors(workspace, n_params, history_size);
  const long long P = (long long)n_params;
  lbfgs_advance_kernel<<<kLbCtas, kLbThreads, 0, st>>>(state, V, flat_params, grad, loss, dev_status, P);
  // the next three return at once unless the kernel above started a new outer iteration (decided on the device)
  const int stage = vl_stage_bytes(history_size);
  PINN_CUDA(vl_allow_stage(stage));
  vl_dots_kernel<<<V.grid, kVlThreads, stage, st>>>(state, V, grad, P, stage);
  vl_combine_kernel<<<V.grid, kVlThreads, 0, st>>>(state, V, grad, dev_status, P);
  lbfgs_trial_kernel<<<V.grid, kVlThreads, 0, st>>>(state, V, flat_params, grad, P);
  PINN_CUDA(cudaGetLastError());
  PINN_CUDA(cudaMemcpyAsync(status_host, dev_status, sizeof(LbStatus), cudaMemcpyDeviceToHost, st));
  return PINN_OK;
}

// Standalone timing hook for bench.py ("L-BFGS direction ... HBM-bound, reported as GB/s"): runs kernels 2-3
// (the two passes over the history) on a workspace whose history is full, without touching the optimiser's state machine.
extern "C" int pinn_lbfgs_direction_probe(void* workspace, int64_t n_params, int32_t history_size, const float* grad,
                                          double* bytes_out, void* stream) {
  if (!workspace || !grad || n_params <= 0) return set_error("lbfgs_direction_probe: bad arguments"), PINN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const LbLayout L = lb_layout(n_params, history_size);
  char* b = reinterpret_cast<char*>(workspace);
  LbState* state = reinterpret_cast<LbState*>(b);
  const LbVectors V = lb_vectors(workspace, n_params, history_size);
  lbfgs_probe_setup_kernel<<<1, 32, 0, st>>>(state, history_size);
  const int stage = vl_stage_bytes(history_size);
  PINN_CUDA(vl_allow_stage(stage));
  vl_dots_kernel<<<V.grid, kVlThreads, stage, st>>>(state, V, grad, (long long)n_params, stage);
  vl_combine_kernel<<<V.grid, kVlThreads, 0, st>>>(state, V, grad, reinterpret_cast<LbStatus*>(b + L.status), (long long)n_params);
  PINN_CUDA(cudaGetLastError());
  // algorithmic bytes: both passes read the 2 m P history once; pass 1 also reads g, prev_g, d and writes the new pair,
  // pass 2 reads g and writes d, prev_g
  if (bytes_out) *bytes_out = 4.0 * (double)n_params * (2.0 * 2.0 * history_size + 5.0 + 3.0);
  return PINN_OK;
}
