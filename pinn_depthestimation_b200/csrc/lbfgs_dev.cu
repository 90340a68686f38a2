// Device-resident L-BFGS (north_star (4): "two-loop recursion and line search run on device").
//
// Everything torch.optim.LBFGS.step does between two closure evaluations -- the strong-Wolfe bracket / zoom transition
// (torch/optim/lbfgs.py:40-209, cubic interpolation :12-37), the clones of the gradient it keeps for the bracket ends,
// the end-of-iteration termination tests (:511-526), the curvature-pair update (:404-421), the two-loop recursion
// (:432-447), the first step length (:454-457) and the next trial point x0 + t d -- is ONE launch of an 8-CTA thread-block
// cluster (`lbfgs_advance_kernel`).  The host launches the closure's kernels, then this kernel, then reads ONE small status
// block (one sync per evaluation) that only says "evaluate again" or "finished".  Call sites in the reference:
// train_newmethod.py:108-117 (construction), :204-209 (the single step(closure) call with max_iter = 50000).
//
// All scalar decisions are taken in double precision by thread 0 of every CTA from the same cluster-reduced inputs
// (fixed summation order), so every CTA -- and every GPU of a sharded run, which all-reduce loss and gradient first --
// takes the same branch.  pinn_depthestimation_b200/lbfgs.py holds the same logic as host code (used for
// line_search_fn=None and as the CPU-tested restatement of torch's functions); tests/test_gpu_optim.py holds both to
// torch's iteration / evaluation counts.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pinn {

constexpr int kLbCtas = 8;
constexpr int kLbThreads = 1024;
constexpr int kLbMaxHistory = 1024;

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
};

struct LbStatus {   // copied to the host after every advance
  int code, n_iter, current_evals, n_iter_total, func_evals_total, ph, used, pad;
  double t, loss, first_loss, gtd, d_norm;
};

struct LbVectors {  // all [P] unless noted; carved from the workspace
  float *d, *prev_g, *x0, *gp, *b0, *b1, *S, *Y, *rho, *h_diag;   // S, Y: [(hist+1), P]; rho: [hist+1]; h_diag: [1]
};

// ---- cluster-wide deterministic reductions: every thread of every CTA gets the same value ----
struct LbShared {
  float warp_buf[32];
  float slot[2][8];     // per-call partials of this CTA (up to 8 quantities), double-buffered
  float bcast[8];
  float al[kLbMaxHistory];
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

__device__ __forceinline__ float lb_dot_sum(cg::cluster_group& cl, LbShared& sh, float part, int& phase) {
  part = lb_block_sum(part, sh.warp_buf);
  __syncthreads();
  if (threadIdx.x == 0) sh.slot[phase][0] = part;
  cl.sync();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (unsigned r = 0; r < cl.num_blocks(); ++r) t += *cl.map_shared_rank(&sh.slot[phase][0], r);
    sh.bcast[0] = t;
  }
  __syncthreads();
  phase ^= 1;
  return sh.bcast[0];
}

// d = -H g by the two-loop recursion (torch/optim/lbfgs.py:432-447); slot(i) = (head + i) % cap, i = 0 oldest
__device__ void lb_two_loop(cg::cluster_group& cl, LbShared& sh, const LbVectors& V, const float* g, int cap, int m, int head,
                            long long P, long long lo, long long hi, int& phase) {
  float* d = V.d;
  const float* s_last = m > 0 ? V.S + (long long)((head + m - 1) % cap) * P : nullptr;
  float part = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) {
    const float q = -g[i];
    d[i] = q;
    if (s_last) part = fmaf(s_last[i], q, part);
  }
  if (m > 0) {
    float a_i = lb_dot_sum(cl, sh, part, phase) * V.rho[(head + m - 1) % cap];
    if (threadIdx.x == 0) sh.al[m - 1] = a_i;
    for (int i = m - 1; i >= 0; --i) {
      const float* y = V.Y + (long long)((head + i) % cap) * P;
      const float* sn = i > 0 ? V.S + (long long)((head + i - 1) % cap) * P : nullptr;
      float p2 = 0.f;
      for (long long e = lo + threadIdx.x; e < hi; e += kLbThreads) {
        const float q = fmaf(-a_i, y[e], d[e]);
        d[e] = q;
        if (sn) p2 = fmaf(sn[e], q, p2);
      }
      if (i > 0) {
        a_i = lb_dot_sum(cl, sh, p2, phase) * V.rho[(head + i - 1) % cap];
        if (threadIdx.x == 0) sh.al[i - 1] = a_i;
      }
    }
  }
  const float hd = *V.h_diag;
  const float* y0 = m > 0 ? V.Y + (long long)(head % cap) * P : nullptr;
  part = 0.f;
  for (long long e = lo + threadIdx.x; e < hi; e += kLbThreads) {
    const float r = d[e] * hd;
    d[e] = r;
    if (y0) part = fmaf(y0[e], r, part);
  }
  for (int i = 0; i < m; ++i) {
    const int sl = (head + i) % cap;
    const float be = lb_dot_sum(cl, sh, part, phase) * V.rho[sl];
    __syncthreads();
    const float coef = sh.al[i] - be;
    const float* s = V.S + (long long)sl * P;
    const float* yn = i + 1 < m ? V.Y + (long long)((head + i + 1) % cap) * P : nullptr;
    part = 0.f;
    for (long long e = lo + threadIdx.x; e < hi; e += kLbThreads) {
      const float r = fmaf(coef, s[e], d[e]);
      d[e] = r;
      if (yn) part = fmaf(yn[e], r, part);
    }
  }
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
  if (threadIdx.x == 0) st = *st_g;
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

  if (sh.do_begin) {
    // ---- start of an outer iteration (lbfgs.py:394-487) ----
    if (threadIdx.x == 0) {
      st.n_iter += 1;
      st.n_iter_total += 1;
    }
    __syncthreads();
    if (st.n_iter_total == 1) {
      if (threadIdx.x == 0) st.used = 0, st.head = 0;
      if (cl.block_rank() == 0 && threadIdx.x == 0) *V.h_diag = 1.f;
    } else {
      // curvature pair into the free ring slot: y = g - prev_g, s = t d; kept only if y.s > 1e-10
      const int slot = (st.head + st.used) % cap;
      float* Ys = V.Y + (long long)slot * P;
      float* Ss = V.S + (long long)slot * P;
      const float tf = (float)st.t;
      for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) {
        Ys[i] = g[i] - V.prev_g[i];
        Ss[i] = V.d[i] * tf;
      }
      __syncthreads();
      lb_stats(cl, sh, Ys, Ss, lo, hi, phase);
      const double ys = (double)sh.bcast[0], yy = (double)sh.bcast[4];
      if (ys > 1e-10) {
        if (threadIdx.x == 0) {
          if (st.used == st.hist) st.head = (st.head + 1) % cap;
          else st.used += 1;
        }
        if (cl.block_rank() == 0 && threadIdx.x == 0) {
          V.rho[slot] = (float)(1.0 / ys);
          *V.h_diag = (float)(ys / yy);
        }
      }
    }
    __threadfence();
    cl.sync();      // rho / h_diag written by rank 0 are visible to every CTA; st.used / st.head settled
    lb_two_loop(cl, sh, V, g, cap, st.used, st.head, P, lo, hi, phase);
    __syncthreads();
    for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) V.prev_g[i] = g[i];
    cl.sync();      // d complete everywhere before the statistics
    lb_stats(cl, sh, g, V.d, lo, hi, phase);
    if (threadIdx.x == 0) {
      st.prev_loss = st.loss;
      st.gtd = (double)sh.bcast[0];
      st.d_norm = (double)sh.bcast[3];
      const double g_l1 = (double)sh.bcast[1];
      double t;
      if (st.n_iter_total == 1) {
        const double inv = 1.0 / g_l1;
        t = (inv < 1.0 ? inv : 1.0) * st.lr;      // min(1., 1. / |g|_1) * lr
      } else {
        t = st.lr;
      }
      st.t = t;
      if (st.gtd > -st.tol_change) {              // lbfgs.py:463
        sh.status = LB_STATUS_DONE;
      } else {
        // line search set-up (lbfgs.py:40-56 with max_ls = max_eval - current_evals, :486)
        st.f0 = st.loss, st.gtd0 = st.gtd, st.ls_t = t;
        st.t_prev = 0.0, st.f_prev = st.loss, st.gtd_prev = st.gtd;
        st.br_n = 0, st.done = 0, st.insuf = 0, st.ls_iter = 0, st.ls_evals = 0;
        st.max_ls = st.max_eval - st.current_evals;
        st.low = 0, st.high = 1;
        st.ph = LB_PH_BRACKET;
        sh.do_trial = 2, sh.trial_t = t;
      }
    }
    __syncthreads();
  }

  if (sh.do_trial) {
    const float tt = (float)sh.trial_t;
    if (sh.do_trial == 2) {   // first trial of a line search: x0 = x, g_prev = g (torch clones both)
      for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) {
        const float x = flat[i];
        V.x0[i] = x;
        V.gp[i] = g[i];
        flat[i] = fmaf(tt, V.d[i], x);
      }
    } else {
      for (long long i = lo + threadIdx.x; i < hi; i += kLbThreads) flat[i] = fmaf(tt, V.d[i], V.x0[i]);
    }
    if (threadIdx.x == 0) sh.status = LB_STATUS_EVAL;
  }
  __syncthreads();
  cl.sync();   // every CTA has taken its copy of the old state long ago; remote shared-memory reads are over
  if (cl.block_rank() == 0 && threadIdx.x == 0) {
    st.status = sh.status;
    if (sh.status == LB_STATUS_DONE) st.ph = LB_PH_DONE;
    *st_g = st;
    LbStatus o;
    o.code = sh.status, o.n_iter = st.n_iter, o.current_evals = st.current_evals, o.n_iter_total = st.n_iter_total;
    o.func_evals_total = st.func_evals_total, o.ph = st.ph, o.used = st.used, o.pad = 0;
    o.t = st.t, o.loss = st.loss, o.first_loss = st.first_loss, o.gtd = st.gtd, o.d_norm = st.d_norm;
    *status_out = o;
  }
}

static size_t lb_align(size_t x) { return (x + 255) & ~size_t(255); }

struct LbLayout {
  size_t state, status, vec, hist, rho, total;
};
static LbLayout lb_layout(long long P, int hist) {
  LbLayout L;
  L.state = 0;
  L.status = lb_align(sizeof(LbState));
  size_t o = L.status + lb_align(sizeof(LbStatus));
  L.vec = o;
  o += 6 * lb_align((size_t)P * 4);
  L.hist = o;
  o += 2 * lb_align((size_t)(hist + 1) * (size_t)P * 4);
  L.rho = o;
  o += lb_align((size_t)(hist + 2) * 4);
  L.total = o;
  return L;
}
static LbVectors lb_vectors(void* ws, long long P, int hist) {
  const LbLayout L = lb_layout(P, hist);
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
  V.Y = reinterpret_cast<float*>(b + L.hist + lb_align((size_t)(hist + 1) * (size_t)P * 4));
  V.rho = reinterpret_cast<float*>(b + L.rho);
  V.h_diag = V.rho + hist + 1;
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
  LbStatus* dev_status = reinterpret_cast<LbStatus*>(b + L.status);
  lbfgs_advance_kernel<<<kLbCtas, kLbThreads, 0, st>>>(reinterpret_cast<LbState*>(b), lb_vectors(workspace, n_params, history_size),
                                                       flat_params, grad, loss, dev_status, (long long)n_params);
  PINN_CUDA(cudaGetLastError());
  PINN_CUDA(cudaMemcpyAsync(status_host, dev_status, sizeof(LbStatus), cudaMemcpyDeviceToHost, st));
  return PINN_OK;
}
