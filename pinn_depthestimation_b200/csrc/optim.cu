// Optimiser kernels on the flat parameter / gradient vectors.
//
// Replaces the vector work inside torch.optim.LBFGS.step (torch/optim/lbfgs.py:386-457, the
// two-loop recursion with its 2m host-synchronising dots) and torch.optim.Adam.step, which the
// reference constructs at train_newmethod.py:95-117 and drives at train_newmethod.py:197-209.
//
// The two-loop recursion is ONE launch of a thread-block cluster: the cluster's CTAs split the
// vector, per-step partial dots are exchanged through distributed shared memory and summed in a
// fixed order (bit-reproducible, identical in every CTA), one cluster barrier per step.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pinn {

constexpr int kClusterCtas = 8;
constexpr int kVecThreads = 1024;
constexpr int kMaxHistory = 1024;

__device__ __forceinline__ float block_sum(float v, float* warp_buf) {
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
// max that propagates NaN like torch's abs().max() (fmaxf drops it: a NaN gradient must not look converged)
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float block_max(float v, float* warp_buf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) warp_buf[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t = nan_max(t, warp_buf[i]);
  return t;
}

// Sum of one float per CTA across the cluster, the same value returned to every thread of every CTA.
// slot / bcast are 2-entry (double-buffered) shared arrays; `phase` alternates per call.
__device__ __forceinline__ float cluster_sum(cg::cluster_group& cl, float mine, float* slot,
                                             float* bcast, int phase) {
  if (threadIdx.x == 0) slot[phase] = mine;
  cl.sync();
  if (threadIdx.x < 32) {
    const unsigned n = cl.num_blocks();
    const float v = threadIdx.x < n ? *cl.map_shared_rank(slot + phase, threadIdx.x) : 0.f;
    float t = 0.f;
    for (unsigned r = 0; r < n; ++r) t += __shfl_sync(0xffffffffu, v, r);  // fixed order
    if (threadIdx.x == 0) bcast[phase] = t;
  }
  __syncthreads();
  return bcast[phase];
}

// d = -H g by the two-loop recursion (torch/optim/lbfgs.py:432-447).
//   slot(i) = (head + i) % hist, i = 0 oldest .. m-1 newest;  old_dirs = y, old_stps = s.
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kVecThreads)
    lbfgs_direction_kernel(const float* __restrict__ S, const float* __restrict__ Y,
                           const float* __restrict__ rho, const float* __restrict__ h_diag,
                           const float* __restrict__ g, float* __restrict__ d, int hist, int m,
                           int head, long long P) {
  cg::cluster_group cl = cg::this_cluster();
  __shared__ float warp_buf[32];
  __shared__ float slot[2], bcast[2];
  __shared__ float al[kMaxHistory];  // identical in every CTA
  const long long per = (P + kClusterCtas - 1) / kClusterCtas;
  const long long lo = per * cl.block_rank();
  const long long hi = lo + per < P ? lo + per : P;
  int phase = 0;

  // q = -g ; partial of s_{m-1}.q
  {
    const float* s = m > 0 ? S + (long long)((head + m - 1) % hist) * P : nullptr;
    float part = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += kVecThreads) {
      const float q = -g[i];
      d[i] = q;
      if (s) part = fmaf(s[i], q, part);
    }
    if (m > 0) {
      part = block_sum(part, warp_buf);
      float dot = cluster_sum(cl, part, slot, bcast, phase);
      phase ^= 1;
      const int sl = (head + m - 1) % hist;
      const float a = dot * rho[sl];
      if (threadIdx.x == 0) al[m - 1] = a;
      // backward loop: q -= al_i*y_i fused with the next dot s_{i-1}.q
      float a_i = a;
      for (int i = m - 1; i >= 0; --i) {
        const float* y = Y + (long long)((head + i) % hist) * P;
        const float* sn = i > 0 ? S + (long long)((head + i - 1) % hist) * P : nullptr;
        float p2 = 0.f;
        for (long long e = lo + threadIdx.x; e < hi; e += kVecThreads) {
          const float q = fmaf(-a_i, y[e], d[e]);
          d[e] = q;
          if (sn) p2 = fmaf(sn[e], q, p2);
        }
        if (i > 0) {
          p2 = block_sum(p2, warp_buf);
          const float dot2 = cluster_sum(cl, p2, slot, bcast, phase);
          phase ^= 1;
          a_i = dot2 * rho[(head + i - 1) % hist];
          if (threadIdx.x == 0) al[i - 1] = a_i;
        }
      }
    }
  }
  // r = q * H_diag ; forward loop: be_i = rho_i * y_i.r ; r += (al_i - be_i) * s_i
  const float hd = *h_diag;
  {
    const float* y0 = m > 0 ? Y + (long long)(head % hist) * P : nullptr;
    float part = 0.f;
    for (long long e = lo + threadIdx.x; e < hi; e += kVecThreads) {
      const float r = d[e] * hd;
      d[e] = r;
      if (y0) part = fmaf(y0[e], r, part);
    }
    for (int i = 0; i < m; ++i) {
      part = block_sum(part, warp_buf);
      const float dot = cluster_sum(cl, part, slot, bcast, phase);
      phase ^= 1;
      const int sl = (head + i) % hist;
      const float be = dot * rho[sl];
      const float coef = al[i] - be;
      const float* s = S + (long long)sl * P;
      const float* yn = i + 1 < m ? Y + (long long)((head + i + 1) % hist) * P : nullptr;
      part = 0.f;
      for (long long e = lo + threadIdx.x; e < hi; e += kVecThreads) {
        const float r = fmaf(coef, s[e], d[e]);
        d[e] = r;
        if (yn) part = fmaf(yn[e], r, part);
      }
    }
  }
  cl.sync();  // keep every CTA's shared memory alive until all remote reads are done
}

// out = [a.b, sum|a|, max|a|, max|b|, a.a, b.b]
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kVecThreads)
    vec_stats_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                     float* __restrict__ out) {
  cg::cluster_group cl = cg::this_cluster();
  __shared__ float warp_buf[32];
  __shared__ float part[6];
  const long long per = (n + kClusterCtas - 1) / kClusterCtas;
  const long long lo = per * cl.block_rank();
  const long long hi = lo + per < n ? lo + per : n;
  float ab = 0.f, l1 = 0.f, ma = 0.f, mb = 0.f, aa = 0.f, bb = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += kVecThreads) {
    const float x = a[i], y = b ? b[i] : 0.f;
    ab = fmaf(x, y, ab);
    l1 += fabsf(x);
    ma = nan_max(ma, fabsf(x));
    mb = nan_max(mb, fabsf(y));
    aa = fmaf(x, x, aa);
    bb = fmaf(y, y, bb);
  }
  ab = block_sum(ab, warp_buf);
  l1 = block_sum(l1, warp_buf);
  aa = block_sum(aa, warp_buf);
  bb = block_sum(bb, warp_buf);
  ma = block_max(ma, warp_buf);
  mb = block_max(mb, warp_buf);
  if (threadIdx.x == 0) {
    part[0] = ab, part[1] = l1, part[2] = ma, part[3] = mb, part[4] = aa, part[5] = bb;
  }
  cl.sync();
  if (cl.block_rank() == 0 && threadIdx.x < 6) {
    const int k = threadIdx.x;
    float t = 0.f;
    for (unsigned r = 0; r < cl.num_blocks(); ++r) {
      const float v = *cl.map_shared_rank(part + k, r);
      t = (k == 2 || k == 3) ? nan_max(t, v) : t + v;
    }
    out[k] = t;
  }
  cl.sync();
}

__global__ void axpy_kernel(float alpha, const float* __restrict__ x, float* __restrict__ y,
                            long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

// torch.optim.Adam (amsgrad=False, maximize=False) single-tensor update on the flat vector
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, long long n,
                            float step_size, float beta1, float beta2, float eps, float wd,
                            float inv_bc2_sqrt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (wd != 0.f) gi = fmaf(wd, p[i], gi);
    const float mi = m[i] + (1.f - beta1) * (gi - m[i]);       // exp_avg.lerp_(grad, 1-beta1)
    const float vi = fmaf(1.f - beta2, gi * gi, v[i] * beta2);  // mul_(beta2).addcmul_(g,g,1-beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

static int grid_for(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : b > 1184 ? 1184 : b);
}

int run_lbfgs_direction(const float* S, const float* Y, const float* rho, const float* h_diag,
                        const float* g, float* d, int hist, int m, int head, long long P,
                        float* scratch, cudaStream_t st) {
  if (m < 0 || m > hist || hist < 1 || head < 0 || head >= hist)
    return set_error("lbfgs_direction: bad history indices (hist=%d m=%d head=%d)", hist, m, head), PINN_E_ARG;
  if (P <= 0) return set_error("lbfgs_direction: n_params <= 0"), PINN_E_ARG;
  if (hist > kMaxHistory) return set_error("lbfgs_direction: history_size > %d", kMaxHistory), PINN_E_UNSUPPORTED;
  (void)scratch;
  lbfgs_direction_kernel<<<kClusterCtas, kVecThreads, 0, st>>>(S, Y, rho, h_diag, g, d, hist, m, head, P);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

int run_vec_stats(const float* a, const float* b, long long n, float* out, cudaStream_t st) {
  if (n <= 0) return set_error("vec_stats: n <= 0"), PINN_E_ARG;
  vec_stats_kernel<<<kClusterCtas, kVecThreads, 0, st>>>(a, b, n, out);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

int run_axpy(float alpha, const float* x, float* y, long long n, cudaStream_t st) {
  if (n <= 0) return PINN_OK;
  axpy_kernel<<<grid_for(n), 256, 0, st>>>(alpha, x, y, n);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

int run_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
             float eps, float wd, long long step, cudaStream_t st) {
  if (n <= 0) return PINN_OK;
  if (step < 1) return set_error("adam: step_count must be >= 1"), PINN_E_ARG;
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  adam_kernel<<<grid_for(n), 256, 0, st>>>(p, g, m, v, n, step_size, b1, b2, eps, wd, inv_bc2_sqrt);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

}  // namespace pinn
