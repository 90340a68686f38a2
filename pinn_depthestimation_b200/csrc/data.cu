// One-off data path on the device: what the reference's __main__ blocks do with numpy before the
// tensors reach the GPU -- min/max normalisation to [-1,1] (operations.py:4-7), nanmin/nanmax ranges (operations.py:26-29),
// hstack of the columns to [N,d] and removal of the rows that hold a NaN (train_newmethod.py:226-255, train.py:274-276).
// Order-preserving stream compaction in three small launches (per-block counts, scan of the counts, scatter).
#include <math.h>

#include "common.cuh"

namespace pinn {

constexpr int kPrepThreads = 256;
constexpr int kPrepMaxCols = 16;

struct PrepArgs {
  const float* cols[kPrepMaxCols];   // n_in input columns followed by n_true true columns, each [n]
  float lo[kPrepMaxCols];            // normalisation range of the input columns
  float hi[kPrepMaxCols];
  int n_in, n_true;
  int nan_policy;                    // 1: drop rows with a NaN true value; 2: drop rows with a NaN input; 3: either
  long long n;
};

__device__ __forceinline__ bool prep_keep(const PrepArgs& a, long long i) {
  bool bad = false;
  if (a.nan_policy & 2)
    for (int c = 0; c < a.n_in; ++c) bad |= isnan(a.cols[c][i]);
  if (a.nan_policy & 1)
    for (int c = a.n_in; c < a.n_in + a.n_true; ++c) bad |= isnan(a.cols[c][i]);
  return !bad;
}

__global__ void __launch_bounds__(kPrepThreads) prep_count_kernel(const __grid_constant__ PrepArgs a, int* __restrict__ block_counts) {
  const long long i = (long long)blockIdx.x * kPrepThreads + threadIdx.x;
  const int keep = (i < a.n && prep_keep(a, i)) ? 1 : 0;
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of the block counts by ONE block (N / 256 entries: a few thousand at most per pass of 1024 * items)
__global__ void __launch_bounds__(1024) prep_scan_kernel(int* __restrict__ block_counts, int n_blocks, long long* __restrict__ n_kept) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? block_counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_tot[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_tot[threadIdx.x] = w;
    }
    __syncthreads();
    const int before = carry_s + (threadIdx.x >= 32 ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + x - v;
    if (i < n_blocks) block_counts[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_kept = carry_s;
}

__global__ void __launch_bounds__(kPrepThreads) prep_scatter_kernel(const __grid_constant__ PrepArgs a, const int* __restrict__ block_offsets,
                                                                    float* __restrict__ inputs, float* __restrict__ trues) {
  __shared__ int warp_tot[kPrepThreads / 32];
  const long long i = (long long)blockIdx.x * kPrepThreads + threadIdx.x;
  const int keep = (i < a.n && prep_keep(a, i)) ? 1 : 0;
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_tot[w] = __popc(m);
  __syncthreads();
  int before = block_offsets[blockIdx.x] + __popc(m & ((1u << lane) - 1u));
  for (int q = 0; q < w; ++q) before += warp_tot[q];
  if (!keep) return;
  for (int c = 0; c < a.n_in; ++c) {
    const float lo = a.lo[c], hi = a.hi[c];
    // operations.py:4-7: zeros when max == min, else 2 (x - min) / (max - min) - 1
    inputs[(long long)before * a.n_in + c] = (hi == lo) ? 0.f : 2.f * (a.cols[c][i] - lo) / (hi - lo) - 1.f;
  }
  for (int c = 0; c < a.n_true; ++c) trues[(long long)before * a.n_true + c] = a.cols[a.n_in + c][i];
}

// out2 = [nanmin, nanmax] of x (operations.py:26-29); +inf / -inf when every entry is NaN
__global__ void nan_minmax_kernel(const float* __restrict__ x, long long n, float* __restrict__ out2) {
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (!isnan(v)) lo = fminf(lo, v), hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // float atomics on min / max through the usual ordered-int trick (values are finite or +-inf, never NaN here)
    auto amin = [](float* addr, float v) { v >= 0.f ? atomicMin((int*)addr, __float_as_int(v)) : atomicMax((unsigned*)addr, __float_as_uint(v)); };
    auto amax = [](float* addr, float v) { v >= 0.f ? atomicMax((int*)addr, __float_as_int(v)) : atomicMin((unsigned*)addr, __float_as_uint(v)); };
    amin(out2, lo);
    amax(out2 + 1, hi);
  }
}

__global__ void minmax_init_kernel(float* out2) {
  out2[0] = INFINITY;
  out2[1] = -INFINITY;
}

}  // namespace pinn

using namespace pinn;

extern "C" int pinn_nan_minmax(const float* x, int64_t n, float* out2, void* stream) {
  if (!out2 || (!x && n > 0)) return set_error("nan_minmax: NULL pointer"), PINN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  minmax_init_kernel<<<1, 1, 0, st>>>(out2);
  if (n > 0) {
    long long b = (n + 255) / 256;
    nan_minmax_kernel<<<(int)(b > 1184 ? 1184 : b), 256, 0, st>>>(x, n, out2);
  }
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

extern "C" int pinn_assemble_points(const float* const* cols, int32_t n_in, int32_t n_true, const float* lo_host,
                                    const float* hi_host, int32_t nan_policy, int64_t n, float* inputs_out,
                                    float* trues_out, int64_t* n_kept, int32_t* scratch, void* stream) {
  if (n_in < 1 || n_true < 0 || n_in + n_true > kPrepMaxCols)
    return set_error("assemble_points: 1 <= n_in, n_in + n_true <= %d", kPrepMaxCols), PINN_E_ARG;
  if (!cols || !lo_host || !hi_host || !inputs_out || (n_true > 0 && !trues_out) || !n_kept || !scratch)
    return set_error("assemble_points: NULL pointer"), PINN_E_ARG;
  if (n < 0 || n > (1LL << 38)) return set_error("assemble_points: bad n"), PINN_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PrepArgs a;
  for (int c = 0; c < n_in + n_true; ++c) {
    if (!cols[c] && n > 0) return set_error("assemble_points: column %d is NULL", c), PINN_E_ARG;
    a.cols[c] = cols[c];
  }
  for (int c = 0; c < n_in; ++c) a.lo[c] = lo_host[c], a.hi[c] = hi_host[c];
  a.n_in = n_in, a.n_true = n_true, a.nan_policy = nan_policy, a.n = n;
  const long long blocks = (n + kPrepThreads - 1) / kPrepThreads;
  if (blocks == 0) {
    PINN_CUDA(cudaMemsetAsync(n_kept, 0, 8, st));
    return PINN_OK;
  }
  prep_count_kernel<<<(unsigned)blocks, kPrepThreads, 0, st>>>(a, scratch);
  prep_scan_kernel<<<1, 1024, 0, st>>>(scratch, (int)blocks, reinterpret_cast<long long*>(n_kept));
  prep_scatter_kernel<<<(unsigned)blocks, kPrepThreads, 0, st>>>(a, scratch, inputs_out, trues_out);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}
