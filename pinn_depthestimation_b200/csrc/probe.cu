// Measurement helper: FP32 FMA throughput probe used by bench.py as the roofline denominator of the
// FP32 path (MEASURED_PEAKS.json only holds HBM and bf16 tensor peaks).
#include "common.cuh"

namespace pinn {

// 8 independent accumulator chains per thread, `iters` x 8 x 4 FMAs each.
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float a, float b) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = fmaf(r[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i];
  if (s == 123.456f) out[0] = s;  // never true; keeps the chains alive
}

}  // namespace pinn

extern "C" int pinn_fma_probe(float* out, int32_t iters, int32_t ctas, double* flops_out, void* stream) {
  if (!out || iters < 1 || ctas < 1) return pinn::set_error("fma_probe: bad arguments"), PINN_E_ARG;
  pinn::fma_probe_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(out, iters, 1.0000001f, 1e-9f);
  PINN_CUDA(cudaGetLastError());
  if (flops_out) *flops_out = 2.0 * 32.0 * (double)iters * 256.0 * (double)ctas;
  return PINN_OK;
}
