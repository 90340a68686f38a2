// Standalone probe: how fast can 148 CTAs accumulate 128x256 fp32 tiles into ONE shared 256x256 matrix?
// (the weight-gradient drain of jet_tc.cu).  Variants:
//   0  red.global.add.v2.f32, quad = 8 consecutive columns of one row (sector-complete), all CTAs same order
//   1  same, CTA-rotated start row/column (spreads the hot lines over time)
//   2  red.global.add.v4.f32, thread = row, 4 consecutive columns (32 rows per instruction, half sectors)
//   3  red.global.add.v4.f32, fully coalesced (warp = 512 contiguous bytes)
//   4  cp.reduce.async.bulk.global.shared::cta.add.f32 of 1 KB rows staged in shared memory
//   5  variant 0 into a CTA-private matrix (no inter-SM contention; raw LSU->L2 RED rate)
//   6  variant 3 into a CTA-private matrix
//   7  variant 4 with 4 KB pieces (4 rows contiguous: needs a [64][1024]-float private layout, here CTA-shared G viewed flat)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void red_v2(float* a, float x, float y) { asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(a), "f"(x), "f"(y) : "memory"); }
__device__ __forceinline__ void red_v4(float* a, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void bulk_red(float* g, const float* s, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(512, 1) red_kernel(float* G, float* Gpriv, int variant, int reps, long long* cycles) {
  extern __shared__ __align__(128) float stage[];   // 128 KB for variant 4/7
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sp = warp & 3, half = warp >> 2;      // as in jet_tc.cu: 4 row groups x 4 column slices of 64
  const int cbase = half * 64;
  if (variant == 4 || variant == 7) {
    for (int i = tid; i < 128 * 256; i += 512) stage[i] = 1.0f;
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  float* Gp = Gpriv + (size_t)blockIdx.x * 65536;
  for (int r = 0; r < reps; ++r) {
    const int h = r & 1;
    if (variant == 0 || variant == 1 || variant == 5) {
      float* base = variant == 5 ? Gp : G;
      const int rot = variant == 1 ? (int)blockIdx.x : 0;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int row = h * 128 + ((sp * 32 + g * 16 + (lane >> 2) + rot * 8) & 127);
        float* grow = base + (size_t)row * 256 + 2 * (lane & 3);
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c0 = (cbase + cb * 32 + 8 * u + rot * 32) & 255;
            red_v2(grow + c0, 1.f, 1.f);
            const int row2 = h * 128 + ((sp * 32 + g * 16 + (lane >> 2) + 8 + rot * 8) & 127);
            red_v2(base + (size_t)row2 * 256 + 2 * (lane & 3) + c0, 1.f, 1.f);
          }
        }
      }
    } else if (variant == 2) {
      const int row = h * 128 + sp * 32 + lane;
      float* grow = G + (size_t)row * 256 + cbase;
#pragma unroll
      for (int c = 0; c < 16; ++c) red_v4(grow + 4 * c, 1.f, 1.f, 1.f, 1.f);
    } else if (variant == 3 || variant == 6) {
      float* base = (variant == 6 ? Gp : G) + (size_t)h * 32768;
      // 32768 floats per half, 512 threads x 4 floats = 2048 floats per sweep, 16 sweeps
#pragma unroll
      for (int s = 0; s < 16; ++s) red_v4(base + (size_t)s * 2048 + tid * 4, 1.f, 1.f, 1.f, 1.f);
    } else if (variant == 8) {
      // scalar red, transposed accumulator: thread = column (f_in), 64 rows (f_out) per thread; a warp instruction = 128 contiguous bytes
      float* base = G + (size_t)(half * 64) * 256 + h * 128 + sp * 32 + lane;
#pragma unroll
      for (int c = 0; c < 64; ++c) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(base + (size_t)c * 256), "f"(1.0f) : "memory");
    } else if (variant == 9) {
      // v2, quad pairs own 16 consecutive columns: 8 lanes x 8 B = 64 B per row, 4 rows per instruction
      const int row = h * 128 + sp * 32 + (lane >> 3);
      float* grow = G + (size_t)row * 256 + cbase + 2 * (lane & 7);
#pragma unroll
      for (int g = 0; g < 8; ++g)
#pragma unroll
        for (int u = 0; u < 4; ++u) red_v2(grow + (size_t)(4 * g) * 256 + 16 * u, 1.f, 1.f);
    } else if (variant == 4) {
      if (tid < 128) bulk_red(G + (size_t)(h * 128 + tid) * 256, stage + tid * 256, 1024);
      if (tid < 128) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (tid < 128) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else if (variant == 7) {
      if (tid < 32) bulk_red(G + (size_t)h * 32768 + (size_t)tid * 1024, stage + tid * 1024, 4096);
      if (tid < 32) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (tid < 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  if (variant == 4 || variant == 7) {
    if (tid < 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  float *G, *Gp; long long* cyc;
  cudaMalloc(&G, 65536 * 4); cudaMalloc(&Gp, (size_t)148 * 65536 * 4); cudaMalloc(&cyc, 148 * 8);
  const int reps = 64;
  cudaFuncSetAttribute(red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  const char* names[] = {"v2 sector-complete, same order", "v2 sector-complete, CTA-rotated", "v4 thread=row (half sectors)", "v4 coalesced",
                         "bulk reduce 1 KB rows", "v2 sector-complete, private", "v4 coalesced, private", "bulk reduce 4 KB pieces",
                         "scalar, warp = 128 contiguous bytes", "v2, 64 B per row x 4 rows"};
  for (int v = 0; v < 10; ++v) {
    cudaMemset(G, 0, 65536 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    red_kernel<<<148, 512, 131072>>>(G, Gp, v, 4, cyc);   // warm-up
    cudaMemset(G, 0, 65536 * 4);
    cudaEventRecord(e0);
    red_kernel<<<148, 512, 131072>>>(G, Gp, v, reps, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", v, cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> c(148); cudaMemcpy(c.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto x : c) mx = x > mx ? x : mx;
    std::vector<float> g(65536); cudaMemcpy(g.data(), G, 65536 * 4, cudaMemcpyDeviceToHost);
    const double bytes = (double)148 * reps * 131072;
    printf("variant %d  %-36s  %8.3f ms   %7.1f cycles per 128x256 drain per CTA   %7.1f GB/s chip   %5.2f B/clk/SM   G[0]=%g G[65535]=%g\n", v, names[v], ms,
           (double)mx / reps, bytes / ms * 1e-6, 131072.0 * reps / (double)mx, g[0], g[65535]);
  }
  return 0;
}
