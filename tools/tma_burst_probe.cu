// Probe: does the TMA engine overlap several outstanding 1-D bulk copies?  One thread issues N copies of B bytes
// into N different shared-memory buffers on ONE mbarrier and waits; prints cycles for N = 1..12 (one CTA alone and all 148).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(64, 1) burst(const char* src, int n, int bytes, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  unsigned char* buf = smem + 1024;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long best = 1ll << 60, tot = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * bytes) : "memory");
      for (int i = 0; i < n; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + (size_t)i * bytes)),
                     "l"(src + ((size_t)(r * n + i) * bytes) % (3u << 20)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
      const long long t1 = clock64();
      uint32_t done;
      do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"((uint32_t)(r & 1)) : "memory");
      } while (!done);
      const long long t2 = clock64();
      if (r > 0) { tot += t2 - t0; if (t2 - t0 < best) best = t2 - t0; }
      if (blockIdx.x == 0 && r == reps - 1) out[2] = t1 - t0;
    }
    if (blockIdx.x == 0) { out[0] = best; out[1] = tot / (reps - 1); }
  }
}
int main() {
  char* W; long long* out; cudaMalloc(&W, 4 << 20); cudaMemset(W, 0, 4 << 20); cudaMalloc(&out, 24);
  cudaFuncSetAttribute(burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 196608);
  for (int grid : {1, 148})
    for (int bytes : {4096, 16384, 32768})
      for (int n : {1, 2, 4, 6, 12}) {
        if ((size_t)n * bytes > 196608) continue;
        burst<<<grid, 64, 1024 + 196608>>>(W, n, bytes, 20, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long h[3]; cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
        printf("grid %3d  %2d copies x %5d B: best %6lld  avg %6lld cycles (issue %4lld)  -> %6.1f B/clk\n", grid, n, bytes, h[0], h[1], h[2], (double)n * bytes / h[1]);
      }
  return 0;
}
