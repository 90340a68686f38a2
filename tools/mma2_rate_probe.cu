// Probe: issue rate of tcgen05.mma.cta_group::2.kind::tf32 (M = 256 over a CTA pair, N = 256, K = 8) with all operands
// resident in shared memory: K-major A (no swizzle) x K-major B half, and MN-major x MN-major (SWIZZLE_128B_BASE32B).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate(int N, int mn, int n_mma, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = tid; i < 196608 / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i & 255);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tptr;
  if (tid == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)mn << 15) | ((uint32_t)mn << 16) | (((uint32_t)N >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 135168;
    uint64_t ad0, bd0; uint32_t astep, bstep, amask = 31, bmask = 3;
    if (mn) { ad0 = desc(b0, 2048, 512, 1); astep = 1024 / 16; amask = 1; bd0 = desc(b0 + 8192, 2048, 512, 1); bstep = 1024 / 16; bmask = 1; }
    else { ad0 = desc(a0, 2112, 128, 0); astep = 2 * 2112 / 16; bd0 = desc(b0, 2048, 128, 0); bstep = 2 * 2048 / 16; }
    const long long t0 = clock64();
    for (int i0 = 0; i0 < n_mma; i0 += 32) {
#pragma unroll
      for (int ks = 0; ks < 32; ++ks) {
        const uint64_t ad = ad0 + (uint64_t)((ks & amask) * astep);
        const uint64_t bd = bd0 + (uint64_t)((ks & bmask) * bstep);
        if (i0 == 0 && ks == 0)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    uint32_t done;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!done);
    out[blockIdx.x >> 1] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 256 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200704);
  struct C { int N, mn, n; } cs[] = {{256, 0, 1024}, {128, 0, 1024}, {256, 1, 1024}, {256, 0, 32}};
  for (int grid : {2, 148})
    for (auto& c : cs) {
      rate<<<grid, 128, 200704>>>(c.N, c.mn, c.n, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[74]; cudaMemcpy(h, out, (grid / 2) * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid / 2; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("grid %3d  cta_group::2 M=256 N=%3d %s  %4d MMAs: %7lld cycles = %6.1f cycles/MMA  (%.0f TFLOP/s chip-equivalent at 1.965 GHz)\n", grid, c.N,
             c.mn ? "MN-major" : "K-major ", c.n, mx, (double)mx / c.n, 2.0 * 256 * c.N * 8 * c.n / (double)mx * 1.965e9 * 74 * 1e-12);
      fflush(stdout);
    }
  return 0;
}
