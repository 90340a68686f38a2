// Probe: can a 1-D cp.async.bulk issued by CTA 1 of a cluster (destination = its OWN shared memory) complete its
// transaction bytes on an mbarrier in CTA 0?  (would remove the follower -> leader "stage has landed" relay of jet_tc.cu)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__global__ void __cluster_dims__(2, 1, 1) probe(const float* src, int* out, long long* cyc) {
  __shared__ __align__(128) float buf[2048];
  __shared__ uint64_t bar;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = -1.f;
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const long long t0 = clock64();
  if (rank == 1 && threadIdx.x == 0) {
    const uint32_t rbar = mapa(smem_u32(&bar), 0);          // the LEADER's barrier
    const uint32_t dst = mapa(smem_u32(buf), 1);            // my own buffer, as a shared::cluster address
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(rbar), "r"(8192u) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(8192u), "r"(rbar) : "memory");
  }
  if (rank == 0 && threadIdx.x == 0) {
    uint32_t done = 0; long long spins = 0;
    while (!done && spins < (1ll << 22)) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      ++spins;
    }
    cyc[0] = clock64() - t0;
    // read the follower's buffer through DSMEM
    const uint32_t rb = mapa(smem_u32(buf), 1);
    float v0, v1;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v0) : "r"(rb));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v1) : "r"(rb + 8188u));
    out[0] = (int)done; out[1] = (int)v0; out[2] = (int)v1;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  float* src; int* out; long long* cyc;
  cudaMalloc(&src, 8192); cudaMalloc(&out, 16); cudaMalloc(&cyc, 8);
  float h[2048]; for (int i = 0; i < 2048; ++i) h[i] = (float)(i + 1);
  cudaMemcpy(src, h, 8192, cudaMemcpyHostToDevice); cudaMemset(out, 0, 16);
  probe<<<2, 64>>>(src, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  int o[3]; long long c; cudaMemcpy(o, out, 12, cudaMemcpyDeviceToHost); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("status %s; leader saw completion: %d after %lld cycles; follower buffer [0] = %d (expect 1), [2047] = %d (expect 2048)\n", cudaGetErrorString(e), o[0], c, o[1], o[2]);
  return 0;
}
