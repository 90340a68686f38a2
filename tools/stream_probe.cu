// Standalone probe: how fast can every SM stream an L2-resident buffer into shared memory with 1-D TMA
// bulk copies through an mbarrier ring?  (the weight / spill streaming of jet_tc.cu).  Sweeps ring depth and
// stage size, with all CTAs reading the SAME 3.6 MB buffer (weights) or CTA-private buffers (spills), and a
// 2-CTA-cluster variant where each CTA loads half of every stage and multicasts it to both.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_1d_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

// mode 0: all CTAs read the same buffer; mode 1: CTA-private buffers; mode 2: same buffer, cluster of 2 with multicast
__global__ void __launch_bounds__(320, 1) stream_kernel(const float* src, size_t src_floats_per_cta, int mode, int stages, int stage_bytes,
                                                        int chunks, long long* cycles, float* sink, int split, int splitw) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  unsigned char* ring = smem + 1024;
  const int tid = threadIdx.x;
  uint32_t rank = 0;
  if (mode == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], mode == 2 ? 1 : split); mbar_init(&empty[s], mode == 2 ? 2 : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (mode == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const float* base = src + (mode == 1 ? (size_t)blockIdx.x * src_floats_per_cta : 0);
  const size_t wrap = src_floats_per_cta * 4 / stage_bytes;   // chunks before wrapping around the buffer
  const long long t0 = clock64();
  float acc = 0.f;
  const int pid = splitw ? (tid >> 5) - 2 : tid;     // producer index: lanes of warp 0, or warps 2..
  const bool is_prod = splitw ? ((tid & 31) == 0 && tid >= 64 && pid < split) : (tid < split);
  if (is_prod && mode != 2) {           // producers: each issues 1/split of every stage
    const int pb = stage_bytes / split;
    for (int c = 0; c < chunks; ++c) {
      const int s = c % stages;
      mbar_wait(&empty[s], (uint32_t)(((c / stages) & 1) ^ 1));
      mbar_expect_tx(&full[s], pb);
      const char* g = reinterpret_cast<const char*>(base) + (size_t)(c % wrap) * stage_bytes + (size_t)pid * pb;
      tma_load_1d(ring + (size_t)s * stage_bytes + (size_t)pid * pb, g, pb, &full[s]);
    }
  } else if (tid == 0 && mode == 2) {
    for (int c = 0; c < chunks; ++c) {
      const int s = c % stages;
      mbar_wait(&empty[s], (uint32_t)(((c / stages) & 1) ^ 1));
      mbar_expect_tx(&full[s], stage_bytes);
      const char* g = reinterpret_cast<const char*>(base) + (size_t)(c % wrap) * stage_bytes;
      const int hb = stage_bytes / 2;
      tma_load_1d_mc(ring + (size_t)s * stage_bytes + rank * hb, g + rank * hb, hb, &full[s], (uint16_t)3);
    }
  } else if (tid == 32) {   // consumer: touch one word, release the stage
    for (int c = 0; c < chunks; ++c) {
      const int s = c % stages;
      mbar_wait(&full[s], (uint32_t)((c / stages) & 1));
      acc += *reinterpret_cast<const float*>(ring + (size_t)s * stage_bytes);
      if (mode == 2) { mbar_arrive_remote(&empty[s], 0); mbar_arrive_remote(&empty[s], 1); }
      else mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (mode == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (tid == 32) { cycles[blockIdx.x] = clock64() - t0; if (acc == 123.456f) *sink = acc; }
}

int main() {
  const size_t wbytes = 7ull * 2 * 256 * 256 * 4;            // the packed weight images: 3.67 MB
  const size_t pbytes = 1310720;                              // per-CTA spill slab: 10 images x 128 KB
  float *W, *P, *sink; long long* cyc;
  cudaMalloc(&W, wbytes); cudaMalloc(&P, pbytes * 148); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  cudaMemset(W, 0, wbytes); cudaMemset(P, 0, pbytes * 148);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 196608);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  struct Cfg { int mode, stages, stage_bytes, split, splitw; };
  Cfg cfgs[] = {{0, 4, 16384, 1, 0}, {0, 4, 16384, 2, 0}, {0, 4, 16384, 4, 0}, {0, 4, 16384, 8, 0}, {0, 4, 16384, 2, 1}, {0, 4, 16384, 4, 1},
                {0, 8, 16384, 4, 0}, {0, 8, 16384, 4, 1}, {0, 2, 65536, 1, 0}, {0, 3, 65536, 1, 0}, {0, 3, 65536, 4, 0}, {0, 4, 32768, 4, 1},
                {1, 4, 16384, 4, 0}, {1, 4, 16384, 4, 1}, {1, 4, 32768, 1, 0}, {1, 8, 16384, 8, 1}};
  const char* mname[] = {"shared 3.6MB buffer", "CTA-private 1.25MB ", "shared, 2-CTA mcast "};
  for (auto& c : cfgs) {
    const size_t total = 64ull << 20;   // 64 MB per CTA
    const int chunks = (int)(total / c.stage_bytes);
    const size_t smem = 1024 + (size_t)c.stages * c.stage_bytes;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(148); lc.blockDim = dim3(320); lc.dynamicSmemBytes = smem; lc.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.mode == 2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    const float* src = c.mode == 1 ? P : W;
    const size_t per = (c.mode == 1 ? pbytes : wbytes) / 4;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cudaError_t le = cudaLaunchKernelEx(&lc, stream_kernel, src, per, c.mode, c.stages, c.stage_bytes, chunks, cyc, sink, c.split, c.splitw);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (le != cudaSuccess || e != cudaSuccess) { printf("mode %d: CUDA error %s / %s\n", c.mode, cudaGetErrorString(le), cudaGetErrorString(e)); return 1; }
      if (rep == 0) continue;
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> cy(148); cudaMemcpy(cy.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (auto x : cy) mx = x > mx ? x : mx;
      printf("%s  %2d stages x %5d B, %d copies/stage by %s (%3d KB in flight): %7.2f B/clk/SM   %8.1f GB/s chip  (%.2f ms)\n", mname[c.mode], c.stages, c.stage_bytes,
             c.split, c.splitw ? "warps" : "lanes", c.stages * c.stage_bytes / 1024, (double)total / (double)mx, 148.0 * total / ms * 1e-6, ms);
    }
  }
  return 0;
}
