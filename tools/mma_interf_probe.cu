// Probe: tcgen05.mma.cta_group::2.kind::tf32 (M = 256 pair, N = 256, K = 8, K-major operands resident in shared memory) rate
// under interference: (t) a TMA warp streaming 16 KB bulk copies into a 5-slot ring, (s) 16 warps parked in mbarrier
// try_wait, (g) 16 warps streaming st.global.v4, (l) 16 warps doing ld/st.shared.v4 on a scratch image, (c) commit every 4.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1)
rate(int n_mma, int f_tma, int f_spin, int f_glob, int f_lds, int commit_every, int f_poll, int cp_bytes, int n_slots, int n_prod, int lanes_mode, const float* src, float* dst, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, done_bar, cbar, pre, full[16];
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = tid; i < 212992 / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i & 255);
  if (tid == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&pre)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&pre)) : "memory");   // phase 0 complete
    for (int s = 0; s < 16; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
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
  // smem map: A image [0, 135168) LBO 2112; B half [135168, 135168+32768) ; TMA ring [168960, +5*8192) (8 KB slots here)
  if (warp == 0) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t a0 = smem_u32(smem), b0 = a0 + 135168;
      const uint64_t ad0 = desc(a0, 2112, 128, 0), bd0 = desc(b0, 2048, 128, 0);
      const uint32_t astep = 2 * 2112 / 16, bstep = 2 * 2048 / 16;
      const long long t0 = clock64();
      for (int i0 = 0; i0 < n_mma; i0 += 32) {
#pragma unroll
        for (int ks = 0; ks < 32; ++ks) {
          if (f_poll && (ks & 3) == 0) { while (!try_wait(&pre, 0)) {} while (!try_wait(&pre, 0)) {} }
          const uint64_t ad = ad0 + (uint64_t)((ks & 31) * astep);
          const uint64_t bd = bd0 + (uint64_t)((ks & 3) * bstep);
          if (i0 == 0 && ks == 0)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc) : "memory");
          else
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc) : "memory");
          if ((ks & (commit_every - 1)) == commit_every - 1)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&cbar)), "h"((uint16_t)3) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      while (!try_wait(&bar, 0)) {}
      out[blockIdx.x >> 1] = clock64() - t0;
    }
    if (lane == 0) {
      if (rank != 0) while (!try_wait(&bar, 0)) {}
      stop = 1;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    }
  } else if (warp >= 1 && warp <= 3) {
    const int pid = lanes_mode ? lane : warp - 1;
    if ((lanes_mode ? (warp == 1 && lane < n_prod) : (lane == 0 && warp - 1 < n_prod)) && f_tma) {
      unsigned char* ring = smem + 168960;
      int i = 0;
      for (; !stop; ++i) {
        const int spp = n_slots / n_prod;                 // slots per producer
        const int s = pid * spp + i % spp;
        if (i >= spp) while (!try_wait(&full[s], (uint32_t)(((i / spp) - 1) & 1))) {}
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)cp_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + s * cp_bytes)),
                     "l"(src + (size_t)((i * 37 + blockIdx.x * 101) & 4095) * 2048), "r"((uint32_t)cp_bytes), "r"(smem_u32(&full[s])) : "memory");
      }
      if (pid == 0) out[128 + blockIdx.x] = (long long)i * n_prod;
      // drain outstanding copies before exit
      for (int s = pid * (n_slots / n_prod); s < (pid + 1) * (n_slots / n_prod); ++s) { int spins = 0; while (!try_wait(&full[s], 0) && !try_wait(&full[s], 1) && ++spins < 100000) {} }
      __nanosleep(20000);
    }
  } else if (warp >= 4) {
    if (f_spin) {
      while (!try_wait(&done_bar, 0)) {}
    } else if (f_glob) {
      float* p = dst + ((size_t)blockIdx.x * 512 + (tid - 128)) * 4;
      size_t off = 0;
      while (!stop) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p + off), "f"(1.0f) : "memory");
          off = (off + (size_t)148 * 512 * 4) & ((size_t)(1u << 27) - 1);
        }
      }
    } else if (f_lds) {
      float4* q = reinterpret_cast<float4*>(smem + 209920 - 8192 * 0) ;   // 3 KB scratch at the top: [192 float4]
      float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
      while (!stop) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          q[(tid + 32 * u) % 192] = v;
          const float4 w = q[(tid + 32 * u + 64) % 192];
          v.x += w.y;
        }
      }
      if (v.x == 12345.f) dst[0] = v.x;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 256 * 8);
  float *src, *dst; cudaMalloc(&src, (size_t)4096 * 2048 * 4 + 65536); cudaMalloc(&dst, (size_t)(1u << 27) * 4 + (1 << 24)); cudaMemset(src, 0, (size_t)4096 * 2048 * 4);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 212992);
  struct C { int tma, spin, glob, lds, ce; const char* name; int poll; int cpb = 8192, ns = 5, np = 1, lm = 0; } cs[] = {
      {1, 0, 0, 0, 4, "TMA 2 KB x 16, 2 producer warps", 0, 2048, 16, 2, 0}, {1, 0, 0, 0, 4, "TMA 2 KB x 16, 3 producer warps", 0, 2048, 15, 3, 0}, {1, 0, 0, 0, 4, "TMA 2 KB x 16, 4 producer lanes", 0, 2048, 16, 4, 1}, {1, 0, 1, 0, 4, "TMA 2 KB x 16, 2 producer warps + st.global", 0, 2048, 16, 2, 0}, {1, 0, 1, 0, 4, "TMA 2 KB x 16, 4 producer lanes + st.global", 0, 2048, 16, 4, 1},
      {1, 0, 0, 0, 4, "TMA 1 KB x 16 slots, quiet", 0, 1024, 16}, {1, 0, 1, 0, 4, "TMA 1 KB x 16 slots + st.global warps", 0, 1024, 16}, {1, 0, 0, 0, 4, "TMA 8 KB x 5", 0, 8192, 5}, {1, 0, 1, 0, 4, "TMA 8 KB x 5 + st.global warps", 0, 8192, 5}, {1, 0, 0, 0, 4, "TMA 2 KB x 16", 0, 2048, 16}, {1, 0, 1, 0, 4, "TMA 2 KB x 16 + st.global", 0, 2048, 16},
      {0, 0, 0, 0, 4, "polls (2 per 4 MMAs), quiet", 1}, {1, 1, 0, 0, 4, "polls + TMA + try_wait warps", 1}, {1, 0, 1, 0, 4, "polls + TMA + st.global warps", 1}, {1, 0, 0, 1, 4, "polls + TMA + ld/st.shared warps", 1},
      {0, 0, 0, 0, 1024, "quiet, one commit"}, {0, 0, 0, 0, 4, "commit every 4"}, {1, 0, 0, 0, 4, "TMA ring + commit/4"},
      {0, 1, 0, 0, 4, "16 warps in try_wait + commit/4"}, {1, 1, 0, 0, 4, "TMA + try_wait + commit/4"},
      {0, 0, 1, 0, 4, "16 warps st.global.v4 + commit/4"}, {1, 0, 1, 0, 4, "TMA + st.global + commit/4"},
      {0, 0, 0, 1, 4, "16 warps ld/st.shared.v4 + commit/4"}, {1, 0, 0, 1, 4, "TMA + ld/st.shared + commit/4"}};
  for (auto& c : cs) {
    rate<<<148, 640, 212992>>>(2048, c.tma, c.spin, c.glob, c.lds, c.ce, c.poll, c.cpb, c.ns, c.np, c.lm, src, dst, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s (%s)\n", cudaGetErrorString(e), c.name); return 1; }
    long long h[74]; cudaMemcpy(h, out, 74 * 8, cudaMemcpyDeviceToHost); long long cp[2]; cudaMemcpy(cp, out + 128, 16, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 74; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-40s 2048 MMAs: %8lld cycles = %6.1f cycles/MMA", c.name, mx, (double)mx / 2048);
    if (c.tma) printf("   TMA: %lld copies of %d B by CTA 0 = %.0f cycles per copy, %.1f B/clk", cp[0], c.cpb, (double)h[0] / cp[0], (double)cp[0] * c.cpb / h[0]);
    printf("\n");
    fflush(stdout);
  }
  return 0;
}
