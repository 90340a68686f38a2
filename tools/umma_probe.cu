// Standalone probe: which smem-descriptor encoding makes tcgen05.mma kind::tf32 read my operand
// images MN-major (contraction over rows)?  Runs D[128 x 256] = sum_m Z[m][nf] * Ain[m][kf] for several
// (LBO, SBO, major) variants and prints the error against the CPU result.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define OP_LBO 2064
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a, uint32_t lbo, uint32_t sbo, uint32_t extra_hi) {
  uint64_t d = 0;
  d |= (uint64_t)((a >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)extra_hi << 32;
  return d;
}
struct Variant { uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_mn, b_mn, a_kstep, b_kstep, extra_hi; int nk; int img; };

__global__ void probe(const float* Aimg, const float* Bimg, float* Dout, Variant v) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_s = smem;                 // 132 KB region (operand image, 128 rows x 256 features)
  unsigned char* b_s = smem + 133120;        // 16 KB chunk image, 1024-aligned
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 133120 / 4; i += blockDim.x) ((float*)a_s)[i] = Aimg[i];
  for (int i = tid; i < 4096; i += blockDim.x) ((float*)b_s)[i] = Bimg[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tptr;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (v.a_mn << 15) | (v.b_mn << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < v.nk; ++ks) {
      const uint64_t ad = umma_desc(smem_u32(a_s) + ks * v.a_kstep, v.a_lbo, v.a_sbo, v.extra_hi);
      const uint64_t bd = umma_desc(smem_u32(b_s) + ks * v.b_kstep, v.b_lbo, v.b_sbo, v.extra_hi);
      const uint32_t acc = ks > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  } while (!done);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < 256; c0 += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tb + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) Dout[row * 256 + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

int main() {
  // Z[m][f] (operand image, 128 rows x 256 features) and Ain chunk [16 rows][256 features]
  std::vector<float> Z(128 * 256), Ain(16 * 256);
  for (int m = 0; m < 128; ++m) for (int f = 0; f < 256; ++f) Z[m * 256 + f] = (float)(((m * 7 + f * 3) % 11) - 5);
  for (int m = 0; m < 16; ++m) for (int f = 0; f < 256; ++f) Ain[m * 256 + f] = (float)(((m * 5 + f * 13) % 7) - 3);
  // img 0: interleaved images (as in jet_tc.cu); img 1: K-major sanity (B = W chunk [k/4][n][4], W[n][k] = Ain-like);
  // img 2: SWIZZLE_128B blocks [fb][rg][r][128 B], chunk16 ^= r
  std::vector<float> Aimg[3], Bimg[3];
  for (int t = 0; t < 3; ++t) { Aimg[t].assign(133120 / 4, 0.f); Bimg[t].assign(4096, 0.f); }
  std::vector<float> Wk(256 * 16);
  for (int n = 0; n < 256; ++n) for (int k = 0; k < 16; ++k) Wk[n * 16 + k] = (float)(((n * 3 + k * 5) % 9) - 4);
  for (int m = 0; m < 128; ++m) for (int f = 0; f < 256; ++f) {
    Aimg[0][((f / 4) * OP_LBO + m * 16 + (f % 4) * 4) / 4] = Z[m * 256 + f];
    Aimg[1][((f / 4) * OP_LBO + m * 16 + (f % 4) * 4) / 4] = Z[m * 256 + f];
    Aimg[2][((f / 32) * 16384 + (m / 8) * 1024 + (m % 8) * 128 + ((((f % 32) / 4) ^ (m % 8)) * 16) + (f % 4) * 4) / 4] = Z[m * 256 + f];
  }
  for (int m = 0; m < 16; ++m) for (int f = 0; f < 256; ++f) {
    Bimg[0][(f / 4) * 64 + m * 4 + (f % 4)] = Ain[m * 256 + f];
    Bimg[2][((f / 32) * 2048 + (m / 8) * 1024 + (m % 8) * 128 + ((((f % 32) / 4) ^ (m % 8)) * 16) + (f % 4) * 4) / 4] = Ain[m * 256 + f];
  }
  for (int n = 0; n < 256; ++n) for (int k = 0; k < 16; ++k) Bimg[1][(k / 4) * 1024 + n * 4 + (k % 4)] = Wk[n * 16 + k];
  // expected MN-major result over the 16 rows of the chunk: D[nf][kf] = sum_{m<16} Z[m][nf] * Ain[m][kf], nf < 128
  std::vector<float> Emn(128 * 256, 0.f), Ek(128 * 256, 0.f);
  for (int nf = 0; nf < 128; ++nf) for (int kf = 0; kf < 256; ++kf) { float s = 0; for (int m = 0; m < 16; ++m) s += Z[m * 256 + nf] * Ain[m * 256 + kf]; Emn[nf * 256 + kf] = s; }
  float *dA, *dB, *dD; cudaMalloc(&dA, 133120); cudaMalloc(&dB, 16384); cudaMalloc(&dD, 128 * 256 * 4);
  const size_t smem = 133120 + 16384 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Named { const char* name; Variant v; };
  Named vs[] = {
      {"K-major sanity (interleaved)", {OP_LBO, 128, 4096, 128, 0, 0, 2 * OP_LBO, 2 * 4096, 0, 2, 1}},
      {"MN interleaved: A(lbo128,sbo2064) B(lbo128,sbo256)", {128, OP_LBO, 128, 256, 1, 1, 128, 128, 0, 2, 0}},
      {"MN SW128: A(lbo16384,sbo1024) B(lbo2048,sbo1024)", {16384, 1024, 2048, 1024, 1, 1, 1024, 1024, 2u << 29, 2, 2}},
      {"MN SW128 swapped lbo/sbo", {1024, 16384, 1024, 2048, 1, 1, 1024, 1024, 2u << 29, 2, 2}},
      {"MN SW128 single k-step", {16384, 1024, 2048, 1024, 1, 1, 1024, 1024, 2u << 29, 1, 2}},
  };
  for (auto& nv : vs) {
    cudaMemcpy(dA, Aimg[nv.v.img].data(), 133120, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bimg[nv.v.img].data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * 256 * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, nv.v);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-60s CUDA error %s\n", nv.name, cudaGetErrorString(e)); return 1; }
    std::vector<float> Dh(128 * 256); cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost);
    // compare with the nk-step expectation
    double err = 0, ref = 0, sumabs = 0; int nk = nv.v.nk;
    for (int nf = 0; nf < 128; ++nf) for (int kf = 0; kf < 256; ++kf) {
      float s = 0;
      if (nv.v.img == 1) { for (int k = 0; k < 8 * nk; ++k) s += Z[nf * 256 + k] * Wk[kf * 16 + k]; }
      else { for (int m = 0; m < 8 * nk; ++m) s += Z[m * 256 + nf] * Ain[m * 256 + kf]; }
      double dlt = Dh[nf * 256 + kf] - s; err += dlt * dlt; ref += (double)s * s; sumabs += fabs(Dh[nf * 256 + kf]);
    }
    printf("%-60s rel err %.3e   sum|D| %.3e   D[0][0..3] %g %g %g %g\n", nv.name, sqrt(err / ref), sumabs, Dh[0], Dh[1], Dh[2], Dh[3]);
  }
  return 0;
}
