// Standalone probe: how does tcgen05.mma kind::tf32 round its FP32 accumulator?  D[128 x 64] = A[128 x K] * B[64 x K]^T
// accumulated over K/8 MMAs in TMEM, compared bit for bit with two host models of the per-instruction update
//   acc <- round(acc + exact sum of the 8 products):   RN (nearest-even)   and   RZ (toward zero)
// and with the exactly rounded result.  Decides how the split-operand (3xTF32) mode has to treat the accumulator.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((a >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
constexpr int M = 128, N = 64, KMAX = 256;
constexpr int A_BYTES = (KMAX / 4) * M * 16, B_BYTES = (KMAX / 4) * N * 16;

__global__ void probe(const float* Aimg, const float* Bimg, float* Dout, int nk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_s = smem;
  unsigned char* b_s = smem + A_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < A_BYTES / 4; i += blockDim.x) ((float*)a_s)[i] = Aimg[i];
  for (int i = tid; i < B_BYTES / 4; i += blockDim.x) ((float*)b_s)[i] = Bimg[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tptr;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < nk; ++ks) {
      const uint64_t ad = umma_desc(smem_u32(a_s) + ks * 2 * M * 16, M * 16, 128);
      const uint64_t bd = umma_desc(smem_u32(b_s) + ks * 2 * N * 16, N * 16, 128);
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
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tb + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) Dout[row * N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64u) : "memory");
}

static float tf32_rna(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x1000u; u &= 0xffffe000u;
  float y; memcpy(&y, &u, 4); return y;
}
static float round_rz(double x) {   // double -> float toward zero
  float f = (float)x;               // RN
  if (std::fabs((double)f) > std::fabs(x)) f = std::nextafterf(f, 0.f);
  return f;
}
static double ulp_of(double x) { int e; std::frexp(x, &e); return std::ldexp(1.0, e - 24); }

int main() {
  const size_t smem = A_BYTES + B_BYTES + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float *dA, *dB, *dD; cudaMalloc(&dA, A_BYTES); cudaMalloc(&dB, B_BYTES); cudaMalloc(&dD, M * N * 4);
  struct Pat { const char* name; int nk; int mode; };
  // mode 0: values in [1,2) (monotone accumulation); 1: signed uniform (-1,1) (random walk); 2: hi part positive [1,2) then the SAME rows
  // against a "lo" operand 2^-12 smaller in the second half of K (what the split-operand kernel does to its accumulator)
  Pat pats[] = {{"K=8   positive [1,2)", 1, 0}, {"K=8   signed (-1,1)", 1, 1}, {"K=16  positive", 2, 0}, {"K=256 positive [1,2)", 32, 0},
                {"K=256 signed (-1,1)", 32, 1}, {"K=256 signed, 2nd half of K scaled 2^-12", 32, 2}};
  srand(1234);
  for (auto& pt : pats) {
    const int K = 8 * pt.nk;
    std::vector<float> A(M * KMAX, 0.f), B(N * KMAX, 0.f), Aimg(A_BYTES / 4, 0.f), Bimg(B_BYTES / 4, 0.f);
    auto rnd = [&]() { return (float)rand() / (float)RAND_MAX; };
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
      float v = pt.mode == 0 ? 1.f + rnd() : 2.f * rnd() - 1.f;
      A[m * KMAX + k] = tf32_rna(v);
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
      float v = pt.mode == 0 ? 1.f + rnd() : 2.f * rnd() - 1.f;
      if (pt.mode == 2 && k >= K / 2) v *= (1.f / 4096.f);
      B[n * KMAX + k] = tf32_rna(v);
    }
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Aimg[(k / 4) * M * 4 + m * 4 + (k % 4)] = A[m * KMAX + k];
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bimg[(k / 4) * N * 4 + n * 4 + (k % 4)] = B[n * KMAX + k];
    cudaMemcpy(dA, Aimg.data(), A_BYTES, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bimg.data(), B_BYTES, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * N * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, pt.nk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", pt.name, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(M * N); cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    long eq_rn = 0, eq_rz = 0, eq_exact_rn = 0, eq_exact_rz = 0, below = 0, above = 0;
    double mean_ulps = 0, mean_abs_ulps = 0, mean_rel = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double exact = 0; float a_rn = 0.f, a_rz = 0.f;
      for (int ks = 0; ks < pt.nk; ++ks) {
        double s8 = 0;
        for (int k = 8 * ks; k < 8 * ks + 8; ++k) s8 += (double)A[m * KMAX + k] * (double)B[n * KMAX + k];   // exact in double
        exact += s8;
        a_rn = (float)((double)a_rn + s8);
        a_rz = round_rz((double)a_rz + s8);
      }
      const float d = D[m * N + n];
      eq_rn += d == a_rn; eq_rz += d == a_rz; eq_exact_rn += d == (float)exact; eq_exact_rz += d == round_rz(exact);
      const double du = ((double)d - exact) / ulp_of(exact);
      mean_ulps += du; mean_abs_ulps += std::fabs(du); mean_rel += ((double)d - exact) / std::fabs(exact);
      below += std::fabs((double)d) < std::fabs(exact); above += std::fabs((double)d) > std::fabs(exact);
    }
    const double T = (double)M * N;
    printf("%-44s match: RN-step %.3f  RZ-step %.3f  RN(exact) %.3f  RZ(exact) %.3f | |D|<|exact| %.3f  |D|>|exact| %.3f | "
           "mean err %+.3f ulp, mean |err| %.3f ulp, mean rel %+.3e\n",
           pt.name, eq_rn / T, eq_rz / T, eq_exact_rn / T, eq_exact_rz / T, below / T, above / T, mean_ulps / T, mean_abs_ulps / T, mean_rel / T);
  }
  return 0;
}
