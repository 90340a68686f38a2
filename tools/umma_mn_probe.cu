// Standalone probe: tcgen05.mma kind::tf32 with MN-major operands in the SWIZZLE_128B_BASE32B layout
// (the only MN-major shared-memory layout CUTLASS offers for 32-bit operands on sm_100).
// Computes D[128 x 256] = sum_{m < 16} Z[m][nf] * Ain[m][kf] (contraction over ROWS) from two images stored
// "row-major by 32-feature panels":  (row m, feature f) at byte
//     (f/32)*PS + m*128 + ((((f%32)/8) ^ (m%4)) * 32) + (f%8)*4 ,   PS = rows_in_chunk * 128
// and prints the error for several (LBO, SBO) assignments.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((a >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
struct Variant { uint32_t a_lbo, a_sbo, b_lbo, b_sbo, kstep, layout_type, a_mn, b_mn; int nk; };

__global__ void probe(const float* Aimg, const float* Bimg, float* Dout, Variant v) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_s = smem;            // 16 KB: Z chunk (16 rows x 256 features)
  unsigned char* b_s = smem + 16384;    // 16 KB: Ain chunk
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4096; i += blockDim.x) ((float*)a_s)[i] = Aimg[i];
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
      const uint64_t ad = umma_desc(smem_u32(a_s) + ks * v.kstep, v.a_lbo, v.a_sbo, v.layout_type);
      const uint64_t bd = umma_desc(smem_u32(b_s) + ks * v.kstep, v.b_lbo, v.b_sbo, v.layout_type);
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
  const int R = 16;  // rows in the chunk
  std::vector<float> Z(R * 256), Ain(R * 256);
  for (int m = 0; m < R; ++m) for (int f = 0; f < 256; ++f) Z[m * 256 + f] = (float)(((m * 7 + f * 3) % 11) - 5);
  for (int m = 0; m < R; ++m) for (int f = 0; f < 256; ++f) Ain[m * 256 + f] = (float)(((m * 5 + f * 13) % 7) - 3);
  // image kinds: 0 = swizzled by (m%4) in 32-byte units; 1 = unswizzled panels; 2 = swizzled by (m%8)>>1 ; 3 = 16B-unit swizzle by m%8 (plain SW128)
  auto build = [&](const std::vector<float>& src, int kind) {
    std::vector<float> img(4096, 0.f);
    const int PS = R * 128;
    for (int m = 0; m < R; ++m) for (int f = 0; f < 256; ++f) {
      int off;
      const int c32 = (f % 32) / 8, c16 = (f % 32) / 4;
      if (kind == 0) off = (f / 32) * PS + m * 128 + ((c32 ^ (m % 4)) * 32) + (f % 8) * 4;
      else if (kind == 1) off = (f / 32) * PS + m * 128 + (f % 32) * 4;
      else if (kind == 2) off = (f / 32) * PS + m * 128 + ((c32 ^ ((m % 8) >> 1)) * 32) + (f % 8) * 4;
      else off = (f / 32) * PS + m * 128 + ((c16 ^ (m % 8)) * 16) + (f % 4) * 4;
      img[off / 4] = src[m * 256 + f];
    }
    return img;
  };
  float *dA, *dB, *dD; cudaMalloc(&dA, 16384); cudaMalloc(&dB, 16384); cudaMalloc(&dD, 128 * 256 * 4);
  const size_t smem = 32768 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Named { const char* name; int kind; Variant v; };
  const uint32_t PS = R * 128;
  Named vs[] = {
      {"BASE32B img0  LBO=panel SBO=512  kstep=1024", 0, {PS, 512, PS, 512, 1024, 1, 1, 1, 2}},
      {"BASE32B img0  LBO=512 SBO=panel  kstep=1024", 0, {512, PS, 512, PS, 1024, 1, 1, 1, 2}},
      {"BASE32B img0  LBO=panel SBO=1024 kstep=1024", 0, {PS, 1024, PS, 1024, 1024, 1, 1, 1, 2}},
      {"BASE32B img2  LBO=panel SBO=1024 kstep=1024", 2, {PS, 1024, PS, 1024, 1024, 1, 1, 1, 2}},
      {"BASE32B img2  LBO=panel SBO=512  kstep=1024", 2, {PS, 512, PS, 512, 1024, 1, 1, 1, 2}},
      {"BASE32B img1  LBO=panel SBO=512  kstep=1024", 1, {PS, 512, PS, 512, 1024, 1, 1, 1, 2}},
      {"BASE32B img0  one k-step LBO=panel SBO=512", 0, {PS, 512, PS, 512, 1024, 1, 1, 1, 1}},
      {"SW128   img3  LBO=panel SBO=1024 kstep=1024", 3, {PS, 1024, PS, 1024, 1024, 2, 1, 1, 2}},
      {"SW128   img3  LBO=1024 SBO=panel kstep=1024", 3, {1024, PS, 1024, PS, 1024, 2, 1, 1, 2}},
      {"NONE    img1  LBO=panel SBO=512  kstep=1024", 1, {PS, 512, PS, 512, 1024, 0, 1, 1, 2}},
  };
  for (auto& nv : vs) {
    auto Ai = build(Z, nv.kind), Bi = build(Ain, nv.kind);
    cudaMemcpy(dA, Ai.data(), 16384, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bi.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * 256 * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, nv.v);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-50s CUDA error %s\n", nv.name, cudaGetErrorString(e)); return 1; }
    std::vector<float> Dh(128 * 256); cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, ref = 0, sumabs = 0; const int nk = nv.v.nk;
    for (int nf = 0; nf < 128; ++nf) for (int kf = 0; kf < 256; ++kf) {
      float s = 0;
      for (int m = 0; m < 8 * nk; ++m) s += Z[m * 256 + nf] * Ain[m * 256 + kf];
      const double dlt = Dh[nf * 256 + kf] - s; err += dlt * dlt; ref += (double)s * s; sumabs += fabs(Dh[nf * 256 + kf]);
    }
    printf("%-50s rel err %.3e   sum|D| %.3e   D[0][0..3] %g %g %g %g   D[1][0] %g D[33][40] %g\n", nv.name, sqrt(err / ref), sumabs,
           Dh[0], Dh[1], Dh[2], Dh[3], Dh[256], Dh[33 * 256 + 40]);
  }
  // reference values for eyeballing
  {
    float s00 = 0, s01 = 0, s10 = 0, s3340 = 0;
    for (int m = 0; m < 16; ++m) { s00 += Z[m * 256] * Ain[m * 256]; s01 += Z[m * 256] * Ain[m * 256 + 1]; s10 += Z[m * 256 + 1] * Ain[m * 256]; s3340 += Z[m * 256 + 33] * Ain[m * 256 + 40]; }
    printf("expected (16 rows): D[0][0] %g D[0][1] %g D[1][0] %g D[33][40] %g\n", s00, s01, s10, s3340);
  }
  return 0;
}
