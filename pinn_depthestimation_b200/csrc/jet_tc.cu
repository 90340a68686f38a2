// Tensor-core (tcgen05 / TMEM) jet-MLP kernel for sm_100a, TF32 operands with FP32 accumulation.
//
// Same contract as the FP32 kernel (jet_fp32.cu): per tile of collocation points, forward value +
// tangent jets through every Linear+tanh layer, fused PDE-residual epilogue, reverse sweep to the
// flat weight gradient.  Here the 256x256 hidden layers -- >99 % of the FLOPs -- run on the 5th-gen
// tensor cores:
//
//   tile            32 points x 4 jet rows = 128 rows (row m = 4*point + jet)  -> UMMA M = 128
//   forward  l      D[128 x 256] = A[128 x 256] * W_l^T          A: smem (K-major), B: W image via TMA
//   adjoint  l      D[128 x 256] = Zbar[128 x 256] * W_l         A: smem (K-major), B: W^T image via TMA
//   weight grad l   D[256 x 256] = Zbar^T * A_in  (two M=128 halves, K = 128 rows)
//                                                                A, B: row-transposed images via TMA (K-major)
//   accumulators live in TMEM (512 columns), read back with tcgen05.ld for the tanh / jet / adjoint
//   epilogues; the d->256 and 256->o edge layers (<1 % of the work) stay on the FP32 pipes.
//
// Warp roles (320 threads, one CTA per SM, persistent over tiles):
//   warps 0-7  workers: edge layers, epilogues (TMEM -> registers -> operand image in smem / slab / RED)
//   warp  8    producer: 1-D TMA bulk copies of weight images and slab chunks into a 4-stage ring
//   warp  9    MMA issuer: one thread issues tcgen05.mma and tcgen05.commit; owns the TMEM allocation
//
// Operand image in shared memory ("interleaved", no swizzle): element (row m, feature f) at byte
//   (f/4)*OP_LBO + m*16 + (f%4)*4, OP_LBO = 128*16 + 16: the canonical K-major UMMA layout with
//   LBO = OP_LBO, SBO = 128; the 16-byte pad makes both thread-per-row and thread-per-feature accesses
//   bank-conflict free.  kind::tf32 returns zeros for MN-major (transposed) operands on this part
//   (tools/umma_probe.cu), so the weight-gradient contraction over ROWS takes both operands from
//   row-transposed images in global memory ("T-image": (m, f) at float (m/16)*4096 + ((m%16)/4)*1024 +
//   f*4 + m%4, i.e. 16-row chunks that are K-major with K = row): the activation slab is stored that way
//   in the forward pass and Zbar is spilled that way in the reverse pass, 128 KB each per layer.
#include "common.cuh"
#include "residual.cuh"

namespace pinn {

constexpr int TC_H = 256;                  // hidden width handled by this kernel
constexpr int TC_M = 128;                  // rows per tile
constexpr int TC_TP = 32;                  // points per tile
#ifndef PINN_TC_WPS
#define PINN_TC_WPS 4
#endif
constexpr int TC_WPS = PINN_TC_WPS;         // worker warps per TMEM subpartition (2 or 4)
constexpr int TC_WORKERS = 128 * TC_WPS;
constexpr int TC_THREADS = TC_WORKERS + 64;
constexpr int TC_WCOLS = TC_H / TC_WPS;    // columns of a 128x256 accumulator owned by one worker warp
constexpr int TC_WBLK = TC_WCOLS / 16;     // 16-column blocks per worker
constexpr int TC_PARTS = TC_WORKERS / TC_H;  // worker threads per feature in the thread-per-feature phases
constexpr int TC_STAGES = 4;
constexpr int TC_STAGE_BYTES = 16384;
constexpr int TC_STAGE_FLOATS = TC_STAGE_BYTES / 4;
constexpr int OP_LBO = TC_M * 16 + 16;     // 2064
constexpr int OP_BYTES = (TC_H / 4) * OP_LBO;
constexpr int TC_WCHUNKS = TC_H * TC_H * 4 / TC_STAGE_BYTES;   // 16 chunks per weight image
constexpr int TC_SCHUNKS = TC_M * TC_H * 4 / TC_STAGE_BYTES;   // 8 chunks per slab entry
constexpr int TC_MAX_HH = 7;               // hidden->hidden layers whose biases are staged in shared memory

struct TcArgs {
  const float* params;
  const float* packed;   // per hidden->hidden layer: Wk image [k/4][n][k%4], then WT image [n/4][k][n%4]
  const float* inputs;
  const float* targets;
  const float* mask_count;
  float* grad;
  double* sums;
  float* out;
  float* dout[PINN_MAX_DIRS];
  float* slab;            // per CTA: (L-2) T-images of layer outputs, then one T-image of Zbar
  long long slab_stride;  // floats per CTA
  long long n_points;
  int n_tiles;
  float inv_n_res;
  float inv_n_fid;
};

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TC_WORKERS) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, FP32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 TMEM lanes x 32 columns; thread T gets, for u = 0..3: v[4u], v[4u+1] = (lane T/4,     cols 8u + 2(T%4) + {0,1})
//                                                         v[4u+2], v[4u+3] = (lane T/4 + 8, same cols)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// UMMA shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" core matrices of 8 x 16 bytes)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for sm_100
  return d;
}
// instruction descriptor: TF32 x TF32 -> F32, M x N, operand majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float warp_sum_tc(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct RowMajorJets {  // output jets of one point in the [128 rows][8] area, row = 4*p + j
  float* outs;
  int p;
  __device__ __forceinline__ float get(int col, int j) const { return outs[(4 * p + j) * 8 + col]; }
  __device__ __forceinline__ void set(int col, int j, float v) { outs[(4 * p + j) * 8 + col] = v; }
  __device__ __forceinline__ void add(int col, int j, float v) { outs[(4 * p + j) * 8 + col] += v; }
};

#ifdef PINN_TC_DEBUG
#define TCT_DECL long long tct[16] = {0}; long long tct0 = clock64();
#define TCT(i) { const long long t_ = clock64(); tct[i] += t_ - tct0; tct0 = t_; }
#else
#define TCT_DECL
#define TCT(i)
#endif

template <bool BWD>
__global__ void __launch_bounds__(TC_THREADS, 1)
    jet_tc_kernel(const __grid_constant__ pinn_desc_t D, const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* op = smem_raw;                                   // operand image (A or Zbar)
  unsigned char* ring = op + OP_BYTES;
  float* w0s = reinterpret_cast<float*>(ring + TC_STAGES * TC_STAGE_BYTES);  // [H][8]  W0[f][c]
  float* wls = w0s + TC_H * 8;                                    // [8][H]  Wlast[c][f]
  float* outs = wls + 8 * TC_H;                                   // [128][8] output jets / seeds
  float* xin = outs + TC_M * 8;                                   // [32][8]
  float* bias_s = xin + TC_TP * 8;                                // [TC_MAX_HH][H] hidden-layer biases
  double* red = reinterpret_cast<double*>(bias_s + TC_MAX_HH * TC_H);  // [16]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + PINN_NSUMS);  // [S]
  uint64_t* empty = full + TC_STAGES;                             // [S]
  uint64_t* op_ready = empty + TC_STAGES;
  uint64_t* mma_done = op_ready + 1;
  uint64_t* slab_ready = mma_done + 1;
  uint64_t* zt_ready = slab_ready + 1;
  uint64_t* mma_done_b = zt_ready + 1;   // weight-gradient accumulator (TMEM columns 256..511) complete
  uint64_t* rb_free = mma_done_b + 1;    // ... and drained by the workers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rb_free + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int L = D.n_linear;
  const int d = D.widths[0], o = D.widths[L];
  const int NHH = L - 2;                      // hidden->hidden layers (tensor-core jobs per direction)
  const int kind = D.residual_kind;
  const int my_tiles = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  float* slab = A.slab + (long long)blockIdx.x * A.slab_stride;
  float* slab_r = slab + (size_t)(L - 2) * TC_M * TC_H;      // thread-private row images (reverse loads)
  float* zt = slab_r + (size_t)(L - 2) * TC_M * TC_H;        // two T-images of Zbar (layer parity), split by feature half
  const long long P0 = (long long)d * TC_H + TC_H;               // params of layer 0
  const long long PH = (long long)TC_H * TC_H + TC_H;            // params of a hidden->hidden layer
  const long long poffL = P0 + (long long)NHH * PH;              // params offset of the last layer

  // ---------------- one-time setup ----------------
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(op_ready, TC_WORKERS);
    mbar_init(mma_done, 1);
    mbar_init(slab_ready, TC_WORKERS);
    mbar_init(zt_ready, TC_WORKERS);
    mbar_init(mma_done_b, 1);
    mbar_init(rb_free, TC_WORKERS);
    mbar_fence_init();
  }
  if (tid < PINN_NSUMS) red[tid] = 0.0;
  for (int i = tid; i < TC_H * 8; i += TC_THREADS) {
    const int f = i >> 3, c = i & 7;
    w0s[i] = c < d ? A.params[(long long)f * d + c] : 0.f;
  }
  for (int i = tid; i < 8 * TC_H; i += TC_THREADS) {
    const int c = i / TC_H, f = i - c * TC_H;
    wls[i] = c < o ? A.params[poffL + (long long)c * TC_H + f] : 0.f;
  }
  for (int i = tid; i < TC_MAX_HH * TC_H; i += TC_THREADS) {
    const int hl = i / TC_H, f = i - hl * TC_H;
    bias_s[i] = hl < NHH ? A.params[P0 + (long long)hl * PH + (long long)TC_H * TC_H + f] : 0.f;
  }
  if (warp == TC_WORKERS / 32 + 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == TC_WORKERS / 32) {
    // =========================================== producer ===========================================
    if (lane == 0) {
      int pc = 0, nzt = 0;
      auto load = [&](const float* src) {
        const int s = pc % TC_STAGES;
        mbar_wait(&empty[s], (uint32_t)(((pc / TC_STAGES) & 1) ^ 1));
        mbar_expect_tx(&full[s], TC_STAGE_BYTES);
        tma_load_1d(ring + s * TC_STAGE_BYTES, src, TC_STAGE_BYTES, &full[s]);
        ++pc;
      };
      for (int it = 0; it < my_tiles; ++it) {
        for (int hl = 0; hl < NHH; ++hl)
          for (int c = 0; c < TC_WCHUNKS; ++c)
            load(A.packed + (size_t)hl * 2 * TC_H * TC_H + (size_t)c * TC_STAGE_FLOATS);
        if (BWD) {
          mbar_wait(slab_ready, (uint32_t)(it & 1));  // this tile's slab entries are written and fenced
          // one weight-gradient half job = 128 Zbar features x 256 A_in features over the tile's 128 rows:
          // per 32 rows one stage of Zbar^T (two 16-row half-chunks) and two stages of A_in^T
          auto load_dw_half = [&](int l, int h) {
            const float* zsrc = zt + (size_t)(l & 1) * TC_M * TC_H + (size_t)h * (TC_M * TC_H / 2);
            const float* asrc = slab + (size_t)(l - 1) * TC_M * TC_H;
            for (int g = 0; g < 4; ++g) {
              load(zsrc + (size_t)g * TC_STAGE_FLOATS);
              load(asrc + (size_t)(2 * g) * TC_STAGE_FLOATS);
              load(asrc + (size_t)(2 * g + 1) * TC_STAGE_FLOATS);
            }
          };
          for (int l = L - 2; l >= 1; --l) {
            for (int c = 0; c < TC_WCHUNKS; ++c)       // adjoint job of layer l
              load(A.packed + (size_t)(l - 1) * 2 * TC_H * TC_H + (size_t)TC_H * TC_H +
                   (size_t)c * TC_STAGE_FLOATS);
            if (l < L - 2) load_dw_half(l + 1, 1);
            mbar_wait(zt_ready, (uint32_t)(nzt & 1));  // Zbar_l has been spilled as a T-image
            ++nzt;
            load_dw_half(l, 0);
          }
          load_dw_half(1, 1);
        }
      }
    }
  } else if (warp == TC_WORKERS / 32 + 1) {
    // =========================================== MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc_k = umma_idesc(TC_M, TC_H, 0, 0);
      const uint32_t op_addr = smem_u32(op);
      const uint32_t ring_addr = smem_u32(ring);
      int cc = 0, jobs = 0, nB = 0;
      auto wait_ready = [&]() {
        mbar_wait(op_ready, (uint32_t)(jobs & 1));
        ++jobs;
        tc_fence_after();
      };
      // D[128 x 256] = OP (K-major, 256 features) * image chunks (K-major)
      auto gemm_k = [&]() {
        for (int c = 0; c < TC_WCHUNKS; ++c) {
          const int s = cc % TC_STAGES;
          mbar_wait(&full[s], (uint32_t)((cc / TC_STAGES) & 1));
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int kstep = c * 2 + kk;  // 8 contraction features per MMA = two 16-byte K chunks
            const uint64_t ad = umma_desc(op_addr + (uint32_t)kstep * 2u * OP_LBO, OP_LBO, 128);
            const uint64_t bd = umma_desc(ring_addr + (uint32_t)s * TC_STAGE_BYTES + (uint32_t)kk * 2u * (TC_H * 16),
                                          TC_H * 16, 128);
            umma_tf32(tmem_base, ad, bd, idesc_k, kstep > 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          ++cc;
        }
        umma_commit(mma_done);
      };
      for (int it = 0; it < my_tiles; ++it) {
        for (int hl = 0; hl < NHH; ++hl) {
          wait_ready();
          gemm_k();
        }
        if (BWD) {
          // weight-gradient half job into TMEM columns 256..511: both operands are K-major with K = row
          auto dw_half = [&]() {
            if (nB > 0) {
              mbar_wait(rb_free, (uint32_t)((nB - 1) & 1));   // the previous half has been drained
              tc_fence_after();
            }
            ++nB;
            for (int g = 0; g < 4; ++g) {
              const int sz = cc % TC_STAGES;
              mbar_wait(&full[sz], (uint32_t)((cc / TC_STAGES) & 1));
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int ca = cc + 1 + i, sa = ca % TC_STAGES;
                mbar_wait(&full[sa], (uint32_t)((ca / TC_STAGES) & 1));
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const int kstep = g * 4 + i * 2 + kk;   // 8 rows per MMA
                  const uint64_t ad = umma_desc(ring_addr + (uint32_t)sz * TC_STAGE_BYTES + (uint32_t)i * 8192u +
                                                    (uint32_t)kk * 2u * (128u * 16u),
                                                128 * 16, 128);
                  const uint64_t bd = umma_desc(ring_addr + (uint32_t)sa * TC_STAGE_BYTES + (uint32_t)kk * 2u * (TC_H * 16),
                                                TC_H * 16, 128);
                  umma_tf32(tmem_base + 256u, ad, bd, idesc_k, kstep > 0 ? 1u : 0u);
                }
                umma_commit(&empty[sa]);
              }
              umma_commit(&empty[sz]);
              cc += 3;
            }
            umma_commit(mma_done_b);
          };
          for (int l = L - 2; l >= 1; --l) {
            wait_ready();                      // Zbar_l is in the operand image, columns 0..255 are drained
            gemm_k();                          // adjoint of the layer input
            if (l < L - 2) dw_half();          // layer l+1, features 128..255
            dw_half();                         // layer l,   features 0..127
          }
          dw_half();                           // layer 1,   features 128..255
        }
      }
    }
  } else {
    // =========================================== workers ============================================
    const int sp = warp & 3, half = warp >> 2;   // `half` = which TC_WCOLS-wide column slice this warp owns
    const int cbase = half * TC_WCOLS;
    const int m = sp * 32 + lane;          // this thread's row of the tile = its TMEM lane
    const int p = m >> 2, j = m & 3;
    const int leader = lane & ~3;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(sp * 32) << 16);
    const float inv_cnt = (kind == PINN_RES_CONT_ONLY && A.mask_count) ? 1.0f / *A.mask_count : 0.f;
    int mj = 0;   // adjoint / forward MMA jobs waited for
    int nbw = 0;  // weight-gradient half jobs drained
    TCT_DECL
    auto wait_mma = [&]() {
      mbar_wait(mma_done, (uint32_t)(mj & 1));
      ++mj;
      tc_fence_after();
    };
    auto signal_ready = [&]() {
      tc_fence_before();
      fence_async_proxy();
      mbar_arrive(op_ready);
    };
    // 16 consecutive features f0.. of row m -> operand image (and slab entry e)
    auto store_op = [&](int f0, const float (&v)[16]) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(op + (f0 / 4 + q) * OP_LBO + m * 16) =
            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    };
    // T-image (row-transposed, 16-row chunks): (row m, feature f) at float (m/16)*4096 + ((m%16)/4)*1024 + f*4 + m%4.
    // Each warp transposes the 32 rows x 128 features it owns out of the operand image: 4 conflict-free
    // LDS.32 (the 4 rows of a quad) -> one coalesced 16-byte store per feature.
    auto t_copy = [&](float* img, bool split) {
      __syncwarp();
#pragma unroll 4
      for (int itc = 0; itc < TC_WCOLS / 4; ++itc) {
        const int idx = itc * 32 + lane;
        const int q = idx / TC_WCOLS, f = cbase + (idx % TC_WCOLS);
        const unsigned char* src = op + (f >> 2) * OP_LBO + (sp * 32 + 4 * q) * 16 + (f & 3) * 4;
        float4 v;
        v.x = *reinterpret_cast<const float*>(src);
        v.y = *reinterpret_cast<const float*>(src + 16);
        v.z = *reinterpret_cast<const float*>(src + 32);
        v.w = *reinterpret_cast<const float*>(src + 48);
        const size_t dst = split ? (size_t)(f >> 7) * (TC_M * TC_H / 2) + (size_t)(2 * sp + (q >> 2)) * (TC_STAGE_FLOATS / 2) +
                                       (size_t)(q & 3) * (TC_H * 2) + (size_t)(f & 127) * 4
                                 : (size_t)(2 * sp + (q >> 2)) * TC_STAGE_FLOATS + (size_t)(q & 3) * (TC_H * 4) + (size_t)f * 4;
        *reinterpret_cast<float4*>(img + dst) = v;
      }
    };
    // R-image: thread-private float4 slots [(block b, q)][worker thread] (coalesced), for the reverse loads;
    // written as a copy of this thread's own row out of the operand image, after the MMA has been released
    const int wtid = tid;  // workers are threads 0..255
    auto r_copy = [&](float* img) {
#pragma unroll 8
      for (int c = 0; c < TC_WCOLS / 4; ++c)
        reinterpret_cast<float4*>(img)[(size_t)c * TC_WORKERS + wtid] =
            *reinterpret_cast<const float4*>(op + (cbase / 4 + c) * OP_LBO + m * 16);
    };
    auto r_load = [&](const float* img, int b, float (&v)[16]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = reinterpret_cast<const float4*>(img)[(size_t)(b * 4 + q) * TC_WORKERS + wtid];
        v[4 * q] = t.x, v[4 * q + 1] = t.y, v[4 * q + 2] = t.z, v[4 * q + 3] = t.w;
      }
    };
    // forward activation on 16 features: z (pre-activation of this row) -> post-activation jets
    auto activate = [&](float (&z)[16], const float* bias) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = tanh_approx(z[i] + (bias ? bias[i] : 0.f));
        const float al = __shfl_sync(0xffffffffu, a, leader);
        z[i] = round_tf32(j == 0 ? a : (1.f - al * al) * z[i]);
      }
    };
    // adjoint through the activation: abar (adjoint of post-activation jets of this row), act (stored
    // post-activation jets of this row) -> adjoint of the pre-activation jets
    auto adjoint = [&](float (&ab)[16], const float (&act)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float al = __shfl_sync(0xffffffffu, act[i], leader);
        const float s = 1.f - al * al;
        float pr = j == 0 ? 0.f : ab[i] * act[i];
        pr += __shfl_xor_sync(0xffffffffu, pr, 1);
        pr += __shfl_xor_sync(0xffffffffu, pr, 2);
        ab[i] = round_tf32(j == 0 ? fmaf(-2.f * al, pr, ab[i] * s) : ab[i] * s);
      }
    };

    for (int it = 0; it < my_tiles; ++it) {
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      const long long p0 = tile * TC_TP;
      for (int i = tid; i < TC_TP * 8; i += TC_WORKERS) {
        const int pp = i >> 3, c = i & 7;
        const long long gp = p0 + pp;
        xin[i] = (c < d && gp < A.n_points) ? A.inputs[gp * d + c] : 0.f;
      }
      worker_bar();
      TCT(0)
      // ---------------- layer 0 (d -> 256) on the FP32 pipes ----------------
      {
        const int dircol = (j >= 1 && j - 1 < D.n_dirs) ? D.dir_cols[j - 1] : -1;
        for (int b = 0; b < TC_WBLK; ++b) {
          const int f0 = cbase + b * 16;
          float z[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float* wr = w0s + (f0 + i) * 8;
            float acc;
            if (j == 0) {
              acc = A.params[(long long)d * TC_H + f0 + i];
              for (int c = 0; c < d; ++c) acc = fmaf(xin[p * 8 + c], wr[c], acc);
            } else {
              acc = dircol >= 0 ? wr[dircol] : 0.f;
            }
            z[i] = acc;
          }
          activate(z, nullptr);
          store_op(f0, z);
        }
      }
      TCT(1)
      signal_ready();
      if (BWD && NHH >= 1) {
        r_copy(slab_r);
        t_copy(slab, false);
      }
      TCT(3)
      // ---------------- hidden layers 1..L-2 on the tensor cores ----------------
      for (int l = 1; l <= L - 2; ++l) {
        wait_mma();
        TCT(2)
        const float* bias_l = (l - 1 < TC_MAX_HH) ? bias_s + (l - 1) * TC_H
                                                  : A.params + P0 + (long long)(l - 1) * PH + (long long)TC_H * TC_H;
        for (int b = 0; b < TC_WBLK; ++b) {
          const int f0 = cbase + b * 16;
          float z[16];
          tmem_ld16(tmem_row + (uint32_t)f0, z);
          // value rows: a = tanh(z + b); tangent rows: s * zdot with s = 1 - a^2 of the point's value row
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = tanh_approx(z[i] + bias_l[f0 + i]);
            const float al = __shfl_sync(0xffffffffu, a, leader);
            z[i] = round_tf32(j == 0 ? a : (1.f - al * al) * z[i]);
          }
          store_op(f0, z);
        }
        TCT(1)
        if (l < L - 2) signal_ready();
        if (BWD && l <= L - 3) {
          r_copy(slab_r + (size_t)l * TC_M * TC_H);
          t_copy(slab + (size_t)l * TC_M * TC_H, false);
        }
        TCT(3)
      }
      if (BWD) {
        // the TMA engine reads the slab at L2: publish the generic-proxy stores at GPU scope first
        __threadfence();
        fence_async_proxy();
        mbar_arrive(slab_ready);
      }
      tc_fence_before();
      worker_bar();
      // ---------------- last layer (256 -> o) on the FP32 pipes ----------------
      {
        constexpr int NCS = TC_WORKERS / 128, NCI = 8 / NCS;
        const int mm = tid & 127, cs = tid >> 7;
        float acc[NCI];
#pragma unroll
        for (int ci = 0; ci < NCI; ++ci) acc[ci] = 0.f;
        for (int q = 0; q < TC_H / 4; ++q) {
          const float4 a = *reinterpret_cast<const float4*>(op + q * OP_LBO + mm * 16);
#pragma unroll
          for (int ci = 0; ci < NCI; ++ci) {
            const float4 w = *reinterpret_cast<const float4*>(wls + (cs + NCS * ci) * TC_H + 4 * q);
            acc[ci] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[ci]))));
          }
        }
#pragma unroll
        for (int ci = 0; ci < NCI; ++ci) {
          const int c = cs + NCS * ci;
          float v = acc[ci];
          if ((mm & 3) == 0 && c < o) v += A.params[poffL + (long long)TC_H * o + c];
          outs[mm * 8 + c] = c < o ? v : 0.f;
        }
      }
      worker_bar();
      // ---------------- residual / misfit epilogue: warp 0, one lane per point ----------------
      if (warp == 0) {
        const long long gp = p0 + lane;
        RowMajorJets acc{outs, lane};
        float ls[PINN_NSUMS];
        residual_epilogue<4>(D, acc, true, gp < A.n_points, gp, xin + lane * 8,
                             EpiArgs{A.targets, nullptr, {nullptr, nullptr, nullptr}, A.out,
                                     {A.dout[0], A.dout[1], A.dout[2]}, A.inv_n_res, A.inv_n_fid, inv_cnt},
                             ls);
#pragma unroll
        for (int i = 0; i < PINN_NSUMS; ++i) {
          const float v = warp_sum_tc(ls[i]);
          if (lane == 0 && v != 0.f) red[i] += (double)v;
        }
      }
      worker_bar();
      TCT(4)
      if (!BWD) continue;

      // =============================== reverse ===============================
      // ---- last layer: dW_last[c][f] = sum_m zbar[m][c] * A[m][f] (thread per feature), db_last ----
      {
        const int f = tid & (TC_H - 1), part = tid / TC_H;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        const unsigned char* ap = op + (f >> 2) * OP_LBO + (f & 3) * 4;
        for (int mm = part * (TC_M / TC_PARTS); mm < (part + 1) * (TC_M / TC_PARTS); ++mm) {
          const float a = *reinterpret_cast<const float*>(ap + mm * 16);
          const float4 z0 = *reinterpret_cast<const float4*>(outs + mm * 8);
          const float4 z1 = *reinterpret_cast<const float4*>(outs + mm * 8 + 4);
          acc[0] = fmaf(z0.x, a, acc[0]), acc[1] = fmaf(z0.y, a, acc[1]);
          acc[2] = fmaf(z0.z, a, acc[2]), acc[3] = fmaf(z0.w, a, acc[3]);
          acc[4] = fmaf(z1.x, a, acc[4]), acc[5] = fmaf(z1.y, a, acc[5]);
          acc[6] = fmaf(z1.z, a, acc[6]), acc[7] = fmaf(z1.w, a, acc[7]);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < o) atomicAdd(A.grad + poffL + (long long)c * TC_H + f, acc[c]);
        if (tid < o) {
          float s = 0.f;
          for (int pp = 0; pp < TC_TP; ++pp) s += outs[(4 * pp) * 8 + tid];
          atomicAdd(A.grad + poffL + (long long)TC_H * o + tid, s);
        }
      }
      worker_bar();
      // ---- abar = zbar_last * W_last, through the activation of layer L-2 -> Zbar_{L-2} in place ----
      {
        float zl[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) zl[c] = outs[m * 8 + c];
        for (int b = 0; b < TC_WBLK; ++b) {
          const int f0 = cbase + b * 16;
          float ab[16], act[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) a = fmaf(zl[c], wls[c * TC_H + f0 + i], a);
            ab[i] = a;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(op + (f0 / 4 + q) * OP_LBO + m * 16);
            act[4 * q] = v.x, act[4 * q + 1] = v.y, act[4 * q + 2] = v.z, act[4 * q + 3] = v.w;
          }
          adjoint(ab, act);
          store_op(f0, ab);
        }
      }
      t_copy(zt + (size_t)((L - 2) & 1) * TC_M * TC_H, true);
      __threadfence();
      fence_async_proxy();
      mbar_arrive(zt_ready);
      worker_bar();
      TCT(5)
      // ---- hidden layers L-2 .. 1, software-pipelined ----
      //   tensor core: adjoint job of layer l (TMEM columns 0..255), weight-gradient halves (columns 256..511)
      //   workers:     adjoint epilogue of layer l | drain half 1 of layer l+1 | spill Zbar_{l-1}^T | drain half 0 of layer l
      // The only serial chain is adjoint epilogue -> adjoint MMA -> adjoint epilogue; the drains hide behind it.
      auto bias_grad = [&](int l) {   // sum over the value rows of Zbar_l (thread per feature)
        const int f = tid & (TC_H - 1), part = tid / TC_H;
        const unsigned char* zp = op + (f >> 2) * OP_LBO + (f & 3) * 4;
        float sacc = 0.f;
#pragma unroll 8
        for (int pp = part * (TC_TP / TC_PARTS); pp < (part + 1) * (TC_TP / TC_PARTS); ++pp)
          sacc += *reinterpret_cast<const float*>(zp + (4 * pp) * 16);
        atomicAdd(A.grad + P0 + (long long)(l - 1) * PH + (long long)TC_H * TC_H + f, sacc);
      };
      // drain TMEM columns 256..511 = dW rows [128h, 128h+128) of layer l.  16x256b loads: the 4 lanes of a
      // quad hold 8 consecutive columns of one row, so every RED instruction updates full 32-byte sectors
      auto drain = [&](int l, int h) {
        mbar_wait(mma_done_b, (uint32_t)(nbw & 1));
        ++nbw;
        tc_fence_after();
        const long long poff = P0 + (long long)(l - 1) * PH;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int row = h * 128 + sp * 32 + g * 16 + (lane >> 2);
          float* grow = A.grad + poff + (long long)row * TC_H + cbase + 2 * (lane & 3);
          const uint32_t ta = tmem_base + ((uint32_t)(sp * 32 + g * 16) << 16) + (uint32_t)(256 + cbase);
#pragma unroll
          for (int cb = 0; cb < TC_WCOLS / 32; ++cb) {
            float v[16];
            tmem_ld_16x256b_x4(ta + (uint32_t)(cb * 32), v);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              red_add_v2(grow + cb * 32 + 8 * u, v[4 * u], v[4 * u + 1]);
              red_add_v2(grow + 8 * TC_H + cb * 32 + 8 * u, v[4 * u + 2], v[4 * u + 3]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(rb_free);
      };
      bias_grad(L - 2);
      signal_ready();                       // adjoint job of layer L-2 may start
      TCT(6)
      for (int l = L - 2; l >= 1; --l) {
        // adjoint through the activation of layer l-1 -> Zbar_{l-1} in place
        {
          const float* rimg = slab_r + (size_t)(l - 1) * TC_M * TC_H;
          float act[16], nxt[16];
          r_load(rimg, 0, act);   // issued before the wait: the loads overlap the adjoint MMA
          wait_mma();
          TCT(9)
          for (int b = 0; b < TC_WBLK; ++b) {
            const int f0 = cbase + b * 16;
            float ab[16];
            if (b < TC_WBLK - 1) r_load(rimg, b + 1, nxt);
            tmem_ld16(tmem_row + (uint32_t)f0, ab);
            adjoint(ab, act);
            store_op(f0, ab);
#pragma unroll
            for (int i = 0; i < 16; ++i) act[i] = nxt[i];
          }
        }
        tc_fence_before();
        worker_bar();                       // Zbar_{l-1} complete in the operand image
        TCT(10)
        if (l > 1) {
          bias_grad(l - 1);
          signal_ready();                   // adjoint job of layer l-1 may start
        }
        TCT(6)
        if (l < L - 2) drain(l + 1, 1);
        TCT(8)
        if (l > 1) {
          // Zbar_{l-1}^T feeds the weight-gradient jobs of layer l-1.  Its buffer (layer parity) was last read
          // by the half-1 job of layer l+1, which the drain above has just seen complete.
          t_copy(zt + (size_t)((l - 1) & 1) * TC_M * TC_H, true);
          __threadfence();
          fence_async_proxy();
          mbar_arrive(zt_ready);
        }
        TCT(11)
        drain(l, 0);
        TCT(7)
      }
      drain(1, 1);
      worker_bar();
      TCT(8)
      // ---- layer 0: dW0[f][c], db0[f] from Zbar_0 (thread per feature) ----
      {
        const int f = tid & (TC_H - 1), part = tid / TC_H;
        const unsigned char* zp = op + (f >> 2) * OP_LBO + (f & 3) * 4;
        float acc[8], tj[3] = {0.f, 0.f, 0.f}, sb = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        for (int pp = part * (TC_TP / TC_PARTS); pp < (part + 1) * (TC_TP / TC_PARTS); ++pp) {
          const float z0 = *reinterpret_cast<const float*>(zp + (4 * pp) * 16);
          sb += z0;
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = fmaf(z0, xin[pp * 8 + c], acc[c]);
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) tj[jj] += *reinterpret_cast<const float*>(zp + (4 * pp + 1 + jj) * 16);
        }
        for (int jj = 0; jj < D.n_dirs; ++jj) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c == D.dir_cols[jj]) acc[c] += tj[jj];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < d) atomicAdd(A.grad + (long long)f * d + c, acc[c]);
        atomicAdd(A.grad + (long long)d * TC_H + f, sb);
      }
      worker_bar();
      TCT(12)
    }
#ifdef PINN_TC_DEBUG
    if (blockIdx.x == 0 && tid == 0) {
      long long tot = 0;
      for (int i = 0; i < 13; ++i) tot += tct[i];
      printf("TC phases (cycles per tile, %d tiles): in+bar %lld | L0+fwd-epi %lld | fwd-wait-mma %lld | fwd signal+copies %lld | last+residual %lld | rev-last %lld | db+signal %lld | wait-dW %lld | drain %lld | signal+wait-adj %lld | adj-epi %lld | tcopy+fence+bar %lld | L0-rev %lld | total %lld\n", my_tiles,
             tct[0] / my_tiles, tct[1] / my_tiles, tct[2] / my_tiles, tct[3] / my_tiles, tct[4] / my_tiles, tct[5] / my_tiles, tct[6] / my_tiles,
             tct[7] / my_tiles, tct[8] / my_tiles, tct[9] / my_tiles, tct[10] / my_tiles, tct[11] / my_tiles, tct[12] / my_tiles, tot / my_tiles);
    }
#endif
  }

  // ---------------- teardown ----------------
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == TC_WORKERS / 32 + 1) tmem_dealloc(tmem_base, 512);
  if (tid < PINN_NSUMS && A.sums && red[tid] != 0.0) atomicAdd(A.sums + tid, red[tid]);
}

// Weight images for the tensor-core jobs, TF32-rounded: layer hl (= linear layer hl+1)
//   Wk[(k/4)*H*4 + n*4 + k%4] = W[n][k]   (forward:  B operand, rows n, contraction k)
//   WT[(n/4)*H*4 + k*4 + n%4] = W[n][k]   (adjoint:  B operand, rows k, contraction n)
__global__ void pack_tc_kernel(const __grid_constant__ pinn_desc_t D, const float* __restrict__ params,
                               float* __restrict__ packed) {
  const int hl = blockIdx.y;
  const int d = D.widths[0];
  const long long poff = (long long)d * TC_H + TC_H + (long long)hl * ((long long)TC_H * TC_H + TC_H);
  float* wk = packed + (size_t)hl * 2 * TC_H * TC_H;
  float* wt = wk + (size_t)TC_H * TC_H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < TC_H * TC_H; i += gridDim.x * blockDim.x) {
    const int n = i / TC_H, k = i - n * TC_H;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(params[poff + i]));
    const float w = __uint_as_float(r);
    wk[(size_t)(k >> 2) * TC_H * 4 + n * 4 + (k & 3)] = w;
    wt[(size_t)(n >> 2) * TC_H * 4 + k * 4 + (n & 3)] = w;
  }
}

// --------------------------------------------------------------------------------------- host side
constexpr size_t tc_smem_bytes() {
  return (size_t)OP_BYTES + (size_t)TC_STAGES * TC_STAGE_BYTES + (size_t)TC_H * 8 * 4 + (size_t)8 * TC_H * 4 +
         (size_t)TC_M * 8 * 4 + (size_t)TC_TP * 8 * 4 + (size_t)TC_MAX_HH * TC_H * 4 + PINN_NSUMS * 8 + (2 * TC_STAGES + 6) * 8 + 16;
}

// Can this description run on the tensor-core kernel?  (otherwise the caller reports UNSUPPORTED)
bool tc_supported(const pinn_desc_t* D, const char** why) {
  const int L = D->n_linear;
  *why = "";
  if (L < 3) return *why = "needs at least two hidden layers", false;
  for (int i = 1; i < L; ++i)
    if (D->widths[i] != TC_H) return *why = "every hidden layer must be 256 wide", false;
  if (D->activation != PINN_ACT_TANH) return *why = "tanh activation only", false;
  const int k = D->residual_kind;
  if (k < PINN_RES_CONT_ONLY || k > PINN_RES_WAVE_AVG)
    return *why = "needs a PDE residual kind (value-only and external-seed passes use the FP32 kernel)", false;
  return true;
}

int tc_workspace(const pinn_desc_t* D, long long n_points, int sms, size_t* packed_bytes, size_t* slab_bytes,
                 long long* slab_stride, int* grid) {
  const int L = D->n_linear;
  long long tiles = (n_points + TC_TP - 1) / TC_TP;
  long long g = tiles < sms ? tiles : sms;
  if (g < 1) g = 1;
  *grid = (int)g;
  *packed_bytes = (size_t)(L - 2) * 2 * TC_H * TC_H * 4;
  *slab_stride = (long long)(2 * (L - 2) + 2) * TC_M * TC_H;   // T- and R-images of (L-2) layer outputs + 2 Zbar spills
  *slab_bytes = (size_t)g * (size_t)(*slab_stride) * 4;
  return PINN_OK;
}

int run_tc_pass(const pinn_desc_t* D, const pinn_eval_args_t* a, bool bwd, void* workspace, size_t ws_bytes,
                cudaStream_t st) {
  int dev = 0, sms = 0;
  PINN_CUDA(cudaGetDevice(&dev));
  PINN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  size_t pk = 0, sl = 0;
  long long stride = 0;
  int grid = 1;
  tc_workspace(D, a->n_points, sms, &pk, &sl, &stride, &grid);
  const size_t pk_al = (pk + 255) & ~size_t(255);
  if (ws_bytes < pk_al + sl) return set_error("workspace too small: %zu < %zu bytes", ws_bytes, pk_al + sl), PINN_E_WORKSPACE;
  if (bwd && ((uintptr_t)a->grad & 15) != 0) return set_error("grad must be 16-byte aligned for the tensor-core path"), PINN_E_ARG;
  float* packed = reinterpret_cast<float*>(workspace);
  float* slab = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + pk_al);
  if (!(a->flags & PINN_FLAG_SKIP_PACK)) {
    dim3 g(64, D->n_linear - 2);
    pack_tc_kernel<<<g, 256, 0, st>>>(*D, a->params, packed);
    PINN_CUDA(cudaGetLastError());
  }
  long long tiles = (a->n_points + TC_TP - 1) / TC_TP;
  if (tiles == 0) return PINN_OK;
  if (tiles > 0x7fffffffLL) return set_error("too many tiles"), PINN_E_UNSUPPORTED;
  TcArgs A;
  A.params = a->params;
  A.packed = packed;
  A.inputs = a->inputs;
  A.targets = D->n_targets > 0 ? a->targets : nullptr;
  A.mask_count = a->mask_count;
  A.grad = a->grad;
  A.sums = a->sums;
  A.out = a->out;
  for (int j = 0; j < PINN_MAX_DIRS; ++j) A.dout[j] = a->dout[j];
  A.slab = slab;
  A.slab_stride = stride;
  A.n_points = a->n_points;
  A.n_tiles = (int)tiles;
  A.inv_n_res = a->n_res_global > 0 ? (float)(1.0 / (double)a->n_res_global) : 0.f;
  A.inv_n_fid = a->n_fid_global > 0 ? (float)(1.0 / (double)a->n_fid_global) : 0.f;
  const size_t smem = tc_smem_bytes();
  if (bwd) {
    PINN_CUDA(cudaFuncSetAttribute(jet_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jet_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(*D, A);
  } else {
    PINN_CUDA(cudaFuncSetAttribute(jet_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jet_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(*D, A);
  }
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

}  // namespace pinn
