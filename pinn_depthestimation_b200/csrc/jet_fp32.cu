// Fused FP32 jet-MLP kernel for sm_100a: forward value+tangent propagation, PDE-residual epilogue,
// reverse sweep to the flat weight gradient.  Persistent CTAs, one tile of collocation points at a
// time; activations of a tile never leave the CTA's (L2-resident) slab.
//
// What it replaces in the reference (files under the reference tree):
//   dnn.py:54-55 (DNN.forward), physics.py:6-15 (compute_gradient: one autograd sweep per call),
//   physics.py:18-120 (the four residual losses), the MSE terms of pinn.loss_func
//   (train_newmethod.py:129-133, train.py:136-141) and loss.backward() (train_newmethod.py:200).
//
// Layout inside a CTA (shared memory, "feature-major"):
//   buffer[f][m], f = feature (row), m = j*TP + p with j = jet component (0 value, 1.. tangents),
//   p = point in tile; row pitch MP = J*TP + 4 floats.  A thread owns 4 points x J jets x 4 features.
//   Weights stream through a 3-stage ring of shared-memory chunks filled by 1-D TMA bulk copies
//   (cp.async.bulk + mbarrier complete_tx) from a packed, zero-padded copy in the workspace.
#include <string.h>

#include "common.cuh"
#include "residual.cuh"

namespace pinn {

constexpr int kStages = 3;

// jets of the network output in the feature-major tile buffer: (col, j) -> buf[col][j*TP + p]
template <int J, int TP, int MP>
struct FeatureMajorJets {
  float* buf;
  int p;
  __device__ __forceinline__ float get(int col, int j) const { return buf[col * MP + j * TP + p]; }
  __device__ __forceinline__ void set(int col, int j, float v) { buf[col * MP + j * TP + p] = v; }
  __device__ __forceinline__ void add(int col, int j, float v) { buf[col * MP + j * TP + p] += v; }
};

struct KArgs {
  const float* params;
  const float* packed;  // per layer: WT [KP][NP] then Wn [NP][KP], zero padded
  const float* inputs;
  const float* targets;
  const float* mask_count;
  const float* seed_out;
  const float* seed_dout[PINN_MAX_DIRS];
  float* grad;
  double* sums;
  float* out;
  float* dout[PINN_MAX_DIRS];
  float* slab;            // per-CTA activation slabs
  long long slab_stride;  // floats per CTA
  long long n_points;
  int n_tiles;
  int stage_floats;
  int wp;  // rows of each activation buffer
  float inv_n_res;
  float inv_n_fid;
};

// Producer (thread 0) and consumers (all threads) walk the same weight-chunk sequence per tile:
// forward layers 0..L-1 (rows of WT), then -- when a gradient is wanted -- layers L-1..1 (rows of Wn).
struct Cursor {
  int phase, l, r0, pk;
  __device__ __forceinline__ void start() { phase = 0, l = 0, r0 = 0, pk = 0; }
  __device__ __forceinline__ void get(const pinn_desc_t& D, int sf, int& src_off, int& rows,
                                      int& rowlen) const {
    const int KP = pad4(D.widths[l]), NP = pad4(D.widths[l + 1]);
    if (phase == 0) {
      rowlen = NP;
      rows = min(max(1, sf / NP), KP - r0);
      src_off = pk + r0 * NP;
    } else {
      rowlen = KP;
      rows = min(max(1, sf / KP), NP - r0);
      src_off = pk + KP * NP + r0 * KP;
    }
  }
  __device__ __forceinline__ bool advance(const pinn_desc_t& D, int sf, bool bwd) {
    const int KP = pad4(D.widths[l]), NP = pad4(D.widths[l + 1]);
    const int total = phase == 0 ? KP : NP;
    const int rowlen = phase == 0 ? NP : KP;
    r0 += max(1, sf / rowlen);
    if (r0 < total) return true;
    r0 = 0;
    if (phase == 0) {
      if (l + 1 < D.n_linear) {
        pk += 2 * KP * NP;
        ++l;
        return true;
      }
      if (!bwd || D.n_linear == 1) return false;
      phase = 1;
      return true;
    }
    if (l - 1 >= 1) {
      --l;
      pk -= 2 * pad4(D.widths[l]) * pad4(D.widths[l + 1]);
      return true;
    }
    return false;
  }
};

// acc[j][pp][c] += sum_k X[k][j*TP + 4pg+pp] * Wc[k][4cg+c]   (xs, wc already offset by pg / cg)
template <int J, int TP, int MP>
__device__ __forceinline__ void gemm_chunk(float (&acc)[J][4][4], const float* __restrict__ xs,
                                           const float* __restrict__ wc, int rowlen, int rows) {
#pragma unroll 4
  for (int k = 0; k < rows; ++k) {
    const float4 w = *reinterpret_cast<const float4*>(wc + k * rowlen);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float4 x = *reinterpret_cast<const float4*>(xs + k * MP + j * TP);
      acc[j][0][0] = fmaf(x.x, w.x, acc[j][0][0]);
      acc[j][0][1] = fmaf(x.x, w.y, acc[j][0][1]);
      acc[j][0][2] = fmaf(x.x, w.z, acc[j][0][2]);
      acc[j][0][3] = fmaf(x.x, w.w, acc[j][0][3]);
      acc[j][1][0] = fmaf(x.y, w.x, acc[j][1][0]);
      acc[j][1][1] = fmaf(x.y, w.y, acc[j][1][1]);
      acc[j][1][2] = fmaf(x.y, w.z, acc[j][1][2]);
      acc[j][1][3] = fmaf(x.y, w.w, acc[j][1][3]);
      acc[j][2][0] = fmaf(x.z, w.x, acc[j][2][0]);
      acc[j][2][1] = fmaf(x.z, w.y, acc[j][2][1]);
      acc[j][2][2] = fmaf(x.z, w.z, acc[j][2][2]);
      acc[j][2][3] = fmaf(x.z, w.w, acc[j][2][3]);
      acc[j][3][0] = fmaf(x.w, w.x, acc[j][3][0]);
      acc[j][3][1] = fmaf(x.w, w.y, acc[j][3][1]);
      acc[j][3][2] = fmaf(x.w, w.z, acc[j][3][2]);
      acc[j][3][3] = fmaf(x.w, w.w, acc[j][3][3]);
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int J, int TP, int NT, bool BWD>
__global__ void __launch_bounds__(NT)
    jet_kernel(const __grid_constant__ pinn_desc_t D, const __grid_constant__ KArgs A) {
  constexpr int M = J * TP;
  constexpr int MP = M + 4;
  constexpr int PG = TP / 4;
  constexpr int CG = NT / PG;
  static_assert(TP % 4 == 0 && TP <= 32 && NT % PG == 0 && NT >= 32, "tile shape");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* buf0 = reinterpret_cast<float*>(smem_raw);
  float* buf1 = buf0 + A.wp * MP;
  float* stage = buf1 + A.wp * MP;
  float* xin = stage + kStages * A.stage_floats;
  double* red = reinterpret_cast<double*>(xin + TP * PINN_MAX_IN);
  uint64_t* full = reinterpret_cast<uint64_t*>(red + PINN_NSUMS);

  const int tid = threadIdx.x;
#ifndef FP32_CG_FAST
  // point group fastest: a quarter-warp's 128-bit stores of the activation write-back (row stride 4 * MP = 16 mod 32 banks
  // between column groups) then cover 32 distinct banks -- with the column group fastest they were 4-way conflicted -- and a
  // weight fetch is one 128-byte wavefront per warp (8 column groups, broadcast over the point groups) instead of four
  const int pg = tid % PG, cg = tid / PG;
#else
  const int cg = tid % CG, pg = tid / CG;
#endif
  const int L = D.n_linear;
  const int d = D.widths[0], o = D.widths[L];
  const int KP0 = pad4(d), NPo = pad4(o);
  const int sf = A.stage_floats;
  const int kind = D.residual_kind;
  const bool leaky = D.activation == PINN_ACT_LEAKY_RELU;
  const int my_tiles = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  float* slab = A.slab + (long long)blockIdx.x * A.slab_stride;

  // ---- producer state (meaningful in thread 0 only) ----
  Cursor pc;
  pc.start();
  int ptiles = my_tiles, pcount = 0;
  bool pvalid = my_tiles > 0;
  auto produce = [&]() {
    if (!pvalid) return;
    int src_off, rows, rowlen;
    pc.get(D, sf, src_off, rows, rowlen);
    const int s = pcount % kStages;
    const uint32_t bytes = (uint32_t)(rows * rowlen) * 4u;
    mbar_expect_tx(&full[s], bytes);
    tma_load_1d(stage + s * sf, A.packed + src_off, bytes, &full[s]);
    ++pcount;
    if (!pc.advance(D, sf, BWD)) {
      if (--ptiles > 0) pc.start();
      else pvalid = false;
    }
  };
  int cc = 0;  // chunks consumed so far (same in every thread)
  auto chunk_top = [&]() -> const float* {
    __syncthreads();  // everyone finished the previous chunk -> its stage may be refilled
    if (tid == 0 && cc > 0) produce();
    const int s = cc % kStages;
    mbar_wait(&full[s], (uint32_t)((cc / kStages) & 1));
    ++cc;
    return stage + s * sf;
  };

  if (tid < PINN_NSUMS) red[tid] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    for (int s = 0; s < kStages; ++s) produce();
  }
  __syncthreads();

  const float inv_cnt = (kind == PINN_RES_CONT_ONLY && A.mask_count) ? 1.0f / *A.mask_count : 0.f;

  auto init_jets = [&](float* dst) {
    for (int i = tid; i < KP0 * M; i += NT) {
      const int k = i / M, m = i - k * M;
      const int j = m / TP, p = m - j * TP;
      float v;
      if (j == 0) v = k < d ? xin[p * PINN_MAX_IN + k] : 0.f;
      else v = (k == D.dir_cols[j - 1]) ? 1.f : 0.f;
      dst[k * MP + m] = v;
    }
  };

  for (int it = 0; it < my_tiles; ++it) {
    const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
    const long long p0 = tile * TP;
    __syncthreads();  // previous tile fully retired
    for (int i = tid; i < TP * d; i += NT) {
      const int p = i / d, c = i - p * d;
      const long long gp = p0 + p;
      xin[p * PINN_MAX_IN + c] = gp < A.n_points ? A.inputs[gp * d + c] : 0.f;
    }
    __syncthreads();
    init_jets(buf0);
    float* cur = buf0;
    float* oth = buf1;
    int poff = 0;

    // =============================== forward ===============================
    for (int l = 0; l < L; ++l) {
      const int K = D.widths[l], Nn = D.widths[l + 1];
      const int KP = pad4(K), NP = pad4(Nn);
      const bool active = 4 * cg < NP;
      const bool last = (l == L - 1);
      float bias[4] = {0.f, 0.f, 0.f, 0.f};
      if (active) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int n = 4 * cg + c;
          if (n < Nn) bias[c] = A.params[poff + K * Nn + n];
        }
      }
      float acc[J][4][4];
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[j][a][c] = 0.f;
      const int rpc = max(1, sf / NP);
      for (int r0 = 0; r0 < KP; r0 += rpc) {
        const int rows = min(rpc, KP - r0);
        const float* wc = chunk_top();
        if (active) gemm_chunk<J, TP, MP>(acc, cur + r0 * MP + 4 * pg, wc + 4 * cg, NP, rows);
      }
      if (active) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float s[4];
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const float z = acc[0][pp][c] + bias[c];
            if (last) {
              acc[0][pp][c] = z;
              s[pp] = 1.f;
            } else if (leaky) {
              s[pp] = z > 0.f ? 1.f : 0.01f;
              acc[0][pp][c] = z > 0.f ? z : 0.01f * z;
            } else {
              const float a = tanhf(z);
              acc[0][pp][c] = a;
              s[pp] = 1.f - a * a;
            }
          }
          float* dst = oth + (4 * cg + c) * MP + 4 * pg;
#pragma unroll
          for (int j = 0; j < J; ++j) {
            float4 v;
            if (j == 0) v = make_float4(acc[0][0][c], acc[0][1][c], acc[0][2][c], acc[0][3][c]);
            else v = make_float4(s[0] * acc[j][0][c], s[1] * acc[j][1][c], s[2] * acc[j][2][c],
                                 s[3] * acc[j][3][c]);
            *reinterpret_cast<float4*>(dst + j * TP) = v;
            if (BWD && !last)
              reinterpret_cast<float4*>(slab)[(size_t)((l * J + j) * 4 + c) * NT + tid] = v;
          }
        }
      }
      float* t = cur;
      cur = oth;
      oth = t;
      poff += K * Nn + Nn;
    }
    __syncthreads();  // network outputs visible to the epilogue warp

    // =============================== epilogue (warp 0; lane p < TP owns point p) ===============
    if (tid < 32) {
      const int p = tid;
      const bool isp = p < TP;
      const long long gp = p0 + p;
      FeatureMajorJets<J, TP, MP> acc{cur, p};
      float ls[PINN_NSUMS];
      residual_epilogue<J>(D, acc, isp, isp && gp < A.n_points, gp, xin + (isp ? p : 0) * PINN_MAX_IN,
                           EpiArgs{A.targets, A.seed_out, {A.seed_dout[0], A.seed_dout[1], A.seed_dout[2]},
                                   A.out, {A.dout[0], A.dout[1], A.dout[2]}, A.inv_n_res, A.inv_n_fid,
                                   inv_cnt},
                           ls);
#pragma unroll
      for (int i = 0; i < PINN_NSUMS; ++i) {
        const float v = warp_sum(ls[i]);
        if (tid == 0 && v != 0.f) red[i] += (double)v;
      }
    }

    // =============================== reverse ===============================
    if (BWD) {
      float* bz = cur;  // adjoint of pre-activations of layer l, feature-major
      float* ba = oth;  // input jets of layer l
      // Wide variant (one 256-thread CTA per SM, registers to spare): the input jets of layer l (= stored output of layer
      // l-1) are fetched from the slab one iteration ahead, so the global-load latency hides behind the previous layer's
      // weight-gradient and adjoint loops (+3 %).  The narrow variants load them in place: the 48-64 extra registers
      // cost them occupancy, which is what their large-N throughput lives on (-9 % measured).
      constexpr bool PF = NT >= 256;
      float4 pre[PF ? 4 : 1][PF ? J : 1];
      auto fetch = [&](int ll) {   // jets that iteration `ll` will copy into `ba`
        if (PF && ll >= 1 && 4 * cg < pad4(D.widths[ll])) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < J; ++j)
              pre[PF ? c : 0][PF ? j : 0] = reinterpret_cast<const float4*>(slab)[(size_t)(((ll - 1) * J + j) * 4 + c) * NT + tid];
        }
      };
      fetch(L - 1);
      for (int l = L - 1; l >= 0; --l) {
        const int K = D.widths[l], Nn = D.widths[l + 1];
        const int KP = pad4(K), NP = pad4(Nn);
        poff -= K * Nn + Nn;
        const bool act_in = 4 * cg < KP;
        __syncthreads();  // bz complete; everybody done with the buffer about to become `ba`
        if (l == 0) {
          init_jets(ba);
        } else if (act_in) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < J; ++j)
              *reinterpret_cast<float4*>(ba + (4 * cg + c) * MP + j * TP + 4 * pg) =
                  PF ? pre[PF ? c : 0][PF ? j : 0]
                     : reinterpret_cast<const float4*>(slab)[(size_t)(((l - 1) * J + j) * 4 + c) * NT + tid];
        }
        fetch(l - 1);
        __syncthreads();
        // ---- weight gradient: dW[n][k] = sum_m bz[n][m] * ba[k][m]; micro-tiles of 4x4 with
        //      interleaved rows so that a quarter-warp's LDS.128 hit 32 distinct banks ----
        {
          const int KQ = KP / 4, NQ = NP / 4;
          if ((KQ & 1) == 0) {
            // 4 x 8 micro-tiles (two interleaved K groups per thread): 12 LDS.128 per 128 FFMA instead of 8 per 64
            const int KH = KQ / 2;
            for (int mt = tid; mt < KH * NQ; mt += NT) {
              const int gn = mt / KH, g = mt - gn * KH;
              float w[4][8];
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) w[a][b] = 0.f;
              const float* zr[4];
              const float* ar[8];
#pragma unroll
              for (int a = 0; a < 4; ++a) zr[a] = bz + (gn + a * NQ) * MP;
#pragma unroll
              for (int b = 0; b < 8; ++b) ar[b] = ba + (g + (b >> 2) * KH + (b & 3) * KQ) * MP;
#pragma unroll 2
              for (int m = 0; m < M; m += 4) {
                float4 zv[4], av[8];
#pragma unroll
                for (int a = 0; a < 4; ++a) zv[a] = *reinterpret_cast<const float4*>(zr[a] + m);
#pragma unroll
                for (int b = 0; b < 8; ++b) av[b] = *reinterpret_cast<const float4*>(ar[b] + m);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                  for (int b = 0; b < 8; ++b) {
                    w[a][b] = fmaf(zv[a].x, av[b].x, w[a][b]);
                    w[a][b] = fmaf(zv[a].y, av[b].y, w[a][b]);
                    w[a][b] = fmaf(zv[a].z, av[b].z, w[a][b]);
                    w[a][b] = fmaf(zv[a].w, av[b].w, w[a][b]);
                  }
              }
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const int n = gn + a * NQ;
                if (n < Nn) {
#pragma unroll
                  for (int b = 0; b < 8; ++b) {
                    const int k = g + (b >> 2) * KH + (b & 3) * KQ;
                    if (k < K) atomicAdd(A.grad + poff + n * K + k, w[a][b]);
                  }
                }
              }
            }
          } else
          for (int mt = tid; mt < KQ * NQ; mt += NT) {
            const int gn = mt / KQ, g = mt - gn * KQ;
            float w[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) w[a][b] = 0.f;
            const float* zr[4];
            const float* ar[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              zr[a] = bz + (gn + a * NQ) * MP;
              ar[a] = ba + (g + a * KQ) * MP;
            }
#pragma unroll 4
            for (int m = 0; m < M; m += 4) {
              float4 zv[4], av[4];
#pragma unroll
              for (int a = 0; a < 4; ++a) zv[a] = *reinterpret_cast<const float4*>(zr[a] + m);
#pragma unroll
              for (int b = 0; b < 4; ++b) av[b] = *reinterpret_cast<const float4*>(ar[b] + m);
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  w[a][b] = fmaf(zv[a].x, av[b].x, w[a][b]);
                  w[a][b] = fmaf(zv[a].y, av[b].y, w[a][b]);
                  w[a][b] = fmaf(zv[a].z, av[b].z, w[a][b]);
                  w[a][b] = fmaf(zv[a].w, av[b].w, w[a][b]);
                }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const int n = gn + a * NQ;
              if (n < Nn) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  const int k = g + b * KQ;
                  if (k < K) atomicAdd(A.grad + poff + n * K + k, w[a][b]);
                }
              }
            }
          }
          for (int n = tid; n < Nn; n += NT) {
            float s = 0.f;
            for (int p = 0; p < TP; ++p) s += bz[n * MP + p];
            atomicAdd(A.grad + poff + K * Nn + n, s);
          }
        }
        if (l == 0) break;
        // ---- adjoint of the layer input: abar[m][k] = sum_n bz[n][m] * Wn[n][k] ----
        float acc[J][4][4];
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[j][a][c] = 0.f;
        const int rpc = max(1, sf / KP);
        for (int r0 = 0; r0 < NP; r0 += rpc) {
          const int rows = min(rpc, NP - r0);
          const float* wc = chunk_top();
          if (act_in) gemm_chunk<J, TP, MP>(acc, bz + r0 * MP + 4 * pg, wc + 4 * cg, KP, rows);
        }
        // ---- through the activation of layer l-1: with a' and adot'_j = s*zdot_j stored,
        //      zbar = abar*s - 2a' * sum_j adotbar_j * adot'_j ;  zdotbar_j = adotbar_j * s ----
        if (act_in) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float* q = ba + (4 * cg + c) * MP + 4 * pg;
            const float4 a4 = *reinterpret_cast<const float4*>(q);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            float s[4], zb[4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              s[pp] = leaky ? (av[pp] > 0.f ? 1.f : 0.01f) : 1.f - av[pp] * av[pp];
              zb[pp] = acc[0][pp][c] * s[pp];
            }
#pragma unroll
            for (int j = 1; j < J; ++j) {
              const float4 t4 = *reinterpret_cast<const float4*>(q + j * TP);
              const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
              if (!leaky) {
#pragma unroll
                for (int pp = 0; pp < 4; ++pp)
                  zb[pp] = fmaf(-2.f * av[pp] * acc[j][pp][c], tv[pp], zb[pp]);
              }
              *reinterpret_cast<float4*>(q + j * TP) =
                  make_float4(acc[j][0][c] * s[0], acc[j][1][c] * s[1], acc[j][2][c] * s[2],
                              acc[j][3][c] * s[3]);
            }
            *reinterpret_cast<float4*>(q) = make_float4(zb[0], zb[1], zb[2], zb[3]);
          }
        }
        float* t = bz;
        bz = ba;
        ba = t;
      }
    }
  }

  __syncthreads();
  if (tid < PINN_NSUMS && A.sums && red[tid] != 0.0) atomicAdd(A.sums + tid, red[tid]);
}

// --------------------------------------------------------------------------- auxiliary kernels
// Packs layer l = blockIdx.y: WT[k][n] = W[n][k] and Wn[n][k] = W[n][k], zero padded to multiples of 4.
__global__ void pack_kernel(const __grid_constant__ pinn_desc_t D, const float* __restrict__ params,
                            float* __restrict__ packed) {
  const int l = blockIdx.y;
  int poff = 0, pk = 0;
  for (int i = 0; i < l; ++i) {
    poff += D.widths[i] * D.widths[i + 1] + D.widths[i + 1];
    pk += 2 * pad4(D.widths[i]) * pad4(D.widths[i + 1]);
  }
  const int K = D.widths[l], Nn = D.widths[l + 1];
  const int KP = pad4(K), NP = pad4(Nn);
  const int total = KP * NP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    {  // WT
      const int k = i / NP, n = i - k * NP;
      packed[pk + i] = (k < K && n < Nn) ? params[poff + n * K + k] : 0.f;
    }
    {  // Wn
      const int n = i / KP, k = i - n * KP;
      packed[pk + total + i] = (k < K && n < Nn) ? params[poff + n * K + k] : 0.f;
    }
  }
}

__global__ void mask_count_kernel(const float* __restrict__ inputs, long long n, int d, int col,
                                  float thr, float* __restrict__ out) {
  float c = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    c += inputs[i * d + col] < thr ? 1.f : 0.f;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c != 0.f) atomicAdd(out, c);  // exact: integer-valued partials
}

__global__ void finalize_kernel(const __grid_constant__ pinn_desc_t D, const double* __restrict__ s,
                                const double* __restrict__ sb, long long n_fid, long long n_res,
                                const float* __restrict__ mask_count, float* __restrict__ parts) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double v[PINN_NSUMS];
  for (int i = 0; i < PINN_NSUMS; ++i) v[i] = s[i] + (sb ? sb[i] : 0.0);
  double fid = 0.0;
  for (int i = 0; i < D.n_targets; ++i)
    fid += (double)D.target_w[i] * (v[PINN_SUM_TARGET0 + i] / (double)n_fid);
  double res = 0.0;
  const int k = D.residual_kind;
  if (k == PINN_RES_CONT_ONLY || k == PINN_RES_CONT_FTEMP) res = v[PINN_SUM_FC] / (double)n_res;
  if (k == PINN_RES_NSWE || k == PINN_RES_WAVE_AVG || k == PINN_RES_BOUSSINESQ || k == PINN_RES_BOUSS_SIMPLE)
    res = v[PINN_SUM_FC] / (double)n_res + v[PINN_SUM_FX] / (double)n_res +
          v[PINN_SUM_FY] / (double)n_res;
  if (k == PINN_RES_CONT_ONLY) {
    const double cnt = mask_count ? (double)*mask_count : v[PINN_SUM_MASKCNT];
    res += v[PINN_SUM_COND] / cnt;  // 0/0 = NaN for an empty mask, like torch.mean of nothing
  }
  parts[0] = (float)fid;
  parts[1] = (float)res;
  parts[2] = (float)((double)D.w_fid * fid + (double)D.w_res * res);
  parts[3] = 0.f;
}

// --------------------------------------------------------------------------- host side
// jet_tc.cu
bool tc_supported(const pinn_desc_t* D, const char** why);
int tc_workspace(const pinn_desc_t* D, long long n_points, int sms, size_t* packed_bytes, size_t* slab_bytes,
                 long long* slab_stride, int* grid);
int run_tc_pass(const pinn_desc_t* D, const pinn_eval_args_t* a, bool bwd, void* workspace, size_t ws_bytes,
                cudaStream_t st);

// jet3.cu
bool jet3_supported(const pinn_desc_t* D, const char** why);
int jet3_workspace(const pinn_desc_t* D, long long n_points, bool bwd, size_t* bytes, int* grid_out, long long* stride_out,
                   int* wmax_out);
int run_jet3_pass(const pinn_desc_t* D, const pinn_eval_args_t* a, bool bwd, cudaStream_t st);

// residual kinds the first-order kernels evaluate (BOUSSINESQ needs third-order jets: jet3.cu)
static bool is_residual_kind(int k) { return (k >= PINN_RES_CONT_ONLY && k <= PINN_RES_WAVE_AVG) || k == PINN_RES_BOUSS_SIMPLE; }

// Which kernel runs this pass?  TF32 is only defined for PDE-residual passes of 256-wide tanh nets; value-only
// (fidelity) and external-seed passes always use the FP32 kernel.  Anything else asked for in TF32 is an error.
static int uses_tc(const pinn_desc_t* D, bool* tc) {
  *tc = false;
  if (D->precision == PINN_PREC_FP32 || !is_residual_kind(D->residual_kind)) return PINN_OK;
  const char* why = "";
  if (!tc_supported(D, &why))
    return set_error("precision %s is not available for this net: %s",
                     D->precision == PINN_PREC_TF32X3 ? "tf32x3" : "tf32", why), PINN_E_UNSUPPORTED;
  *tc = true;
  return PINN_OK;
}

struct Config {
  int J, TP, NT;
  int wp, stage_floats;
  size_t smem;
  long long slab_stride;  // floats per CTA
  long long packed_floats;
  int max_ctas;
};

static int jets_for(const pinn_desc_t* D) {
  switch (D->residual_kind) {
    case PINN_RES_NONE: return 1;
    case PINN_RES_CONT_ONLY:
    case PINN_RES_CONT_FTEMP:
    case PINN_RES_WAVE_AVG: return 3;
    case PINN_RES_NSWE:
    case PINN_RES_BOUSSINESQ:
    case PINN_RES_BOUSS_SIMPLE: return 4;
    case PINN_RES_EXTERNAL: return 1 + D->n_dirs;
  }
  return -1;
}

int validate_desc(const pinn_desc_t* D) {
  if (!D) return set_error("desc is NULL"), PINN_E_ARG;
  const int L = D->n_linear;
  if (L < 1 || L > PINN_MAX_LINEAR) return set_error("n_linear %d out of [1,%d]", L, PINN_MAX_LINEAR), PINN_E_ARG;
  for (int i = 0; i <= L; ++i)
    if (D->widths[i] < 1 || D->widths[i] > PINN_MAX_WIDTH)
      return set_error("layer width %d at %d out of [1,%d]", D->widths[i], i, PINN_MAX_WIDTH), PINN_E_UNSUPPORTED;
  if (D->widths[0] > PINN_MAX_IN) return set_error("more than %d input features", PINN_MAX_IN), PINN_E_UNSUPPORTED;
  if (D->widths[L] > PINN_MAX_OUT) return set_error("more than %d output features", PINN_MAX_OUT), PINN_E_UNSUPPORTED;
  if (D->activation != PINN_ACT_TANH && D->activation != PINN_ACT_LEAKY_RELU)
    return set_error("unknown activation %d", D->activation), PINN_E_ARG;
  const int k = D->residual_kind;
  if (k < PINN_RES_NONE || k > PINN_RES_BOUSS_SIMPLE) return set_error("unknown residual kind %d", k), PINN_E_ARG;
  const bool txy = k == PINN_RES_NSWE || k == PINN_RES_BOUSSINESQ || k == PINN_RES_BOUSS_SIMPLE;
  const int want_dirs = (k == PINN_RES_NONE) ? 0 : txy ? 3 : (k == PINN_RES_EXTERNAL) ? D->n_dirs : 2;
  if (D->n_dirs != want_dirs || D->n_dirs < 0 || D->n_dirs > PINN_MAX_DIRS)
    return set_error("residual kind %d needs %d differentiated directions, got %d", k, want_dirs, D->n_dirs), PINN_E_ARG;
  for (int j = 0; j < D->n_dirs; ++j)
    if (D->dir_cols[j] < 0 || D->dir_cols[j] >= D->widths[0])
      return set_error("dir_cols[%d]=%d is not an input column", j, D->dir_cols[j]), PINN_E_ARG;
  const int nf = (k == PINN_RES_CONT_ONLY || k == PINN_RES_CONT_FTEMP) ? 3 : txy ? 4 : (k == PINN_RES_WAVE_AVG) ? 6 : 0;
  for (int f = 0; f < nf; ++f) {
    if (D->field_cols[f] < 0 || D->field_cols[f] >= D->widths[L])
      return set_error("field_cols[%d]=%d is not an output column", f, D->field_cols[f]), PINN_E_ARG;
    for (int g = 0; g < f; ++g)
      if (D->field_cols[g] == D->field_cols[f]) return set_error("field_cols must be distinct"), PINN_E_ARG;
  }
  if (k == PINN_RES_CONT_ONLY && (D->mask_col < 0 || D->mask_col >= D->widths[0]))
    return set_error("mask_col %d is not an input column", D->mask_col), PINN_E_ARG;
  if (D->n_targets < 0 || D->n_targets > PINN_MAX_OUT) return set_error("n_targets out of range"), PINN_E_ARG;
  for (int i = 0; i < D->n_targets; ++i)
    if (D->target_cols[i] < 0 || D->target_cols[i] >= D->widths[L])
      return set_error("target_cols[%d]=%d is not an output column", i, D->target_cols[i]), PINN_E_ARG;
  if (D->precision != PINN_PREC_FP32 && D->precision != PINN_PREC_TF32 && D->precision != PINN_PREC_TF32X3)
    return set_error("unknown precision %d", D->precision), PINN_E_ARG;
  return PINN_OK;
}

// Opt this instantiation in to `smem` bytes of dynamic shared memory on the current device.  The limit is only ever
// RAISED (per instantiation and device), so a cached configuration of a larger net stays launchable after a smaller one.
template <int J, int TP, int NT, bool BWD>
static int ensure_smem(size_t smem) {
  static size_t allowed[64] = {0};
  int dev = 0;
  PINN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > allowed[dev]) {
    PINN_CUDA(cudaFuncSetAttribute(jet_kernel<J, TP, NT, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) allowed[dev] = smem;
  }
  return PINN_OK;
}

template <int J, int TP, int NT, bool BWD>
static int launch_t(const pinn_desc_t* D, const KArgs& A, const Config& c, int grid, cudaStream_t st) {
  auto kern = jet_kernel<J, TP, NT, BWD>;
  int rc = ensure_smem<J, TP, NT, BWD>(c.smem);
  if (rc) return rc;
  kern<<<grid, NT, c.smem, st>>>(*D, A);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

template <int J, int TP, int NT, bool BWD>
static int occupancy_t(size_t smem, int* blocks) {
  auto kern = jet_kernel<J, TP, NT, BWD>;
  int rc = ensure_smem<J, TP, NT, BWD>(smem);
  if (rc) return rc;
  PINN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kern, NT, smem));
  return PINN_OK;
}

// tile shape by the widest (padded) layer: narrow nets keep more points per CTA, wide ones more
// column threads.  (J, TP, NT)
#define PINN_DISPATCH(J_, WCLASS, BWD_, FN, ...)                                   \
  do {                                                                             \
    if (WCLASS == 0) return FN<J_, 32, 64, BWD_>(__VA_ARGS__);                      \
    if (WCLASS == 1) return FN<J_, 32, 128, BWD_>(__VA_ARGS__);                     \
    if (WCLASS == 2) return FN<J_, 16, 128, BWD_>(__VA_ARGS__);                     \
    return FN<J_, 16, 256, BWD_>(__VA_ARGS__);                                      \
  } while (0)

static int wclass_of(int wp) { return wp <= 32 ? 0 : wp <= 64 ? 1 : wp <= 128 ? 2 : 3; }

static int occupancy(int J, int wclass, bool bwd, size_t smem, int* blocks) {
  if (bwd) {
    if (J == 1) PINN_DISPATCH(1, wclass, true, occupancy_t, smem, blocks);
    if (J == 2) PINN_DISPATCH(2, wclass, true, occupancy_t, smem, blocks);
    if (J == 3) PINN_DISPATCH(3, wclass, true, occupancy_t, smem, blocks);
    PINN_DISPATCH(4, wclass, true, occupancy_t, smem, blocks);
  }
  if (J == 1) PINN_DISPATCH(1, wclass, false, occupancy_t, smem, blocks);
  if (J == 2) PINN_DISPATCH(2, wclass, false, occupancy_t, smem, blocks);
  if (J == 3) PINN_DISPATCH(3, wclass, false, occupancy_t, smem, blocks);
  PINN_DISPATCH(4, wclass, false, occupancy_t, smem, blocks);
}

static int launch(int J, int wclass, bool bwd, const pinn_desc_t* D, const KArgs& A, const Config& c,
                  int grid, cudaStream_t st) {
  if (bwd) {
    if (J == 1) PINN_DISPATCH(1, wclass, true, launch_t, D, A, c, grid, st);
    if (J == 2) PINN_DISPATCH(2, wclass, true, launch_t, D, A, c, grid, st);
    if (J == 3) PINN_DISPATCH(3, wclass, true, launch_t, D, A, c, grid, st);
    PINN_DISPATCH(4, wclass, true, launch_t, D, A, c, grid, st);
  }
  if (J == 1) PINN_DISPATCH(1, wclass, false, launch_t, D, A, c, grid, st);
  if (J == 2) PINN_DISPATCH(2, wclass, false, launch_t, D, A, c, grid, st);
  if (J == 3) PINN_DISPATCH(3, wclass, false, launch_t, D, A, c, grid, st);
  PINN_DISPATCH(4, wclass, false, launch_t, D, A, c, grid, st);
}

static int make_config_uncached(const pinn_desc_t* D, bool bwd, Config* c);

// The configuration (tile shape, shared memory, occupancy query) depends only on the descriptor and the device: keep the
// last few so that the per-evaluation host path is a memcmp, not three runtime-API queries.
int make_config(const pinn_desc_t* D, bool bwd, Config* c) {
  if (!D) return set_error("desc is NULL"), PINN_E_ARG;
  struct Entry {
    pinn_desc_t d;
    int dev, bwd, valid;
    Config c;
  };
  static thread_local Entry cache[8];
  static thread_local int next = 0;
  int dev = 0;
  PINN_CUDA(cudaGetDevice(&dev));
  for (int i = 0; i < 8; ++i)
    if (cache[i].valid && cache[i].dev == dev && cache[i].bwd == (int)bwd && memcmp(&cache[i].d, D, sizeof(pinn_desc_t)) == 0) {
      *c = cache[i].c;
      return PINN_OK;
    }
  int rc = make_config_uncached(D, bwd, c);
  if (rc) return rc;
  Entry& e = cache[next];
  next = (next + 1) % 8;
  memcpy(&e.d, D, sizeof(pinn_desc_t));
  e.dev = dev, e.bwd = (int)bwd, e.valid = 1, e.c = *c;
  return PINN_OK;
}

static int make_config_uncached(const pinn_desc_t* D, bool bwd, Config* c) {
  int rc = validate_desc(D);
  if (rc) return rc;
  const int J = jets_for(D);
  if (J < 1 || J > 4) return set_error("unsupported jet count %d", J), PINN_E_UNSUPPORTED;
  int wp = 0;
  long long packed = 0;
  for (int i = 0; i <= D->n_linear; ++i) wp = wp > pad4(D->widths[i]) ? wp : pad4(D->widths[i]);
  for (int i = 0; i < D->n_linear; ++i) packed += 2LL * pad4(D->widths[i]) * pad4(D->widths[i + 1]);
  const int wc = wclass_of(wp);
  c->J = J;
  c->TP = wc <= 1 ? 32 : 16;
  c->NT = wc == 0 ? 64 : wc == 3 ? 256 : 128;
  c->wp = wp;
  // weight rows per ring stage: whole matrices for narrow nets, 16 rows for widths up to 128, 28 rows for the 256-wide
  // nets (fewer chunk barriers: +4 % measured; 3 x 28 KB is what still fits beside the two jet buffers)
  c->stage_floats = wp <= 32 ? wp * wp : wp > 128 ? wp * 28 : wp * 16;
  const int MP = J * c->TP + 4;
  size_t smem = size_t(2) * wp * MP * 4 + size_t(kStages) * c->stage_floats * 4 +
                size_t(c->TP) * PINN_MAX_IN * 4 + PINN_NSUMS * 8 + kStages * 8;
  c->smem = smem;
  c->slab_stride = bwd ? (long long)(D->n_linear - 1) * J * 4 * c->NT * 4 : 0;
  c->packed_floats = packed;
  int dev = 0, sms = 0, blocks = 0;
  PINN_CUDA(cudaGetDevice(&dev));
  PINN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  rc = occupancy(J, wc, bwd, smem, &blocks);
  if (rc) return rc;
  if (blocks < 1) return set_error("kernel does not fit an SM (smem %zu B)", smem), PINN_E_UNSUPPORTED;
  c->max_ctas = sms * blocks;
  return PINN_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int workspace_bytes(const pinn_desc_t* D, long long n_points, bool bwd, size_t* bytes) {
  Config c;
  int rc = validate_desc(D);
  if (rc) return rc;
  if (D->residual_kind == PINN_RES_BOUSSINESQ) {
    const char* why = "";
    if (!jet3_supported(D, &why)) return set_error("Boussinesq residual: %s", why), PINN_E_UNSUPPORTED;
    return jet3_workspace(D, n_points, bwd, bytes, nullptr, nullptr, nullptr);
  }
  bool tc = false;
  rc = uses_tc(D, &tc);
  if (rc) return rc;
  if (tc) {
    int dev = 0, sms = 0, grid = 1;
    PINN_CUDA(cudaGetDevice(&dev));
    PINN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    size_t pk = 0, sl = 0;
    long long stride = 0;
    tc_workspace(D, n_points, sms, &pk, &sl, &stride, &grid);
    *bytes = align256(pk) + align256(sl) + 256;
    return PINN_OK;
  }
  rc = make_config(D, bwd, &c);   // (forward-only: no activation slabs)
  if (rc) return rc;
  long long tiles = (n_points + c.TP - 1) / c.TP;
  long long grid = tiles < c.max_ctas ? tiles : c.max_ctas;
  if (grid < 1) grid = 1;
  *bytes = align256((size_t)c.packed_floats * 4) + align256((size_t)grid * c.slab_stride * 4) + 256;
  return PINN_OK;
}

int run_pass(const pinn_desc_t* D, const pinn_eval_args_t* a, bool bwd, cudaStream_t st) {
  if (!a) return set_error("args is NULL"), PINN_E_ARG;
  Config c;
  int rc = make_config(D, bwd, &c);
  if (rc) return rc;
  if (a->n_points < 0) return set_error("n_points < 0"), PINN_E_ARG;
  if (!a->params || (!a->inputs && a->n_points > 0) || !a->workspace)
    return set_error("params / inputs / workspace must be non-NULL"), PINN_E_ARG;
  if (bwd && !a->grad) return set_error("grad must be non-NULL for fwdbwd"), PINN_E_ARG;
  if (D->n_targets > 0 && a->targets && a->n_fid_global <= 0)
    return set_error("n_fid_global must be positive when targets are given"), PINN_E_ARG;
  if ((is_residual_kind(D->residual_kind) || D->residual_kind == PINN_RES_BOUSSINESQ) && a->n_res_global <= 0)
    return set_error("n_res_global must be positive"), PINN_E_ARG;
  if (D->residual_kind == PINN_RES_CONT_ONLY && !a->mask_count)
    return set_error("continuity_only needs mask_count (see pinn_mask_count)"), PINN_E_ARG;
  if (((uintptr_t)a->workspace & 255) != 0) return set_error("workspace must be 256-byte aligned"), PINN_E_ARG;
  bool tc = false;
  rc = uses_tc(D, &tc);
  if (rc) return rc;
  if (((uintptr_t)a->params & 15) != 0) return set_error("params must be 16-byte aligned"), PINN_E_ARG;

  long long P = 0;
  for (int i = 0; i < D->n_linear; ++i) P += (long long)D->widths[i] * D->widths[i + 1] + D->widths[i + 1];
  if (D->residual_kind == PINN_RES_BOUSSINESQ) {
    const char* why = "";
    if (!jet3_supported(D, &why)) return set_error("Boussinesq residual: %s", why), PINN_E_UNSUPPORTED;
    if (!(a->flags & PINN_FLAG_ACCUMULATE)) {
      if (bwd) PINN_CUDA(cudaMemsetAsync(a->grad, 0, (size_t)P * 4, st));
      if (a->sums) PINN_CUDA(cudaMemsetAsync(a->sums, 0, PINN_NSUMS * 8, st));
    }
    return run_jet3_pass(D, a, bwd, st);
  }
  if (tc) {
    if (!(a->flags & PINN_FLAG_ACCUMULATE)) {
      if (bwd) PINN_CUDA(cudaMemsetAsync(a->grad, 0, (size_t)P * 4, st));
      if (a->sums) PINN_CUDA(cudaMemsetAsync(a->sums, 0, PINN_NSUMS * 8, st));
    }
    return run_tc_pass(D, a, bwd, a->workspace, a->workspace_bytes, st);
  }
  long long tiles = (a->n_points + c.TP - 1) / c.TP;
  long long grid = tiles < c.max_ctas ? tiles : c.max_ctas;
  const size_t need = align256((size_t)c.packed_floats * 4) +
                      align256((size_t)(grid > 0 ? grid : 1) * c.slab_stride * 4);
  if (a->workspace_bytes < need)
    return set_error("workspace too small: %zu < %zu bytes", a->workspace_bytes, need), PINN_E_WORKSPACE;

  float* packed = reinterpret_cast<float*>(a->workspace);
  float* slab = reinterpret_cast<float*>(reinterpret_cast<char*>(a->workspace) +
                                         align256((size_t)c.packed_floats * 4));
  if (!(a->flags & PINN_FLAG_ACCUMULATE)) {
    if (bwd) PINN_CUDA(cudaMemsetAsync(a->grad, 0, (size_t)P * 4, st));
    if (a->sums) PINN_CUDA(cudaMemsetAsync(a->sums, 0, PINN_NSUMS * 8, st));
  }
  if (!(a->flags & PINN_FLAG_SKIP_PACK)) {
    int maxl = 0;
    for (int i = 0; i < D->n_linear; ++i) {
      const int t = pad4(D->widths[i]) * pad4(D->widths[i + 1]);
      maxl = maxl > t ? maxl : t;
    }
    dim3 g((maxl + 255) / 256, D->n_linear);
    pack_kernel<<<g, 256, 0, st>>>(*D, a->params, packed);
    PINN_CUDA(cudaGetLastError());
  }
  if (tiles == 0) return PINN_OK;

  KArgs A;
  A.params = a->params;
  A.packed = packed;
  A.inputs = a->inputs;
  A.targets = D->n_targets > 0 ? a->targets : nullptr;
  A.mask_count = a->mask_count;
  A.seed_out = a->seed_out;
  for (int j = 0; j < PINN_MAX_DIRS; ++j) {
    A.seed_dout[j] = a->seed_dout[j];
    A.dout[j] = a->dout[j];
  }
  A.grad = a->grad;
  A.sums = a->sums;
  A.out = a->out;
  A.slab = slab;
  A.slab_stride = c.slab_stride;
  A.n_points = a->n_points;
  A.n_tiles = (int)tiles;
  A.stage_floats = c.stage_floats;
  A.wp = c.wp;
  A.inv_n_res = a->n_res_global > 0 ? (float)(1.0 / (double)a->n_res_global) : 0.f;
  A.inv_n_fid = a->n_fid_global > 0 ? (float)(1.0 / (double)a->n_fid_global) : 0.f;
  if (tiles > 0x7fffffffLL) return set_error("too many tiles"), PINN_E_UNSUPPORTED;
  return launch(c.J, wclass_of(c.wp), bwd, D, A, c, (int)grid, st);
}

int run_mask_count(const pinn_desc_t* D, const float* inputs, long long n, float* out, cudaStream_t st) {
  PINN_CUDA(cudaMemsetAsync(out, 0, 4, st));
  if (n == 0) return PINN_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  mask_count_kernel<<<(int)blocks, 256, 0, st>>>(inputs, n, D->widths[0], D->mask_col, D->cond_threshold, out);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

int run_finalize(const pinn_desc_t* D, const double* s, const double* sb, long long n_fid, long long n_res,
                 const float* mask_count, float* parts, cudaStream_t st) {
  finalize_kernel<<<1, 32, 0, st>>>(*D, s, sb, n_fid, n_res, mask_count, parts);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

}  // namespace pinn
