// Third-order Taylor-jet kernel for the reference's historical fully-nonlinear Boussinesq residual
// (__pycache__/physics_functions.cpython-38.pyc, `Boussinesq`, source lines 55-130 -- decompiled by tools/pyc38_decompile.py and
// re-typed in oracle/boussinesq_oracle.py).  The reference obtains u_2, V1, V2, V3 ... by NESTING
// torch.autograd.grad three levels deep; here every field is carried as its truncated Taylor polynomial in (t, x, y):
//
//   jets        20 normalised Taylor coefficients c_alpha = d^alpha f / alpha!, |alpha| <= 3, per feature and point
//   Linear      acts on every coefficient plane separately (the bias only on c_0): one [20 TP x K] x [K x N] product per layer
//   tanh        polynomial composition y = T0 + T1 d + T2 d^2 + T3 d^3 (d = z - z_0, T_k = tanh^(k)(z_0) / k!), truncated
//   residual    the decompiled formulas, evaluated in POLYNOMIAL ARITHMETIC by a small register machine (one warp per point, one
//               lane per coefficient): products are truncated convolutions, compute_gradient(., x) is a coefficient shift; only
//               the constant terms of f_cont, f_momx, f_momy enter the loss
//   reverse     the same register program walked backwards gives d loss / d (output coefficients); through tanh the adjoint is
//               the correlation with p = f'(z) = 1 - y^2 (because dy = p dz as truncated polynomials); dW += sum_alpha zbar_alpha
//               a_alpha^T, abar_alpha = W^T zbar_alpha plane by plane -- no autograd graph, no nested differentiation
//
// Scope: FP32, every layer width <= 64, exactly three differentiated directions (t, x, y), four fields (h, z, u, v).  This is a
// "next" row of the scope table: written for correctness and to keep the whole evaluation on the device (one launch); the
// per-layer products are plain shared-memory FMA loops, not the register-tiled / tensor-core contractions of the first-order
// kernels.
#include "common.cuh"

namespace pinn {

constexpr int J3 = 20;            // coefficients per jet
constexpr int J3_TP = 8;          // points per tile
constexpr int J3_M = J3 * J3_TP;  // columns of a layer product
constexpr int J3_NT = 160;        // threads (= columns)
constexpr int J3_MAXW = 64;
constexpr int J3_NREG = 116;      // polynomial registers of the residual program

struct J3Tables {
  signed char up[3][J3];          // index of alpha + e_d (or -1)
  signed char down[3][J3];        // index of alpha - e_d (or -1)
  unsigned char ex[3][J3];        // exponents
  unsigned char fwd_n[J3];        // products contributing to coefficient k: pairs (i, j), alpha_i + alpha_j = alpha_k
  unsigned char fwd_i[J3][8], fwd_j[J3][8];
  unsigned char rev_n[J3];        // for operand coefficient i: pairs (j, k) with alpha_k = alpha_i + alpha_j
  unsigned char rev_j[J3][J3], rev_k[J3][J3];
};
__constant__ J3Tables kJ3;

// residual program: one instruction per line of the decompiled function
enum { J3_DER = 0, J3_MUL = 1, J3_LIN = 2 };
struct J3Ins {
  unsigned char op, dst, a, b;    // DER: b = direction (0 t, 1 x, 2 y); LIN: b = 255 -> single operand
  float fa, fb;
};
constexpr int J3_NINS = J3_NREG - 4;
__constant__ J3Ins kJ3Prog[J3_NINS];

struct J3Args {
  const float* params;
  const float* inputs;
  const float* targets;
  float* grad;
  double* sums;
  float* out;
  float* slab;            // per CTA: (L-1) x [W][J3_M] post-activation jets
  long long slab_stride;
  long long n_points;
  int n_tiles;
  int wmax;
  float inv_n_res, inv_n_fid;
};

__device__ __forceinline__ float j3_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// c = a (x) b truncated at degree 3 (per-thread arrays)
__device__ __forceinline__ void j3_pmul(const float* a, const float* b, float* c) {
  for (int k = 0; k < J3; ++k) {
    float s = 0.f;
    for (int q = 0; q < kJ3.fwd_n[k]; ++q) s = fmaf(a[kJ3.fwd_i[k][q]], b[kJ3.fwd_j[k][q]], s);
    c[k] = s;
  }
}
// abar_i += sum_{(j,k)} b_j cbar_k   (adjoint of c = a (x) b with respect to a)
__device__ __forceinline__ void j3_pcorr(const float* b, const float* cbar, float* abar) {
  for (int i = 0; i < J3; ++i) {
    float s = 0.f;
    for (int q = 0; q < kJ3.rev_n[i]; ++q) s = fmaf(b[kJ3.rev_j[i][q]], cbar[kJ3.rev_k[i][q]], s);
    abar[i] += s;
  }
}

__global__ void __launch_bounds__(J3_NT)
    jet3_kernel(const __grid_constant__ pinn_desc_t D, const __grid_constant__ J3Args A, int bwd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int WM = A.wmax;
  float* bufA = reinterpret_cast<float*>(smem_raw);     // [WM][J3_M]
  float* bufB = bufA + WM * J3_M;                       // [WM][J3_M]
  float* wsm = bufB + WM * J3_M;                        // [WM][WM] weights of the current layer, row n = output feature
  float* xin = wsm + WM * WM;                           // [TP][8]
  double* red = reinterpret_cast<double*>(xin + J3_TP * PINN_MAX_IN);   // [16]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = D.n_linear;
  const int d = D.widths[0], o = D.widths[L];
  float* slab = A.slab + (long long)blockIdx.x * A.slab_stride;
  if (tid < PINN_NSUMS) red[tid] = 0.0;
  __syncthreads();

  for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
    const long long p0 = (long long)tile * J3_TP;
    for (int i = tid; i < J3_TP * PINN_MAX_IN; i += J3_NT) {
      const int p = i / PINN_MAX_IN, c = i - p * PINN_MAX_IN;
      const long long gp = p0 + p;
      xin[i] = (c < d && gp < A.n_points) ? A.inputs[gp * d + c] : 0.f;
    }
    __syncthreads();
    // input jets: value, and a unit first-degree coefficient for the differentiated columns (t, x, y)
    auto init_jets = [&](float* dst) {
      for (int i = tid; i < d * J3_M; i += J3_NT) {
        const int k = i / J3_M, m = i - k * J3_M;
        const int c = m / J3_TP, p = m - c * J3_TP;
        float v = 0.f;
        if (c == 0) v = xin[p * PINN_MAX_IN + k];
        else if (c <= 3 && D.dir_cols[c - 1] == k) v = 1.f;
        dst[k * J3_M + m] = v;
      }
    };
    init_jets(bufA);
    float* cur = bufA;
    float* oth = bufB;
    long long poff = 0;

    // =============================== forward ===============================
    for (int l = 0; l < L; ++l) {
      const int K = D.widths[l], Nn = D.widths[l + 1];
      __syncthreads();
      for (int i = tid; i < Nn * K; i += J3_NT) wsm[i] = A.params[poff + i];
      __syncthreads();
      {   // thread = column m (coefficient plane c, point p): z[n][m] = sum_k W[n][k] a[k][m] (+ b[n] on plane 0)
        const int m = tid;
        const bool plane0 = m < J3_TP;
        for (int n = 0; n < Nn; ++n) {
          float acc = plane0 ? A.params[poff + (long long)K * Nn + n] : 0.f;
          const float* wr = wsm + n * K;
          for (int k = 0; k < K; ++k) acc = fmaf(wr[k], cur[k * J3_M + m], acc);
          oth[n * J3_M + m] = acc;
        }
      }
      __syncthreads();
      if (l < L - 1) {
        // tanh as a polynomial composition, thread per (feature, point)
        for (int i = tid; i < Nn * J3_TP; i += J3_NT) {
          const int n = i / J3_TP, p = i - n * J3_TP;
          float* zp = oth + n * J3_M + p;
          float dl[J3], d2[J3], d3[J3];
          const float a = tanhf(zp[0]);
          const float s = 1.f - a * a;
          const float T1 = s, T2 = -a * s, T3 = s * (2.f * a * a - s) * (1.f / 3.f);
          dl[0] = 0.f;
          for (int c = 1; c < J3; ++c) dl[c] = zp[c * J3_TP];
          j3_pmul(dl, dl, d2);
          j3_pmul(d2, dl, d3);
          zp[0] = a;
          for (int c = 1; c < J3; ++c) zp[c * J3_TP] = fmaf(T1, dl[c], fmaf(T2, d2[c], T3 * d3[c]));
        }
        __syncthreads();
        if (bwd) {   // keep the post-activation jets of this layer for the reverse sweep
          float* sl = slab + (size_t)l * WM * J3_M;
          for (int i = tid; i < Nn * J3_M; i += J3_NT) sl[i] = oth[i];
        }
      }
      float* t_ = cur;
      cur = oth;
      oth = t_;
      poff += (long long)K * Nn + Nn;
    }
    __syncthreads();

    // =============================== residual: one warp per point, one lane per coefficient ===============================
    for (int p = warp; p < J3_TP; p += J3_NT / 32) {
      const long long gp = p0 + p;
      const bool valid = gp < A.n_points;
      const int c = lane < J3 ? lane : 0;
      const bool act = lane < J3;
      float R[J3_NREG];
      // fields in the argument order of the residual: h, z, u, v
      for (int f = 0; f < 4; ++f) R[f] = act ? cur[D.field_cols[f] * J3_M + c * J3_TP + p] : 0.f;
      if (A.out && valid && lane < o) A.out[gp * o + lane] = cur[lane * J3_M + p];
      // data misfit on the values (plane 0)
      float terr = 0.f;
      if (A.targets && valid && lane < D.n_targets) {
        terr = cur[D.target_cols[lane] * J3_M + p] - A.targets[gp * D.n_targets + lane];
        const float sq = terr * terr;
        // (one lane per target)
        atomicAdd(&red[PINN_SUM_TARGET0 + lane], (double)sq);
      }
      auto run_fwd = [&](const J3Ins& I) {
        float v;
        if (I.op == J3_DER) {
          const int dir = I.b;
          const int src = kJ3.up[dir][c];
          const float x = __shfl_sync(0xffffffffu, R[I.a], src < 0 ? 0 : src);
          v = src < 0 ? 0.f : x * (float)(kJ3.ex[dir][c] + 1);
        } else if (I.op == J3_MUL) {
          v = 0.f;
          const float ra = R[I.a], rb = R[I.b];
          for (int q = 0; q < 8; ++q) {
            const bool on = q < kJ3.fwd_n[c];
            const float x = __shfl_sync(0xffffffffu, ra, on ? kJ3.fwd_i[c][q] : 0);
            const float y = __shfl_sync(0xffffffffu, rb, on ? kJ3.fwd_j[c][q] : 0);
            if (on) v = fmaf(x, y, v);
          }
        } else {
          v = I.fa * R[I.a];
          if (I.b != 255) v = fmaf(I.fb, R[I.b], v);
        }
        R[I.dst] = act ? v : 0.f;
      };
      for (int ip = 0; ip < J3_NINS; ++ip) run_fwd(kJ3Prog[ip]);
      const float fc = __shfl_sync(0xffffffffu, R[99], 0), fx = __shfl_sync(0xffffffffu, R[107], 0),
                  fy = __shfl_sync(0xffffffffu, R[115], 0);
      if (lane == 0 && valid) {
        atomicAdd(&red[PINN_SUM_FC], (double)(fc * fc));
        atomicAdd(&red[PINN_SUM_FX], (double)(fx * fx));
        atomicAdd(&red[PINN_SUM_FY], (double)(fy * fy));
        atomicAdd(&red[PINN_SUM_NPOINTS], 1.0);
      }
      if (!bwd) continue;
      // ---- reverse of the register program: adjoints of the three residuals' constant terms ----
      float G[J3_NREG];
#pragma unroll 1
      for (int r = 0; r < J3_NREG; ++r) G[r] = 0.f;
      const float wr = valid ? 2.f * D.w_res * A.inv_n_res : 0.f;
      if (lane == 0) G[99] = wr * fc, G[107] = wr * fx, G[115] = wr * fy;
      for (int ip = J3_NINS - 1; ip >= 0; --ip) {
        const J3Ins I = kJ3Prog[ip];
        const float gd = G[I.dst];
        if (I.op == J3_DER) {
          // dst_alpha = (alpha_d + 1) src_{alpha + e_d}  =>  srcbar_beta += beta_d dstbar_{beta - e_d}
          const int dir = I.b;
          const int from = kJ3.down[dir][c];
          const float x = __shfl_sync(0xffffffffu, gd, from < 0 ? 0 : from);
          if (act && from >= 0) G[I.a] += x * (float)kJ3.ex[dir][c];
        } else if (I.op == J3_MUL) {
          const float ra = R[I.a], rb = R[I.b];
          float sa = 0.f, sb = 0.f;
          for (int q = 0; q < J3; ++q) {
            const bool on = q < kJ3.rev_n[c];
            const int jj = on ? kJ3.rev_j[c][q] : 0, kk = on ? kJ3.rev_k[c][q] : 0;
            const float gk = __shfl_sync(0xffffffffu, gd, kk);
            const float bj = __shfl_sync(0xffffffffu, rb, jj);
            const float aj = __shfl_sync(0xffffffffu, ra, jj);
            if (on) sa = fmaf(bj, gk, sa), sb = fmaf(aj, gk, sb);
          }
          if (act) {
            if (I.a == I.b) G[I.a] += sa + sb;
            else G[I.a] += sa, G[I.b] += sb;
          }
        } else {
          if (act) {
            G[I.a] += I.fa * gd;
            if (I.b != 255) G[I.b] += I.fb * gd;
          }
        }
      }
      // seeds = d loss / d (output coefficients), written over the output jets; data misfit adds to the values
      __syncwarp();
      for (int n = 0; n < o; ++n) {
        float sd = 0.f;
        for (int f = 0; f < 4; ++f)
          if (D.field_cols[f] == n) sd = G[f];
        if (lane == 0 && A.targets && valid) {
          for (int i = 0; i < D.n_targets; ++i)
            if (D.target_cols[i] == n) {
              const float e = cur[n * J3_M + p] - A.targets[gp * D.n_targets + i];
              sd += 2.f * D.w_fid * A.inv_n_fid * D.target_w[i] * e;
            }
        }
        __syncwarp();
        if (act) cur[n * J3_M + c * J3_TP + p] = sd;
      }
      (void)terr;
    }
    __syncthreads();
    if (!bwd) continue;

    // =============================== reverse through the layers ===============================
    // cur holds zbar of layer L-1 (the last layer is linear); oth is free
    float* bz = cur;
    float* ba = oth;
    for (int l = L - 1; l >= 0; --l) {
      const int K = D.widths[l], Nn = D.widths[l + 1];
      poff -= (long long)K * Nn + Nn;
      __syncthreads();
      // input jets of layer l: the stored post-activation jets of layer l-1 (or the input jets)
      if (l == 0) {
        init_jets(ba);
      } else {
        const float* sl = slab + (size_t)(l - 1) * WM * J3_M;
        for (int i = tid; i < K * J3_M; i += J3_NT) ba[i] = sl[i];
      }
      for (int i = tid; i < Nn * K; i += J3_NT) wsm[i] = A.params[poff + i];
      __syncthreads();
      // dW[n][k] += sum_m zbar[n][m] a[k][m];  db[n] += sum_p zbar[n][0][p]
      for (int i = tid; i < Nn * K; i += J3_NT) {
        const int n = i / K, k = i - n * K;
        const float* zr = bz + n * J3_M;
        const float* ar = ba + k * J3_M;
        float acc = 0.f;
        for (int m = 0; m < J3_M; ++m) acc = fmaf(zr[m], ar[m], acc);
        atomicAdd(A.grad + poff + i, acc);
      }
      for (int n = tid; n < Nn; n += J3_NT) {
        float acc = 0.f;
        for (int p = 0; p < J3_TP; ++p) acc += bz[n * J3_M + p];
        atomicAdd(A.grad + poff + (long long)K * Nn + n, acc);
      }
      if (l == 0) break;
      // abar[k][m] = sum_n W[n][k] zbar[n][m], then through tanh of layer l-1: zbar_{l-1} = corr(1 - y^2, abar)
      float ab_col[J3_MAXW];
      {
        const int m = tid;
        for (int k = 0; k < K; ++k) {
          float acc = 0.f;
          for (int n = 0; n < Nn; ++n) acc = fmaf(wsm[n * K + k], bz[n * J3_M + m], acc);
          ab_col[k] = acc;
        }
      }
      __syncthreads();           // everyone has read bz: it becomes the abar / zbar buffer of layer l-1
      {
        const int m = tid;
        for (int k = 0; k < K; ++k) bz[k * J3_M + m] = ab_col[k];
      }
      __syncthreads();
      for (int i = tid; i < K * J3_TP; i += J3_NT) {
        const int k = i / J3_TP, p = i - k * J3_TP;
        float y[J3], pp[J3], ab[J3], zb[J3];
        for (int c = 0; c < J3; ++c) y[c] = ba[k * J3_M + c * J3_TP + p], ab[c] = bz[k * J3_M + c * J3_TP + p], zb[c] = 0.f;
        j3_pmul(y, y, pp);
        for (int c = 0; c < J3; ++c) pp[c] = -pp[c];
        pp[0] += 1.f;                       // p = f'(z) = 1 - y^2 as a truncated polynomial
        j3_pcorr(pp, ab, zb);               // zbar_alpha = sum_{beta >= alpha} p_{beta - alpha} ybar_beta
        for (int c = 0; c < J3; ++c) bz[k * J3_M + c * J3_TP + p] = zb[c];
      }
      // (ba now holds y_{l-1}, which the next iteration overwrites with y_{l-2}; bz holds zbar_{l-1})
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid < PINN_NSUMS && A.sums && red[tid] != 0.0) atomicAdd(A.sums + tid, red[tid]);
}

// --------------------------------------------------------------------------------------------------------- host side
static void j3_build_tables(J3Tables* T, J3Ins* prog) {
  int e[J3][3], n = 0;
  for (int deg = 0; deg <= 3; ++deg)
    for (int a = deg; a >= 0; --a)
      for (int b = deg - a; b >= 0; --b) {
        e[n][0] = a, e[n][1] = b, e[n][2] = deg - a - b;
        ++n;
      }
  auto find = [&](int a, int b, int c) {
    for (int i = 0; i < J3; ++i)
      if (e[i][0] == a && e[i][1] == b && e[i][2] == c) return i;
    return -1;
  };
  for (int i = 0; i < J3; ++i)
    for (int dd = 0; dd < 3; ++dd) {
      int u[3] = {e[i][0], e[i][1], e[i][2]};
      T->ex[dd][i] = (unsigned char)u[dd];
      u[dd] += 1;
      T->up[dd][i] = (signed char)(u[0] + u[1] + u[2] <= 3 ? find(u[0], u[1], u[2]) : -1);
      u[dd] -= 2;
      T->down[dd][i] = (signed char)(u[dd] >= 0 ? find(u[0], u[1], u[2]) : -1);
    }
  for (int k = 0; k < J3; ++k) T->fwd_n[k] = 0, T->rev_n[k] = 0;
  for (int i = 0; i < J3; ++i)
    for (int j = 0; j < J3; ++j) {
      const int s[3] = {e[i][0] + e[j][0], e[i][1] + e[j][1], e[i][2] + e[j][2]};
      if (s[0] + s[1] + s[2] > 3) continue;
      const int k = find(s[0], s[1], s[2]);
      T->fwd_i[k][T->fwd_n[k]] = (unsigned char)i, T->fwd_j[k][T->fwd_n[k]] = (unsigned char)j, ++T->fwd_n[k];
      T->rev_j[i][T->rev_n[i]] = (unsigned char)j, T->rev_k[i][T->rev_n[i]] = (unsigned char)k, ++T->rev_n[i];
    }
  // ---- the residual program: physics_functions.py:55-130, one instruction per assignment (registers 0..3 = h, z, u, v) ----
  int ip = 0;
  auto DER = [&](int dst, int a, int dir) { prog[ip++] = J3Ins{J3_DER, (unsigned char)dst, (unsigned char)a, (unsigned char)dir, 0.f, 0.f}; };
  auto MUL = [&](int dst, int a, int b) { prog[ip++] = J3Ins{J3_MUL, (unsigned char)dst, (unsigned char)a, (unsigned char)b, 0.f, 0.f}; };
  auto LIN = [&](int dst, int a, float fa, int b, float fb) {
    prog[ip++] = J3Ins{J3_LIN, (unsigned char)dst, (unsigned char)a, (unsigned char)(b < 0 ? 255 : b), fa, fb};
  };
  enum { h = 0, z = 1, u = 2, v = 3, T_ = 0, X_ = 1, Y_ = 2 };
  DER(4, u, T_), DER(5, u, X_), DER(6, u, Y_);            // u_t u_x u_y
  DER(7, v, T_), DER(8, v, X_), DER(9, v, Y_);            // v_t v_x v_y
  DER(10, z, T_), DER(11, z, X_), DER(12, z, Y_);         // z_t z_x z_y
  MUL(13, h, u), MUL(14, h, v), DER(15, 13, X_), DER(16, 14, Y_);     // hu hv hu_x hv_y
  LIN(17, 15, 1.f, 16, 1.f), LIN(18, 5, 1.f, 9, 1.f);                  // A B
  DER(19, 17, T_), DER(20, 17, X_), DER(21, 17, Y_);                  // A_t A_x A_y
  DER(22, 18, T_), DER(23, 18, X_), DER(24, 18, Y_);                  // B_t B_x B_y
  LIN(25, h, -0.53f, z, 0.47f), DER(26, 25, X_), DER(27, 25, Y_);     // z_alpha, z_alpha_x, z_alpha_y
  MUL(28, 25, 25), MUL(29, h, h), MUL(30, h, z), MUL(31, z, z);       // z_alpha^2 h^2 hz z^2
  LIN(32, 29, 1.f, 30, -1.f), LIN(33, 32, 1.f, 31, 1.f);
  LIN(34, 28, 0.5f, 33, -(float)0.16666666666666666);                  // temp1
  LIN(35, h, 1.f, z, -1.f), LIN(36, 25, 1.f, 35, 0.5f);               // temp2
  MUL(37, 34, 23), MUL(38, 36, 20), LIN(39, 37, 1.f, 38, 1.f);        // u_2
  MUL(40, 34, 24), MUL(41, 36, 21), LIN(42, 40, 1.f, 41, 1.f);        // v_2
  LIN(43, u, 1.f, 39, 1.f), LIN(44, v, 1.f, 42, 1.f), LIN(45, h, 1.f, z, 1.f);   // u_surface v_surface H
  MUL(46, 45, 43), MUL(47, 45, 44), DER(48, 46, X_), DER(49, 47, Y_); // Hu_x Hv_y
  LIN(50, 28, 0.5f, -1, 0.f);                                          // z_alpha^2 / 2
  MUL(51, 50, 23), MUL(52, 25, 20), LIN(53, 51, 1.f, 52, 1.f), DER(54, 53, T_);   // V1Ax, V1Ax_t
  MUL(55, 50, 24), MUL(56, 25, 21), LIN(57, 55, 1.f, 56, 1.f), DER(58, 57, T_);   // V1Ay, V1Ay_t
  LIN(59, 31, 0.5f, -1, 0.f), MUL(60, 59, 22), MUL(61, z, 19), LIN(62, 60, 1.f, 61, 1.f);   // V1B
  DER(63, 62, X_), DER(64, 62, Y_);                                   // V1Bx V1By
  LIN(65, 54, 1.f, 63, -1.f), LIN(66, 58, 1.f, 64, -1.f);             // V1x V1y
  LIN(67, 25, 1.f, z, -1.f), MUL(68, u, 20), MUL(69, v, 21), LIN(70, 68, 1.f, 69, 1.f), MUL(71, 67, 70);
  LIN(72, 28, 0.5f, 31, -0.5f), MUL(73, u, 23), MUL(74, v, 24), LIN(75, 73, 1.f, 74, 1.f), MUL(76, 72, 75);
  MUL(77, z, 18), LIN(78, 17, 1.f, 77, 1.f), MUL(79, 78, 78), LIN(80, 71, 1.f, 76, 1.f), LIN(81, 80, 1.f, 79, 0.5f);   // V2
  DER(82, 81, X_), DER(83, 81, Y_);                                   // V2x V2y
  LIN(84, 8, 1.f, 6, -1.f);                                            // omega0 = v_x - u_y
  MUL(85, 25, 24), LIN(86, 21, 1.f, 85, 1.f), MUL(87, 26, 86);
  MUL(88, 25, 23), LIN(89, 20, 1.f, 88, 1.f), MUL(90, 27, 89), LIN(91, 87, 1.f, 90, -1.f);   // omega2
  MUL(92, 84, 42), MUL(93, 91, v), LIN(94, 92, -1.f, 93, -1.f);       // V3x
  MUL(95, 84, 39), MUL(96, 91, u), LIN(97, 95, 1.f, 96, 1.f);         // V3y
  LIN(98, 10, 1.f, 48, 1.f), LIN(99, 98, 1.f, 49, 1.f);               // f_cont
  MUL(100, u, 5), MUL(101, v, 6), LIN(102, 4, 1.f, 100, 1.f), LIN(103, 102, 1.f, 101, 1.f), LIN(104, 103, 1.f, 11, 9.81f);
  LIN(105, 104, 1.f, 65, 1.f), LIN(106, 105, 1.f, 82, 1.f), LIN(107, 106, 1.f, 94, 1.f);    // f_momx
  MUL(108, u, 8), MUL(109, v, 9), LIN(110, 7, 1.f, 108, 1.f), LIN(111, 110, 1.f, 109, 1.f), LIN(112, 111, 1.f, 12, 9.81f);
  LIN(113, 112, 1.f, 66, 1.f), LIN(114, 113, 1.f, 83, 1.f), LIN(115, 114, 1.f, 97, 1.f);    // f_momy
  if (ip != J3_NINS) abort();
}

bool jet3_supported(const pinn_desc_t* D, const char** why) {
  *why = "";
  for (int i = 0; i <= D->n_linear; ++i)
    if (D->widths[i] > J3_MAXW) return *why = "the third-order jet kernel takes layer widths up to 64", false;
  if (D->activation != PINN_ACT_TANH) return *why = "tanh activation only", false;
  if (D->precision != PINN_PREC_FP32) return *why = "FP32 only", false;
  return true;
}

static size_t j3_smem(int wmax) {
  return (size_t)(2 * wmax * J3_M + wmax * wmax + J3_TP * PINN_MAX_IN) * 4 + PINN_NSUMS * 8 + 16;
}

int jet3_workspace(const pinn_desc_t* D, long long n_points, bool bwd, size_t* bytes, int* grid_out, long long* stride_out,
                   int* wmax_out) {
  int wmax = 4;
  for (int i = 0; i <= D->n_linear; ++i) wmax = wmax > D->widths[i] ? wmax : D->widths[i];
  int dev = 0, sms = 148;
  PINN_CUDA(cudaGetDevice(&dev));
  PINN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long tiles = (n_points + J3_TP - 1) / J3_TP;
  const size_t smem = j3_smem(wmax);
  int per_sm = (int)(200 * 1024 / smem);
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long grid = tiles < (long long)sms * per_sm ? tiles : (long long)sms * per_sm;
  if (grid < 1) grid = 1;
  const long long stride = bwd ? (long long)(D->n_linear - 1) * wmax * J3_M : 0;
  if (bytes) *bytes = (size_t)grid * (size_t)stride * 4 + 256;
  if (grid_out) *grid_out = (int)grid;
  if (stride_out) *stride_out = stride;
  if (wmax_out) *wmax_out = wmax;
  return PINN_OK;
}

int run_jet3_pass(const pinn_desc_t* D, const pinn_eval_args_t* a, bool bwd, cudaStream_t st) {
  static bool tables_done[64] = {false};
  int dev = 0;
  PINN_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !tables_done[dev]) {
    static J3Tables T;
    static J3Ins prog[J3_NINS];
    j3_build_tables(&T, prog);
    PINN_CUDA(cudaMemcpyToSymbolAsync(kJ3, &T, sizeof(T), 0, cudaMemcpyHostToDevice, st));
    PINN_CUDA(cudaMemcpyToSymbolAsync(kJ3Prog, prog, sizeof(J3Ins) * J3_NINS, 0, cudaMemcpyHostToDevice, st));
    PINN_CUDA(cudaStreamSynchronize(st));     // once per device: the host arrays are static, but keep the first use simple
    tables_done[dev] = true;
  }
  size_t need = 0;
  int grid = 1, wmax = 4;
  long long stride = 0;
  int rc = jet3_workspace(D, a->n_points, bwd, &need, &grid, &stride, &wmax);
  if (rc) return rc;
  if (a->workspace_bytes < need) return set_error("workspace too small: %zu < %zu bytes", a->workspace_bytes, need), PINN_E_WORKSPACE;
  long long tiles = (a->n_points + J3_TP - 1) / J3_TP;
  if (tiles == 0) return PINN_OK;
  if (tiles > 0x7fffffffLL) return set_error("too many tiles"), PINN_E_UNSUPPORTED;
  J3Args A;
  A.params = a->params;
  A.inputs = a->inputs;
  A.targets = D->n_targets > 0 ? a->targets : nullptr;
  A.grad = a->grad;
  A.sums = a->sums;
  A.out = a->out;
  A.slab = reinterpret_cast<float*>(a->workspace);
  A.slab_stride = stride;
  A.n_points = a->n_points;
  A.n_tiles = (int)tiles;
  A.wmax = wmax;
  A.inv_n_res = a->n_res_global > 0 ? (float)(1.0 / (double)a->n_res_global) : 0.f;
  A.inv_n_fid = a->n_fid_global > 0 ? (float)(1.0 / (double)a->n_fid_global) : 0.f;
  const size_t smem = j3_smem(wmax);
  PINN_CUDA(cudaFuncSetAttribute(jet3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  jet3_kernel<<<grid, J3_NT, smem, st>>>(*D, A, bwd ? 1 : 0);
  PINN_CUDA(cudaGetLastError());
  return PINN_OK;
}

}  // namespace pinn
