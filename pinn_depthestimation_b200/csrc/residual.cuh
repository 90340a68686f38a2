// Per-point epilogue shared by the FP32 and the tensor-core jet kernels: data misfit, the PDE residual
// of the selected physics.py function, partial loss sums and the adjoint seeds of SURVEY.md
// Appendix A, written back in place over the output jets.
//
// `Acc` exposes the network-output jets of ONE point: get/set/add(col, j), j = 0 value, 1.. the
// derivative w.r.t. direction j-1 (direction order = argument order of the physics function).
#pragma once
#include "common.cuh"

namespace pinn {

constexpr float kG = 9.81f;
constexpr float kCb = (float)(3.0 / 16.0 * 9.81 * 0.78 * 0.78);  // physics.py:77-78
constexpr float kCd = 0.002f;                                     // physics.py:100

struct EpiArgs {
  const float* targets;
  const float* seed_out;
  const float* seed_dout[PINN_MAX_DIRS];
  float* out;
  float* dout[PINN_MAX_DIRS];
  float inv_n_res, inv_n_fid, inv_cnt;
};

// J: jets carried (1 + directions).  isp: this lane owns a point slot of the tile; valid: the slot holds
// a real point (gp < n_points).  xin: the point's raw input row.  ls: this lane's contribution to the
// PINN_NSUMS raw sums (caller reduces).
template <int J, class Acc>
__device__ __forceinline__ void residual_epilogue(const pinn_desc_t& D, Acc& acc, bool isp, bool valid,
                                                  long long gp, const float* xin, const EpiArgs& E,
                                                  float (&ls)[PINN_NSUMS]) {
#pragma unroll
  for (int i = 0; i < PINN_NSUMS; ++i) ls[i] = 0.f;
  if (!isp) return;
  const int kind = D.residual_kind;
  const int o = D.widths[D.n_linear];
  const int NPo = pad4(o);
  if (E.out && valid)
    for (int c = 0; c < o; ++c) E.out[gp * o + c] = acc.get(c, 0);
  for (int j = 1; j < J; ++j)
    if (E.dout[j - 1] && valid)
      for (int c = 0; c < o; ++c) E.dout[j - 1][gp * o + c] = acc.get(c, j);
  // data misfit (train_newmethod.py:129-133 / train.py:136-141)
  float terr[PINN_MAX_OUT];
#pragma unroll
  for (int i = 0; i < PINN_MAX_OUT; ++i) {
    terr[i] = 0.f;
    if (E.targets && i < D.n_targets && valid) {
      terr[i] = acc.get(D.target_cols[i], 0) - E.targets[gp * D.n_targets + i];
      ls[PINN_SUM_TARGET0 + i] = terr[i] * terr[i];
    }
  }
  const float vf = valid ? 1.f : 0.f;
  ls[PINN_SUM_NPOINTS] = vf;
  const float wr = 2.f * D.w_res * E.inv_n_res * vf;
  auto clear = [&]() {
    for (int c = 0; c < NPo; ++c)
      for (int j = 0; j < J; ++j) acc.set(c, j, 0.f);
  };
  if (kind == PINN_RES_CONT_ONLY || kind == PINN_RES_CONT_FTEMP) {
    if constexpr (J >= 3) {
      const int ch = D.field_cols[0], cU = D.field_cols[1], cV = D.field_cols[2];
      const float h = acc.get(ch, 0), hx = acc.get(ch, 1), hy = acc.get(ch, 2);
      const float U = acc.get(cU, 0), Ux = acc.get(cU, 1), V = acc.get(cV, 0), Vy = acc.get(cV, 2);
      const float fc = hx * U + h * Ux + hy * V + h * Vy;  // physics.py:20-23
      ls[PINN_SUM_FC] = fc * fc * vf;
      const float r = wr * fc;
      float sh = r * (Ux + Vy);
      if (kind == PINN_RES_CONT_ONLY) {  // physics.py:27-28
        const bool m = valid && xin[D.mask_col] < D.cond_threshold;
        const float dev = h - D.cond_value;
        if (m) {
          ls[PINN_SUM_COND] = dev * dev;
          ls[PINN_SUM_MASKCNT] = 1.f;
          sh += 2.f * D.w_res * dev * E.inv_cnt;
        }
      }
      clear();
      acc.add(cU, 0, r * hx);
      acc.add(cV, 0, r * hy);
      acc.add(ch, 0, sh);
      acc.add(cU, 1, r * h);
      acc.add(cV, 2, r * h);
      acc.add(ch, 1, r * U);
      acc.add(ch, 2, r * V);
    }
  } else if (kind == PINN_RES_NSWE) {
    if constexpr (J >= 4) {
      const int ch = D.field_cols[0], cz = D.field_cols[1], cu = D.field_cols[2], cv = D.field_cols[3];
      const float h = acc.get(ch, 0), hx = acc.get(ch, 2), hy = acc.get(ch, 3);
      const float z = acc.get(cz, 0), zt = acc.get(cz, 1), zx = acc.get(cz, 2), zy = acc.get(cz, 3);
      const float u = acc.get(cu, 0), ut = acc.get(cu, 1), ux = acc.get(cu, 2), uy = acc.get(cu, 3);
      const float v = acc.get(cv, 0), vt = acc.get(cv, 1), vx = acc.get(cv, 2), vy = acc.get(cv, 3);
      const float H = h + z, Hx = hx + zx, Hy = hy + zy;
      const float fc = zt + Hx * u + H * ux + Hy * v + H * vy;         // physics.py:81
      const float fx = ut + u * ux + v * uy + kG * zx + kCb * Hx * H;  // physics.py:82
      const float fy = vt + u * vx + v * vy + kG * zy + kCb * Hy * H;  // physics.py:83
      ls[PINN_SUM_FC] = fc * fc * vf;
      ls[PINN_SUM_FX] = fx * fx * vf;
      ls[PINN_SUM_FY] = fy * fy * vf;
      const float rc = wr * fc, rx = wr * fx, ry = wr * fy;
      clear();
      const float shz = rc * (ux + vy) + rx * kCb * Hx + ry * kCb * Hy;
      acc.add(ch, 0, shz);
      acc.add(cz, 0, shz);
      acc.add(cu, 0, rc * Hx + rx * ux + ry * vx);
      acc.add(cv, 0, rc * Hy + rx * uy + ry * vy);
      acc.add(ch, 2, rc * u + rx * kCb * H);
      acc.add(ch, 3, rc * v + ry * kCb * H);
      acc.add(cz, 1, rc);
      acc.add(cz, 2, rc * u + rx * (kG + kCb * H));
      acc.add(cz, 3, rc * v + ry * (kG + kCb * H));
      acc.add(cu, 1, rx);
      acc.add(cu, 2, rc * H + rx * u);
      acc.add(cu, 3, rx * v);
      acc.add(cv, 1, ry);
      acc.add(cv, 2, ry * u);
      acc.add(cv, 3, rc * H + ry * v);
    }
  } else if (kind == PINN_RES_BOUSS_SIMPLE) {
    if constexpr (J >= 4) {
      // physics_functions.py:18-52 (decompiled bytecode): f_cont = z_t + (hu)_x + (hv)_y, momentum without the breaking term
      const int ch = D.field_cols[0], cz = D.field_cols[1], cu = D.field_cols[2], cv = D.field_cols[3];
      const float h = acc.get(ch, 0), hx = acc.get(ch, 2), hy = acc.get(ch, 3);
      const float zt = acc.get(cz, 1), zx = acc.get(cz, 2), zy = acc.get(cz, 3);
      const float u = acc.get(cu, 0), ut = acc.get(cu, 1), ux = acc.get(cu, 2), uy = acc.get(cu, 3);
      const float v = acc.get(cv, 0), vt = acc.get(cv, 1), vx = acc.get(cv, 2), vy = acc.get(cv, 3);
      const float fc = zt + hx * u + h * ux + hy * v + h * vy;
      const float fx = ut + u * ux + v * uy + kG * zx;
      const float fy = vt + u * vx + v * vy + kG * zy;
      ls[PINN_SUM_FC] = fc * fc * vf;
      ls[PINN_SUM_FX] = fx * fx * vf;
      ls[PINN_SUM_FY] = fy * fy * vf;
      const float rc = wr * fc, rx = wr * fx, ry = wr * fy;
      clear();
      acc.add(ch, 0, rc * (ux + vy));
      acc.add(ch, 2, rc * u);
      acc.add(ch, 3, rc * v);
      acc.add(cz, 1, rc);
      acc.add(cz, 2, rx * kG);
      acc.add(cz, 3, ry * kG);
      acc.add(cu, 0, rc * hx + rx * ux + ry * vx);
      acc.add(cu, 1, rx);
      acc.add(cu, 2, rc * h + rx * u);
      acc.add(cu, 3, rx * v);
      acc.add(cv, 0, rc * hy + rx * uy + ry * vy);
      acc.add(cv, 1, ry);
      acc.add(cv, 2, ry * u);
      acc.add(cv, 3, rc * h + ry * v);
    }
  } else if (kind == PINN_RES_WAVE_AVG) {
    if constexpr (J >= 3) {
      const int ch = D.field_cols[0], cU = D.field_cols[1], cV = D.field_cols[2], ce = D.field_cols[3],
                cH = D.field_cols[4], ck = D.field_cols[5];
      const float h = acc.get(ch, 0), U = acc.get(cU, 0), V = acc.get(cV, 0), eta = acc.get(ce, 0);
      const float Hr = acc.get(cH, 0), kk = acc.get(ck, 0);
      const float Ux = acc.get(cU, 1), Uy = acc.get(cU, 2), Vx = acc.get(cV, 1), Vy = acc.get(cV, 2);
      const float ex = acc.get(ce, 1), ey = acc.get(ce, 2);
      const float Dp = eta + h;
      // physics.py:106: E = 1/8**rho*g*Hrms**2 is exactly 0*Hrms^2; the radiation-stress terms only
      // propagate NaN where sinh(2kh) is 0 or overflows.
      const float kh2 = 2.f * kk * h, sh_ = sinhf(kh2);
      const float poison = (0.f * Hr * Hr) * (kh2 / sh_ + 0.5f) + 0.f * (coshf(kh2) / (sh_ * sh_));
      const float Fx = kCd * U * fabsf(U) / Dp, Fy = kCd * V * fabsf(V) / Dp;
      const float fc = Ux + Vy;
      const float fx = U * Ux + V * Uy + kG * ex + Fx + poison;
      const float fy = U * Vx + V * Vy + kG * ey + Fy + poison;
      ls[PINN_SUM_FC] = fc * fc * vf;
      ls[PINN_SUM_FX] = fx * fx * vf;
      ls[PINN_SUM_FY] = fy * fy * vf;
      const float rc = wr * fc, rx = wr * fx, ry = wr * fy;
      clear();
      const float sD = -(rx * Fx + ry * Fy) / Dp;
      acc.add(ch, 0, sD);
      acc.add(ce, 0, sD);
      acc.add(cU, 0, rx * (Ux + 2.f * kCd * fabsf(U) / Dp) + ry * Vx);
      acc.add(cV, 0, rx * Uy + ry * (Vy + 2.f * kCd * fabsf(V) / Dp));
      acc.add(cU, 1, rc + rx * U);
      acc.add(cU, 2, rx * V);
      acc.add(cV, 1, ry * U);
      acc.add(cV, 2, rc + ry * V);
      acc.add(ce, 1, rx * kG);
      acc.add(ce, 2, ry * kG);
    }
  } else if (kind == PINN_RES_EXTERNAL) {
    for (int c = 0; c < NPo; ++c) {
      acc.set(c, 0, (E.seed_out && valid && c < o) ? E.seed_out[gp * o + c] : 0.f);
      for (int j = 1; j < J; ++j)
        acc.set(c, j, (E.seed_dout[j - 1] && valid && c < o) ? E.seed_dout[j - 1][gp * o + c] : 0.f);
    }
  } else {  // PINN_RES_NONE
    clear();
  }
  if (E.targets && valid) {
    const float wf = 2.f * D.w_fid * E.inv_n_fid;
#pragma unroll
    for (int i = 0; i < PINN_MAX_OUT; ++i)
      if (i < D.n_targets) acc.add(D.target_cols[i], 0, wf * D.target_w[i] * terr[i]);
  }
}

}  // namespace pinn
