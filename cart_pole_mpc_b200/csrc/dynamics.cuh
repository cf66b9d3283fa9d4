// Device-side cart-pole dynamics and RK4 for sm_100a.
//
// What it computes (reference citations; the code is a re-derivation organised for registers, not a port):
//   single_pendulum_dynamics  optimization/single_pendulum_dynamics.hpp:13-186  -> dyn_eval / dyn_eval_jac
//   runge_kutta_4th_order*    optimization/integration.hpp:13-62                -> rk4_step / rk4_step_jac
//   mod_pi                    optimization/integration.hpp:65-73                -> mod_pi
//
// Differences from the generated reference code that matter for speed and are exact or ulp-level:
//   * every sub-expression that depends only on SingleCartPoleParams (1/l, M/l^2, -c_d/6, -(M mu_b) g, 1/v_mu ...)
//     is hoisted into DynConst once per instance: 3 of the 6 divisions per evaluation disappear;
//   * the state Jacobian of the dynamics has two constant rows ([0 0 1 0], [0 0 0 1]); only rows 2-3 (8 numbers) and
//     the two non-zero entries of the control Jacobian are produced, and the RK4 chain rule uses that structure
//     (32 instead of 64 FMAs per stage product);
//   * sin/cos, tanh, sqrt and the reciprocals come from fastmath.cuh (branch-free, ~1 ulp) so that one evaluation is
//     a single basic block whose four long dependent chains overlap;
//   * 1/(det*det) is computed as (1/det)^2 (1 ulp).
#pragma once
#include <cuda_runtime.h>

#include "fastmath.cuh"

namespace cpmpc {

struct DynConst {
  double m1;         // pole mass
  double m_total;    // m_1 + m_b
  double len;        // l_1
  double inv_len;    // 1 / l_1
  double len2;       // 2 l_1
  double m_over_l2;  // (m_1 + m_b) / l_1^2
  double drag_k;     // -c_d / 6
  double fric_k;     // -(m_total mu_b) g
  double inv_v_cut;  // 1 / max(v_mu_b, 1e-6)
  double k_s;        // bumper spring constant
  double x_s;        // bumper position
  double grav;       // g
};

__device__ __forceinline__ DynConst make_dyn_const(const double* __restrict__ p) {
  DynConst c;
  const double m_b = p[0];
  c.m1 = p[1];
  c.len = p[2];
  c.grav = p[3];
  const double mu_b = p[4];
  const double v_mu = p[5];
  const double c_d = p[6];
  c.x_s = p[7];
  c.k_s = p[8];
  c.m_total = c.m1 + m_b;
  c.inv_len = 1.0 / c.len;
  c.len2 = c.len * 2.0;
  c.m_over_l2 = c.m_total * (1.0 / (c.len * c.len));
  c.drag_k = (-0.16666666666666666) * c_d;
  c.fric_k = (c.m_total * mu_b) * -c.grav;
  c.inv_v_cut = 1.0 / ((1e-06 < v_mu) ? v_mu : 1e-06);
  return c;
}

// DynConst parked in / fetched from shared memory (12 doubles).
__device__ __forceinline__ void store_dyn_const(double* dst, const DynConst& c) {
  dst[0] = c.m1;
  dst[1] = c.m_total;
  dst[2] = c.len;
  dst[3] = c.inv_len;
  dst[4] = c.len2;
  dst[5] = c.m_over_l2;
  dst[6] = c.drag_k;
  dst[7] = c.fric_k;
  dst[8] = c.inv_v_cut;
  dst[9] = c.k_s;
  dst[10] = c.x_s;
  dst[11] = c.grav;
}
__device__ __forceinline__ void load_dyn_const(DynConst& c, const double* src) {
  c.m1 = src[0];
  c.m_total = src[1];
  c.len = src[2];
  c.inv_len = src[3];
  c.len2 = src[4];
  c.m_over_l2 = src[5];
  c.drag_k = src[6];
  c.fric_k = src[7];
  c.inv_v_cut = src[8];
  c.k_s = src[9];
  c.x_s = src[10];
  c.grav = src[11];
}

// Map an angle to (-pi, pi].  fmod(a, 2pi) == a exactly whenever |a| < 2pi, which skips CUDA's software fmod on
// the overwhelmingly common path.
__device__ __forceinline__ double mod_pi(double a) {
  constexpr double pi = 3.141592653589793;
  constexpr double two_pi = 6.283185307179586;
  if (!(fabs(a) < two_pi)) a = fmod(a, two_pi);
  a += (a < 0.0) ? two_pi : 0.0;
  a -= (a > pi) ? two_pi : 0.0;
  return a;
}

// Accelerations (b_x'', th'') at state (b_x, th, b_x', th') and control u.  FORCES selects the simulator variant that
// takes external forces on the base (x component only is used by the model) and on the pole mass.
template <bool FORCES>
__device__ __forceinline__ void dyn_eval(const DynConst& P, const double b_x, const double th, const double b_xd,
                                         const double thd, const double u, const double fb_x, const double fm_x,
                                         const double fm_y, double& acc_x, double& acc_th) {
  // four independent chains start here: sincos(th), tanh(friction), bumper springs, and (after sin) 1/det
  double s, c;
  fm::sincos_bf(th, s, c);
  const double fric = fm::tanh_bf(b_xd * P.inv_v_cut);
  const double over_r = b_x - P.x_s;
  const double pen_r = (0.0 < over_r) ? over_r : 0.0;
  const double over_l = -(b_x + P.x_s);
  const double pen_l = (0.0 < over_l) ? over_l : 0.0;

  const double s_m1 = s * P.m1;
  const double det = P.m1 * (P.m_total - s_m1 * s);
  const double inv_det = fm::rcp(det);

  const double lw = thd * P.len;
  const double cc = c * c;
  const double vx = b_xd - lw * s;
  const double lw_cc = lw * cc;
  const double speed_sq = lw * lw_cc + vx * vx;
  double speed, inv_speed;
  fm::sqrt_rsqrt(speed_sq, speed, inv_speed);
  const double speed15 = 1.5 * speed;
  const double proj = lw_cc - vx * s;
  const bool moving = 0.0 < speed_sq;
  const double drag_x = moving ? (vx * 2.0) * speed15 : 0.0;
  const double drag_th = moving ? proj * P.len2 * speed15 : 0.0;

  double q_th, q_x;
  if (FORCES) {
    q_th = drag_th * P.drag_k + P.len * (c * (fm_y - P.grav * P.m1) - fm_x * s);
    q_x = u + fb_x + fm_x + fric * P.fric_k + drag_x * P.drag_k + P.m1 * lw * (thd * c) + P.k_s * (pen_l - pen_r);
  } else {
    q_th = drag_th * P.drag_k + P.len * (c * (-(P.grav * P.m1)));
    q_x = u + fric * P.fric_k + drag_x * P.drag_k + P.m1 * lw * (thd * c) + P.k_s * (pen_l - pen_r);
  }
  const double num_th = q_x * (P.inv_len * s_m1) + q_th * P.m_over_l2;
  const double num_x = q_x + q_th * (s * P.inv_len);
  acc_x = num_x * (P.m1 * inv_det);
  acc_th = inv_det * num_th;
}

// Accelerations plus rows 2-3 of d(x_dot)/dx (fx[4], fth[4]) and the two non-zero entries of d(x_dot)/du.
// External forces are zero (the optimiser never passes any, optimization.cc:124-126).
__device__ __forceinline__ void dyn_eval_jac(const DynConst& P, const double b_x, const double th, const double b_xd,
                                             const double thd, const double u, double& acc_x, double& acc_th,
                                             double fx[4], double fth[4], double& gx, double& gth) {
  double s, c;
  fm::sincos_bf(th, s, c);
  const double fric = fm::tanh_bf(b_xd * P.inv_v_cut);
  const double over_r = b_x - P.x_s;
  const bool act_r = 0.0 < over_r;
  const double pen_r = act_r ? over_r : 0.0;
  const double over_l = -(b_x + P.x_s);
  const bool act_l = 0.0 < over_l;
  const double pen_l = act_l ? over_l : 0.0;

  const double s_m1 = s * P.m1;
  const double det = P.m1 * (P.m_total - s_m1 * s);
  const double inv_det = fm::rcp(det);

  const double lw = thd * P.len;
  const double cc = c * c;
  const double vx = b_xd - lw * s;
  const double lw_cc = lw * cc;
  const double speed_sq = lw * lw_cc + vx * vx;
  double speed, inv_speed;
  fm::sqrt_rsqrt(speed_sq, speed, inv_speed);
  const double speed15 = 1.5 * speed;
  const double vx2 = vx * 2.0;
  const double proj = lw_cc - vx * s;
  const bool moving = 0.0 < speed_sq;
  const double drag_x = moving ? vx2 * speed15 : 0.0;
  const double drag_th = moving ? proj * P.len2 * speed15 : 0.0;

  const double gm1 = P.grav * P.m1;
  const double q_th = drag_th * P.drag_k - P.len * (c * gm1);
  const double q_x = u + fric * P.fric_k + drag_x * P.drag_k + P.m1 * lw * (thd * c) + P.k_s * (pen_l - pen_r);
  const double s_m1_over_l = P.inv_len * s_m1;
  const double s_over_l = s * P.inv_len;
  const double inv_det_l = inv_det * P.inv_len;
  const double num_th = q_x * s_m1_over_l + q_th * P.m_over_l2;
  const double num_x = q_x + q_th * s_over_l;
  gth = s_m1 * inv_det_l;
  gx = P.m1 * inv_det;
  acc_x = num_x * gx;
  acc_th = inv_det * num_th;

  // derivatives of the drag terms (zero at rest): computed unconditionally, selected at the end
  const double proj_n = 1.5 * (inv_speed * proj);
  const double vx_rel = vx + s * lw;
  const double speed3 = speed * 3.0;
  const double d_th_q = moving ? -((lw * proj_n + speed15) * (c * P.len) * (2.0 * vx_rel)) : 0.0;
  const double d_th_x = moving ? -((inv_speed * vx_rel * (vx2 * 1.5) + speed3) * lw * c) : 0.0;
  const double d_v_l = moving ? P.len * (vx2 * proj_n - speed3 * s) : 0.0;
  const double d_v_x = moving ? vx2 * vx2 * (0.75 * inv_speed) + speed3 : 0.0;
  const double d_w_th = moving ? P.len * (proj * proj_n + (cc + s * s) * speed15) * P.len2 : 0.0;

  const double dq_th_dth = d_th_q * P.drag_k + P.len * (s * gm1);
  const double dq_x_dth = d_th_x * P.drag_k - lw * s_m1 * thd;
  const double m1_c = P.m1 * c;
  const double dv_l_k = d_v_l * P.drag_k;
  const double inv_det_sq = inv_det * inv_det;
  const double dq_x_dw = 2.0 * lw * m1_c + dv_l_k;
  const double dw_th_k = d_w_th * P.drag_k;
  const double dq_x_dv = P.inv_v_cut * (1.0 - fric * fric) * P.fric_k + d_v_x * P.drag_k;
  const double spring_sw = P.k_s * ((act_l ? -1.0 : 0.0) - (act_r ? 1.0 : 0.0));

  fx[0] = gx * spring_sw;
  fx[1] = P.m1 * (num_x * s_m1 * m1_c * (2.0 * inv_det_sq) +
                  inv_det * (dq_x_dth + P.inv_len * (s * dq_th_dth + c * q_th)));
  fx[2] = (dq_x_dv + s_over_l * dv_l_k) * gx;
  fx[3] = (dq_x_dw + s_over_l * dw_th_k) * gx;
  fth[0] = spring_sw * gth;
  fth[1] = inv_det * dq_th_dth * P.m_over_l2 +
           P.m1 * (inv_det_sq * num_th * s_m1 * (c * 2.0) + (c * q_x + s * dq_x_dth) * inv_det_l);
  fth[2] = inv_det * (P.m_over_l2 * dv_l_k + dq_x_dv * s_m1_over_l);
  fth[3] = inv_det * (P.m_over_l2 * dw_th_k + dq_x_dw * s_m1_over_l);
}

// One classic RK4 step, in place, no Jacobians (integration.hpp:53-62).  The four stages run as a rolled loop around
// ONE inlined copy of the dynamics so the hot code stays inside the instruction cache.
template <bool FORCES>
__device__ __forceinline__ void rk4_step(const DynConst& P, double x[4], const double u, const double h,
                                         const double fb_x = 0.0, const double fm_x = 0.0, const double fm_y = 0.0) {
  const double hh = h * 0.5;
  double a0 = x[0], a1 = x[1], a2 = x[2], a3 = x[3];  // stage argument
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;      // k1 + 2 k2 + 2 k3 + k4
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    double ax, at;
    dyn_eval<FORCES>(P, a0, a1, a2, a3, u, fb_x, fm_x, fm_y, ax, at);
    const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
    s0 += a2 * wgt;
    s1 += a3 * wgt;
    s2 += ax * wgt;
    s3 += at * wgt;
    const double cst = (st == 2) ? h : hh;
    const double n0 = x[0] + a2 * cst, n1 = x[1] + a3 * cst, n2 = x[2] + ax * cst, n3 = x[3] + at * cst;
    a0 = n0;
    a1 = n1;
    a2 = n2;
    a3 = n3;
  }
  const double h6 = h / 6.0;
  x[0] += h6 * s0;
  x[1] += h6 * s1;
  x[2] += h6 * s2;
  x[3] += h6 * s3;
}

// One RK4 step with the Jacobians of the step map (integration.hpp:13-49): x is advanced in place,
// A = d x+ / d x (4x4 row-major), Bv = d x+ / d u.  Stage Jacobians have the structure
// F = [[0 0 1 0],[0 0 0 1],[fx],[fth]], Fu = [0 0 gx gth]; the chain rule K_i = F_i (I + c_i K_{i-1}) and
// ku_i = F_i ku_{i-1} c_i + Fu_i is carried through a rolled stage loop.
__device__ __forceinline__ void rk4_step_jac(const DynConst& P, double x[4], const double u, const double h,
                                             double A[16], double Bv[4]) {
  const double hh = h * 0.5;
  double a0 = x[0], a1 = x[1], a2 = x[2], a3 = x[3];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  double K[16], ku[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    K[i] = 0.0;
    A[i] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ku[i] = 0.0;
    Bv[i] = 0.0;
  }
  double cprev = 0.0;  // coefficient that multiplied the previous stage's k in this stage's argument
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    double ax, at, fx[4], fth[4], gx, gth;
    dyn_eval_jac(P, a0, a1, a2, a3, u, ax, at, fx, fth, gx, gth);
    // Pm = I + K_prev * cprev  (identity at stage 0 because cprev = 0 and K = 0)
    double Pm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) Pm[i] = ((i % 5 == 0) ? 1.0 : 0.0) + K[i] * cprev;
    // K = F * Pm
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      K[j] = Pm[8 + j];
      K[4 + j] = Pm[12 + j];
      K[8 + j] = fx[0] * Pm[j] + fx[1] * Pm[4 + j] + fx[2] * Pm[8 + j] + fx[3] * Pm[12 + j];
      K[12 + j] = fth[0] * Pm[j] + fth[1] * Pm[4 + j] + fth[2] * Pm[8 + j] + fth[3] * Pm[12 + j];
    }
    // ku = (F * ku_prev) * cprev + Fu
    const double k0 = ku[2] * cprev, k1 = ku[3] * cprev;
    const double k2 = (fx[0] * ku[0] + fx[1] * ku[1] + fx[2] * ku[2] + fx[3] * ku[3]) * cprev + gx;
    const double k3 = (fth[0] * ku[0] + fth[1] * ku[1] + fth[2] * ku[2] + fth[3] * ku[3]) * cprev + gth;
    ku[0] = k0;
    ku[1] = k1;
    ku[2] = k2;
    ku[3] = k3;
    const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] += K[i] * wgt;
#pragma unroll
    for (int i = 0; i < 4; ++i) Bv[i] += ku[i] * wgt;
    s0 += a2 * wgt;
    s1 += a3 * wgt;
    s2 += ax * wgt;
    s3 += at * wgt;
    const double cst = (st == 2) ? h : hh;
    const double n0 = x[0] + a2 * cst, n1 = x[1] + a3 * cst, n2 = x[2] + ax * cst, n3 = x[3] + at * cst;
    a0 = n0;
    a1 = n1;
    a2 = n2;
    a3 = n3;
    cprev = cst;
  }
  const double h6 = h / 6.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) A[i] = ((i % 5 == 0) ? 1.0 : 0.0) + h6 * A[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) Bv[i] = h6 * Bv[i];
  x[0] += h6 * s0;
  x[1] += h6 * s1;
  x[2] += h6 * s2;
  x[3] += h6 * s3;
}

}  // namespace cpmpc
