// Device-side cart-pole dynamics and RK4 for sm_100a.
//
// What it computes (reference citations; the code is a re-derivation organised for registers, not a port):
//   single_pendulum_dynamics  optimization/single_pendulum_dynamics.hpp:13-186  -> dyn_eval / dyn_eval_jac
//   runge_kutta_4th_order*    optimization/integration.hpp:13-62                -> rk4_step / rk4_step_jac
//   mod_pi                    optimization/integration.hpp:65-73                -> mod_pi
//
// Differences from the generated reference code that matter for speed and are exact or ulp-level:
//   * every product that depends only on SingleCartPoleParams (1/l, (m_1 + m_b)/(m_1 l^2), -c_d/2, -(M mu_b) g, 1/v_mu,
//     -l g m_1 ...) is hoisted into DynConst once per instance, and the factor m_1 common to the numerators and the
//     mass-matrix determinant is cancelled: one reciprocal per evaluation (the generated code has six divisions);
//   * the state Jacobian of the dynamics has two constant rows ([0 0 1 0], [0 0 0 1]); only rows 2-3 (8 numbers) and
//     the two non-zero entries of the control Jacobian are produced, from the closed-form derivatives of the drag
//     terms, and the RK4 chain rule uses that structure (32 instead of 64 FMAs per stage product);
//   * sin/cos, tanh, sqrt and the reciprocals come from fastmath.cuh (branch-free, ~1 ulp) so that one evaluation is
//     a single basic block whose four long dependent chains overlap.
// The file also compiles for the host (tests/native/dynamics_check.cc).
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define CPMPC_DEV __device__ __forceinline__
#else
// host build of the same code (tests/native/dynamics_check.cc compares it with the oracle's restatement of the
// generated dynamics, which is bit-equal to the reference)
#include <cmath>
#define CPMPC_DEV inline
#endif

#include "fastmath.cuh"

namespace cpmpc {

constexpr int DYN_CONST_DOUBLES = 16;

// Everything of an evaluation that depends on SingleCartPoleParams only, computed once per instance.
struct DynConst {
  double m1;         // pole mass
  double m_total;    // m_1 + m_b
  double len;        // l_1
  double inv_len;    // 1 / l_1
  double kappa;      // (m_1 + m_b) / (m_1 l_1^2)
  double k30;        // 3 * (-c_d / 6): drag force = k30 * |v_mass| * v_mass
  double k30len;     // k30 * l_1
  double k30len2;    // k30 * l_1^2
  double fric_k;     // -(m_total mu_b) g
  double inv_v_cut;  // 1 / max(v_mu_b, 1e-6)
  double fric_kv;    // fric_k * inv_v_cut
  double k_s;        // bumper spring constant
  double x_s;        // bumper position
  double nlgm;       // -l_1 g m_1
  double m1len;      // m_1 l_1
  double m1x2;       // 2 m_1
};

CPMPC_DEV DynConst make_dyn_const(const double* __restrict__ p) {
  DynConst c;
  const double m_b = p[0];
  c.m1 = p[1];
  c.len = p[2];
  const double grav = p[3];
  const double mu_b = p[4];
  const double v_mu = p[5];
  const double c_d = p[6];
  c.x_s = p[7];
  c.k_s = p[8];
  c.m_total = c.m1 + m_b;
  c.inv_len = 1.0 / c.len;
  c.kappa = c.m_total / (c.m1 * (c.len * c.len));
  c.k30 = -0.5 * c_d;
  c.k30len = c.k30 * c.len;
  c.k30len2 = c.k30len * c.len;
  c.fric_k = (c.m_total * mu_b) * -grav;
  c.inv_v_cut = 1.0 / ((1e-06 < v_mu) ? v_mu : 1e-06);
  c.fric_kv = c.fric_k * c.inv_v_cut;
  c.nlgm = -(c.len * (grav * c.m1));
  c.m1len = c.m1 * c.len;
  c.m1x2 = c.m1 + c.m1;
  return c;
}

// DynConst parked in / fetched from shared memory (DYN_CONST_DOUBLES doubles).
CPMPC_DEV void store_dyn_const(double* dst, const DynConst& c) {
  dst[0] = c.m1;
  dst[1] = c.m_total;
  dst[2] = c.len;
  dst[3] = c.inv_len;
  dst[4] = c.kappa;
  dst[5] = c.k30;
  dst[6] = c.k30len;
  dst[7] = c.k30len2;
  dst[8] = c.fric_k;
  dst[9] = c.inv_v_cut;
  dst[10] = c.fric_kv;
  dst[11] = c.k_s;
  dst[12] = c.x_s;
  dst[13] = c.nlgm;
  dst[14] = c.m1len;
  dst[15] = c.m1x2;
}
CPMPC_DEV void load_dyn_const(DynConst& c, const double* src) {
  c.m1 = src[0];
  c.m_total = src[1];
  c.len = src[2];
  c.inv_len = src[3];
  c.kappa = src[4];
  c.k30 = src[5];
  c.k30len = src[6];
  c.k30len2 = src[7];
  c.fric_k = src[8];
  c.inv_v_cut = src[9];
  c.fric_kv = src[10];
  c.k_s = src[11];
  c.x_s = src[12];
  c.nlgm = src[13];
  c.m1len = src[14];
  c.m1x2 = src[15];
}

// The constants of a harmless plant (unit masses and length, no friction, drag or bumpers) for lane groups that carry
// no instance but still walk through the phases with their warp.
CPMPC_DEV DynConst dummy_dyn_const() {
  const double p[9] = {1.0, 1.0, 1.0, 9.81, 0.0, 1.0, 0.0, 1.0, 0.0};
  return make_dyn_const(p);
}

// Map an angle to (-pi, pi].  fmod(a, 2pi) == a exactly whenever |a| < 2pi, which skips CUDA's software fmod on
// the overwhelmingly common path.  On the device the fmod is an out-of-line call: inlined, its expansion was 1249 of
// the solve kernel's 11640 instructions (a copy of about 70 in every rollout loop, jumped over every time).
#if defined(__CUDACC__)
static __device__ __noinline__ double fmod_two_pi(double a) { return fmod(a, 6.283185307179586); }
#else
inline double fmod_two_pi(double a) { return fmod(a, 6.283185307179586); }
#endif
CPMPC_DEV double mod_pi(double a) {
  constexpr double pi = 3.141592653589793;
  constexpr double two_pi = 6.283185307179586;
  if (!(fabs(a) < two_pi)) a = fmod_two_pi(a);
  a += (a < 0.0) ? two_pi : 0.0;
  a -= (a > pi) ? two_pi : 0.0;
  return a;
}

// The equations of motion in the form evaluated here (b_x, th, v = b_x', w = th'; s = sin th, c = cos th):
//     D      = m_total - m_1 s^2                                  (mass-matrix determinant over m_1 l^2)
//     v_x    = v - l w s,   proj = l w c^2 - v_x s,   |v_m| = sqrt((l w c)^2 + v_x^2)        (pole-mass velocity)
//     q_x    = u + fric_k tanh(v / v_cut) + k30 |v_m| v_x + m_1 l w^2 c + spring(b_x)         (generalised forces)
//     q_th   = k30 l |v_m| proj - l g m_1 c
//     b_x''  = (q_x + q_th s / l) / D
//     th''   = (q_x s / l + q_th kappa) / D
// which is single_pendulum_dynamics.hpp:13-186 with every parameter-only product folded into DynConst and the common
// factor m_1 cancelled from numerator and determinant: 100 FP64 instructions per evaluation instead of 110, 141 instead
// of 172 with the Jacobian rows, one reciprocal.  Against the generated code the operations are re-associated (the
// results agree to a few ulp; tests/native/dynamics_check.cc holds both forms against the oracle's restatement, which is
// bit-equal to the reference).
//
// Accelerations (b_x'', th'') at state (b_x, th, b_x', th') and control u.  FORCES selects the simulator variant that
// takes external forces on the base (x component only is used by the model) and on the pole mass.
template <bool FORCES>
CPMPC_DEV void dyn_eval(const DynConst& P, const double b_x, const double th, const double b_xd, const double thd,
                        const double u, const double fb_x, const double fm_x, const double fm_y, double& acc_x,
                        double& acc_th) {
  // four independent chains start here: sincos(th), tanh(friction), bumper springs, and (after sin) 1/D
  double s, c;
  fm::sincos_bf(th, s, c);
  const double fric = fm::tanh_bf(b_xd * P.inv_v_cut);
  // one-sided bumper springs at +-x_s: force -sign(b_x) k_s max(|b_x| - x_s, 0)
  const double over = fabs(b_x) - P.x_s;
  const double pen = (0.0 < over) ? over : 0.0;
  const double ks_signed = copysign(P.k_s, -b_x);

  const double D = fma(-P.m1, s * s, P.m_total);
  const double inv_d = fm::rcp(D);

  const double lw = thd * P.len;
  const double cc = c * c;
  const double vx = b_xd - lw * s;
  const double lw_cc = lw * cc;
  const double speed_sq = lw * lw_cc + vx * vx;
  double speed, inv_speed;
  fm::sqrt_rsqrt(speed_sq, speed, inv_speed);
  const double spd = (0.0 < speed_sq) ? speed : 0.0;  // sqrt_rsqrt(0) is NaN
  const double proj = lw_cc - vx * s;

  double q_th = proj * (spd * P.k30len) + c * P.nlgm;
  double q_x = u + fric * P.fric_k + vx * (spd * P.k30) + (thd * P.m1len) * (thd * c) + ks_signed * pen;
  if (FORCES) {
    q_th += P.len * (c * fm_y - fm_x * s);
    q_x += fb_x + fm_x;
  }
  const double s_over_l = s * P.inv_len;
  acc_x = (q_x + q_th * s_over_l) * inv_d;
  acc_th = (q_x * s_over_l + q_th * P.kappa) * inv_d;
}

// Accelerations plus rows 2-3 of d(x_dot)/dx (fx[4], fth[4]) and the two non-zero entries of d(x_dot)/du.
// External forces are zero (the optimiser never passes any, optimization.cc:124-126).
//   d|v_m|/d(th, v, w) = (-l w c v_rel, v_x, l proj) / |v_m|   with v_rel = v_x + l w s
//   d v_x /d(th, v, w) = (-l w c, 1, -l s),   d proj/d(th, v, w) = (-c v_rel, -s, l)
// and d(b_x'')/dxi = (d num_x/dxi - b_x'' dD/dxi) / D with dD/dth = -2 m_1 s c (likewise th'').
CPMPC_DEV void dyn_eval_jac(const DynConst& P, const double b_x, const double th, const double b_xd, const double thd,
                            const double u, double& acc_x, double& acc_th, double fx[4], double fth[4], double& gx,
                            double& gth) {
  double s, c;
  fm::sincos_bf(th, s, c);
  const double fric = fm::tanh_bf(b_xd * P.inv_v_cut);
  const double over = fabs(b_x) - P.x_s;
  const bool act = 0.0 < over;
  const double pen = act ? over : 0.0;
  const double ks_signed = copysign(P.k_s, -b_x);

  const double D = fma(-P.m1, s * s, P.m_total);
  const double inv_d = fm::rcp(D);

  const double lw = thd * P.len;
  const double cc = c * c;
  const double vx = b_xd - lw * s;
  const double lw_cc = lw * cc;
  const double speed_sq = lw * lw_cc + vx * vx;
  double speed, inv_speed;
  fm::sqrt_rsqrt(speed_sq, speed, inv_speed);
  const bool moving = 0.0 < speed_sq;
  const double spd = moving ? speed : 0.0;       // every drag term and derivative below vanishes at rest
  const double ispd = moving ? inv_speed : 0.0;
  const double proj = lw_cc - vx * s;

  const double m1lw = thd * P.m1len;
  const double q_th = proj * (spd * P.k30len) + c * P.nlgm;
  const double q_x = u + fric * P.fric_k + vx * (spd * P.k30) + m1lw * (thd * c) + ks_signed * pen;
  const double s_over_l = s * P.inv_len;
  acc_x = (q_x + q_th * s_over_l) * inv_d;
  acc_th = (q_x * s_over_l + q_th * P.kappa) * inv_d;
  gx = inv_d;
  gth = s_over_l * inv_d;

  // derivatives of the two drag terms T_x = k30 |v_m| v_x and T_th = k30 l |v_m| proj
  const double vx_rel = vx + lw * s;
  const double ivx = ispd * vx;
  const double ipr = ispd * proj;
  const double tx_th = -(((lw * c) * P.k30) * (spd + ivx * vx_rel));
  const double tx_v = P.k30 * (spd + ivx * vx);
  const double x_w = P.k30len * (ivx * proj - s * spd);  // dT_x/dw == dT_th/dv
  const double tth_th = -(((c * vx_rel) * P.k30len) * (spd + ipr * lw));
  const double tth_w = P.k30len2 * (spd + ipr * proj);

  const double dq_x_dv = P.fric_kv * (1.0 - fric * fric) + tx_v;
  const double dq_x_dth = tx_th - (m1lw * thd) * s;
  const double dq_x_dw = (m1lw + m1lw) * c + x_w;
  const double dq_th_dth = tth_th - s * P.nlgm;
  const double spring_sw = act ? -P.k_s : 0.0;
  const double c_over_l = c * P.inv_len;
  const double msc2 = (P.m1x2 * s) * c;  // -dD/dth

  fx[0] = spring_sw * inv_d;
  fth[0] = spring_sw * gth;
  fx[1] = (dq_x_dth + dq_th_dth * s_over_l + q_th * c_over_l + acc_x * msc2) * inv_d;
  fth[1] = (dq_x_dth * s_over_l + q_x * c_over_l + dq_th_dth * P.kappa + acc_th * msc2) * inv_d;
  fx[2] = (dq_x_dv + x_w * s_over_l) * inv_d;
  fth[2] = (dq_x_dv * s_over_l + x_w * P.kappa) * inv_d;
  fx[3] = (dq_x_dw + tth_w * s_over_l) * inv_d;
  fth[3] = (dq_x_dw * s_over_l + tth_w * P.kappa) * inv_d;
}

// The branch-free sincos of fastmath.cuh is valid for |x| < 2^20.  Angles the integrator produces stay far inside (every
// step ends in mod_pi), but a caller may hand in any finite pole angle: states are passed through this on the way into
// a kernel.  sin / cos and the wrapped angle differences of the problem do not see the wrap (the reference evaluates
// libm sin / cos on the raw angle: the same value up to the rounding of 2 pi in the reduction, 1e-10 rad at 2^21).  For
// angles inside the range - every one a controller will ever see - nothing changes, not even the bits.
// The wrap itself is an out-of-line call behind a branch that is never taken in practice: inlined into the prologue of
// the solve kernel its software fmod moves that kernel's register allocation (1 % of the step, measured).
#if defined(__CUDACC__)
static __device__ __noinline__ double wrap_far_angle_slow(double th) { return mod_pi(th); }
__device__ __forceinline__ double wrap_far_angle(double th) {
  if (!(fabs(th) < 1048576.0)) th = wrap_far_angle_slow(th);
  return th;
}
#else
inline double wrap_far_angle(double th) { return (fabs(th) < 1048576.0) ? th : mod_pi(th); }
#endif

// One classic RK4 step, in place, no Jacobians (integration.hpp:53-62).  The four stages run as a rolled loop around
// ONE inlined copy of the dynamics so the hot code stays inside the instruction cache.
template <bool FORCES>
CPMPC_DEV void rk4_step(const DynConst& P, double x[4], const double u, const double h,
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
CPMPC_DEV void rk4_step_jac(const DynConst& P, double x[4], const double u, const double h,
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
