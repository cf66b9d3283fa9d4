// Plant-generic batched shooting-method MPC solve for sm_100a: the algorithm of solver_kernel.cuh (one sub-warp group of
// L lanes per instance, forward-mode linearisation per shooting segment, segment-wise Riccati QP, L1-merit polynomial
// line search and the Levenberg-Marquardt state machine of mini_opt, nonlinear.cc:75-605) written over a `Plant`
// policy with NX states instead of the single pendulum's four.  It exists for the cart-mounted DOUBLE pendulum
// (BASELINE.json configs[4]; physics spec symbolic/dynamics_double.py:25-148, for which the reference ships neither
// generated dynamics nor an optimiser: optimization/optimization.cc:197-199 carries the TODO) and is also instantiated
// for the single pendulum so that the generic code can be checked against the single-pendulum oracle.
//
// What is kept from the reference's controller (optimization.cc:39-371), with NX states per node:
//   variable layout [x(0) .. x(NS-1) | u(0) .. u(W-1)], costs (u, u-rate, u_0 - u_prev, terminal), terminal rule
//   (negative weight = hard equality), initial-state equalities, shooting defects with angle wrapping, retraction with
//   angle wrapping and clamps, warm start by shifting the controls, FillInitialGuess, ComputePredictedStates.
// What is not here (solver_kernel.cuh only): chunked linearisation, the fused closed loop.
//
// A Plant provides
//   NX, NQ = NX/2 (x = [q, q'], so rows 0..NQ-1 of the state Jacobian are [0 I]), NP (doubles of parameters per
//   instance), NC (doubles of hoisted constants), Const, make()/store()/load()/dummy(), is_angle(i),
//   eval<JAC>(P, a[NX], u, acc[NQ], Jq[NQ*NX], gu[NQ]): accelerations, rows NQ.. of d(x')/dx, d(acc)/du.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cpmpc.h"
#include "double_pendulum.cuh"
#include "dynamics.cuh"

namespace cpmpc {
namespace gp {

// mod_pi with the fmod expanded in place, as this file has always been compiled (cpmpc::mod_pi calls it out of line
// since round 2, which is worth 2 % to the single-pendulum kernel at large batches).  The value is the same function
// (profiles/microbench/fmod_probe.cu: 2^20 arguments, inline == call == host libm), but this translation unit's
// dynamics are plain `a * b + c` expressions whose multiply-add contraction is the compiler's choice, and that choice
// moved with the surrounding code: the double pendulum's results changed in the last bits (then grew, as they do on a
// chaotic plant).  Keeping the old form keeps this kernel's bits from build to build.
CPMPC_DEV double mod_pi(double a) {
  constexpr double pi = 3.141592653589793;
  constexpr double two_pi = 6.283185307179586;
  if (!(fabs(a) < two_pi)) a = fmod(a, two_pi);
  a += (a < 0.0) ? two_pi : 0.0;
  a -= (a > pi) ? two_pi : 0.0;
  return a;
}

// ---------------------------------------------------------------------------------------------------------------------
// Plants
// ---------------------------------------------------------------------------------------------------------------------
struct DoublePlant {
  static constexpr int NX = 6, NQ = 3, NP = CPMPC_DP_DYN_STRIDE, NC = 8;
  using Const = DpConst;
  __device__ __forceinline__ static Const make(const double* __restrict__ p) { return make_dp_const(p); }
  __device__ __forceinline__ static Const dummy() { return Const{3.0, 0.5, 0.25, 0.125, 0.0625, 0.0625, 4.9, 2.45}; }
  __device__ __forceinline__ static void store(double* d, const Const& c) {
    d[0] = c.mt, d[1] = c.h1, d[2] = c.h2, d[3] = c.a11, d[4] = c.a12, d[5] = c.a22, d[6] = c.g1, d[7] = c.g2;
  }
  __device__ __forceinline__ static void load(Const& c, const double* s) {
    c.mt = s[0], c.h1 = s[1], c.h2 = s[2], c.a11 = s[3], c.a12 = s[4], c.a22 = s[5], c.g1 = s[6], c.g2 = s[7];
  }
  __device__ __forceinline__ static constexpr bool is_angle(int i) { return i == 1 || i == 2; }
  template <bool JAC>
  __device__ __forceinline__ static void eval(const Const& P, const double a[NX], const double u, double acc[NQ],
                                              double* Jq, double* gu) {
    double j1[3], j2[3], j4[3], j5[3], ju[3];
    dp_eval<JAC>(P, a[1], a[2], a[4], a[5], u, acc, j1, j2, j4, j5, ju);
    if (JAC) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Jq[6 * r] = 0.0;
        Jq[6 * r + 1] = j1[r];
        Jq[6 * r + 2] = j2[r];
        Jq[6 * r + 3] = 0.0;
        Jq[6 * r + 4] = j4[r];
        Jq[6 * r + 5] = j5[r];
        gu[r] = ju[r];
      }
    }
  }
};

struct SinglePlant {
  static constexpr int NX = 4, NQ = 2, NP = CPMPC_DYN_STRIDE, NC = DYN_CONST_DOUBLES;
  using Const = DynConst;
  __device__ __forceinline__ static Const make(const double* __restrict__ p) { return make_dyn_const(p); }
  __device__ __forceinline__ static Const dummy() {
    return dummy_dyn_const();
  }
  __device__ __forceinline__ static void store(double* d, const Const& c) { store_dyn_const(d, c); }
  __device__ __forceinline__ static void load(Const& c, const double* s) { load_dyn_const(c, s); }
  __device__ __forceinline__ static constexpr bool is_angle(int i) { return i == 1; }
  template <bool JAC>
  __device__ __forceinline__ static void eval(const Const& P, const double a[NX], const double u, double acc[NQ],
                                              double* Jq, double* gu) {
    if (JAC) {
      double fx[4], fth[4], gx, gth;
      dyn_eval_jac(P, a[0], a[1], a[2], a[3], u, acc[0], acc[1], fx, fth, gx, gth);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Jq[j] = fx[j];
        Jq[4 + j] = fth[j];
      }
      gu[0] = gx;
      gu[1] = gth;
    } else {
      dyn_eval<false>(P, a[0], a[1], a[2], a[3], u, 0.0, 0.0, 0.0, acc[0], acc[1]);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// RK4 over a plant (integration.hpp:13-62)
// ---------------------------------------------------------------------------------------------------------------------
template <class PL>
__device__ __forceinline__ void rk4_step(const typename PL::Const& P, double x[PL::NX], const double u,
                                         const double h) {
  constexpr int NX = PL::NX, NQ = PL::NQ;
  const double hh = h * 0.5;
  double a[NX], s[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    a[i] = x[i];
    s[i] = 0.0;
  }
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    double acc[NQ];
    PL::template eval<false>(P, a, u, acc, nullptr, nullptr);
    const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
    const double cst = (st == 2) ? h : hh;
    double k[NX];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      k[i] = a[NQ + i];
      k[NQ + i] = acc[i];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      s[i] += k[i] * wgt;
      a[i] = x[i] + k[i] * cst;
    }
  }
  const double h6 = h / 6.0;
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] += h6 * s[i];
}

// One RK4 step with the Jacobians of the step map: A = d x+ / d x (row-major NX x NX), Bv = d x+ / d u.  Stage chain
// rule K_i = F_i (I + c_i K_{i-1}), ku_i = F_i ku_{i-1} c_i + Fu_i with F = [[0 I],[Jq]]: column j of the new K depends
// only on column j of the old one, so K is updated in place column by column.
template <class PL>
__device__ __forceinline__ void rk4_step_jac(const typename PL::Const& P, double x[PL::NX], const double u,
                                             const double h, double A[PL::NX * PL::NX], double Bv[PL::NX]) {
  constexpr int NX = PL::NX, NQ = PL::NQ;
  const double hh = h * 0.5;
  double a[NX], s[NX], K[NX * NX], ku[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    a[i] = x[i];
    s[i] = 0.0;
    ku[i] = 0.0;
    Bv[i] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) {
    K[i] = 0.0;
    A[i] = 0.0;
  }
  double cprev = 0.0;
#pragma unroll 1
  for (int st = 0; st < 4; ++st) {
    double acc[NQ], Jq[NQ * NX], gu[NQ];
    PL::template eval<true>(P, a, u, acc, Jq, gu);
    const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      double pm[NX];
#pragma unroll
      for (int r = 0; r < NX; ++r) pm[r] = ((r == j) ? 1.0 : 0.0) + K[NX * r + j] * cprev;
#pragma unroll
      for (int r = 0; r < NQ; ++r) {
        K[NX * r + j] = pm[NQ + r];
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < NX; ++q) v += Jq[NX * r + q] * pm[q];
        K[NX * (NQ + r) + j] = v;
      }
#pragma unroll
      for (int r = 0; r < NX; ++r) A[NX * r + j] += K[NX * r + j] * wgt;
    }
    {
      double kn[NX];
#pragma unroll
      for (int r = 0; r < NQ; ++r) {
        kn[r] = ku[NQ + r] * cprev;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < NX; ++q) v += Jq[NX * r + q] * ku[q];
        kn[NQ + r] = v * cprev + gu[r];
      }
#pragma unroll
      for (int r = 0; r < NX; ++r) {
        ku[r] = kn[r];
        Bv[r] += kn[r] * wgt;
      }
    }
    const double cst = (st == 2) ? h : hh;
    double k[NX];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      k[i] = a[NQ + i];
      k[NQ + i] = acc[i];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      s[i] += k[i] * wgt;
      a[i] = x[i] + k[i] * cst;
    }
    cprev = cst;
  }
  const double h6 = h / 6.0;
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) A[i] = ((i % (NX + 1) == 0) ? 1.0 : 0.0) + h6 * A[i];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    Bv[i] = h6 * Bv[i];
    x[i] += h6 * s[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Solver
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GP_MAX_NX = 6;

struct GpCfg {
  int W, S, NS, nseg, N, max_iter, max_ls;
  int condensed;  // 1: lambda == 0 iterations use the condensed QP (short horizons), 0: always Riccati
  int stride;
  int off_xs, off_u, off_dxs, off_du, off_phi, off_e, off_gam, off_aug, off_vbuf, off_invd, off_dyn, off_xc, off_uc;
  int kk_stride;
  double dt, rel_tol, first_tol, abs_tol, pen_init, pen_scale, rho, armijo_c1;
  double lam_fail_init, lam_dec_success, lam_dec_restore, lam_max, lam_min;
  double w_u, w_r, wt[GP_MAX_NX];
  double x_clamp, u_clamp;
};

struct GpIO {
  long long B;
  const double* states;      // [B,NX]
  const double* dyn;         // [B,NP]
  const double* set_points;  // [B]
  const double* weights;     // [B,2+NX] or nullptr
  const double* u_guess;     // [W]
  double* prev;              // [B,N]
  unsigned char* has_prev;   // [B]
  double* u_out;             // [B,W]
  double* pred_out;          // [B,W,NX] or nullptr
  cpmpc_stats* stats;        // [B] or nullptr
  double* iter_log;          // [B,max_iter,27] or nullptr
  double* kk_scratch;        // [n_groups, kk_stride]
  unsigned long long* queue; // [2]
  // split launch (cpmpc_plant_capi.cu Launch), as in solver_kernel.cuh: the serial rollouts in their own kernels
  const double* guess;       // [B, NX NS] shooting states of the initial guess, or nullptr: integrate them here
  int split;                 // 1: the predicted states are plant_post_kernel's
};

enum : int { SR_SUCCESS = 0, SR_MAX_ITERATIONS = 1, SR_FIRST_ORDER = 2, SR_POSITIVE_DERIVATIVE = 3,
             SR_NON_FINITE = 4, SR_INVALID_ALPHA = 5 };
enum : int { TS_MAX_ITERATIONS = 0, TS_ABS_TOL = 1, TS_REL_TOL = 2, TS_FIRST_ORDER = 3, TS_MAX_LAMBDA = 4,
             TS_QP_INDEFINITE = 5, TS_REJECTED = -1 };
enum : int { OS_NOMINAL = 0, OS_RESTORE_LM = 1 };

template <int L>
__device__ __forceinline__ double group_sum(double v, unsigned gmask) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return v < lo ? lo : (hi < v ? hi : v); }

// a[i] for a run-time i without forcing the array into local memory
template <int N>
__device__ __forceinline__ double sel(const double (&a)[N], const int i) {
  double v = a[0];
#pragma unroll
  for (int k = 1; k < N; ++k) v = (i == k) ? a[k] : v;
  return v;
}
template <int N>
__device__ __forceinline__ int seli(const int (&a)[N], const int i) {
  int v = a[0];
#pragma unroll
  for (int k = 1; k < N; ++k) v = (i == k) ? a[k] : v;
  return v;
}

template <int NX>
struct Weights {
  double w_u, w_r;
  double wt[NX];
};

struct Ctx {
  const GpCfg* cfg;
  double* sm;
  double* kk;
  int gl;
  unsigned gmask;
};

#define GP_GSYNC() __syncwarp(c.gmask)
#define GP_PARK() asm volatile("" ::: "memory")

template <class PL>
__device__ __forceinline__ double target_of(const int i, const double set_point) {
  constexpr double half_pi = 1.5707963267948966;
  return (i == 0) ? set_point : (PL::is_angle(i) ? half_pi : 0.0);
}

template <int L, class PL>
__device__ __forceinline__ double control_cost_partial(const Ctx& c, const Weights<PL::NX>& w, const double* uu,
                                                       const double u_prev) {
  const GpCfg& k = *c.cfg;
  double f = 0.0;
  for (int s = c.gl; s < k.nseg; s += L) {
    for (int i = 0; i < k.S; ++i) {
      const int col = s * k.S + i;
      const double uk = uu[col];
      if (w.w_r > 0.0) {
        if (col + 1 < k.W) {
          const double r = (uk - uu[col + 1]) * w.w_r;
          f += 0.5 * (r * r);
        }
        if (col == 0) {
          const double r = (uk - u_prev) * w.w_r;
          f += 0.5 * (r * r);
        }
      }
      if (w.w_u > 0.0) {
        const double r = uk * w.w_u;
        f += 0.5 * (r * r);
      }
    }
  }
  return f;
}

// initial-state equalities and terminal costs / equalities (optimization.cc:228-267)
template <class PL>
__device__ __forceinline__ void boundary_terms(const GpCfg& k, const Weights<PL::NX>& w, const double* xx,
                                               const double (&x_cur)[PL::NX], const double set_point, double& f_term,
                                               double& eq_bnd) {
  constexpr int NX = PL::NX;
  f_term = 0.0;
  eq_bnd = 0.0;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double d = xx[i] - x_cur[i];
    if (PL::is_angle(i)) d = mod_pi(d);
    eq_bnd += fabs(d);
  }
  const double* xN = xx + NX * (k.NS - 1);
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double d = xN[i] - target_of<PL>(i, set_point);
    if (PL::is_angle(i)) d = mod_pi(d);
    if (w.wt[i] >= 0.0) {
      const double r = d * w.wt[i];
      f_term += 0.5 * (r * r);
    } else {
      eq_bnd += fabs(d);
    }
  }
}

// EvaluateNonlinearErrors (nonlinear.cc:279-293): every lane integrates its own segments without Jacobians.
template <int L, class PL>
__device__ __forceinline__ void evaluate_errors(const Ctx& c, const Weights<PL::NX>& w, const double* xx,
                                                const double* uu, const double (&x_cur)[PL::NX],
                                                const double set_point, const double u_prev, double& f_out,
                                                double& eq_out) {
  constexpr int NX = PL::NX;
  const GpCfg& k = *c.cfg;
  double eq = 0.0;
  {
    typename PL::Const P;
    PL::load(P, c.sm + k.off_dyn);
    for (int s = c.gl; s < k.nseg; s += L) {
      double x[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xx[NX * s + i];
      for (int i = 0; i < k.S; ++i) rk4_step<PL>(P, x, uu[s * k.S + i], k.dt);
      const double* xn = xx + NX * (s + 1);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double e;
        if (PL::is_angle(i)) {
          e = mod_pi(mod_pi(x[i]) - xn[i]);
        } else {
          e = x[i] - xn[i];
        }
        eq += fabs(e);
      }
    }
  }
  double f = control_cost_partial<L, PL>(c, w, uu, u_prev);
  if (c.gl == 0) {
    double ft, eb;
    boundary_terms<PL>(k, w, xx, x_cur, set_point, ft, eb);
    f += ft;
    eq += eb;
  }
  f_out = group_sum<L>(f, c.gmask);
  eq_out = group_sum<L>(eq, c.gmask);
}

__device__ __forceinline__ int cbase(const int D, const int col) { return col * D - ((col * (col - 1)) >> 1) - col; }

__device__ __forceinline__ void axpy_neg(double* y, const double* x, const double a, const int n) {
  int r = 0;
  for (; r + 4 <= n; r += 4) {
    const double x0 = x[r], x1 = x[r + 1], x2 = x[r + 2], x3 = x[r + 3];
    const double y0 = y[r], y1 = y[r + 1], y2 = y[r + 2], y3 = y[r + 3];
    y[r] = fma(-a, x0, y0);
    y[r + 1] = fma(-a, x1, y1);
    y[r + 2] = fma(-a, x2, y2);
    y[r + 3] = fma(-a, x3, y3);
  }
  for (; r < n; ++r) y[r] = fma(-a, x[r], y[r]);
}

// ---------------------------------------------------------------------------------------------------------------------
// Equality-constrained QP (reference: QPNullSpaceSolver::Solve, qp.cc:679-729) by the segment-wise Riccati recursion of
// solver_kernel.cuh::riccati_qp, with NX states: stage state xi = (node step dx_s [NX], previous control), rho = (affine
// coordinate, one multiplier per hard terminal row).  Stage matrix over (S controls | xi | rho), column-major packed
// lower triangle, LDL^T elimination of the controls; see solver_kernel.cuh for the derivation.
// ---------------------------------------------------------------------------------------------------------------------
template <int L, class PL>
__device__ __forceinline__ bool riccati_qp(const Ctx& c, const Weights<PL::NX>& w, const double lambda,
                                           const double (&x_cur)[PL::NX], const double set_point, double& cdx) {
  constexpr int NX = PL::NX;
  constexpr int XI = NX + 1;    // stage state
  constexpr int NRM = NX + 1;   // rho coordinates at most
  constexpr int NCOL = XI + NRM;
  const GpCfg& k = *c.cfg;
  const int W = k.W, S = k.S, NS = k.NS, nseg = k.nseg;
  const double* xs = c.sm + k.off_xs;
  double* dxs = c.sm + k.off_dxs;
  double* du = c.sm + k.off_du;
  const double* Phi = c.sm + k.off_phi;
  const double* ev = c.sm + k.off_e;
  const double* Gam = c.sm + k.off_gam;
  double* aug = c.sm + k.off_aug;
  double* vb = c.sm + k.off_vbuf;  // P (XI x XI, full) then q [XI][NRM]
  double* qb = vb + XI * XI;
  double* invd = c.sm + k.off_invd;
  const double wu2 = w.w_u * w.w_u, wr2 = w.w_r * w.w_r;

  int nh = 0;
  int hcomp[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) hcomp[i] = 0;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    if (w.wt[i] < 0.0) {
#pragma unroll
      for (int j = 0; j < NX; ++j)
        if (j == nh) hcomp[j] = i;
      ++nh;
    }
  }
  const int nr = 1 + nh;
  const int n2 = XI + nr;
  const int D = S + n2;
  double dT[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double d = xs[NX * (NS - 1) + i] - target_of<PL>(i, set_point);
    if (PL::is_angle(i)) d = mod_pi(d);
    dT[i] = d;
  }

  // terminal value function over (xi | rho) in rows/cols S..D-1
  for (int b = c.gl; b < n2; b += L) {
    double* colp = aug + cbase(D, S + b) + S;
    for (int a = b; a < n2; ++a) {
      double v = 0.0;
      if (a < NX) {
        if (a == b) {
          const double wa = sel(w.wt, a);
          v = lambda + ((wa >= 0.0) ? wa * wa : 0.0);
        }
      } else if (a == XI && b < NX) {
        const double wb = sel(w.wt, b);
        v = (wb >= 0.0) ? (wb * wb) * sel(dT, b) : 0.0;
      } else if (a > XI && b < NX) {
        v = (b == seli(hcomp, a - XI - 1)) ? 1.0 : 0.0;
      }
      colp[a] = v;
    }
  }
  GP_GSYNC();

#define GP_TR(a, b) aug[cbase(D, S + (b)) + S + (a)]
#define GP_TRS(a, b) (((a) >= (b)) ? GP_TR(a, b) : GP_TR(b, a))
  bool ok = true;
  for (int s = nseg - 1; s >= 0; --s) {
    const double* Ph = Phi + NX * NX * s;
    double fv[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) fv[i] = ev[NX * s + i];
    // (i) snapshot P and q = p + (P f) e_0; fold f . p_r into the rho-rho column 0
    for (int t = c.gl; t < XI * XI + XI * nr + nh; t += L) {
      if (t < XI * XI) {
        const int a = t / XI, b = t - XI * a;
        vb[t] = GP_TRS(a, b);
      } else if (t < XI * XI + XI * nr) {
        const int t2 = t - XI * XI;
        const int r = t2 / XI, a = t2 - XI * r;
        double v = GP_TR(XI + r, a);
        if (r == 0) {
#pragma unroll
          for (int q = 0; q < NX; ++q) v += GP_TRS(a, q) * fv[q];
        }
        qb[a * NRM + r] = v;
      } else {
        const int r = t - XI * XI - XI * nr + 1;  // 1..nh
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < NX; ++q) acc += GP_TR(XI + r, q) * fv[q];
        GP_TR(XI + r, XI) += acc;
      }
    }
    GP_GSYNC();

    // (ii) assemble the stage matrix, whole columns per lane
    for (int t = c.gl; t < S + XI; t += L) {
      double* colp = aug + cbase(D, t);
      if (t < S) {
        const int j = t, col = s * S + j;
        double g[XI];
#pragma unroll
        for (int q = 0; q < NX; ++q) g[q] = Gam[q * W + col];
        g[NX] = (j == S - 1) ? 1.0 : 0.0;
        double PG[XI];
#pragma unroll
        for (int a = 0; a < XI; ++a) {
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < XI; ++q) v += vb[XI * a + q] * g[q];
          PG[a] = v;
        }
        for (int i = j; i < S; ++i) {
          const int ci = s * S + i;
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < NX; ++q) v += Gam[q * W + ci] * PG[q];
          if (i == S - 1) v += PG[NX];
          if (i == j) v += wu2 + lambda + wr2 * ((col + 1 < W) ? 2.0 : 1.0);
          if (i == j + 1) v -= wr2;
          colp[i] = v;
        }
#pragma unroll
        for (int cc = 0; cc < NX; ++cc) {
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < NX; ++q) v += Ph[NX * q + cc] * PG[q];
          colp[S + cc] = v;
        }
        colp[S + NX] = (j == 0 && s > 0) ? -wr2 : 0.0;
        for (int r = 0; r < nr; ++r) {
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < XI; ++q) v += g[q] * qb[q * NRM + r];
          if (r == 0) v += du[col];
          colp[S + XI + r] = v;
        }
      } else {
        const int cc = t - S;
        if (cc < NX) {
          double hcol[NX], PF[NX];
#pragma unroll
          for (int q = 0; q < NX; ++q) hcol[q] = Ph[NX * q + cc];
#pragma unroll
          for (int a = 0; a < NX; ++a) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < NX; ++q) v += vb[XI * a + q] * hcol[q];
            PF[a] = v;
          }
          for (int c2 = cc; c2 < NX; ++c2) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < NX; ++q) v += Ph[NX * q + c2] * PF[q];
            if (c2 == cc && s > 0) v += lambda;
            colp[S + c2] = v;
          }
          colp[S + NX] = 0.0;
          for (int r = 0; r < nr; ++r) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < NX; ++q) v += hcol[q] * qb[q * NRM + r];
            colp[S + XI + r] = v;
          }
        } else {
          colp[S + NX] = 0.0;
          for (int r = 0; r < nr; ++r) colp[S + XI + r] = 0.0;
        }
      }
    }
    GP_GSYNC();

    // (iii) eliminate the S controls.  Same two bit-identical forms as solver_kernel.cuh: right-looking for wide lane
    // groups, left-looking (register accumulation, phase A control columns / phase B trailing block) for L <= 4.
    if constexpr (L >= 8) {
    if (c.gl == 0) {
      const double piv = aug[0];
      invd[0] = (piv > 0.0) ? fm::rcp(piv) : -1.0;
    }
    GP_GSYNC();
    for (int j = 0; j < S; ++j) {
      const double rd = invd[j];
      if (!(rd > 0.0)) {
        ok = false;
        break;
      }
      const double* colj = aug + cbase(D, j);
      for (int cc = j + 1 + c.gl; cc < D; cc += L) {
        double* colc = aug + cbase(D, cc);
        const double l = colj[cc] * rd;
        if (cc == j + 1 && cc < S) {
          const double piv = fma(-l, colj[cc], colc[cc]);
          colc[cc] = piv;
          invd[cc] = (piv > 0.0) ? fm::rcp(piv) : -1.0;
          axpy_neg(colc + cc + 1, colj + cc + 1, l, D - cc - 1);
        } else {
          axpy_neg(colc + cc, colj + cc, l, D - cc);
        }
      }
      GP_GSYNC();
    }
    } else {
      constexpr int R = 4;  // rows per register chunk
      for (int j = 0; j < S; ++j) {
        const int cbj = cbase(D, j);
        for (int r0 = j + R * c.gl; r0 < D; r0 += R * L) {
          int rr[R];
          double y[R];
#pragma unroll
          for (int t = 0; t < R; ++t) {
            rr[t] = min(r0 + t, D - 1);
            y[t] = aug[cbj + rr[t]];
          }
          int cb = 0;
          for (int i = 0; i < j; ++i) {
            const double* ci = aug + cb;  // column i: entry (r, i) at ci[r]
            const double l = ci[j] * invd[i];
#pragma unroll
            for (int t = 0; t < R; ++t) y[t] = fma(-l, ci[rr[t]], y[t]);
            cb += D - i - 1;
          }
#pragma unroll
          for (int t = 0; t < R; ++t)
            if (r0 + t < D) aug[cbj + r0 + t] = y[t];
          if (r0 == j) invd[j] = (y[0] > 0.0) ? fm::rcp(y[0]) : -1.0;
        }
        GP_GSYNC();
        if (!(invd[j] > 0.0)) {
          ok = false;
          break;
        }
      }
      if (ok) {
        for (int cc = S + c.gl; cc < D; cc += L) {
          const int cbc = cbase(D, cc);
          for (int r0 = cc; r0 < D; r0 += R) {
            int rr[R];
            double y[R];
#pragma unroll
            for (int t = 0; t < R; ++t) {
              rr[t] = min(r0 + t, D - 1);
              y[t] = aug[cbc + rr[t]];
            }
            int cb = 0;
            for (int i = 0; i < S; ++i) {
              const double* ci = aug + cb;
              const double l = ci[cc] * invd[i];
#pragma unroll
              for (int t = 0; t < R; ++t) y[t] = fma(-l, ci[rr[t]], y[t]);
              cb += D - i - 1;
            }
#pragma unroll
            for (int t = 0; t < R; ++t)
              if (r0 + t < D) aug[cbc + r0 + t] = y[t];
          }
        }
      }
      GP_GSYNC();
    }
    if (!ok) break;

    // (iv) gains: row S + q of the panel holds Z[:, q]; x_i = rd_i (z_i - sum_{j > i} a(j, i) x_j).
    {
      // Left-looking form: the sum of one unknown is accumulated in a register in the order j = S-1 .. i+1 (the order
      // in which a right-looking sweep applies the same FMAs to z_i: bit-identical to the first implementation,
      // profiles/ab_compare.py); solved unknowns are parked in the dead z slots of their columns, the latest one stays
      // in a register, and two right-hand sides are interleaved per lane for instruction-level parallelism.
      for (int q0 = c.gl; q0 < n2; q0 += 2 * L) {
        const int q1 = q0 + L;
        const bool two = q1 < n2;
        const int ra = S + q0, rb = two ? S + q1 : ra;
        double* dsta = c.kk + ((size_t)s * NCOL + q0) * S;
        double* dstb = c.kk + ((size_t)s * NCOL + (two ? q1 : q0)) * S;
        double xa = 0.0, xb = 0.0;
        int cbi = cbase(D, S - 1);
        const int cb_last = cbi;
        for (int i = S - 1; i >= 0; --i) {
          double* ci = aug + cbi;  // column i: a(j, i) at ci[j]
          double acca = ci[ra], accb = ci[rb];
          int ob = cb_last;  // column j = S-1
          for (int j = S - 1; j > i + 1; --j) {
            const double a = ci[j];
            acca = fma(-a, aug[ob + ra], acca);
            accb = fma(-a, aug[ob + rb], accb);
            ob -= (D - j);  // cbase(D, j - 1)
          }
          if (i < S - 1) {
            const double a = ci[i + 1];
            acca = fma(-a, xa, acca);
            accb = fma(-a, xb, accb);
          }
          const double rd = invd[i];
          xa = acca * rd;
          xb = accb * rd;
          __stcg(dsta + i, xa);
          ci[ra] = xa;
          if (two) {
            __stcg(dstb + i, xb);
            ci[rb] = xb;
          }
          cbi -= (D - i);  // cbase(D, i - 1)
        }
      }
    }
    GP_GSYNC();
  }
  ok = __all_sync(c.gmask, ok);
  if (!ok) return false;

  // multipliers of the hard terminal rows
  double m0[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double d = xs[i] - x_cur[i];
    if (PL::is_angle(i)) d = mod_pi(d);
    m0[i] = -d;
  }
  double coef[NRM];
  coef[0] = 1.0;
#pragma unroll
  for (int i = 1; i < NRM; ++i) coef[i] = 0.0;
  if (nh > 0) {
    double Mh[NX][NX], rhs[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      rhs[j] = 0.0;
#pragma unroll
      for (int q = 0; q < NX; ++q) Mh[j][q] = (j == q) ? 1.0 : 0.0;
      if (j < nh) {
        const int rj = XI + 1 + j;
        double acc = GP_TR(rj, XI) + sel(dT, seli(hcomp, j));
#pragma unroll
        for (int q = 0; q < NX; ++q) acc += GP_TR(rj, q) * m0[q];
        rhs[j] = acc;
#pragma unroll
        for (int q = 0; q < NX; ++q)
          if (q <= j) Mh[j][q] = -GP_TR(rj, XI + 1 + q);
      }
    }
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      double d = Mh[j][j];
#pragma unroll
      for (int q = 0; q < j; ++q) d -= Mh[j][q] * Mh[j][q];
      if (!(d > 0.0)) ok = false;
      const double inv = 1.0 / sqrt(d);
      Mh[j][j] = inv;
#pragma unroll
      for (int i = j + 1; i < NX; ++i) {
        double sacc = Mh[i][j];
#pragma unroll
        for (int q = 0; q < j; ++q) sacc -= Mh[i][q] * Mh[j][q];
        Mh[i][j] = sacc * inv;
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double sacc = rhs[i];
#pragma unroll
      for (int q = 0; q < i; ++q) sacc -= Mh[i][q] * rhs[q];
      rhs[i] = sacc * Mh[i][i];
    }
#pragma unroll
    for (int i = NX - 1; i >= 0; --i) {
      double sacc = rhs[i];
#pragma unroll
      for (int q = i + 1; q < NX; ++q) sacc -= Mh[q][i] * rhs[q];
      rhs[i] = sacc * Mh[i][i];
    }
#pragma unroll
    for (int j = 0; j < NX; ++j) coef[1 + j] = (j < nh) ? rhs[j] : 0.0;
    if (!ok) return false;
  }

  // forward sweep
  double xi[XI];
#pragma unroll
  for (int i = 0; i < NX; ++i) xi[i] = m0[i];
  xi[NX] = 0.0;
  if (c.gl == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) dxs[i] = m0[i];
  }
  cdx = 0.0;
  for (int s = 0; s < nseg; ++s) {
    const double* kks = c.kk + (size_t)s * NCOL * S;
    for (int j = c.gl; j < S; j += L) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < XI; ++q) acc += __ldcg(kks + q * S + j) * xi[q];
#pragma unroll
      for (int r = 0; r < NRM; ++r)
        if (r < nr) acc += __ldcg(kks + (XI + r) * S + j) * coef[r];
      const int col = s * S + j;
      const double v = -acc;
      cdx += du[col] * v;
      du[col] = v;
    }
    GP_GSYNC();
    const double* Ph = Phi + NX * NX * s;
    for (int a = c.gl; a < NX; a += L) {
      double n0 = ev[NX * s + a];
#pragma unroll
      for (int q = 0; q < NX; ++q) n0 += Ph[NX * a + q] * xi[q];
      const double* ga = Gam + a * W + s * S;
      const double* vv = du + s * S;
      double n1 = 0.0;
      for (int j = 0; j < S; ++j) n1 += ga[j] * vv[j];
      dxs[NX * (s + 1) + a] = n0 + n1;
    }
    GP_GSYNC();
#pragma unroll
    for (int a = 0; a < NX; ++a) xi[a] = dxs[NX * (s + 1) + a];
    xi[NX] = du[s * S + S - 1];
  }
  return true;
#undef GP_TR
#undef GP_TRS
}

// ---------------------------------------------------------------------------------------------------------------------
// The same QP in condensed form for undamped iterations on short horizons (cfg.condensed), NX-generic version of
// condensed_qp in solver_kernel.cuh: with the shooting states eliminated forward through the linearised dynamics the
// Hessian over the W controls is T + U^T D U, T the tridiagonal control / control-rate Hessian (factored once per
// solve) and U the NX rows of terminal-state sensitivities (soft rows carry weight d = w^2, hard rows are
// constraints).  With Z = L^-1 U^T the multipliers come from an NX x NX system and du from one more bidiagonal solve.
// Scratch (aliases the Riccati stage matrix): Z [NX][W], tl [W], tinv [W].
// ---------------------------------------------------------------------------------------------------------------------
template <int L, class PL>
__device__ __forceinline__ bool condensed_qp(const Ctx& c, const Weights<PL::NX>& w, bool& tri_valid,
                                             const double (&x_cur)[PL::NX], const double set_point, double& cdx) {
  constexpr int NX = PL::NX;
  constexpr int NP = NX * (NX + 1) / 2;  // Gram pairs (i <= j), then NX gradient entries
  const GpCfg& k = *c.cfg;
  const int W = k.W, S = k.S, NS = k.NS, nseg = k.nseg;
  const double* xs = c.sm + k.off_xs;
  double* dxs = c.sm + k.off_dxs;
  double* du = c.sm + k.off_du;
  const double* Phi = c.sm + k.off_phi;
  const double* ev = c.sm + k.off_e;
  const double* Gam = c.sm + k.off_gam;
  double* Z = c.sm + k.off_aug;  // [NX][W]
  double* tl = Z + NX * W;       // sub-diagonal multipliers of the unit bidiagonal factor
  double* tinv = tl + W;         // reciprocal pivots
  const double wu2 = w.w_u * w.w_u, wr2 = w.w_r * w.w_r;
  bool ok = true;

  // (1) factor T once per solve (lane L-1, while the other lanes build the rows of U)
  if (!tri_valid && c.gl == L - 1) {
    double dprev_inv = 0.0;
    #pragma unroll 1
    for (int j = 0; j < W; ++j) {
      const double a = wu2 + wr2 * ((j + 1 < W) ? 2.0 : 1.0);
      double l = 0.0, d = a;
      if (j > 0) {
        l = -wr2 * dprev_inv;
        d = fma(l, wr2, a);
      }
      ok = ok && (d > 0.0);
      dprev_inv = fm::rcp(d);
      tl[j] = l;
      tinv[j] = dprev_inv;
    }
  }
  tri_valid = true;
  // (2) rows of U: sensitivity of the terminal node to every control, through the later segments
  #pragma unroll 1
  for (int col = c.gl; col < W; col += L) {
    const int s = col / S;
    double v[NX];
#pragma unroll
    for (int a = 0; a < NX; ++a) v[a] = Gam[a * W + col];
    #pragma unroll 1
    for (int t = s + 1; t < nseg; ++t) {
      const double* Ph = Phi + NX * NX * t;
      double n[NX];
#pragma unroll
      for (int r = 0; r < NX; ++r) {
        double acc = Ph[NX * r] * v[0];
#pragma unroll
        for (int q = 1; q < NX; ++q) acc += Ph[NX * r + q] * v[q];
        n[r] = acc;
      }
#pragma unroll
      for (int r = 0; r < NX; ++r) v[r] = n[r];
    }
#pragma unroll
    for (int a = 0; a < NX; ++a) Z[a * W + col] = v[a];
  }
  // node steps for du = 0: m_0 = -(x_0 - x_cur), m_{s+1} = Phi_s m_s + e_s (every lane, registers)
  double m[NX], m0[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double d = xs[i] - x_cur[i];
    if (PL::is_angle(i)) d = mod_pi(d);
    m[i] = -d;
    m0[i] = -d;
  }
  #pragma unroll 1
  for (int s = 0; s < nseg; ++s) {
    const double* Ph = Phi + NX * NX * s;
    double n[NX];
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      double acc = Ph[NX * r] * m[0];
#pragma unroll
      for (int q = 1; q < NX; ++q) acc += Ph[NX * r + q] * m[q];
      n[r] = acc + ev[NX * s + r];
    }
#pragma unroll
    for (int r = 0; r < NX; ++r) m[r] = n[r];
  }
  GP_GSYNC();
  ok = __all_sync(c.gmask, ok);
  if (!ok) return false;

  // (3) forward substitution with the unit bidiagonal factor: the NX rows of Z and the gradient (in du[]), one
  // right-hand side per lane, the running value carried in a register
  #pragma unroll 1
  for (int r = c.gl; r < NX + 1; r += L) {
    double* row = (r < NX) ? (Z + r * W) : du;
    double prev = row[0];
    #pragma unroll 1
    for (int j = 1; j < W; ++j) {
      prev = fma(-tl[j], prev, row[j]);
      row[j] = prev;
    }
  }
  GP_GSYNC();

  // (4) Gram entries sum_q Z_i[q] Z_j[q] / d_q (i <= j) and sum_q Z_i[q] g[q] / d_q: every lane accumulates over its
  // own columns, a group sum finishes each entry (all lanes need all of them)
  double gm[NP + NX];
#pragma unroll
  for (int e = 0; e < NP + NX; ++e) gm[e] = 0.0;
  #pragma unroll 1
  for (int q = c.gl; q < W; q += L) {
    const double ti = tinv[q];
    const double zg = du[q];
    double z[NX], sc[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      z[i] = Z[i * W + q];
      sc[i] = z[i] * ti;
    }
    int e = 0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int j = i; j < NX; ++j) {
        gm[e] = fma(sc[i], z[j], gm[e]);
        ++e;
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) gm[NP + i] = fma(sc[i], zg, gm[NP + i]);
  }
#pragma unroll
  for (int e = 0; e < NP + NX; ++e) gm[e] = group_sum<L>(gm[e], c.gmask);

  // (5) NX x NX system for the row multipliers, every lane redundantly in registers
  double A[NX][NX];
  {
    int e = 0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int j = i; j < NX; ++j) {
        A[i][j] = gm[e];
        A[j][i] = gm[e];
        ++e;
      }
    }
  }
  double sig[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double d = xs[NX * (NS - 1) + i] - target_of<PL>(i, set_point);
    if (PL::is_angle(i)) d = mod_pi(d);
    double invd, off;
    if (w.wt[i] < 0.0) {
      invd = 0.0;  // hard row
      off = m[i] + d;
    } else {
      const double dd = w.wt[i] * w.wt[i];
      if (dd > 0.0) {
        invd = 1.0 / dd;
        off = m[i] + d;
      } else {
        invd = -1.0;  // zero weight: the row does not exist
        off = 0.0;
      }
    }
    if (invd < 0.0) {
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        A[i][j] = 0.0;
        A[j][i] = 0.0;
      }
      A[i][i] = 1.0;
      sig[i] = 0.0;
    } else {
      A[i][i] += invd;
      sig[i] = off - gm[NP + i];
    }
  }
#pragma unroll
  for (int j = 0; j < NX; ++j) {
    double d = A[j][j];
#pragma unroll
    for (int q = 0; q < j; ++q) d -= A[j][q] * A[j][q];
    if (!(d > 0.0)) ok = false;
    const double inv = 1.0 / sqrt(d);
    A[j][j] = inv;
#pragma unroll
    for (int i = j + 1; i < NX; ++i) {
      double sacc = A[i][j];
#pragma unroll
      for (int q = 0; q < j; ++q) sacc -= A[i][q] * A[j][q];
      A[i][j] = sacc * inv;
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double sacc = sig[i];
#pragma unroll
    for (int q = 0; q < i; ++q) sacc -= A[i][q] * sig[q];
    sig[i] = sacc * A[i][i];
  }
#pragma unroll
  for (int i = NX - 1; i >= 0; --i) {
    double sacc = sig[i];
#pragma unroll
    for (int q = i + 1; q < NX; ++q) sacc -= A[q][i] * sig[q];
    sig[i] = sacc * A[i][i];
  }

  // (6) du = -L^-T D^-1 (z_g + Z sigma); c_u . du = -(L^-1 g)^T D^-1 (z_g + Z sigma)
  double part = 0.0;
  #pragma unroll 1
  for (int q = c.gl; q < W; q += L) {
    const double zg = du[q];
    double acc = zg;
#pragma unroll
    for (int i = 0; i < NX; ++i) acc += sig[i] * Z[i * W + q];
    const double wq = acc * tinv[q];
    part = fma(zg, wq, part);
    du[q] = wq;
  }
  cdx = -part;  // this lane's share of c_u . du
  GP_GSYNC();
  if (c.gl == 0) {
    double nxt = du[W - 1];
    du[W - 1] = -nxt;
    #pragma unroll 1
    for (int j = W - 2; j >= 0; --j) {
      nxt = fma(-tl[j + 1], nxt, du[j]);
      du[j] = -nxt;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) dxs[i] = m0[i];
  }
  GP_GSYNC();

  // (7) node steps by forward propagation
  double xi[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) xi[i] = m0[i];
  #pragma unroll 1
  for (int s = 0; s < nseg; ++s) {
    const double* Ph = Phi + NX * NX * s;
    #pragma unroll 1
    for (int a = c.gl; a < NX; a += L) {
      double n0 = ev[NX * s + a];
#pragma unroll
      for (int q = 0; q < NX; ++q) n0 += Ph[NX * a + q] * xi[q];
      const double* ga = Gam + a * W + s * S;
      const double* vv = du + s * S;
      double n1 = 0.0;
      #pragma unroll 1
      for (int j = 0; j < S; ++j) n1 = fma(ga[j], vv[j], n1);
      dxs[NX * (s + 1) + a] = n0 + n1;
    }
    GP_GSYNC();
#pragma unroll
    for (int a = 0; a < NX; ++a) xi[a] = dxs[NX * (s + 1) + a];
  }
  return true;
}

struct SolveResult {
  int termination, n_iter, n_ls, n_failed, last_step_result, opt_state;
  double f, eq, penalty, lambda;
};

// The whole Optimization::Step for one instance (see solver_kernel.cuh::solve_group for the commentary).
template <int L, class PL>
__device__ __forceinline__ void solve_group(const Ctx& c, const bool active_in, const Weights<PL::NX>& w,
                                            const double (&x_cur)[PL::NX], const double& set_point,
                                            const double& u_prev, double* iter_log, SolveResult& res) {
  constexpr int NX = PL::NX;
  const GpCfg& k = *c.cfg;
  double* xs = c.sm + k.off_xs;
  double* u = c.sm + k.off_u;
  double* xc = c.sm + k.off_xc;
  double* uc = c.sm + k.off_uc;
  double* dxs = c.sm + k.off_dxs;
  double* du = c.sm + k.off_du;
  double* Phi = c.sm + k.off_phi;
  double* ev = c.sm + k.off_e;
  double* Gam = c.sm + k.off_gam;
  const int W = k.W, S = k.S, NS = k.NS, nseg = k.nseg;

  struct Cold {
    double lambda, penalty, cur_f, cur_eq;
    double f0, eq0, d_f, d_eq, lambda_in;
    double alpha, a_prev, phi_prev, a_prev2, phi_prev2, step_f, step_eq, phi0, dphi0, linf;
    int opt_state, step_result, n_steps;
    bool active, have_cur, ls_done, tri_valid;
  } cold;
  asm volatile("" : : "l"(&cold) : "memory");
  bool& active = cold.active;
  double& lambda = cold.lambda;
  double& penalty = cold.penalty;
  int& opt_state = cold.opt_state;
  double& cur_f = cold.cur_f;
  double& cur_eq = cold.cur_eq;
  bool& have_cur = cold.have_cur;
  bool& tri_valid = cold.tri_valid;
  tri_valid = false;  // factorisation of T (condensed QP) is cached across the iterations of one solve
  active = active_in;
  lambda = 0.0;
  penalty = k.pen_init;
  opt_state = OS_NOMINAL;
  cur_f = 0.0;
  cur_eq = 0.0;
  have_cur = false;
  res.termination = TS_MAX_ITERATIONS;
  res.n_iter = 0;
  res.n_ls = 0;
  res.n_failed = 0;
  res.last_step_result = SR_SUCCESS;

  // gradient of the control costs at the current controls into du[] (the QP routines consume and overwrite it)
  auto fill_gradient = [&]() {
    #pragma unroll 1
    for (int col = c.gl; col < W; col += L) {
      double g = 0.0;
      const double uk = u[col];
      if (w.w_r > 0.0) {
        const double wr2 = w.w_r * w.w_r;
        if (col + 1 < W) g += wr2 * (uk - u[col + 1]);
        if (col > 0) g -= wr2 * (u[col - 1] - uk);
        if (col == 0) g += wr2 * (uk - u_prev);
      }
      if (w.w_u > 0.0) g += (w.w_u * w.w_u) * uk;
      du[col] = g;
    }
    GP_GSYNC();
  };

  #pragma unroll 1
  for (int iter = 0; iter < k.max_iter; ++iter) {
    if (!__any_sync(0xffffffffu, active)) break;
    double& f0 = cold.f0;
    double& eq0 = cold.eq0;
    double& d_f = cold.d_f;
    double& d_eq = cold.d_eq;
    double& lambda_in = cold.lambda_in;
    f0 = eq0 = d_f = d_eq = 0.0;
    lambda_in = lambda;

    if (active) {
      // ---- linearise: one shooting segment per lane, Jacobian chain folded in forward mode -----------------------
      double eq_part = 0.0;
      GP_PARK();
      {
        typename PL::Const P;
        PL::load(P, c.sm + k.off_dyn);
        for (int s = c.gl; s < nseg; s += L) {
          double x[NX];
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = xs[NX * s + i];
          double* Ph = Phi + NX * NX * s;
          for (int i = 0; i < S; ++i) {
            double A[NX * NX], Bv[NX];
            const int col = s * S + i;
            rk4_step_jac<PL>(P, x, u[col], k.dt, A, Bv);
            if (i == 0) {
#pragma unroll
              for (int q = 0; q < NX * NX; ++q) Ph[q] = A[q];
            } else {
              // Phi <- A Phi, column by column in place
              for (int j = 0; j < NX; ++j) {
                double pc[NX];
#pragma unroll
                for (int r = 0; r < NX; ++r) pc[r] = Ph[NX * r + j];
#pragma unroll
                for (int r = 0; r < NX; ++r) {
                  double v = 0.0;
#pragma unroll
                  for (int q = 0; q < NX; ++q) v += A[NX * r + q] * pc[q];
                  Ph[NX * r + j] = v;
                }
              }
              for (int j = s * S; j < col; ++j) {
                double gc[NX];
#pragma unroll
                for (int r = 0; r < NX; ++r) gc[r] = Gam[r * W + j];
#pragma unroll
                for (int r = 0; r < NX; ++r) {
                  double v = 0.0;
#pragma unroll
                  for (int q = 0; q < NX; ++q) v += A[NX * r + q] * gc[q];
                  Gam[r * W + j] = v;
                }
              }
            }
#pragma unroll
            for (int r = 0; r < NX; ++r) Gam[r * W + col] = Bv[r];
          }
          const double* xn = xs + NX * (s + 1);
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            double e;
            if (PL::is_angle(i)) {
              e = mod_pi(mod_pi(x[i]) - xn[i]);
            } else {
              e = x[i] - xn[i];
            }
            ev[NX * s + i] = e;
            eq_part += fabs(e);
          }
        }
      }
      GP_PARK();
      double f_part = control_cost_partial<L, PL>(c, w, u, u_prev);
      if (c.gl == 0) {
        double ft, eb;
        boundary_terms<PL>(k, w, xs, x_cur, set_point, ft, eb);
        f_part += ft;
        eq_part += eb;
      }
      f0 = group_sum<L>(f_part, c.gmask);
      eq0 = group_sum<L>(eq_part, c.gmask);
      if (have_cur) {
        f0 = cur_f;
        eq0 = cur_eq;
      } else {
        cur_f = f0;
        cur_eq = eq0;
        have_cur = true;
      }
      // gradient of the control costs (optimization.cc:270-301), parked in du[]
      fill_gradient();
      if (!isfinite(f0) || !isfinite(eq0)) {
        res.termination = TS_REJECTED;  // the reference throws here (nonlinear.cc:356-357)
        active = false;
      }
    }

    if (active) {
      // ---- QP step ------------------------------------------------------------------------------------------------
      double cdx = 0.0;
      const bool try_condensed = k.condensed && !(lambda > 0.0);
      bool qp_ok = false;
      if (try_condensed) {
        qp_ok = condensed_qp<L, PL>(c, w, tri_valid, x_cur, set_point, cdx);
        // The NX x NX multiplier system of the condensed form lost positive definiteness (0.1 % of double-pendulum
        // instances: open-loop sensitivity rows nearly dependent).  That is a property of the formulation, not of
        // the QP: restore the gradient and let the closed-loop Riccati recursion decide.
        if (!qp_ok) fill_gradient();
      }
      if (!qp_ok) {
        tri_valid = false;  // the Riccati stage matrix shares the scratch of the tridiagonal factor
        qp_ok = riccati_qp<L, PL>(c, w, lambda, x_cur, set_point, cdx);
      }
      if (!qp_ok) {
        res.termination = TS_QP_INDEFINITE;
        active = false;
      } else {
        // dx^T G dx and the directional derivatives (nonlinear.cc:457-502)
        double quad = 0.0;
        const double wu2 = w.w_u * w.w_u, wr2 = w.w_r * w.w_r;
        #pragma unroll 1
        for (int col = c.gl; col < W; col += L) {
          const double a = wu2 + lambda + wr2 * ((col + 1 < W) ? 2.0 : 1.0);
          const double v = du[col];
          quad += a * v * v;
          if (col + 1 < W) quad -= 2.0 * wr2 * v * du[col + 1];
        }
        #pragma unroll 1
        for (int q = c.gl; q < NS; q += L) {
          const double* m = dxs + NX * q;
          double ss = 0.0;
#pragma unroll
          for (int i = 0; i < NX; ++i) ss += m[i] * m[i];
          quad += lambda * ss;
        }
        if (c.gl == 0) {
          const double* m = dxs + NX * (NS - 1);
          double deq_term = 0.0;
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            double d = xs[NX * (NS - 1) + i] - target_of<PL>(i, set_point);
            if (PL::is_angle(i)) d = mod_pi(d);
            if (w.wt[i] >= 0.0) {
              const double w2 = w.wt[i] * w.wt[i];
              cdx += (w2 * d) * m[i];
              quad += w2 * m[i] * m[i];
            } else {
              deq_term += ((d > 0.0) ? 1.0 : ((d < 0.0) ? -1.0 : 0.0)) * m[i];
            }
          }
          double l1_other = 0.0;
          #pragma unroll 1
          for (int q = 0; q < NX * nseg; ++q) l1_other += fabs(ev[q]);
#pragma unroll
          for (int i = 0; i < NX; ++i) l1_other += fabs(dxs[i]);
          d_eq = deq_term - l1_other;
        }
        GP_GSYNC();
        d_f = group_sum<L>(cdx, c.gmask);
        quad = group_sum<L>(quad, c.gmask);
        d_eq = group_sum<L>(d_eq, c.gmask);
        const double l1_eq = fmax(eq0, 2.220446049250313e-16);
        const double new_penalty = (d_f + 0.5 * fmax(0.0, quad)) / ((1.0 - k.rho) * l1_eq);
        if (new_penalty > penalty) penalty = new_penalty * k.pen_scale;
      }
    }

    // ---- line search (nonlinear.cc:351-443) --------------------------------------------------------------------------
    double& phi0 = cold.phi0;
    double& dphi0 = cold.dphi0;
    double& linf = cold.linf;
    phi0 = f0 + penalty * eq0;
    dphi0 = d_f + penalty * d_eq;
    linf = fmax(fabs(d_f), fabs(d_eq));
    int& step_result = cold.step_result;
    bool& ls_done = cold.ls_done;
    double& alpha = cold.alpha;
    double& a_prev = cold.a_prev;
    double& phi_prev = cold.phi_prev;
    double& a_prev2 = cold.a_prev2;
    double& phi_prev2 = cold.phi_prev2;
    double& step_f = cold.step_f;
    double& step_eq = cold.step_eq;
    int& n_steps = cold.n_steps;
    step_result = SR_MAX_ITERATIONS;
    ls_done = !active;
    alpha = 1.0;
    a_prev = phi_prev = a_prev2 = phi_prev2 = 0.0;
    step_f = step_eq = 0.0;
    n_steps = 0;
    double* lrow = (iter_log != nullptr && active) ? iter_log + (size_t)iter * CPMPC_ITERLOG_STRIDE : nullptr;
    #pragma unroll 1
    for (int ls = 0; ls <= k.max_ls; ++ls) {
      if (__all_sync(0xffffffffu, ls_done)) break;
      if (!ls_done && ls > 0) {
        double new_alpha = 0.0;
        bool valid;
        if (ls == 1) {
          const double numerator = phi_prev - dphi0 * a_prev - phi0;
          valid = !(dphi0 > 0.0 || numerator <= 0.0);
          new_alpha = -dphi0 * a_prev * a_prev / (2.0 * numerator);
        } else {
          const double a0 = a_prev2, a1 = a_prev;
          const double m00 = a0 * a0 * a0, m01 = a0 * a0, m10 = a1 * a1 * a1, m11 = a1 * a1;
          const double r0 = phi_prev2 - phi0 - dphi0 * a0;
          const double r1 = phi_prev - phi0 - dphi0 * a1;
          const double inv_det = 1.0 / (m00 * m11 - m10 * m01);
          const double ca = (m11 * inv_det) * r0 + (-m01 * inv_det) * r1;
          const double cb = (-m10 * inv_det) * r0 + (m00 * inv_det) * r1;
          const double arg_sqrt = cb * cb - 3.0 * ca * dphi0;
          valid = !(ca == 0.0 || arg_sqrt < -1.0e-12);
          new_alpha = (-cb + sqrt(fmax(arg_sqrt, 0.0))) / (3.0 * ca);
        }
        if (!valid || !isfinite(new_alpha) || new_alpha <= 0.0 || new_alpha >= alpha) {
          step_result = SR_INVALID_ALPHA;
          ls_done = true;
        } else {
          alpha = new_alpha;
        }
      }
      if (!ls_done) {
        // retraction (optimization.cc:309-329)
        for (int q = c.gl; q < NS; q += L) {
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            double v = __dadd_rn(xs[NX * q + i], __dmul_rn(dxs[NX * q + i], alpha));
            if (i == 0) v = clampd(v, -k.x_clamp, k.x_clamp);
            if (PL::is_angle(i)) v = mod_pi(v);
            xc[NX * q + i] = v;
          }
        }
        for (int col = c.gl; col < W; col += L)
          uc[col] = clampd(__dadd_rn(u[col], __dmul_rn(du[col], alpha)), -k.u_clamp, k.u_clamp);
        GP_GSYNC();
        GP_PARK();
        evaluate_errors<L, PL>(c, w, xc, uc, x_cur, set_point, u_prev, step_f, step_eq);
        GP_PARK();
        GP_GSYNC();
        if (lrow != nullptr && c.gl == 0 && n_steps < 6) {
          lrow[9 + 3 * n_steps] = alpha;
          lrow[10 + 3 * n_steps] = step_f;
          lrow[11 + 3 * n_steps] = step_eq;
        }
        ++n_steps;
        const double phi = step_f + penalty * step_eq;
        if (!isfinite(step_f) || !isfinite(step_eq)) {
          step_result = SR_NON_FINITE;
          ls_done = true;
        } else if (linf < k.first_tol) {
          step_result = SR_FIRST_ORDER;
          ls_done = true;
        } else if (dphi0 > 0.0) {
          step_result = SR_POSITIVE_DERIVATIVE;
          ls_done = true;
        } else if (phi <= phi0 + dphi0 * alpha * k.armijo_c1) {
          step_result = SR_SUCCESS;
          ls_done = true;
        } else {
          a_prev2 = a_prev;
          phi_prev2 = phi_prev;
          a_prev = alpha;
          phi_prev = phi;
        }
      }
    }

    // ---- lambda update and exit tests (nonlinear.cc:296-343) ----------------------------------------------------------
    if (active) {
      bool exit_now = false;
      if (step_result == SR_SUCCESS) {
        #pragma unroll 1
        for (int q = c.gl; q < NX * NS; q += L) xs[q] = xc[q];
        #pragma unroll 1
        for (int col = c.gl; col < W; col += L) u[col] = uc[col];
        GP_GSYNC();
        lambda = fmax(lambda * ((opt_state == OS_RESTORE_LM) ? k.lam_dec_restore : k.lam_dec_success), k.lam_min);
        opt_state = OS_NOMINAL;
        cur_f = step_f;
        cur_eq = step_eq;
        if (fmax(step_f, step_eq) < k.abs_tol) {
          res.termination = TS_ABS_TOL;
          exit_now = true;
        } else if (step_f + penalty * step_eq > phi0 * (1.0 - k.rel_tol)) {
          res.termination = TS_REL_TOL;
          exit_now = true;
        }
      } else if (step_result == SR_FIRST_ORDER) {
        res.termination = TS_FIRST_ORDER;
        exit_now = true;
      } else if (step_result == SR_MAX_ITERATIONS || step_result == SR_POSITIVE_DERIVATIVE) {
        if (opt_state == OS_NOMINAL) {
          lambda = fmax(k.lam_fail_init, lambda * 10.0);
          opt_state = OS_RESTORE_LM;
        } else {
          lambda *= 10.0;
        }
        if (lambda > k.lam_max) {
          res.termination = TS_MAX_LAMBDA;
          exit_now = true;
        }
      }
      res.n_iter += 1;
      res.n_ls += n_steps;
      res.n_failed += (step_result != SR_SUCCESS && step_result != SR_FIRST_ORDER) ? 1 : 0;
      res.last_step_result = step_result;
      if (lrow != nullptr && c.gl == 0) {
        lrow[0] = (double)opt_state;
        lrow[1] = lambda_in;
        lrow[2] = f0;
        lrow[3] = eq0;
        lrow[4] = d_f;
        lrow[5] = d_eq;
        lrow[6] = penalty;
        lrow[7] = (double)step_result;
        lrow[8] = (double)n_steps;
      }
      if (exit_now) active = false;
    }
  }
  res.opt_state = opt_state;
  res.f = cur_f;
  res.eq = cur_eq;
  res.penalty = penalty;
  res.lambda = lambda;
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel: work queue over instances, 32/L instances per warp at a time.
// ---------------------------------------------------------------------------------------------------------------------
template <int L, class PL>
__global__ void __launch_bounds__(256, 1) plant_solve_kernel(const GpCfg cfg, const GpIO io) {
  constexpr int NX = PL::NX;
  const int WARPS_PER_BLOCK = blockDim.x >> 5;
  extern __shared__ double smem[];
  constexpr int G = 32 / L;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane / L;
  Ctx c;
  c.cfg = &cfg;
  c.gl = lane % L;
  c.gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (g * L));
  const long long group_id = ((long long)blockIdx.x * WARPS_PER_BLOCK + warp) * G + g;
  c.sm = smem + (size_t)(warp * G + g) * cfg.stride;
  c.kk = io.kk_scratch + (size_t)group_id * cfg.kk_stride;
  const int W = cfg.W, S = cfg.S, NS = cfg.NS, N = cfg.N;
  double* xs = c.sm + cfg.off_xs;
  double* u = c.sm + cfg.off_u;

  #pragma unroll 1
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(io.queue, (unsigned long long)G);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned long long)io.B) break;
    const long long inst = (long long)base + g;
    const bool have = inst < io.B;

    struct KCold {
      Weights<NX> w;
      double x_cur[NX];
      double set_point, u_prev;
      SolveResult res;
    } kc;
    asm volatile("" : : "l"(&kc) : "memory");
    Weights<NX>& w = kc.w;
    double(&x_cur)[NX] = kc.x_cur;
    double& set_point = kc.set_point;
#pragma unroll
    for (int i = 0; i < NX; ++i) x_cur[i] = 0.0;
    set_point = 0.0;
    if (have) {
      if (c.gl == 0) PL::store(c.sm + cfg.off_dyn, PL::make(io.dyn + inst * PL::NP));
      if (io.weights != nullptr) {
        const double* pw = io.weights + inst * (2 + NX);
        w.w_u = pw[0];
        w.w_r = pw[1];
#pragma unroll
        for (int i = 0; i < NX; ++i) w.wt[i] = pw[2 + i];
      } else {
        w.w_u = cfg.w_u;
        w.w_r = cfg.w_r;
#pragma unroll
        for (int i = 0; i < NX; ++i) w.wt[i] = cfg.wt[i];
      }
      if (!(w.w_u > 0.0)) w.w_u = 0.0;
      if (!(w.w_r > 0.0)) w.w_r = 0.0;
      const double* ps = io.states + inst * NX;
#pragma unroll
      for (int i = 0; i < NX; ++i) x_cur[i] = ps[i];
      set_point = io.set_points[inst];
    } else {
      if (c.gl == 0) PL::store(c.sm + cfg.off_dyn, PL::dummy());
      w.w_u = w.w_r = 1.0;
#pragma unroll
      for (int i = 0; i < NX; ++i) w.wt[i] = 1.0;
    }
    GP_GSYNC();

    double& u_prev = kc.u_prev;
    u_prev = 0.0;
    if (have) {
      // BuildProblem + initial guess (optimization.cc:44-71)
      double* prev = io.prev + inst * N;
      const bool warm = io.has_prev[inst] != 0;
      if (warm) u_prev = prev[NX * NS];
      #pragma unroll 1
      for (int q = c.gl; q < W; q += L) u[q] = warm ? prev[NX * NS + ((q + 1 < W) ? q + 1 : q)] : io.u_guess[q];
      GP_GSYNC();
      if (io.guess != nullptr) {
        const double* gx = io.guess + inst * (NX * NS);
        #pragma unroll 1
        for (int q = c.gl; q < NX * NS; q += L) xs[q] = gx[q];
      } else if (c.gl == 0) {
        // FillInitialGuess (optimization.cc:333-351)
        typename PL::Const Pl;
        PL::load(Pl, c.sm + cfg.off_dyn);
        double x[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          x[i] = x_cur[i];
          xs[i] = x[i];
        }
        #pragma unroll 1
        for (int s = 1; s < NS; ++s) {
          #pragma unroll 1
          for (int i = 0; i < S; ++i) {
            rk4_step<PL>(Pl, x, u[(s - 1) * S + i], cfg.dt);
#pragma unroll
            for (int q = 0; q < NX; ++q)
              if (PL::is_angle(q)) x[q] = mod_pi(x[q]);
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) xs[NX * s + i] = x[i];
        }
      }
      GP_GSYNC();
    }

    SolveResult& res = kc.res;
    double* ilog = (io.iter_log != nullptr && have)
                       ? io.iter_log + (size_t)inst * cfg.max_iter * CPMPC_ITERLOG_STRIDE
                       : nullptr;
    solve_group<L, PL>(c, have, w, x_cur, set_point, u_prev, ilog, res);

    if (have) {
      double* prev = io.prev + inst * N;
      #pragma unroll 1
      for (int q = c.gl; q < NX * NS; q += L) prev[q] = xs[q];
      #pragma unroll 1
      for (int q = c.gl; q < W; q += L) {
        prev[NX * NS + q] = u[q];
        if (io.u_out != nullptr) io.u_out[inst * W + q] = u[q];
      }
      if (c.gl == 0) {
        io.has_prev[inst] = 1;
        if (io.pred_out != nullptr && !io.split) {
          // ComputePredictedStates (optimization.cc:353-371)
          typename PL::Const Pl;
          PL::load(Pl, c.sm + cfg.off_dyn);
          double x[NX];
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = x_cur[i];
          double* po = io.pred_out + (size_t)inst * W * NX;
          #pragma unroll 1
          for (int q = 0; q < W; ++q) {
            rk4_step<PL>(Pl, x, u[q], cfg.dt);
#pragma unroll
            for (int i = 0; i < NX; ++i) {
              if (PL::is_angle(i)) x[i] = mod_pi(x[i]);
              po[NX * q + i] = x[i];
            }
          }
        }
        if (io.stats != nullptr) {
          cpmpc_stats st;
          st.termination_state = res.termination;
          st.n_iterations = res.n_iter;
          st.n_line_search_steps = res.n_ls;
          st.n_failed_line_searches = res.n_failed;
          st.last_step_result = res.last_step_result;
          st.optimizer_state = res.opt_state;
          st.f = res.f;
          st.equality = res.eq;
          st.penalty = res.penalty;
          st.lambda = res.lambda;
          io.stats[inst] = st;
        }
      }
      GP_GSYNC();
    }
  }
  if (lane == 0) {
    __threadfence();
    const unsigned long long total = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    if (atomicAdd(io.queue + 1, 1ull) == total - 1) {
      io.queue[0] = 0ull;
      io.queue[1] = 0ull;
    }
  }
}

// The plant constants exactly as a lane of plant_solve_kernel sees them: every one a finished double that went through
// shared memory.  Without the barrier the compiler may contract a multiplication inside PL::make with an addition of the
// dynamics into one FMA, and the rollouts of the split launch would round differently from the fused kernel's.
template <class PL>
__device__ __forceinline__ typename PL::Const make_plant_const_opaque(const double* __restrict__ p) {
  double tmp[PL::NC];
  PL::store(tmp, PL::make(p));
#pragma unroll
  for (int i = 0; i < PL::NC; ++i) asm volatile("" : "+d"(tmp[i]));
  typename PL::Const c;
  PL::load(c, tmp);
  return c;
}

// Split launch: FillInitialGuess (optimization.cc:333-351) and ComputePredictedStates (optimization.cc:353-371) with one
// THREAD per instance, before and after plant_solve_kernel (see cpmpc_guess_kernel / cpmpc_post_kernel in
// solver_kernel.cuh).  Same controls, same rk4_step, same order: bit-identical to the fused path.
template <class PL>
__global__ void __launch_bounds__(128) plant_guess_kernel(const GpCfg cfg, const GpIO io, double* __restrict__ guess) {
  constexpr int NX = PL::NX;
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= io.B) return;
  const int W = cfg.W, S = cfg.S, NS = cfg.NS, N = cfg.N;
  const typename PL::Const P = make_plant_const_opaque<PL>(io.dyn + inst * PL::NP);
  const double* ps = io.states + inst * NX;
  double x[NX];
  double* gx = guess + inst * (NX * NS);
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x[i] = ps[i];
    gx[i] = x[i];
  }
  const bool warm = io.has_prev[inst] != 0;
  const double* pu = warm ? io.prev + inst * N + NX * NS : io.u_guess;
  #pragma unroll 1
  for (int s = 1; s < NS; ++s) {
    #pragma unroll 1
    for (int i = 0; i < S; ++i) {
      const int q = (s - 1) * S + i;
      const double uq = warm ? pu[(q + 1 < W) ? q + 1 : q] : pu[q];
      rk4_step<PL>(P, x, uq, cfg.dt);
#pragma unroll
      for (int a = 0; a < NX; ++a)
        if (PL::is_angle(a)) x[a] = mod_pi(x[a]);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) gx[NX * s + i] = x[i];
  }
}

template <class PL>
__global__ void __launch_bounds__(128) plant_post_kernel(const GpCfg cfg, const GpIO io, const double* __restrict__ u) {
  constexpr int NX = PL::NX;
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= io.B || io.pred_out == nullptr) return;
  const int W = cfg.W;
  const typename PL::Const P = make_plant_const_opaque<PL>(io.dyn + inst * PL::NP);
  const double* ps = io.states + inst * NX;
  double x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = ps[i];
  const double* pu = u + inst * W;
  double* po = io.pred_out + (size_t)inst * W * NX;
  #pragma unroll 1
  for (int q = 0; q < W; ++q) {
    rk4_step<PL>(P, x, pu[q], cfg.dt);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      if (PL::is_angle(i)) x[i] = mod_pi(x[i]);
      po[NX * q + i] = x[i];
    }
  }
}

#undef GP_GSYNC
#undef GP_PARK

}  // namespace gp
}  // namespace cpmpc
