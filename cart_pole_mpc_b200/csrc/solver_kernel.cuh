// Batched shooting-method MPC solve for sm_100a: one sub-warp "group" of L lanes per cart-pole instance.
//
// What one group does is what one call of pendulum::Optimization::Step does in the reference
// (optimization/optimization.cc:39-97 -> mini_opt nonlinear.cc:75-158), re-organised around the structure of THIS
// problem instead of a generic dense QP:
//
//   * linearisation (rows a1-a3 of SURVEY.md section 8a): every shooting segment is cut into `cps` chunks of
//     consecutive knots; one lane integrates one chunk through RK4 with the Jacobian chain folded in forward mode,
//     starting from the chunk-start state that the last accepted merit rollout left behind, and the chunk blocks are
//     then multiplied together.  The dependent chain of a linearisation is S/cps RK4 steps instead of S;
//   * the equality-constrained QP (reference: dense column-pivoted QR + explicit Q + 37x37 Cholesky, qp.cc:679-729)
//     is solved by a segment-wise Riccati recursion.  Stage = one shooting segment, stage state xi = (node step dx_s,
//     previous control), stage control = the S knots of the segment.  Per stage a symmetric matrix over
//     (controls | xi | rho) is assembled in shared memory (rho = the affine coordinate plus one multiplier coordinate
//     per hard terminal row) and S steps of right-looking Cholesky eliminate the controls: the trailing block IS the
//     next value function, the sub-diagonal panel gives the feedback gains.  The multipliers of the hard terminal
//     rows come from a (<= 4x4) Cholesky on the rho block, one forward sweep yields du and the node steps.
//     Levenberg-Marquardt damping (lambda on every diagonal entry of G, nonlinear.cc:187-189) is just a stage cost.
//     The step is the reference's null-space solution (the QP is strictly convex on the constraint null space);
//     closed-loop propagation keeps it accurate for long, unstable horizons where open-loop condensing is not
//     (tests/prototypes/qp_riccati_proto.py);
//   * the L1-merit polynomial line search, penalty update, LM state machine and exit tests follow
//     nonlinear.cc:296-605 branch for branch, run in lock step by the groups of a warp.
//
// All per-instance state lives in shared memory (about 5.5 KB at W=40,S=10) except the Riccati gains, which go
// through an L2-resident global scratch; HBM is touched only for the inputs, the warm start and the results.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cpmpc.h"
#include "dynamics.cuh"

namespace cpmpc {

constexpr int NR_MAX = 5;              // rho coordinates: affine + up to 4 hard terminal rows
constexpr int XI_DIM = 5;              // stage state: node step (4) + previous control (1)
constexpr int NCOL_MAX = XI_DIM + NR_MAX;

struct KernelCfg {
  int W, S, NS, nseg, N, max_iter, max_ls;
  int condensed;  // 1: lambda == 0 iterations use the condensed QP (short horizons), 0: always Riccati
  int cps;     // chunks per shooting segment in the linearisation
  int nchunk;  // nseg * cps
  int stride;  // doubles of shared memory per instance
  int off_xs, off_u, off_dxs, off_du, off_phi, off_e, off_gam, off_xk, off_phic, off_aug, off_vbuf, off_invd;
  int off_dyn;  // this instance's DynConst (12 doubles): phase functions reload it instead of pinning registers
  int off_xc, off_uc, off_xkc;  // line-search candidates: alias the Gamma block (dead after the QP)
  int kk_stride;                // doubles of global scratch per group: nseg * NCOL_MAX * S Riccati gains
  double dt, rel_tol, first_tol, abs_tol, pen_init, pen_scale, rho, armijo_c1;
  double lam_fail_init, lam_dec_success, lam_dec_restore, lam_max, lam_min;
  double w_u, w_r, wt[4];
  double x_clamp, u_clamp;
};

struct KernelIO {
  long long B;
  const double* states;      // [B,4]
  const double* dyn;         // [B,9]
  const double* set_points;  // [B]
  const double* weights;     // [B,6] or nullptr
  const double* u_guess;     // [W] cold-start control guess (host libm, optimization.cc:63-67)
  double* prev;              // [B,N] warm start, read then overwritten with the new solution
  unsigned char* has_prev;   // [B]
  double* u_out;             // [B,W]
  double* pred_out;          // [B,W,4] or nullptr
  cpmpc_stats* stats;        // [B] or nullptr
  double* iter_log;          // [B,max_iter,27] or nullptr
  double* kk_scratch;        // [n_groups, kk_stride]
  unsigned long long* queue; // [2] work queue: next instance, warps finished (self-resetting)
  // closed loop (n_sim_steps > 0): the plant state is integrated forward with u[0] after every solve
  long long n_steps;         // number of Step()+Simulator::Step() rounds (1 and no plant update when sim == 0)
  int sim;                   // 1 = closed loop
  double* states_io;         // [B,4] in/out plant state when sim == 1
  double* controls_out;      // [B,n_steps] or nullptr
  long long* stats_sum;      // [4] or nullptr
  // debug view of the first SQP iteration (cpmpc_debug_first_qp): when dbg != nullptr every instance linearises once,
  // solves the QP with damping dbg_lambda, writes its blocks to dbg[inst * dbg_stride ...] and stops
  double* dbg;               // [B, dbg_stride]: Phi [nseg*16] | Gam [4*W] | e [4*nseg] | dxs [4*NS] | du [W] | 8 scalars
  long long dbg_stride;
  double dbg_lambda;
  int dbg_qp;                // 0: the handle's choice, 1: Riccati recursion, 2: condensed form
  // [grid warps] SM cycles every warp spent between its first and its last instance (clock64), or nullptr: names
  // the straggler warp of a launch (bench.py per_rank)
  unsigned long long* warp_cycles;
};

// mini_opt enums (dependencies/mini_opt/include/mini_opt/structs.hpp:159-248), by declaration order.
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

// Per-instance weights resolved from the batch defaults or the per-instance override.
struct Weights {
  double w_u, w_r;  // 0 when the corresponding cost family is absent (optimization.cc:270,296)
  double wt[4];     // terminal weights; negative = hard equality (optimization.cc:236-267)
};

// Everything a group needs to address its instance.
struct Ctx {
  const KernelCfg* cfg;
  double* sm;  // this instance's shared memory
  double* kk;  // this group's global scratch for the Riccati gains
  int gl;      // lane within the group
  unsigned gmask;
};

#define CPMPC_GSYNC() __syncwarp(c.gmask)
// Compiler-level fence around the hot rollout loops: everything whose address has escaped (the cold solver scalars
// below) is written back to local memory before and re-read after, so it does not occupy registers inside.
#define CPMPC_PARK() asm volatile("" ::: "memory")

// ---------------------------------------------------------------------------------------------------------------------
// Cost terms over the controls (optimization.cc:270-301): 0.5 * sum of squared residuals for the columns owned by
// this lane's segments.  `uu` is the control array in shared memory.
// ---------------------------------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ double control_cost_partial(const Ctx& c, const Weights& w, const double* uu,
                                                       const double u_prev) {
  const KernelCfg& k = *c.cfg;
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

// Terminal cost / terminal equality / initial-state equality residuals (optimization.cc:228-267), evaluated by
// one lane from the shooting states `xx` in shared memory.
__device__ __forceinline__ void boundary_terms(const KernelCfg& k, const Weights& w, const double* xx,
                                               const double x_cur[4], const double set_point, double& f_term,
                                               double& eq_bnd) {
  constexpr double half_pi = 1.5707963267948966;
  const double target[4] = {set_point, half_pi, 0.0, 0.0};
  f_term = 0.0;
  eq_bnd = 0.0;
  // initial state rows (weight 1)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double d = xx[i] - x_cur[i];
    if (i == 1) d = mod_pi(d);
    eq_bnd += fabs(d);
  }
  const double* xN = xx + 4 * (k.NS - 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double d = xN[i] - target[i];
    if (i == 1) d = mod_pi(d);
    if (w.wt[i] >= 0.0) {
      const double r = d * w.wt[i];
      f_term += 0.5 * (r * r);
    } else {
      eq_bnd += fabs(d);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// One merit evaluation == EvaluateNonlinearErrors (nonlinear.cc:279-293) on the shooting states `xx` / controls `uu`.
// Each lane integrates its own segments without Jacobians; the defects against the next shooting state are L1-summed.
// The states at the chunk boundaries inside a segment are left in `xk_out` for the next linearisation.
// ---------------------------------------------------------------------------------------------------------------------
// The rollouts of one merit evaluation for this lane's segments, as a real (not inlined) function: inside the solver
// loop some 230 registers are live across the call site and ptxas schedules an inlined copy of the dynamics with 11
// double temporaries, i.e. its four dependent chains strictly one after the other.  Behind a call boundary the body
// gets the whole register file (the caller's live values are saved around the call: a few hundred cycles per
// 10-step rollout).  Returns this lane's share of the L1 norm of the shooting defects.
template <int L>
__device__ __forceinline__ double rollout_defects(const double* __restrict__ dynp, const double* __restrict__ xx,
                                               const double* __restrict__ uu, double* __restrict__ xk_out,
                                               const int gl, const int nseg, const int S, const int cps,
                                               const double dt) {
  DynConst P;
  load_dyn_const(P, dynp);
  double eq = 0.0;
  for (int s = gl; s < nseg; s += L) {
    double x[4] = {xx[4 * s], xx[4 * s + 1], xx[4 * s + 2], xx[4 * s + 3]};
    int cc = 1;
    int nb = (cps > 1) ? (S / cps) : S;  // first step of chunk cc
    for (int i = 0; i < S; ++i) {
      if (i == nb) {
        double* dst = xk_out + 4 * (s * cps + cc);
        dst[0] = x[0];
        dst[1] = x[1];
        dst[2] = x[2];
        dst[3] = x[3];
        ++cc;
        nb = (cc < cps) ? (cc * S) / cps : S;
      }
      rk4_step<false>(P, x, uu[s * S + i], dt);
    }
    x[1] = mod_pi(x[1]);
    const double* xn = xx + 4 * (s + 1);
    eq += fabs(x[0] - xn[0]);
    eq += fabs(mod_pi(x[1] - xn[1]));
    eq += fabs(x[2] - xn[2]);
    eq += fabs(x[3] - xn[3]);
  }
  return eq;
}

template <int L>
__device__ __forceinline__ void evaluate_errors(const Ctx& c, const Weights& w, const double* xx,
                                                const double* uu, double* xk_out, const double x_cur[4],
                                                const double set_point, const double u_prev, double& f_out,
                                                double& eq_out) {
  const KernelCfg& k = *c.cfg;
  double eq = rollout_defects<L>(c.sm + k.off_dyn, xx, uu, xk_out, c.gl, k.nseg, k.S, k.cps, k.dt);
  double f = control_cost_partial<L>(c, w, uu, u_prev);
  if (c.gl == 0) {
    double ft, eb;
    boundary_terms(k, w, xx, x_cur, set_point, ft, eb);
    f += ft;
    eq += eb;
  }
  f_out = group_sum<L>(f, c.gmask);
  eq_out = group_sum<L>(eq, c.gmask);
}

// Column-major packed lower-triangular index of a D x D matrix: column c holds rows c..D-1 contiguously.
__device__ __forceinline__ int cbase(const int D, const int col) { return col * D - ((col * (col - 1)) >> 1) - col; }
// aug[cbase(D, col) + r] is entry (r, col), r >= col.

// y[0..n) -= a * x[0..n) on shared memory.  Loads are issued in batches ahead of the stores so that the LDS latency
// is paid once per batch (x and y never overlap, which the compiler cannot know).
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
  if (r + 2 <= n) {
    const double x0 = x[r], x1 = x[r + 1];
    const double y0 = y[r], y1 = y[r + 1];
    y[r] = fma(-a, x0, y0);
    y[r + 1] = fma(-a, x1, y1);
    r += 2;
  }
  if (r < n) y[r] = fma(-a, x[r], y[r]);
}

// ---------------------------------------------------------------------------------------------------------------------
// Equality-constrained QP of one SQP iteration (reference: QPNullSpaceSolver::Solve, qp.cc:679-729) by a segment-wise
// Riccati recursion; see the header comment.  On entry Phi / Gam / ev hold the segment blocks, du[] the gradient of
// the control costs.  On exit du[] and dxs[] hold the step; returns false when a pivot is not positive (the reduced
// Hessian is not positive definite: QP_INDEFINITE).  `cdx` returns this lane's share of c_u . du.
//
// Stage matrix (column-major packed lower triangle, D = S + 5 + nr):
//        [ H_vv                     ]   v   : the S controls of the segment
//    A = [ H_xv   H_xx              ]   xi  : (dx_s, previous control)
//        [ h_v^T  h_x^T    C        ]   rho : (1, mu_1 .. mu_nh)
// S steps of right-looking LDL^T elimination (no square roots: column j stays unscaled, its pivot reciprocal goes to
// invd[j]) leave the next value function in the trailing (xi, rho) block and Z = L^-1 [H_vx h_v] in rows S.. of the
// first S columns; the gains K = L^-T D^-1 Z follow by back-substitution.
// ---------------------------------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ bool riccati_qp(const Ctx& c, const Weights& w, const double lambda,
                                           const double x_cur[4], const double target[4], double& cdx) {
  const KernelCfg& k = *c.cfg;
  const int W = k.W, S = k.S, NS = k.NS, nseg = k.nseg;
  const double* xs = c.sm + k.off_xs;
  double* dxs = c.sm + k.off_dxs;
  double* du = c.sm + k.off_du;
  const double* Phi = c.sm + k.off_phi;
  const double* ev = c.sm + k.off_e;
  const double* Gam = c.sm + k.off_gam;
  double* aug = c.sm + k.off_aug;
  double* vb = c.sm + k.off_vbuf;  // [25] P (full, symmetric) then [5][NR_MAX] q
  double* qb = vb + 25;
  double* invd = c.sm + k.off_invd;
  const double wu2 = w.w_u * w.w_u, wr2 = w.w_r * w.w_r;

  // hard terminal rows (negative weight, optimization.cc:236-267) and their rho coordinates
  int nh = 0;
  int hcomp[4] = {0, 0, 0, 0};  // state component of hard row j
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (w.wt[i] < 0.0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == nh) hcomp[j] = i;
      ++nh;
    }
  }
  const int nr = 1 + nh;
  const int n2 = XI_DIM + nr;
  const int D = S + n2;
  double dT[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double d = xs[4 * (NS - 1) + i] - target[i];
    if (i == 1) d = mod_pi(d);
    dT[i] = d;
  }

  // terminal value function V_N over (xi | rho), stored where a stage elimination leaves it: rows/cols S..D-1
  for (int b = c.gl; b < n2; b += L) {
    double* colp = aug + cbase(D, S + b) + S;  // entry (S + a, S + b) at colp[a]
    for (int a = b; a < n2; ++a) {
      double v = 0.0;
      if (a < 4) {
        if (a == b) {
          const double wa = (a == 0) ? w.wt[0] : (a == 1) ? w.wt[1] : (a == 2) ? w.wt[2] : w.wt[3];
          v = lambda + ((wa >= 0.0) ? wa * wa : 0.0);
        }
      } else if (a == XI_DIM && b < 4) {
        const double wb = (b == 0) ? w.wt[0] : (b == 1) ? w.wt[1] : (b == 2) ? w.wt[2] : w.wt[3];
        const double db = (b == 0) ? dT[0] : (b == 1) ? dT[1] : (b == 2) ? dT[2] : dT[3];
        v = (wb >= 0.0) ? (wb * wb) * db : 0.0;
      } else if (a > XI_DIM && b < 4) {
        const int j = a - XI_DIM - 1;
        const int hj = (j == 0) ? hcomp[0] : (j == 1) ? hcomp[1] : (j == 2) ? hcomp[2] : hcomp[3];
        v = (b == hj) ? 1.0 : 0.0;
      }
      colp[a] = v;
    }
  }
  CPMPC_GSYNC();

  // trailing-block accessor: entry (S + a, S + b), a >= b
#define CPMPC_TR(a, b) aug[cbase(D, S + (b)) + S + (a)]
  bool ok = true;
  for (int s = nseg - 1; s >= 0; --s) {
    const double* Ph = Phi + 16 * s;
    const double f0 = ev[4 * s], f1 = ev[4 * s + 1], f2 = ev[4 * s + 2], f3 = ev[4 * s + 3];
    // (i) snapshot P (full 5x5) and q = p + (P f) e_0; fold f . p_r into the rho-rho column 0 (in place)
    for (int t = c.gl; t < 25 + XI_DIM * nr + nh; t += L) {
      if (t < 25) {
        const int a = t / 5, b = t - 5 * a;
        vb[t] = (a >= b) ? CPMPC_TR(a, b) : CPMPC_TR(b, a);
      } else if (t < 25 + XI_DIM * nr) {
        const int t2 = t - 25;
        const int r = t2 / XI_DIM, a = t2 - XI_DIM * r;
        double v = CPMPC_TR(XI_DIM + r, a);
        if (r == 0) {
          const double p0 = CPMPC_TR(a, 0);
          const double p1 = (a >= 1) ? CPMPC_TR(a, 1) : CPMPC_TR(1, a);
          const double p2 = (a >= 2) ? CPMPC_TR(a, 2) : CPMPC_TR(2, a);
          const double p3 = (a >= 3) ? CPMPC_TR(a, 3) : CPMPC_TR(3, a);
          v += p0 * f0 + p1 * f1 + p2 * f2 + p3 * f3;
        }
        qb[a * NR_MAX + r] = v;
      } else {
        const int r = t - 25 - XI_DIM * nr + 1;  // 1..nh
        CPMPC_TR(XI_DIM + r, XI_DIM) += CPMPC_TR(XI_DIM + r, 0) * f0 + CPMPC_TR(XI_DIM + r, 1) * f1 +
                                        CPMPC_TR(XI_DIM + r, 2) * f2 + CPMPC_TR(XI_DIM + r, 3) * f3;
      }
    }
    CPMPC_GSYNC();

    // (ii) assemble the stage matrix: every lane writes whole columns (contiguous)
    for (int t = c.gl; t < S + XI_DIM; t += L) {
      double* colp = aug + cbase(D, t);  // entry (r, t) at colp[r]
      if (t < S) {
        const int j = t, col = s * S + j;
        const double g0 = Gam[col], g1 = Gam[W + col], g2 = Gam[2 * W + col], g3 = Gam[3 * W + col];
        const double g4 = (j == S - 1) ? 1.0 : 0.0;
        double PG[XI_DIM];
#pragma unroll
        for (int a = 0; a < XI_DIM; ++a)
          PG[a] = vb[5 * a] * g0 + vb[5 * a + 1] * g1 + vb[5 * a + 2] * g2 + vb[5 * a + 3] * g3 + vb[5 * a + 4] * g4;
        for (int i = j; i < S; ++i) {
          const int ci = s * S + i;
          double v = Gam[ci] * PG[0] + Gam[W + ci] * PG[1] + Gam[2 * W + ci] * PG[2] + Gam[3 * W + ci] * PG[3];
          if (i == S - 1) v += PG[4];
          if (i == j) v += wu2 + lambda + wr2 * ((col + 1 < W) ? 2.0 : 1.0);
          if (i == j + 1) v -= wr2;
          colp[i] = v;
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          colp[S + cc] = Ph[cc] * PG[0] + Ph[4 + cc] * PG[1] + Ph[8 + cc] * PG[2] + Ph[12 + cc] * PG[3];
        colp[S + 4] = (j == 0 && s > 0) ? -wr2 : 0.0;
        for (int r = 0; r < nr; ++r) {
          double v = g0 * qb[r] + g1 * qb[NR_MAX + r] + g2 * qb[2 * NR_MAX + r] + g3 * qb[3 * NR_MAX + r] +
                     g4 * qb[4 * NR_MAX + r];
          if (r == 0) v += du[col];
          colp[S + XI_DIM + r] = v;
        }
      } else {
        const int cc = t - S;  // xi column
        if (cc < 4) {
          const double h0 = Ph[cc], h1 = Ph[4 + cc], h2 = Ph[8 + cc], h3 = Ph[12 + cc];
          double PF[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) PF[a] = vb[5 * a] * h0 + vb[5 * a + 1] * h1 + vb[5 * a + 2] * h2 + vb[5 * a + 3] * h3;
          for (int c2 = cc; c2 < 4; ++c2) {
            double v = Ph[c2] * PF[0] + Ph[4 + c2] * PF[1] + Ph[8 + c2] * PF[2] + Ph[12 + c2] * PF[3];
            if (c2 == cc && s > 0) v += lambda;
            colp[S + c2] = v;
          }
          colp[S + 4] = 0.0;
          for (int r = 0; r < nr; ++r)
            colp[S + XI_DIM + r] = h0 * qb[r] + h1 * qb[NR_MAX + r] + h2 * qb[2 * NR_MAX + r] + h3 * qb[3 * NR_MAX + r];
        } else {
          colp[S + 4] = 0.0;
          for (int r = 0; r < nr; ++r) colp[S + XI_DIM + r] = 0.0;
        }
      }
    }
    CPMPC_GSYNC();

    // Two implementations of the same elimination, bit-identical in their results (profiles/ab_compare.py).  Wide
    // lane groups (L >= 8: long horizons, one column per lane, several warps per scheduler) are instruction-bound and
    // keep the right-looking form; narrow groups (L <= 4: the default horizon, Levenberg-Marquardt iterations only,
    // one warp per scheduler) are latency-bound and use the left-looking form below, 5-8 % faster end to end on
    // LM-heavy batches.
    if constexpr (L >= 8) {
    // (iii) eliminate the S controls: right-looking LDL^T steps over the whole D x D matrix, one trailing column per
    // lane.  The lane that owns column j + 1 handles it first and publishes the reciprocal of the next pivot.
    if (c.gl == 0) {
      const double piv = aug[0];
      invd[0] = (piv > 0.0) ? fm::rcp(piv) : -1.0;
    }
    CPMPC_GSYNC();
    for (int j = 0; j < S; ++j) {
      const double rd = invd[j];
      if (!(rd > 0.0)) {
        ok = false;
        break;
      }
      const double* colj = aug + cbase(D, j);  // (r, j) at colj[r]
      for (int cc = j + 1 + c.gl; cc < D; cc += L) {
        double* colc = aug + cbase(D, cc);
        const double l = colj[cc] * rd;
        if (cc == j + 1 && cc < S) {
          // next pivot first, so that its reciprocal is ready when the others reach the barrier
          const double piv = fma(-l, colj[cc], colc[cc]);
          colc[cc] = piv;
          invd[cc] = (piv > 0.0) ? fm::rcp(piv) : -1.0;
          axpy_neg(colc + cc + 1, colj + cc + 1, l, D - cc - 1);
        } else {
          axpy_neg(colc + cc, colj + cc, l, D - cc);
        }
      }
      CPMPC_GSYNC();
    }
    } else {
    // (iii) eliminate the S controls: LEFT-looking LDL^T.  Every entry (r, c) of the factor / Schur complement is
    //   a(r, c) - sum_{i < min(c, S)} (a(c, i) / d_i) a(r, i),   accumulated in increasing i with one FMA per term
    // - exactly the updates, in exactly the order, that a right-looking elimination applies to that entry (so the
    // result is bit-identical to it) - but the running value stays in a register instead of making a shared-memory
    // round trip per elimination step, and the loads of a chain do not depend on each other.  Phase A finalises the
    // control columns one after the other (rows split across the lanes, one barrier per column, the pivot reciprocal
    // published with it); phase B then updates the trailing (xi, rho) block, one column per lane, without barriers.
    {
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
        CPMPC_GSYNC();
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
      CPMPC_GSYNC();
    }
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
        double* dsta = c.kk + ((size_t)s * NCOL_MAX + q0) * S;
        double* dstb = c.kk + ((size_t)s * NCOL_MAX + (two ? q1 : q0)) * S;
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
    CPMPC_GSYNC();
  }
  ok = __all_sync(c.gmask, ok);
  if (!ok) return false;

  // multipliers of the hard terminal rows: x_N[h_j](mu) + d_j = 0 with x_N[h_j] = p_j . xi_0 + C_j0 + sum_k C_jk mu_k
  double m0[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double d = xs[i] - x_cur[i];
    if (i == 1) d = mod_pi(d);
    m0[i] = -d;
  }
  double coef[NR_MAX] = {1.0, 0.0, 0.0, 0.0, 0.0};
  if (nh > 0) {
    double Mh[4][4], rhs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      rhs[j] = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) Mh[j][q] = (j == q) ? 1.0 : 0.0;
      if (j < nh) {
        const int rj = XI_DIM + 1 + j;
        const int hj = hcomp[j];
        const double dj = (hj == 0) ? dT[0] : (hj == 1) ? dT[1] : (hj == 2) ? dT[2] : dT[3];
        rhs[j] = CPMPC_TR(rj, 0) * m0[0] + CPMPC_TR(rj, 1) * m0[1] + CPMPC_TR(rj, 2) * m0[2] +
                 CPMPC_TR(rj, 3) * m0[3] + CPMPC_TR(rj, XI_DIM) + dj;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q <= j) Mh[j][q] = -CPMPC_TR(rj, XI_DIM + 1 + q);
      }
    }
    // Cholesky of the (nh x nh) lower triangle, rows >= nh are identity
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double d = Mh[j][j];
#pragma unroll
      for (int q = 0; q < j; ++q) d -= Mh[j][q] * Mh[j][q];
      if (!(d > 0.0)) ok = false;
      const double inv = 1.0 / sqrt(d);
      Mh[j][j] = inv;  // keep the reciprocal
#pragma unroll
      for (int i = j + 1; i < 4; ++i) {
        double sacc = Mh[i][j];
#pragma unroll
        for (int q = 0; q < j; ++q) sacc -= Mh[i][q] * Mh[j][q];
        Mh[i][j] = sacc * inv;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double sacc = rhs[i];
#pragma unroll
      for (int q = 0; q < i; ++q) sacc -= Mh[i][q] * rhs[q];
      rhs[i] = sacc * Mh[i][i];
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
      double sacc = rhs[i];
#pragma unroll
      for (int q = i + 1; q < 4; ++q) sacc -= Mh[q][i] * rhs[q];
      rhs[i] = sacc * Mh[i][i];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) coef[1 + j] = (j < nh) ? rhs[j] : 0.0;
    if (!ok) return false;
  }

  // forward sweep: v = -(K xi + k . coef), xi+ = Phi xi + Gam v + e
  double xi[XI_DIM] = {m0[0], m0[1], m0[2], m0[3], 0.0};
  if (c.gl == 0) {
    dxs[0] = m0[0];
    dxs[1] = m0[1];
    dxs[2] = m0[2];
    dxs[3] = m0[3];
  }
  cdx = 0.0;
  for (int s = 0; s < nseg; ++s) {
    const double* kks = c.kk + (size_t)s * NCOL_MAX * S;
    for (int j = c.gl; j < S; j += L) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < XI_DIM; ++q) acc += __ldcg(kks + q * S + j) * xi[q];
#pragma unroll
      for (int r = 0; r < NR_MAX; ++r)
        if (r < nr) acc += __ldcg(kks + (XI_DIM + r) * S + j) * coef[r];
      const int col = s * S + j;
      const double v = -acc;
      cdx += du[col] * v;  // du[] still holds c_u
      du[col] = v;
    }
    CPMPC_GSYNC();
    const double* Ph = Phi + 16 * s;
    for (int a = c.gl; a < 4; a += L) {
      double n0 = Ph[4 * a] * xi[0] + Ph[4 * a + 1] * xi[1] + ev[4 * s + a];
      double n1 = Ph[4 * a + 2] * xi[2] + Ph[4 * a + 3] * xi[3];
      const double* ga = Gam + a * W + s * S;
      const double* vv = du + s * S;
      int j = 0;
      for (; j + 1 < S; j += 2) {
        n0 += ga[j] * vv[j];
        n1 += ga[j + 1] * vv[j + 1];
      }
      if (j < S) n0 += ga[j] * vv[j];
      dxs[4 * (s + 1) + a] = n0 + n1;
    }
    CPMPC_GSYNC();
#pragma unroll
    for (int a = 0; a < 4; ++a) xi[a] = dxs[4 * (s + 1) + a];
    xi[4] = du[s * S + S - 1];
  }
  return true;
#undef CPMPC_TR
}

// ---------------------------------------------------------------------------------------------------------------------
// The same QP in condensed form, for lambda == 0 on short horizons (cfg.condensed): with the shooting states
// eliminated forward through the linearised dynamics the Hessian over the W controls is T + U^T D U, T the
// tridiagonal control / control-rate Hessian and U the 4 rows of the terminal-state sensitivities (soft rows carry
// weight d = w^2, hard rows are constraints).  T = L D_T L^T is factored once per solve (O(W)); with Z = L^-1 U^T
// the multipliers come from the 4x4 system (Z^T D_T^-1 Z + diag(1/d)) sigma = rhs, and
// du = -L^-T D_T^-1 (z_g + Z sigma).  About 3 kflop instead of 15 kflop for the Riccati recursion, but the rows of U
// are open-loop products of the segment transition matrices: accurate while the horizon is short (W dt of a fraction
// of a second), useless on long unstable horizons (tests/prototypes/qp_riccati_proto.py) - the host picks.
// Scratch (aliases the Riccati stage matrix): Z [4][W], tl [W], tinv [W].
// ---------------------------------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ bool condensed_qp(const Ctx& c, const Weights& w, bool& tri_valid, const double x_cur[4],
                                             const double target[4], double& cdx) {
  const KernelCfg& k = *c.cfg;
  const int W = k.W, S = k.S, NS = k.NS, nseg = k.nseg;
  const double* xs = c.sm + k.off_xs;
  double* dxs = c.sm + k.off_dxs;
  double* du = c.sm + k.off_du;
  const double* Phi = c.sm + k.off_phi;
  const double* ev = c.sm + k.off_e;
  const double* Gam = c.sm + k.off_gam;
  double* Z = c.sm + k.off_aug;  // [4][W]
  double* tl = Z + 4 * W;        // sub-diagonal multipliers of the unit bidiagonal factor
  double* tinv = tl + W;         // reciprocal pivots
  const double wu2 = w.w_u * w.w_u, wr2 = w.w_r * w.w_r;
  bool ok = true;

  // (1) factor T once per solve (lane L-1, while the other lanes build the rows of U)
  if (!tri_valid && c.gl == L - 1) {
    double dprev_inv = 0.0;
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
  for (int col = c.gl; col < W; col += L) {
    const int s = col / S;
    double v0 = Gam[col], v1 = Gam[W + col], v2 = Gam[2 * W + col], v3 = Gam[3 * W + col];
    for (int t = s + 1; t < nseg; ++t) {
      const double* Ph = Phi + 16 * t;
      const double n0 = Ph[0] * v0 + Ph[1] * v1 + Ph[2] * v2 + Ph[3] * v3;
      const double n1 = Ph[4] * v0 + Ph[5] * v1 + Ph[6] * v2 + Ph[7] * v3;
      const double n2 = Ph[8] * v0 + Ph[9] * v1 + Ph[10] * v2 + Ph[11] * v3;
      const double n3 = Ph[12] * v0 + Ph[13] * v1 + Ph[14] * v2 + Ph[15] * v3;
      v0 = n0;
      v1 = n1;
      v2 = n2;
      v3 = n3;
    }
    Z[col] = v0;
    Z[W + col] = v1;
    Z[2 * W + col] = v2;
    Z[3 * W + col] = v3;
  }
  // node steps for du = 0: m_0 = -(x_0 - x_cur), m_{s+1} = Phi_s m_s + e_s (every lane, registers)
  double m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double d = xs[i] - x_cur[i];
    if (i == 1) d = mod_pi(d);
    m[i] = -d;
  }
  const double m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3];
  for (int s = 0; s < nseg; ++s) {
    const double* Ph = Phi + 16 * s;
    double n[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
      n[r] = Ph[4 * r] * m[0] + Ph[4 * r + 1] * m[1] + Ph[4 * r + 2] * m[2] + Ph[4 * r + 3] * m[3] + ev[4 * s + r];
#pragma unroll
    for (int r = 0; r < 4; ++r) m[r] = n[r];
  }
  CPMPC_GSYNC();
  ok = __all_sync(c.gmask, ok);
  if (!ok) return false;

  // (3) forward substitution with the unit bidiagonal factor: rows 0..3 of Z and the gradient (in du[]), one
  // right-hand side per lane, the running value carried in a register
  for (int r = c.gl; r < 5; r += L) {
    double* row = (r < 4) ? (Z + r * W) : du;
    double prev = row[0];
    // four elements per trip with the loads issued ahead of the dependent FMAs and the stores behind them: the
    // compiler cannot move a load of row[j + 1] above the store of row[j] on its own (same array), and without that
    // every element pays a shared-memory round trip on top of the FMA.  Same operations in the same order.
    int j = 1;
    for (; j + 4 <= W; j += 4) {
      const double t0 = tl[j], t1 = tl[j + 1], t2 = tl[j + 2], t3 = tl[j + 3];
      const double r0 = row[j], r1 = row[j + 1], r2 = row[j + 2], r3 = row[j + 3];
      const double p0 = fma(-t0, prev, r0);
      const double p1 = fma(-t1, p0, r1);
      const double p2 = fma(-t2, p1, r2);
      prev = fma(-t3, p2, r3);
      row[j] = p0;
      row[j + 1] = p1;
      row[j + 2] = p2;
      row[j + 3] = prev;
    }
    for (; j < W; ++j) {
      prev = fma(-tl[j], prev, row[j]);
      row[j] = prev;
    }
  }
  CPMPC_GSYNC();

  // (4) Gram entries G[i][j] = sum_q Z_i[q] Z_j[q] / d_q for i <= j <= 4 (j == 4: the gradient row): 14 dot products
  // spread over the lanes, results through shuffles-free shared scratch (the first 16 slots of tl's tail are free?)
  // -> kept in registers: every lane needs all 14, so each lane computes its share and the group exchanges them.
  double gm[14];
#pragma unroll
  for (int e = 0; e < 14; ++e) gm[e] = 0.0;
  {
    // pair list: (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3)(0,4)(1,4)(2,4)(3,4)
    // every lane accumulates over its own columns q = gl, gl + L, ...; a group sum finishes each entry
    for (int q = c.gl; q < W; q += L) {
      const double ti = tinv[q];
      const double z0 = Z[q], z1 = Z[W + q], z2 = Z[2 * W + q], z3 = Z[3 * W + q], zg = du[q];
      const double s0 = z0 * ti, s1 = z1 * ti, s2 = z2 * ti, s3 = z3 * ti;
      gm[0] = fma(s0, z0, gm[0]);
      gm[1] = fma(s0, z1, gm[1]);
      gm[2] = fma(s0, z2, gm[2]);
      gm[3] = fma(s0, z3, gm[3]);
      gm[4] = fma(s1, z1, gm[4]);
      gm[5] = fma(s1, z2, gm[5]);
      gm[6] = fma(s1, z3, gm[6]);
      gm[7] = fma(s2, z2, gm[7]);
      gm[8] = fma(s2, z3, gm[8]);
      gm[9] = fma(s3, z3, gm[9]);
      gm[10] = fma(s0, zg, gm[10]);
      gm[11] = fma(s1, zg, gm[11]);
      gm[12] = fma(s2, zg, gm[12]);
      gm[13] = fma(s3, zg, gm[13]);
    }
#pragma unroll
    for (int e = 0; e < 14; ++e) gm[e] = group_sum<L>(gm[e], c.gmask);
  }

  // (5) 4x4 system for the row multipliers, every lane redundantly in registers
  double A4[4][4] = {{gm[0], gm[1], gm[2], gm[3]}, {gm[1], gm[4], gm[5], gm[6]}, {gm[2], gm[5], gm[7], gm[8]},
                     {gm[3], gm[6], gm[8], gm[9]}};
  double sig[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double d = xs[4 * (NS - 1) + i] - target[i];
    if (i == 1) d = mod_pi(d);
    double invd, off;
    if (w.wt[i] < 0.0) {
      invd = 0.0;  // hard row
      off = m[i] + d;
    } else {
      const double dd = w.wt[i] * w.wt[i];
      if (dd > 0.0) {
        invd = 1.0 / dd;
        off = m[i] + d;  // (w^2 d) / w^2
      } else {
        invd = -1.0;  // zero weight: the row does not exist
        off = 0.0;
      }
    }
    if (invd < 0.0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        A4[i][j] = 0.0;
        A4[j][i] = 0.0;
      }
      A4[i][i] = 1.0;
      sig[i] = 0.0;
    } else {
      A4[i][i] += invd;
      sig[i] = off - gm[10 + i];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double d = A4[j][j];
#pragma unroll
    for (int q = 0; q < j; ++q) d -= A4[j][q] * A4[j][q];
    if (!(d > 0.0)) ok = false;
    const double inv = 1.0 / sqrt(d);
    A4[j][j] = inv;
#pragma unroll
    for (int i = j + 1; i < 4; ++i) {
      double sacc = A4[i][j];
#pragma unroll
      for (int q = 0; q < j; ++q) sacc -= A4[i][q] * A4[j][q];
      A4[i][j] = sacc * inv;
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double sacc = sig[i];
#pragma unroll
    for (int q = 0; q < i; ++q) sacc -= A4[i][q] * sig[q];
    sig[i] = sacc * A4[i][i];
  }
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    double sacc = sig[i];
#pragma unroll
    for (int q = i + 1; q < 4; ++q) sacc -= A4[q][i] * sig[q];
    sig[i] = sacc * A4[i][i];
  }

  // (6) du = -L^-T D^-1 (z_g + Z sigma); c_u . du = -(L^-1 g)^T D^-1 (z_g + Z sigma) needs no second pass over g
  double part = 0.0;
  for (int q = c.gl; q < W; q += L) {
    const double zg = du[q];
    const double wq = (zg + sig[0] * Z[q] + sig[1] * Z[W + q] + sig[2] * Z[2 * W + q] + sig[3] * Z[3 * W + q]) * tinv[q];
    part = fma(zg, wq, part);
    du[q] = wq;
  }
  cdx = -part;  // this lane's share of c_u . du
  CPMPC_GSYNC();
  if (c.gl == 0) {
    double nxt = du[W - 1];
    du[W - 1] = -nxt;
    int j = W - 2;
    for (; j >= 3; j -= 4) {  // same batching as the forward substitution, descending
      const double t0 = tl[j + 1], t1 = tl[j], t2 = tl[j - 1], t3 = tl[j - 2];
      const double d0 = du[j], d1 = du[j - 1], d2 = du[j - 2], d3 = du[j - 3];
      const double n0 = fma(-t0, nxt, d0);
      const double n1 = fma(-t1, n0, d1);
      const double n2 = fma(-t2, n1, d2);
      nxt = fma(-t3, n2, d3);
      du[j] = -n0;
      du[j - 1] = -n1;
      du[j - 2] = -n2;
      du[j - 3] = -nxt;
    }
    for (; j >= 0; --j) {
      nxt = fma(-tl[j + 1], nxt, du[j]);
      du[j] = -nxt;
    }
    dxs[0] = m00;
    dxs[1] = m01;
    dxs[2] = m02;
    dxs[3] = m03;
  }
  CPMPC_GSYNC();

  // (7) node steps by forward propagation
  double xi[4] = {m00, m01, m02, m03};
  for (int s = 0; s < nseg; ++s) {
    const double* Ph = Phi + 16 * s;
    for (int a = c.gl; a < 4; a += L) {
      double n0 = Ph[4 * a] * xi[0] + Ph[4 * a + 1] * xi[1] + ev[4 * s + a];
      double n1 = Ph[4 * a + 2] * xi[2] + Ph[4 * a + 3] * xi[3];
      const double* ga = Gam + a * W + s * S;
      const double* vv = du + s * S;
      int j = 0;
      for (; j + 1 < S; j += 2) {
        n0 = fma(ga[j], vv[j], n0);
        n1 = fma(ga[j + 1], vv[j + 1], n1);
      }
      if (j < S) n0 = fma(ga[j], vv[j], n0);
      dxs[4 * (s + 1) + a] = n0 + n1;
    }
    CPMPC_GSYNC();
#pragma unroll
    for (int a = 0; a < 4; ++a) xi[a] = dxs[4 * (s + 1) + a];
  }
  return true;
}

struct SolveResult {
  int termination;
  int n_iter;
  int n_ls;
  int n_failed;
  int last_step_result;
  int opt_state;
  double f, eq, penalty, lambda;
};

// ---------------------------------------------------------------------------------------------------------------------
// The whole Optimization::Step for one instance, executed by the L lanes of a group.  `active` is group-uniform;
// inactive groups only take part in the warp-wide votes that keep the groups of a warp in lock step.
// On entry the shooting states, the chunk-start states and the control guess are in shared memory (xs, xk, u); on
// exit xs/u hold the solution.
// ---------------------------------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ void solve_group(const Ctx& c, const bool active_in, const Weights& w,
                                            const double x_cur[4], const double& set_point, const double& u_prev,
                                            double* iter_log, SolveResult& res, double* dbg = nullptr,
                                            const double dbg_lambda = 0.0, const int dbg_qp = 0) {
  const KernelCfg& k = *c.cfg;
  double* xs = c.sm + k.off_xs;
  double* u = c.sm + k.off_u;
  double* xc = c.sm + k.off_xc;
  double* uc = c.sm + k.off_uc;
  double* xk = c.sm + k.off_xk;
  double* xkc = c.sm + k.off_xkc;
  double* dxs = c.sm + k.off_dxs;
  double* du = c.sm + k.off_du;
  double* Phi = c.sm + k.off_phi;
  double* PhiC = c.sm + k.off_phic;
  double* ev = c.sm + k.off_e;
  double* Gam = c.sm + k.off_gam;
  constexpr double half_pi = 1.5707963267948966;
  const double target[4] = {set_point, half_pi, 0.0, 0.0};
  const int W = k.W, S = k.S, NS = k.NS, nseg = k.nseg, cps = k.cps;

  // Cold solver scalars: kept in a struct whose address escapes, so that CPMPC_PARK() moves them out of the
  // register file for the duration of the rollout loops (see CPMPC_PARK).
  struct Cold {
    double lambda, penalty, cur_f, cur_eq;
    double f0, eq0, d_f, d_eq, lambda_in;
    double alpha, a_prev, phi_prev, a_prev2, phi_prev2, step_f, step_eq, phi0, dphi0, linf;
    int opt_state, step_result, n_steps;
    bool active, have_cur, tri_valid, ls_done;
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
  active = active_in;
  lambda = (dbg != nullptr) ? dbg_lambda : 0.0;  // params.lambda_initial (nonlinear.hpp:104)
  penalty = k.pen_init;
  opt_state = OS_NOMINAL;
  cur_f = 0.0;
  cur_eq = 0.0;  // Errors at the current variables
  have_cur = false;
  tri_valid = false;  // factorisation of T (condensed QP) is cached across the iterations of one solve
  res.termination = TS_MAX_ITERATIONS;
  res.n_iter = 0;
  res.n_ls = 0;
  res.n_failed = 0;
  res.last_step_result = SR_SUCCESS;

  // gradient of the control costs at the current controls into du[] (the QP routines consume and overwrite it)
  auto fill_gradient = [&]() {
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
    CPMPC_GSYNC();
  };

  for (int iter = 0; iter < k.max_iter; ++iter) {
    if (!__any_sync(0xffffffffu, active)) break;

    double& f0 = cold.f0;
    double& eq0 = cold.eq0;
    double& d_f = cold.d_f;
    double& d_eq = cold.d_eq;
    double& lambda_in = cold.lambda_in;
    f0 = 0.0;
    eq0 = 0.0;
    d_f = 0.0;
    d_eq = 0.0;
    lambda_in = lambda;

    if (active) {
      // ---- linearise: one chunk of consecutive knots per lane, RK4 with the Jacobian chain in forward mode -------
      double eq_part = 0.0;
      CPMPC_PARK();
      DynConst P;
      load_dyn_const(P, c.sm + k.off_dyn);
      for (int id = c.gl; id < k.nchunk; id += L) {
        const int s = id / cps, cc = id - s * cps;
        const int i0 = (cc * S) / cps, i1 = ((cc + 1) * S) / cps;
        const double* x0 = (cc == 0) ? (xs + 4 * s) : (xk + 4 * id);
        double x[4] = {x0[0], x0[1], x0[2], x0[3]};
        double* Ph = (cps == 1) ? (Phi + 16 * s) : (PhiC + 16 * id);
        for (int i = i0; i < i1; ++i) {
          double A[16], Bv[4];
          const int col = s * S + i;
          rk4_step_jac(P, x, u[col], k.dt, A, Bv);
          if (i == i0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) Ph[q] = A[q];
          } else {
            // Phi <- A Phi
            double T[16];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                T[4 * r + j] = A[4 * r] * Ph[j] + A[4 * r + 1] * Ph[4 + j] + A[4 * r + 2] * Ph[8 + j] +
                               A[4 * r + 3] * Ph[12 + j];
#pragma unroll
            for (int q = 0; q < 16; ++q) Ph[q] = T[q];
            // Gamma_j <- A Gamma_j for the earlier knots of this chunk
            for (int j = s * S + i0; j < col; ++j) {
              const double g0 = Gam[j], g1 = Gam[W + j], g2 = Gam[2 * W + j], g3 = Gam[3 * W + j];
#pragma unroll
              for (int r = 0; r < 4; ++r)
                Gam[r * W + j] = A[4 * r] * g0 + A[4 * r + 1] * g1 + A[4 * r + 2] * g2 + A[4 * r + 3] * g3;
            }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) Gam[r * W + col] = Bv[r];
        }
        if (cc == cps - 1) {
          x[1] = mod_pi(x[1]);
          const double* xn = xs + 4 * (s + 1);
          const double e0 = x[0] - xn[0], e1 = mod_pi(x[1] - xn[1]), e2 = x[2] - xn[2], e3 = x[3] - xn[3];
          ev[4 * s] = e0;
          ev[4 * s + 1] = e1;
          ev[4 * s + 2] = e2;
          ev[4 * s + 3] = e3;
          eq_part += fabs(e0) + fabs(e1) + fabs(e2) + fabs(e3);
        }
      }
      CPMPC_PARK();
      if (cps > 1) {
        // stitch the chunks of a segment: Phi_s = Phi_{cps-1} ... Phi_0, Gamma_j <- (Phi_{cps-1} ... Phi_{cc+1}) Gamma_j
        CPMPC_GSYNC();
        for (int base = 0; base < k.nchunk; base += L) {  // uniform trip count: the loop body synchronises
          const int id = base + c.gl;
          const int s = id / cps, cc = id - s * cps;
          const bool mine = id < k.nchunk && cc < cps - 1;
          double Sx[16], Pr[16];
          if (mine) {
            const double* last = PhiC + 16 * (s * cps + cps - 1);
#pragma unroll
            for (int q = 0; q < 16; ++q) Sx[q] = last[q];
            for (int m = cps - 2; m > cc; --m) {
              const double* Pm = PhiC + 16 * (s * cps + m);
              double T[16];
#pragma unroll
              for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  T[4 * r + j] = Sx[4 * r] * Pm[j] + Sx[4 * r + 1] * Pm[4 + j] + Sx[4 * r + 2] * Pm[8 + j] +
                                 Sx[4 * r + 3] * Pm[12 + j];
#pragma unroll
              for (int q = 0; q < 16; ++q) Sx[q] = T[q];
            }
            if (cc == 0) {
              const double* P0 = PhiC + 16 * id;
#pragma unroll
              for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  Pr[4 * r + j] = Sx[4 * r] * P0[j] + Sx[4 * r + 1] * P0[4 + j] + Sx[4 * r + 2] * P0[8 + j] +
                                  Sx[4 * r + 3] * P0[12 + j];
            }
          }
          CPMPC_GSYNC();
          if (mine) {
            double* dst = PhiC + 16 * id;
#pragma unroll
            for (int q = 0; q < 16; ++q) dst[q] = Sx[q];
            if (cc == 0) {
              double* ph = Phi + 16 * s;
#pragma unroll
              for (int q = 0; q < 16; ++q) ph[q] = Pr[q];
            }
          }
        }
        CPMPC_GSYNC();
        for (int col = c.gl; col < W; col += L) {
          const int s = col / S, i = col - s * S;
          const int cc = ((i + 1) * cps - 1) / S;
          if (cc < cps - 1) {
            const double* Sx = PhiC + 16 * (s * cps + cc);
            const double g0 = Gam[col], g1 = Gam[W + col], g2 = Gam[2 * W + col], g3 = Gam[3 * W + col];
#pragma unroll
            for (int r = 0; r < 4; ++r)
              Gam[r * W + col] = Sx[4 * r] * g0 + Sx[4 * r + 1] * g1 + Sx[4 * r + 2] * g2 + Sx[4 * r + 3] * g3;
          }
        }
      }
      double f_part = control_cost_partial<L>(c, w, u, u_prev);
      if (c.gl == 0) {
        double ft, eb;
        boundary_terms(k, w, xs, x_cur, set_point, ft, eb);
        f_part += ft;
        eq_part += eb;
      }
      f0 = group_sum<L>(f_part, c.gmask);
      eq0 = group_sum<L>(eq_part, c.gmask);
      if (have_cur) {
        // Same variables as the accepted line-search candidate: the reference gets bit-identical numbers from both
        // code paths; keep the line-search values so the merit comparison is self-consistent.
        f0 = cur_f;
        eq0 = cur_eq;
      } else {
        cur_f = f0;
        cur_eq = eq0;
        have_cur = true;
      }

      // gradient of the control costs, c_u = J^T r (optimization.cc:270-301), parked in du[]
      fill_gradient();

      if (!isfinite(f0) || !isfinite(eq0)) {
        // The reference throws here (nonlinear.cc:356-357); the batch reports it per instance instead.
        res.termination = TS_REJECTED;
        active = false;
      }
    }

    // ---- QP step (qp.cc:679-729) ---------------------------------------------------------------------------------
    double cdx = 0.0;
    bool qp_ok = false;
    bool need_riccati = false;
    if (active) {
      const bool try_condensed = (dbg != nullptr && dbg_qp != 0) ? (dbg_qp == 2) : (k.condensed && !(lambda > 0.0));
      if (try_condensed) {
        qp_ok = condensed_qp<L>(c, w, tri_valid, x_cur, target, cdx);
        // A failed 4x4 multiplier system is a property of the condensed formulation (open-loop sensitivity rows nearly
        // dependent), not necessarily of the QP: restore the gradient and let the Riccati recursion decide.
        if (!qp_ok) fill_gradient();
      }
      need_riccati = !qp_ok && !(dbg != nullptr && dbg_qp == 2);
      if (need_riccati) tri_valid = false;  // the Riccati stage matrix shares the scratch of the tridiagonal factor
    }
    if constexpr (L < 32 && L <= 4) {
      // Narrow lane groups (short horizons) reach the Riccati recursion only for Levenberg-Marquardt iterations and as
      // the fallback of the condensed form: a handful of instances per batch.  While one group runs it the other
      // groups of the warp could only wait, so the WHOLE WARP works on that one instance instead (its blocks are in its
      // own shared memory; the per-instance scalars are broadcast from its first lane): about 6 x shorter than L
      // lanes on it alone.  This is what bounded the slowest rank of an 8-GPU run.
      unsigned need_mask = __ballot_sync(0xffffffffu, need_riccati);
      const int lane = (int)(threadIdx.x & 31);
      const int g_mine = lane / L;
      while (need_mask != 0u) {
        const int src = __ffs(need_mask) - 1;  // first lane of the first group that needs it
        const int g_src = src / L;
        Weights wb;
        wb.w_u = __shfl_sync(0xffffffffu, w.w_u, src);
        wb.w_r = __shfl_sync(0xffffffffu, w.w_r, src);
#pragma unroll
        for (int i = 0; i < 4; ++i) wb.wt[i] = __shfl_sync(0xffffffffu, w.wt[i], src);
        double xb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xb[i] = __shfl_sync(0xffffffffu, x_cur[i], src);
        const double lam_b = __shfl_sync(0xffffffffu, lambda, src);
        const double sp_b = __shfl_sync(0xffffffffu, set_point, src);
        const double target_b[4] = {sp_b, half_pi, 0.0, 0.0};
        Ctx cw;
        cw.cfg = c.cfg;
        cw.sm = c.sm + (ptrdiff_t)(g_src - g_mine) * k.stride;
        cw.kk = c.kk + (ptrdiff_t)(g_src - g_mine) * k.kk_stride;
        cw.gl = lane;
        cw.gmask = 0xffffffffu;
        double cdx_w = 0.0;
        const bool ok_w = riccati_qp<32>(cw, wb, lam_b, xb, target_b, cdx_w);
        cdx_w = group_sum<32>(cdx_w, 0xffffffffu);
        if (g_mine == g_src) {
          qp_ok = ok_w;
          cdx = (c.gl == 0) ? cdx_w : 0.0;
        }
        need_mask &= ~((((L == 32) ? 0u : (1u << L)) - 1u) << (g_src * L));
        __syncwarp(0xffffffffu);
      }
    } else {
      if (need_riccati) qp_ok = riccati_qp<L>(c, w, lambda, x_cur, target, cdx);
    }
    if (active) {
      if (dbg != nullptr) {
        // debug view: the blocks of this linearisation and the QP step, then stop (the candidates would overwrite Gam)
        CPMPC_GSYNC();
        double* o = dbg;
        for (int q = c.gl; q < 16 * nseg; q += L) o[q] = Phi[q];
        o += 16 * nseg;
        for (int q = c.gl; q < 4 * W; q += L) o[q] = Gam[q];
        o += 4 * W;
        for (int q = c.gl; q < 4 * nseg; q += L) o[q] = ev[q];
        o += 4 * nseg;
        for (int q = c.gl; q < 4 * NS; q += L) o[q] = dxs[q];
        o += 4 * NS;
        for (int q = c.gl; q < W; q += L) o[q] = du[q];
        o += W;
        if (c.gl == 0) {
          o[0] = qp_ok ? 1.0 : 0.0;
          o[1] = f0;
          o[2] = eq0;
          o[3] = lambda;
        }
        active = false;
        qp_ok = true;
      } else if (!qp_ok) {
        // QP_INDEFINITE: Solve() returns before this iteration is logged (nonlinear.cc:103-105)
        res.termination = TS_QP_INDEFINITE;
        active = false;
      } else {
        // quadratic form dx^T G dx and the directional derivatives (nonlinear.cc:457-502)
        double quad = 0.0;
        const double wu2 = w.w_u * w.w_u, wr2 = w.w_r * w.w_r;
        for (int col = c.gl; col < W; col += L) {
          const double a = wu2 + lambda + wr2 * ((col + 1 < W) ? 2.0 : 1.0);
          const double v = du[col];
          quad += a * v * v;
          if (col + 1 < W) quad -= 2.0 * wr2 * v * du[col + 1];
        }
        for (int q = c.gl; q < NS; q += L) {
          const double* m = dxs + 4 * q;
          quad += lambda * (m[0] * m[0] + m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
        }
        if (c.gl == 0) {
          const double* m = dxs + 4 * (NS - 1);
          // terminal rows: cost gradient / Hessian, and the equality rows' A_i.dx
          double deq_term = 0.0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double d = xs[4 * (NS - 1) + i] - target[i];
            if (i == 1) d = mod_pi(d);
            if (w.wt[i] >= 0.0) {
              const double w2 = w.wt[i] * w.wt[i];
              cdx += (w2 * d) * m[i];
              quad += w2 * m[i] * m[i];
            } else {
              deq_term += ((d > 0.0) ? 1.0 : ((d < 0.0) ? -1.0 : 0.0)) * m[i];
            }
          }
          // every other equality row satisfies A_i.dx = -b_i by construction of dx
          double l1_other = 0.0;
          for (int q = 0; q < 4 * nseg; ++q) l1_other += fabs(ev[q]);
#pragma unroll
          for (int i = 0; i < 4; ++i) l1_other += fabs(dxs[i]);
          d_eq = deq_term - l1_other;
        }
        CPMPC_GSYNC();
        d_f = group_sum<L>(cdx, c.gmask);
        quad = group_sum<L>(quad, c.gmask);
        d_eq = group_sum<L>(d_eq, c.gmask);

        // penalty update, inequality (18.36) (nonlinear.cc:112-118,484-502)
        const double l1_eq = fmax(eq0, 2.220446049250313e-16);
        const double new_penalty = (d_f + 0.5 * fmax(0.0, quad)) / ((1.0 - k.rho) * l1_eq);
        if (new_penalty > penalty) penalty = new_penalty * k.pen_scale;
      }
    }

    // ---- line search (nonlinear.cc:351-443), groups of the warp in lock step ---------------------------------------
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
    for (int ls = 0; ls <= k.max_ls; ++ls) {
      if (__all_sync(0xffffffffu, ls_done)) break;
      if (!ls_done && ls > 0) {
        // polynomial step-length selection (nonlinear.cc:418-443,531-605)
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
        // retraction (optimization.cc:309-329): x + dx*alpha, wrap angles, clamp b_x and u
        for (int q = c.gl; q < NS; q += L) {
          const double b = __dadd_rn(xs[4 * q], __dmul_rn(dxs[4 * q], alpha));
          const double t = __dadd_rn(xs[4 * q + 1], __dmul_rn(dxs[4 * q + 1], alpha));
          xc[4 * q] = clampd(b, -k.x_clamp, k.x_clamp);
          xc[4 * q + 1] = mod_pi(t);
          xc[4 * q + 2] = __dadd_rn(xs[4 * q + 2], __dmul_rn(dxs[4 * q + 2], alpha));
          xc[4 * q + 3] = __dadd_rn(xs[4 * q + 3], __dmul_rn(dxs[4 * q + 3], alpha));
        }
        for (int col = c.gl; col < W; col += L)
          uc[col] = clampd(__dadd_rn(u[col], __dmul_rn(du[col], alpha)), -k.u_clamp, k.u_clamp);
        CPMPC_GSYNC();
        CPMPC_PARK();
        evaluate_errors<L>(c, w, xc, uc, xkc, x_cur, set_point, u_prev, step_f, step_eq);
        CPMPC_PARK();
        CPMPC_GSYNC();
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

    // ---- lambda update and exit tests (nonlinear.cc:296-343) -------------------------------------------------------
    if (active) {
      bool exit_now = false;
      if (step_result == SR_SUCCESS) {
        for (int q = c.gl; q < 4 * NS; q += L) xs[q] = xc[q];
        for (int col = c.gl; col < W; col += L) u[col] = uc[col];
        if (cps > 1)
          for (int q = c.gl; q < 4 * k.nchunk; q += L) xk[q] = xkc[q];
        CPMPC_GSYNC();
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
// Kernel: persistent grid-stride over instances; each warp carries 32/L instances at a time.
// ---------------------------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(256, 1) cpmpc_solve_kernel(const KernelCfg cfg, const KernelIO io) {
  const int WARPS_PER_BLOCK = blockDim.x >> 5;  // up to 8: two warps per SM sub-partition
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
  const long long n_groups = (long long)gridDim.x * WARPS_PER_BLOCK * G;
  c.sm = smem + (size_t)(warp * G + g) * cfg.stride;
  c.kk = io.kk_scratch + (size_t)group_id * cfg.kk_stride;
  const int W = cfg.W, S = cfg.S, NS = cfg.NS, N = cfg.N;
  double* xs = c.sm + cfg.off_xs;
  double* u = c.sm + cfg.off_u;

  // Work queue: a warp takes G consecutive instances at a time, so warps that draw cheap instances (few line-search
  // trials) go on instead of idling until the slowest warp of a static round is done.  Warp-wide votes stay
  // convergent because all groups of a warp enter and leave a batch together.
  const long long t_begin = clock64();
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(io.queue, (unsigned long long)G);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned long long)io.B) break;
    const long long inst = (long long)base + g;
    const bool have = inst < io.B;

    // per-instance scalars that are only touched between the phases: parked like solve_group's (CPMPC_PARK)
    struct KCold {
      Weights w;
      double x_cur[4];
      double set_point, u_prev;
      long long solves, sum_iter, sum_ls, sum_bad;
      SolveResult res;
    } kc;
    asm volatile("" : : "l"(&kc) : "memory");
    Weights& w = kc.w;
    double(&x_cur)[4] = kc.x_cur;
    double& set_point = kc.set_point;
    x_cur[0] = x_cur[1] = x_cur[2] = x_cur[3] = 0.0;
    set_point = 0.0;
    if (have) {
      if (c.gl == 0) store_dyn_const(c.sm + cfg.off_dyn, make_dyn_const(io.dyn + inst * CPMPC_DYN_STRIDE));
      if (io.weights != nullptr) {
        const double* pw = io.weights + inst * CPMPC_WEIGHTS_STRIDE;
        w.w_u = pw[0];
        w.w_r = pw[1];
#pragma unroll
        for (int i = 0; i < 4; ++i) w.wt[i] = pw[2 + i];
      } else {
        w.w_u = cfg.w_u;
        w.w_r = cfg.w_r;
#pragma unroll
        for (int i = 0; i < 4; ++i) w.wt[i] = cfg.wt[i];
      }
      if (!(w.w_u > 0.0)) w.w_u = 0.0;
      if (!(w.w_r > 0.0)) w.w_r = 0.0;
      const double* ps = (io.sim ? io.states_io : io.states) + inst * CPMPC_STATE_STRIDE;
#pragma unroll
      for (int i = 0; i < 4; ++i) x_cur[i] = ps[i];
      set_point = io.set_points[inst];
    } else {
      if (c.gl == 0) store_dyn_const(c.sm + cfg.off_dyn, DynConst{1.0, 2.0, 1.0, 1.0, 2.0, 2.0, 0.0, 0.0, 1.0, 0.0, 1.0, 9.81});
      w.w_u = w.w_r = 1.0;
      w.wt[0] = w.wt[1] = w.wt[2] = w.wt[3] = 1.0;
    }

    CPMPC_GSYNC();

    long long& solves = kc.solves;
    long long& sum_iter = kc.sum_iter;
    long long& sum_ls = kc.sum_ls;
    long long& sum_bad = kc.sum_bad;
    solves = sum_iter = sum_ls = sum_bad = 0;
    for (long long step = 0; step < io.n_steps; ++step) {
      // ---- BuildProblem + initial guess (optimization.cc:44-71) ---------------------------------------------------
      double& u_prev = kc.u_prev;
      u_prev = 0.0;
      if (have) {
        double* prev = io.prev + inst * N;
        const bool warm = io.has_prev[inst] != 0;
        if (warm) u_prev = prev[4 * NS];  // last step's applied control (optimization.cc:289-291)
        for (int q = c.gl; q < W; q += L) {
          u[q] = warm ? prev[4 * NS + ((q + 1 < W) ? q + 1 : q)] : io.u_guess[q];
        }
        CPMPC_GSYNC();
        if (c.gl == 0) {
          // FillInitialGuess: all shooting states re-integrated from x(0) (optimization.cc:333-351); the states at
          // the chunk boundaries inside a segment are kept for the first linearisation
          DynConst Pl;
          load_dyn_const(Pl, c.sm + cfg.off_dyn);
          double x[4] = {x_cur[0], x_cur[1], x_cur[2], x_cur[3]};
          double* xk = c.sm + cfg.off_xk;
#pragma unroll
          for (int i = 0; i < 4; ++i) xs[i] = x[i];
          for (int s = 1; s < NS; ++s) {
            int cc = 1;
            int nb = (cfg.cps > 1) ? (S / cfg.cps) : S;
            for (int i = 0; i < S; ++i) {
              if (i == nb) {
                double* dst = xk + 4 * ((s - 1) * cfg.cps + cc);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = x[q];
                ++cc;
                nb = (cc < cfg.cps) ? (cc * S) / cfg.cps : S;
              }
              rk4_step<false>(Pl, x, u[(s - 1) * S + i], cfg.dt);
              x[1] = mod_pi(x[1]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) xs[4 * s + i] = x[i];
          }
        }
        CPMPC_GSYNC();
      }

      SolveResult& res = kc.res;
      double* ilog = (io.iter_log != nullptr && have)
                         ? io.iter_log + (size_t)inst * cfg.max_iter * CPMPC_ITERLOG_STRIDE
                         : nullptr;
      double* dbg = (io.dbg != nullptr && have) ? io.dbg + (size_t)inst * io.dbg_stride : nullptr;
      solve_group<L>(c, have, w, x_cur, set_point, u_prev, ilog, res, dbg, io.dbg_lambda, io.dbg_qp);

      if (have) {
        // ---- outputs (optimization.cc:83-96) ------------------------------------------------------------------
        // A rejected instance keeps its previous warm start: the reference throws inside Solve()
        // (nonlinear.cc:356-357) before previous_solution_ is assigned (optimization.cc:84-85).  The debug view
        // never touches the warm start either.
        const bool keep = res.termination != TS_REJECTED && io.dbg == nullptr;
        double* prev = io.prev + inst * N;
        if (keep)
          for (int q = c.gl; q < 4 * NS; q += L) prev[q] = xs[q];
        for (int q = c.gl; q < W; q += L) {
          if (keep) prev[4 * NS + q] = u[q];
          if (io.u_out != nullptr) io.u_out[inst * W + q] = u[q];
        }
        if (c.gl == 0) {
          if (keep) io.has_prev[inst] = 1;
          if (io.pred_out != nullptr) {
            // ComputePredictedStates (optimization.cc:353-371)
            DynConst Pl;
            load_dyn_const(Pl, c.sm + cfg.off_dyn);
            double x[4] = {x_cur[0], x_cur[1], x_cur[2], x_cur[3]};
            double* po = io.pred_out + (size_t)inst * W * 4;
            for (int q = 0; q < W; ++q) {
              rk4_step<false>(Pl, x, u[q], cfg.dt);
              x[1] = mod_pi(x[1]);
              po[4 * q] = x[0];
              po[4 * q + 1] = x[1];
              po[4 * q + 2] = x[2];
              po[4 * q + 3] = x[3];
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
          solves += 1;
          sum_iter += res.n_iter;
          sum_ls += res.n_ls;
          sum_bad += (res.termination == TS_QP_INDEFINITE || res.termination == TS_MAX_LAMBDA ||
                      res.termination == TS_REJECTED)
                         ? 1
                         : 0;
          if (io.sim) {
            // Simulator::Step(control_dt, u[0]) with 1 ms sub-steps (simulator.cc:11-36)
            const double u0 = u[0];
            if (io.controls_out != nullptr) io.controls_out[inst * io.n_steps + step] = u0;
            double dt = cfg.dt;
            DynConst Pl;
            load_dyn_const(Pl, c.sm + cfg.off_dyn);
            double x[4] = {x_cur[0], x_cur[1], x_cur[2], x_cur[3]};
            if (isfinite(u0)) {
              while (dt > 0.0) {
                rk4_step<false>(Pl, x, u0, fmin(dt, 0.001));
                x[1] = mod_pi(x[1]);
                dt -= 0.001;
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) x_cur[i] = x[i];
          }
        }
        if (io.sim) {
          const int src = (c.gmask == 0xffffffffu) ? 0 : (__ffs(c.gmask) - 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) x_cur[i] = __shfl_sync(c.gmask, x_cur[i], src);
        }
        CPMPC_GSYNC();
      }
    }
    if (have && c.gl == 0) {
      if (io.sim) {
        double* ps = io.states_io + inst * CPMPC_STATE_STRIDE;
#pragma unroll
        for (int i = 0; i < 4; ++i) ps[i] = x_cur[i];
      }
      if (io.stats_sum != nullptr) {
        atomicAdd((unsigned long long*)&io.stats_sum[0], (unsigned long long)solves);
        atomicAdd((unsigned long long*)&io.stats_sum[1], (unsigned long long)sum_iter);
        atomicAdd((unsigned long long*)&io.stats_sum[2], (unsigned long long)sum_ls);
        atomicAdd((unsigned long long*)&io.stats_sum[3], (unsigned long long)sum_bad);
      }
    }
  }
  // the last warp to drain the queue re-arms it for the next launch on this handle (launches are stream-ordered)
  if (lane == 0) {
    if (io.warp_cycles != nullptr) {
      io.warp_cycles[(size_t)blockIdx.x * WARPS_PER_BLOCK + warp] = (unsigned long long)(clock64() - t_begin);
    }
    __threadfence();
    const unsigned long long total = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    if (atomicAdd(io.queue + 1, 1ull) == total - 1) {
      io.queue[0] = 0ull;
      io.queue[1] = 0ull;
    }
  }
}

#undef CPMPC_GSYNC
#undef CPMPC_PARK

// ---------------------------------------------------------------------------------------------------------------------
// Batched plant simulator: Simulator::Step (optimization/simulator.cc:11-36), one thread per plant.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void cpmpc_sim_kernel(const long long B, double* __restrict__ states, const double* __restrict__ dyn,
                                 const double* __restrict__ u, const double dt_total,
                                 const double* __restrict__ f_base, const double* __restrict__ f_mass) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const DynConst P = make_dyn_const(dyn + i * CPMPC_DYN_STRIDE);
  double x[4] = {states[4 * i], states[4 * i + 1], states[4 * i + 2], states[4 * i + 3]};
  const double ui = u[i];
  const double fbx = f_base ? f_base[2 * i] : 0.0;
  const double fmx = f_mass ? f_mass[2 * i] : 0.0;
  const double fmy = f_mass ? f_mass[2 * i + 1] : 0.0;
  double dt = dt_total;
  while (dt > 0.0) {
    rk4_step<true>(P, x, ui, fmin(dt, 0.001), fbx, fmx, fmy);
    x[1] = mod_pi(x[1]);
    dt -= 0.001;
  }
  states[4 * i] = x[0];
  states[4 * i + 1] = x[1];
  states[4 * i + 2] = x[2];
  states[4 * i + 3] = x[3];
}

// ---------------------------------------------------------------------------------------------------------------------
// FP64 FMA throughput probe: 8 independent register-resident DFMA chains per thread.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void cpmpc_dfma_probe_kernel(double* out, const int iters, const double a, const double b) {
  double v0 = threadIdx.x * 1e-3, v1 = v0 + 1.0, v2 = v0 + 2.0, v3 = v0 + 3.0, v4 = v0 + 4.0, v5 = v0 + 5.0,
         v6 = v0 + 6.0, v7 = v0 + 7.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      v0 = fma(v0, a, b);
      v1 = fma(v1, a, b);
      v2 = fma(v2, a, b);
      v3 = fma(v3, a, b);
      v4 = fma(v4, a, b);
      v5 = fma(v5, a, b);
      v6 = fma(v6, a, b);
      v7 = fma(v7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

}  // namespace cpmpc
