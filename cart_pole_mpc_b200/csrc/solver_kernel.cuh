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

#include <type_traits>

#include "../../include/cpmpc.h"
#include "dynamics.cuh"

namespace cpmpc {

// Per-warp phase profile written next to the warp cycle counters (cpmpc_get_warp_cycles): SM cycles attributed to
//   0 instance set-up + initial guess rollout   1 linearisation   2 condensed QP   3 Riccati recursion
//   4 directional derivatives + penalty         5 line search (merit rollouts)   6 update / exit tests
//   7 outputs (warm start, predicted-state rollout, statistics, plant step)
// The default problem shape (OptimizationParams{}: window_length 40, state_spacing 10), for the instantiations that have
// it at compile time (template parameter FX): loop bounds and the row strides of the Gamma / Z blocks become
// immediates instead of registers and integer multiply-adds.
constexpr int FX_W = 40, FX_S = 10, FX_NSEG = 4, FX_NS = 5, FX_N = 4 * FX_NS + FX_W;
constexpr int N_PHASES = 8;
constexpr int WARP_REC = 2 + N_PHASES;  // cycles, first instance of the last batch, then the phases

constexpr int NR_MAX = 5;              // rho coordinates: affine + up to 4 hard terminal rows
constexpr int XI_DIM = 5;              // stage state: node step (4) + previous control (1)
constexpr int NCOL_MAX = XI_DIM + NR_MAX;

struct KernelCfg {
  int W, S, NS, nseg, N, max_iter, max_ls;
  int condensed;  // 1: lambda == 0 iterations use the condensed QP (short horizons), 0: always Riccati
  int z_global;   // 1: the condensed QP's Z rows live in the global scratch instead of shared memory
  int stage_knots;  // knots of Riccati feedback rows staged in shared memory at a time (S, or fewer to save shared memory)
  int cps;     // chunks per shooting segment in the linearisation
  int nchunk;  // nseg * cps
  int stride;  // doubles of shared memory per instance
  int off_xs, off_u, off_dxs, off_du, off_phi, off_e, off_gam, off_xk, off_phic, off_aug;
  int off_dyn;  // this instance's DynConst (DYN_CONST_DOUBLES doubles): phase functions reload it instead of pinning registers
  int off_xc, off_uc, off_xkc;  // line-search candidates: alias the Gamma block (dead after the QP)
  int kk_stride;                // doubles of global scratch per group: W * GAIN_STRIDE Riccati feedback rows
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
  // [grid warps][WARP_REC] SM cycles every warp spent between its first and its last instance (clock64), the first
  // instance of the last batch it drew from the queue, and the cycles per solver phase; or nullptr.  Names the straggler
  // warp of a launch and, when a launch is one wave, the instances it carried and where its time went
  // (bench.py per_rank, profiles/shard_times.py)
  unsigned long long* warp_cycles;
  // Split launch (large batches, cpmpc_capi.cu Launch): the two serial rollouts of a Step - FillInitialGuess before the
  // solve, ComputePredictedStates (and the plant step of the closed loop) after it - run in their own kernels with one
  // THREAD per instance (cpmpc_guess_kernel / cpmpc_post_kernel) instead of on one lane of every group in here.
  const double* guess;       // [B, 4 NS] shooting states of the initial guess, or nullptr: integrate them here
  int split;                 // 1: predicted states, plant step and controls_out are cpmpc_post_kernel's
  long long step_index;      // closed loop through split launches: the control tick this launch belongs to ...
  long long n_steps_total;   // ... of this many (row length of controls_out)
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
template <int L, bool FX = false>
__device__ __forceinline__ double control_cost_partial(const Ctx& c, const Weights& w, const double* uu,
                                                       const double u_prev) {
  const KernelCfg& k = *c.cfg;
  const int W = FX ? FX_W : k.W, S = FX ? FX_S : k.S, nseg = FX ? FX_NSEG : k.nseg;
  // Branch-free: an absent cost family has weight 0 (its residuals add +0), the rate residual behind the last knot is
  // formed against the knot itself (0), and the one against the previous step's control is added on the lane that owns
  // knot 0.  The squares of the two families go to separate accumulators (two short dependent chains instead of one
  // chain through every term and a branch per term).
  double f_r = 0.0, f_u = 0.0;
  for (int s = c.gl; s < nseg; s += L) {
    const int c0 = s * S;
    #pragma unroll 2
    for (int i = 0; i < S; ++i) {
      const int col = c0 + i;
      const double uk = uu[col];
      const double un = uu[(col + 1 < W) ? col + 1 : col];
      const double rr = (uk - un) * w.w_r;
      const double ru = uk * w.w_u;
      f_r = fma(rr, rr, f_r);
      f_u = fma(ru, ru, f_u);
    }
  }
  if (c.gl == 0) {
    const double r0 = (uu[0] - u_prev) * w.w_r;
    f_r = fma(r0, r0, f_r);
  }
  return 0.5 * (f_r + f_u);
}

// Terminal cost / terminal equality / initial-state equality residuals (optimization.cc:228-267), evaluated by
// one lane from the shooting states `xx` in shared memory.
template <bool FX = false>
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
  const double* xN = xx + 4 * ((FX ? FX_NS : k.NS) - 1);
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
// CHUNKED: the mapping splits segments into chunks (cps > 1, the latency mappings) and the chunk-start states are
// recorded on the way; the throughput mapping's copy of the loop carries none of that.
template <int L, bool CHUNKED>
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
    int nb = CHUNKED ? (S / cps) : S;  // first step of chunk cc
    for (int i = 0; i < S; ++i) {
      if (CHUNKED && i == nb) {
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

template <int L, bool LEAN = false, bool FX = false>
__device__ __forceinline__ void evaluate_errors(const Ctx& c, const Weights& w, const double* xx,
                                                const double* uu, double* xk_out, const double x_cur[4],
                                                const double set_point, const double u_prev, double& f_out,
                                                double& eq_out) {
  const KernelCfg& k = *c.cfg;
  double eq = (!LEAN && k.cps > 1)
                  ? rollout_defects<L, true>(c.sm + k.off_dyn, xx, uu, xk_out, c.gl, k.nseg, k.S, k.cps, k.dt)
                  : rollout_defects<L, false>(c.sm + k.off_dyn, xx, uu, xk_out, c.gl, FX ? FX_NSEG : k.nseg,
                                              FX ? FX_S : k.S, 1, k.dt);
  double f = control_cost_partial<L, FX>(c, w, uu, u_prev);
  if (c.gl == 0) {
    double ft, eb;
    boundary_terms<FX>(k, w, xx, x_cur, set_point, ft, eb);
    f += ft;
    eq += eb;
  }
  f_out = group_sum<L>(f, c.gmask);
  eq_out = group_sum<L>(eq, c.gmask);
}

// ---------------------------------------------------------------------------------------------------------------------
// Equality-constrained QP of one SQP iteration (reference: QPNullSpaceSolver::Solve, qp.cc:679-729) by a KNOT-level
// Riccati recursion in end-of-segment coordinates: one SCALAR control per elimination step.
//
// Inside shooting segment s the recursion state is z = (y, p): y = Phi_s dx_s + e_s + sum_{j<i} Gamma_j v_j — the end
// state of the segment as far as the controls already applied determine it — and p = the previous control (the
// control-rate cost couples neighbours).  One knot i of the segment is then the scalar-input stage
//     z+ = A z + b v,   A = diag(1, 1, 1, 1, 0),   b = (Gamma_i, 1),
//     cost = 1/2 (w_u^2 + lambda) v^2 + 1/2 w_r^2 (v - p)^2 + g_i v
// and the value function over (z | rho) — rho = the affine coordinate plus one multiplier coordinate per hard terminal
// row — stays a (5 + nr) x (5 + nr) symmetric matrix (P 5x5, Q 5 x nr, C nr x nr: at most 55 numbers, in registers of
// the group's first lane).  Every elimination is a scalar pivot piv = w_u^2 + lambda + w_r^2 + b^T P b; a non-positive
// pivot is QP_INDEFINITE.  At a shooting node the value function is pulled back through y = Phi dx + e rho_0, and
// Levenberg-Marquardt damping (lambda on every diagonal entry of G, nonlinear.cc:187-189) adds lambda to the node's own
// diagonal.  The feedback rows (5 + nr numbers per knot) go to an L2-resident global scratch; the multipliers of the
// hard terminal rows come from a (<= 4x4) Cholesky on the rho block, one forward sweep yields du and the node steps.
//
// Compared with eliminating the S controls of a segment from a dense (S + 10)^2 stage matrix in shared memory (round
// 1): the same exact QP step (tests/prototypes/qp_riccati_proto.py: both within 1e-13 of an extended-precision KKT
// solve), about 110 FP64 instructions per knot instead of several thousand shared-memory round trips per segment, no
// stage matrix in shared memory at all, and no barrier inside the recursion.
//
// On entry Phi / Gam / ev hold the segment blocks and du[] the gradient of the control costs; on exit du[] and dxs[]
// hold the step.  `cdx` returns c_u . du on the group's first lane (0 elsewhere).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GAIN_STRIDE = XI_DIM + NR_MAX;  // doubles of feedback per knot: 5 for z, up to 5 for rho

template <int L, bool FX = false>
__device__ __forceinline__ bool riccati_qp(const Ctx& c, const Weights& w, const double lambda,
                                           const double x_cur[4], const double target[4], double& cdx) {
  const KernelCfg& k = *c.cfg;
  const int W = FX ? FX_W : k.W, S = FX ? FX_S : k.S, NS = FX ? FX_NS : k.NS, nseg = FX ? FX_NSEG : k.nseg;
  const double* xs = c.sm + k.off_xs;
  double* dxs = c.sm + k.off_dxs;
  double* du = c.sm + k.off_du;
  const double* Phi = c.sm + k.off_phi;
  const double* ev = c.sm + k.off_e;
  const double* Gam = c.sm + k.off_gam;
  double* stage = c.sm + k.off_aug;  // [S][GAIN_STRIDE] feedback rows of one segment, staged for the forward sweep
  // Squares of the weights as explicit, separately rounded products: written `w * w`, whether the compiler contracts one
  // of them into the sum it feeds (the pivot's `wu2 + lambda + wr2`, the terminal `lambda + w2`) depends on what else
  // it sees of the weights in the instantiation at hand, and the instantiations of this kernel must agree bit for bit.
  const double wu2 = __dmul_rn(w.w_u, w.w_u), wr2 = __dmul_rn(w.w_r, w.w_r);
  cdx = 0.0;

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
  double dT[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double d = xs[4 * (NS - 1) + i] - target[i];
    if (i == 1) d = mod_pi(d);
    dT[i] = d;
  }
  double m0[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double d = xs[i] - x_cur[i];
    if (i == 1) d = mod_pi(d);
    m0[i] = -d;
  }

  bool ok = true;
  double coef[NR_MAX] = {1.0, 0.0, 0.0, 0.0, 0.0};
  // The recursion is instantiated for NRC = 4 rho coordinates (the default problem: three hard terminal rows) and for
  // the general NR_MAX = 5: the Q / C updates are a third of its work and an unused column costs as much as a used one.
  auto backward = [&](auto nrc_tag) {
    constexpr int NRC = decltype(nrc_tag)::value;
    // ---- backward recursion, registers of the first lane ------------------------------------------------------------
    // P symmetric 5x5 (upper triangle used: P[i][j], i <= j), Q 5 x NRC, C symmetric (upper triangle)
    double P[XI_DIM][XI_DIM], Q[XI_DIM][NRC], C[NRC][NRC];
#pragma unroll
    for (int i = 0; i < XI_DIM; ++i) {
#pragma unroll
      for (int j = 0; j < XI_DIM; ++j) P[i][j] = 0.0;
#pragma unroll
      for (int r = 0; r < NRC; ++r) Q[i][r] = 0.0;
    }
#pragma unroll
    for (int r = 0; r < NRC; ++r)
#pragma unroll
      for (int q = 0; q < NRC; ++q) C[r][q] = 0.0;
    // terminal value function (optimization.cc:236-267): soft rows are costs, hard rows couple to their multiplier
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool soft = w.wt[i] >= 0.0;
      const double w2 = soft ? __dmul_rn(w.wt[i], w.wt[i]) : 0.0;
      P[i][i] = lambda + w2;
      Q[i][0] = w2 * dT[i];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nh && 1 + j < NRC && hcomp[j] == i) Q[i][1 + j] = 1.0;
    }
    for (int s = nseg - 1; s >= 0 && ok; --s) {
      for (int i = S - 1; i >= 0; --i) {
        const int col = s * S + i;
        const double b0 = Gam[col], b1 = Gam[W + col], b2 = Gam[2 * W + col], b3 = Gam[3 * W + col];
        const double gk = du[col];
        // P b (b4 = 1), symmetric storage
        double Pb[XI_DIM];
#pragma unroll
        for (int a = 0; a < XI_DIM; ++a) {
          const double pa0 = (a <= 0) ? P[a][0] : P[0][a], pa1 = (a <= 1) ? P[a][1] : P[1][a];
          const double pa2 = (a <= 2) ? P[a][2] : P[2][a], pa3 = (a <= 3) ? P[a][3] : P[3][a];
          Pb[a] = fma(pa0, b0, fma(pa1, b1, fma(pa2, b2, fma(pa3, b3, P[a][4]))));
        }
        const double bPb = fma(Pb[0], b0, fma(Pb[1], b1, fma(Pb[2], b2, fma(Pb[3], b3, Pb[4]))));
        const double piv = (wu2 + lambda + wr2) + bPb;
        if (!(piv > 0.0)) {
          ok = false;
          break;
        }
        const double inv = fm::rcp(piv);
        const double rz[XI_DIM] = {Pb[0], Pb[1], Pb[2], Pb[3], (col > 0) ? -wr2 : 0.0};
        double rr[NRC], kz[XI_DIM], kr[NRC];
#pragma unroll
        for (int r = 0; r < NRC; ++r) {
          rr[r] = fma(Q[0][r], b0, fma(Q[1][r], b1, fma(Q[2][r], b2, fma(Q[3][r], b3, Q[4][r]))));
          if (r == 0) rr[r] += gk;
          kr[r] = rr[r] * inv;
        }
#pragma unroll
        for (int a = 0; a < XI_DIM; ++a) kz[a] = rz[a] * inv;
        double* dst = c.kk + (size_t)col * GAIN_STRIDE;
#pragma unroll
        for (int a = 0; a < XI_DIM; ++a) __stcg(dst + a, kz[a]);
#pragma unroll
        for (int r = 0; r < NR_MAX; ++r) __stcg(dst + XI_DIM + r, (r < NRC) ? kr[r < NRC ? r : 0] : 0.0);
        // value function after the knot: the previous-control coordinate is replaced by this knot's own
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
          for (int bb = a; bb < 4; ++bb) P[a][bb] = fma(-rz[a], kz[bb], P[a][bb]);
          P[a][4] = -rz[a] * kz[4];
#pragma unroll
          for (int r = 0; r < NRC; ++r) Q[a][r] = fma(-rz[a], kr[r], Q[a][r]);
        }
        P[4][4] = fma(-rz[4], kz[4], (col > 0) ? wr2 : 0.0);
#pragma unroll
        for (int r = 0; r < NRC; ++r) {
          Q[4][r] = -rz[4] * kr[r];
#pragma unroll
          for (int q = r; q < NRC; ++q) C[r][q] = fma(-rr[r], kr[q], C[r][q]);
        }
      }
      if (!ok) break;
      // shooting node s: y = Phi dx_s + e_s rho_0; then lambda on dx_s (dx_0 is fixed by the initial-state rows)
      const double* Ph = Phi + 16 * s;
      const double f0 = ev[4 * s], f1 = ev[4 * s + 1], f2 = ev[4 * s + 2], f3 = ev[4 * s + 3];
      double Pf[XI_DIM];
#pragma unroll
      for (int a = 0; a < XI_DIM; ++a) {
        const double pa0 = (a <= 0) ? P[a][0] : P[0][a], pa1 = (a <= 1) ? P[a][1] : P[1][a];
        const double pa2 = (a <= 2) ? P[a][2] : P[2][a], pa3 = (a <= 3) ? P[a][3] : P[3][a];
        Pf[a] = pa0 * f0 + pa1 * f1 + pa2 * f2 + pa3 * f3;
      }
      // C picks up f . Q (row / column 0) and f . P f; only the rows of the multipliers and column 0 are read later
      {
        const double fq0 = Q[0][0] * f0 + Q[1][0] * f1 + Q[2][0] * f2 + Q[3][0] * f3;
        C[0][0] += 2.0 * fq0 + (Pf[0] * f0 + Pf[1] * f1 + Pf[2] * f2 + Pf[3] * f3);
#pragma unroll
        for (int r = 1; r < NRC; ++r) C[0][r] += Q[0][r] * f0 + Q[1][r] * f1 + Q[2][r] * f2 + Q[3][r] * f3;
      }
#pragma unroll
      for (int a = 0; a < XI_DIM; ++a) Q[a][0] += Pf[a];
      // Q <- blockdiag(Phi, 1)^T Q
#pragma unroll
      for (int r = 0; r < NRC; ++r) {
        const double q0 = Q[0][r], q1 = Q[1][r], q2 = Q[2][r], q3 = Q[3][r];
#pragma unroll
        for (int a = 0; a < 4; ++a) Q[a][r] = Ph[a] * q0 + Ph[4 + a] * q1 + Ph[8 + a] * q2 + Ph[12 + a] * q3;
      }
      // P <- blockdiag(Phi, 1)^T P blockdiag(Phi, 1)
      {
        double T[4][4];  // T = P44 Phi  (P44 = the 4x4 block, symmetric)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double pa0 = (a <= 0) ? P[a][0] : P[0][a], pa1 = (a <= 1) ? P[a][1] : P[1][a];
            const double pa2 = (a <= 2) ? P[a][2] : P[2][a], pa3 = (a <= 3) ? P[a][3] : P[3][a];
            T[a][j] = pa0 * Ph[j] + pa1 * Ph[4 + j] + pa2 * Ph[8 + j] + pa3 * Ph[12 + j];
          }
        const double p04 = P[0][4], p14 = P[1][4], p24 = P[2][4], p34 = P[3][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
          for (int j = a; j < 4; ++j)
            P[a][j] = Ph[a] * T[0][j] + Ph[4 + a] * T[1][j] + Ph[8 + a] * T[2][j] + Ph[12 + a] * T[3][j];
          P[a][4] = Ph[a] * p04 + Ph[4 + a] * p14 + Ph[8 + a] * p24 + Ph[12 + a] * p34;
          if (s > 0) P[a][a] += lambda;
        }
      }
    }

    // ---- multipliers of the hard terminal rows: stationarity in mu_j ------------------------------------------------
    //   Q[:4][1+j] . m0 + C[0][1+j] + sum_q C[1+j][1+q] mu_q + d_j = 0   (the rho-rho block is negative definite)
    if (ok && nh > 0) {
      double Mh[4][4], rhs[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rhs[j] = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) Mh[j][q] = (j == q) ? 1.0 : 0.0;
        if (j < nh && 1 + j < NRC) {
          const int hj = hcomp[j];
          const double dj = (hj == 0) ? dT[0] : (hj == 1) ? dT[1] : (hj == 2) ? dT[2] : dT[3];
          constexpr int jj_max = NRC - 1;
          const int jj = (1 + j <= jj_max) ? 1 + j : jj_max;
          rhs[j] = Q[0][jj] * m0[0] + Q[1][jj] * m0[1] + Q[2][jj] * m0[2] + Q[3][jj] * m0[3] + C[0][jj] + dj;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q <= j && q < nh) Mh[j][q] = -C[(1 + q < NRC) ? 1 + q : 0][(1 + j < NRC) ? 1 + j : 0];
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
    }
  };
  if (c.gl == 0) {
    if (nr <= 4) {
      backward(std::integral_constant<int, 4>{});
    } else {
      backward(std::integral_constant<int, NR_MAX>{});
    }
  }
  ok = __all_sync(c.gmask, ok);
  if (!ok) return false;

  // ---- forward sweep: v = -(k_z . z + k_rho . coef), y += Gamma v.  The feedback rows of a segment are copied from
  // the global scratch to shared memory by the whole group (one L2 round trip per segment instead of one per knot).
  double y0 = m0[0], y1 = m0[1], y2 = m0[2], y3 = m0[3], pv = 0.0;
  if (c.gl == 0) {
    dxs[0] = m0[0];
    dxs[1] = m0[1];
    dxs[2] = m0[2];
    dxs[3] = m0[3];
  }
  // rows of a segment: S * GAIN_STRIDE doubles, staged k.stage_knots knots at a time (a whole segment unless shared
  // memory is short); when a whole segment fits NPF per lane it is prefetched into registers one segment ahead, so the L2
  // round trip of segment s + 1 overlaps the first lane's work on segment s
  // (S = 10 knots are 100 doubles: 13 per lane of an 8-lane group, the W = 160 mapping, where the recursion runs on
  // every iteration: 8 % of that step.  Not for 4-lane groups: 25 more live doubles per lane cost the default problem,
  // which takes the recursion on damped iterations only, 9 % of its step through the register allocation.)
  constexpr int NPF = (L == 16) ? 7 : (L == 8) ? 13 : 4;
  const int n_rows = S * GAIN_STRIDE;
  const int SK = k.stage_knots;
  const bool fits = n_rows <= NPF * L && SK == S;
  double pf[NPF];
  auto fetch = [&](const int seg) {
    const double* src = c.kk + (size_t)seg * n_rows;
#pragma unroll
    for (int q = 0; q < NPF; ++q) {
      const int t = c.gl + q * L;
      pf[q] = (t < n_rows) ? __ldcg(src + t) : 0.0;
    }
  };
  CPMPC_GSYNC();  // the backward sweep's stores are visible to the group
  if (fits) fetch(0);
  for (int s = 0; s < nseg; ++s) {
    #pragma unroll 1
    for (int i0 = 0; i0 < S; i0 += SK) {
      const int i1 = (i0 + SK < S) ? i0 + SK : S;
      if (fits) {
#pragma unroll
        for (int q = 0; q < NPF; ++q) {
          const int t = c.gl + q * L;
          if (t < n_rows) stage[t] = pf[q];
        }
        if (s + 1 < nseg) fetch(s + 1);
      } else {
        const double* src = c.kk + (size_t)s * n_rows + i0 * GAIN_STRIDE;
        const int nb = (i1 - i0) * GAIN_STRIDE;
        for (int t = c.gl; t < nb; t += L) stage[t] = __ldcg(src + t);
      }
      CPMPC_GSYNC();
      if (c.gl == 0) {
        if (i0 == 0) {
          const double* Ph = Phi + 16 * s;
          const double n0 = Ph[0] * y0 + Ph[1] * y1 + Ph[2] * y2 + Ph[3] * y3 + ev[4 * s];
          const double n1 = Ph[4] * y0 + Ph[5] * y1 + Ph[6] * y2 + Ph[7] * y3 + ev[4 * s + 1];
          const double n2 = Ph[8] * y0 + Ph[9] * y1 + Ph[10] * y2 + Ph[11] * y3 + ev[4 * s + 2];
          const double n3 = Ph[12] * y0 + Ph[13] * y1 + Ph[14] * y2 + Ph[15] * y3 + ev[4 * s + 3];
          y0 = n0;
          y1 = n1;
          y2 = n2;
          y3 = n3;
        }
        for (int i = i0; i < i1; ++i) {
          const int col = s * S + i;
          const double* g = stage + (i - i0) * GAIN_STRIDE;
          double acc = g[0] * y0 + g[1] * y1 + g[2] * y2 + g[3] * y3 + g[4] * pv;
          double acc2 = g[XI_DIM] * coef[0] + g[XI_DIM + 1] * coef[1] + g[XI_DIM + 2] * coef[2] + g[XI_DIM + 3] * coef[3] +
                        g[XI_DIM + 4] * coef[4];
          const double v = -(acc + acc2);
          cdx = fma(du[col], v, cdx);  // du[] still holds c_u
          du[col] = v;
          y0 = fma(Gam[col], v, y0);
          y1 = fma(Gam[W + col], v, y1);
          y2 = fma(Gam[2 * W + col], v, y2);
          y3 = fma(Gam[3 * W + col], v, y3);
          pv = v;
        }
        if (i1 == S) {
          dxs[4 * (s + 1)] = y0;
          dxs[4 * (s + 1) + 1] = y1;
          dxs[4 * (s + 1) + 2] = y2;
          dxs[4 * (s + 1) + 3] = y3;
        }
      }
      CPMPC_GSYNC();  // the staged rows have been consumed
    }
  }
  return true;
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
// Scratch (shared with the staging buffer of the Riccati recursion): Z [4][W], tinv [W].
// ---------------------------------------------------------------------------------------------------------------------
template <int L, bool FX = false>
__device__ __forceinline__ bool condensed_qp(const Ctx& c, const Weights& w, bool& tri_valid, const double x_cur[4],
                                             const double target[4], double& cdx) {
  const KernelCfg& k = *c.cfg;
  const int W = FX ? FX_W : k.W, S = FX ? FX_S : k.S, NS = FX ? FX_NS : k.NS, nseg = FX ? FX_NSEG : k.nseg;
  const double* xs = c.sm + k.off_xs;
  double* dxs = c.sm + k.off_dxs;
  double* du = c.sm + k.off_du;
  const double* Phi = c.sm + k.off_phi;
  const double* ev = c.sm + k.off_e;
  const double* Gam = c.sm + k.off_gam;
  // Z [4][W] in shared memory behind tinv, or (k.z_global) in the group's L2-resident global scratch, which the Riccati
  // recursion - never active in the same iteration - uses for its feedback rows
  double* tinv = c.sm + k.off_aug;  // reciprocal pivots of T = L D L^T
  double* Z = k.z_global ? c.kk : tinv + W;
  // sub-diagonal multipliers of the unit bidiagonal factor: l_j = -w_r^2 / d_{j-1}, recomputed from the pivots where they
  // are used (one multiply off the dependent chain) instead of being stored: W doubles less shared memory per instance,
  // which is what lets a sixth warp fit on an SM at the default problem size
  const double wu2 = w.w_u * w.w_u, wr2 = w.w_r * w.w_r;
  const double nwr2 = -wr2;
#define CPMPC_TL(j) (nwr2 * tinv[(j) - 1])
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
      tinv[j] = dprev_inv;
    }
  }
  tri_valid = true;
  // (2) rows of U: sensitivity of the terminal node to every control, through the later segments.  Two columns of
  // this lane per trip (col and col + L: the same or neighbouring segments), so that the products of one column fill
  // the FMA latency of the other; the segment index is carried along instead of divided out.  Per column the same
  // operations in the same order as one column at a time.
  {
    int col = c.gl, sa = 0, sa_end = S;  // sa = col / S
    #pragma unroll 1
    for (; col < W; col += 2 * L) {
      while (col >= sa_end) {
        ++sa;
        sa_end += S;
      }
      const int colb = col + L;
      const bool hb = colb < W;
      const int cb = hb ? colb : col;
      int sb = sa;
      if (cb >= sa_end) ++sb;  // L <= S in every mapping: at most one boundary between the two columns
      if (cb >= sa_end + S) sb = cb / S;  // (not reached while L <= S; kept for safety)
      double v0 = Gam[col], v1 = Gam[W + col], v2 = Gam[2 * W + col], v3 = Gam[3 * W + col];
      double y0 = Gam[cb], y1 = Gam[W + cb], y2 = Gam[2 * W + cb], y3 = Gam[3 * W + cb];
      #pragma unroll 1
      for (int t = sa + 1; t < nseg; ++t) {
        const double* Ph = Phi + 16 * t;
        const double p0 = Ph[0], p1 = Ph[1], p2 = Ph[2], p3 = Ph[3], p4 = Ph[4], p5 = Ph[5], p6 = Ph[6], p7 = Ph[7];
        const double p8 = Ph[8], p9 = Ph[9], p10 = Ph[10], p11 = Ph[11], p12 = Ph[12], p13 = Ph[13], p14 = Ph[14],
                     p15 = Ph[15];
        const double n0 = p0 * v0 + p1 * v1 + p2 * v2 + p3 * v3;
        const double n1 = p4 * v0 + p5 * v1 + p6 * v2 + p7 * v3;
        const double n2 = p8 * v0 + p9 * v1 + p10 * v2 + p11 * v3;
        const double n3 = p12 * v0 + p13 * v1 + p14 * v2 + p15 * v3;
        const double m0 = p0 * y0 + p1 * y1 + p2 * y2 + p3 * y3;
        const double m1 = p4 * y0 + p5 * y1 + p6 * y2 + p7 * y3;
        const double m2 = p8 * y0 + p9 * y1 + p10 * y2 + p11 * y3;
        const double m3 = p12 * y0 + p13 * y1 + p14 * y2 + p15 * y3;
        v0 = n0;
        v1 = n1;
        v2 = n2;
        v3 = n3;
        const bool ub = t > sb;
        y0 = ub ? m0 : y0;
        y1 = ub ? m1 : y1;
        y2 = ub ? m2 : y2;
        y3 = ub ? m3 : y3;
      }
      Z[col] = v0;
      Z[W + col] = v1;
      Z[2 * W + col] = v2;
      Z[3 * W + col] = v3;
      if (hb) {
        Z[colb] = y0;
        Z[W + colb] = y1;
        Z[2 * W + colb] = y2;
        Z[3 * W + colb] = y3;
      }
    }
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
  #pragma unroll 1
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
  #pragma unroll 1
  for (int r = c.gl; r < 5; r += L) {
    double* row = (r < 4) ? (Z + r * W) : du;
    double prev = row[0];
    // four elements per trip with the loads issued ahead of the dependent FMAs and the stores behind them: the
    // compiler cannot move a load of row[j + 1] above the store of row[j] on its own (same array), and without that
    // every element pays a shared-memory round trip on top of the FMA.  Same operations in the same order.
    int j = 1;
    for (; j + 4 <= W; j += 4) {
      const double t0 = CPMPC_TL(j), t1 = CPMPC_TL(j + 1), t2 = CPMPC_TL(j + 2), t3 = CPMPC_TL(j + 3);
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
    #pragma unroll 1
    for (; j < W; ++j) {
      prev = fma(-CPMPC_TL(j), prev, row[j]);
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
    auto gram_col = [&](const double z0, const double z1, const double z2, const double z3, const double zg,
                        const double ti) {
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
    };
    int q = c.gl;
    #pragma unroll 1
    for (; q + L < W; q += 2 * L) {  // two columns per trip: the loads of the second ahead of the sums of the first
      const int qb = q + L;
      const double tia = tinv[q], tib = tinv[qb], zga = du[q], zgb = du[qb];
      const double a0 = Z[q], a1 = Z[W + q], a2 = Z[2 * W + q], a3 = Z[3 * W + q];
      const double b0 = Z[qb], b1 = Z[W + qb], b2 = Z[2 * W + qb], b3 = Z[3 * W + qb];
      gram_col(a0, a1, a2, a3, zga, tia);
      gram_col(b0, b1, b2, b3, zgb, tib);
    }
    if (q < W) gram_col(Z[q], Z[W + q], Z[2 * W + q], Z[3 * W + q], du[q], tinv[q]);
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
        invd = fm::rcp(dd);
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
    double root_, inv;
    fm::sqrt_rsqrt(d, root_, inv);  // straight-line 1 / sqrt(d): libdevice's division + square root carry slow paths
    (void)root_;
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
  {
    // two columns per trip with all loads ahead of the stores; the sum `part` runs over the columns in the same order
    int q = c.gl;
    #pragma unroll 1
    for (; q + L < W; q += 2 * L) {
      const int qb = q + L;
      const double zga = du[q], zgb = du[qb], tia = tinv[q], tib = tinv[qb];
      const double a0 = Z[q], a1 = Z[W + q], a2 = Z[2 * W + q], a3 = Z[3 * W + q];
      const double b0 = Z[qb], b1 = Z[W + qb], b2 = Z[2 * W + qb], b3 = Z[3 * W + qb];
      const double wa = (zga + sig[0] * a0 + sig[1] * a1 + sig[2] * a2 + sig[3] * a3) * tia;
      const double wb = (zgb + sig[0] * b0 + sig[1] * b1 + sig[2] * b2 + sig[3] * b3) * tib;
      part = fma(zga, wa, part);
      part = fma(zgb, wb, part);
      du[q] = wa;
      du[qb] = wb;
    }
    if (q < W) {
      const double zg = du[q];
      const double wq = (zg + sig[0] * Z[q] + sig[1] * Z[W + q] + sig[2] * Z[2 * W + q] + sig[3] * Z[3 * W + q]) * tinv[q];
      part = fma(zg, wq, part);
      du[q] = wq;
    }
  }
  cdx = -part;  // this lane's share of c_u . du
  CPMPC_GSYNC();
  if (c.gl == 0) {
    double nxt = du[W - 1];
    du[W - 1] = -nxt;
    int j = W - 2;
    #pragma unroll 1
    for (; j >= 3; j -= 4) {  // same batching as the forward substitution, descending
      const double t0 = CPMPC_TL(j + 1), t1 = CPMPC_TL(j), t2 = CPMPC_TL(j - 1), t3 = CPMPC_TL(j - 2);
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
    #pragma unroll 1
    for (; j >= 0; --j) {
      nxt = fma(-CPMPC_TL(j + 1), nxt, du[j]);
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
  #pragma unroll 1
  for (int s = 0; s < nseg; ++s) {
    const double* Ph = Phi + 16 * s;
    #pragma unroll 1
    for (int a = c.gl; a < 4; a += L) {
      double n0 = Ph[4 * a] * xi[0] + Ph[4 * a + 1] * xi[1] + ev[4 * s + a];
      double n1 = Ph[4 * a + 2] * xi[2] + Ph[4 * a + 3] * xi[3];
      const double* ga = Gam + a * W + s * S;
      const double* vv = du + s * S;
      int j = 0;
      #pragma unroll 1
      for (; j + 3 < S; j += 4) {  // four products per trip, their eight loads first; the two sums keep their order
        const double g0 = ga[j], g1 = ga[j + 1], g2 = ga[j + 2], g3 = ga[j + 3];
        const double w0 = vv[j], w1 = vv[j + 1], w2 = vv[j + 2], w3 = vv[j + 3];
        n0 = fma(g0, w0, n0);
        n1 = fma(g1, w1, n1);
        n0 = fma(g2, w2, n0);
        n1 = fma(g3, w3, n1);
      }
      #pragma unroll 1
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
#undef CPMPC_TL
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
// LEAN: the throughput mapping with nothing attached - one chunk per segment (cfg.cps == 1), no iteration log, no debug
// view.  The host launches this instantiation whenever that holds (cpmpc_capi.cu SelectKernel); the paths it drops are
// 40 % of the general kernel's instructions, never executed there but in the way of its register allocation.
template <int L, bool LEAN = false, bool FX = false>
__device__ __forceinline__ void solve_group(const Ctx& c, const bool active_in, const Weights& w,
                                            const double x_cur[4], const double& set_point, const double& u_prev,
                                            double* iter_log, SolveResult& res, long long* ph, double* dbg = nullptr,
                                            const double dbg_lambda = 0.0, const int dbg_qp = 0) {
  // ph[0..N_PHASES) accumulate cycles per phase, ph[N_PHASES] is the time of the last mark
  // (not in the lean instantiations: eight clock reads and local-memory updates per iteration are 1.5 % of a step;
  // cpmpc_enable_phase_clocks selects the general kernel, which keeps them)
#define CPMPC_MARK(idx)                   \
  do {                                    \
    if (!LEAN) {                          \
      const long long now_ = clock64();   \
      ph[idx] += now_ - ph[N_PHASES];     \
      ph[N_PHASES] = now_;                \
    }                                     \
  } while (0)
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
  const int W = FX ? FX_W : k.W, S = FX ? FX_S : k.S, NS = FX ? FX_NS : k.NS, nseg = FX ? FX_NSEG : k.nseg;
  const int cps = LEAN ? 1 : k.cps, nchunk = LEAN ? nseg : k.nchunk;

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
    // branch-free (weights of absent cost families are 0): the rate residual behind the last knot is formed against the
    // knot itself, the one in front of knot 0 against the previous step's control
    const double wr2 = w.w_r * w.w_r, wu2 = w.w_u * w.w_u;
    int col = c.gl;
    #pragma unroll 1
    for (; col + L < W; col += 2 * L) {  // two columns per trip, loads ahead of the stores
      const int cb = col + L;
      const double uk = u[col], un = u[col + 1], up = (col > 0) ? u[col - 1] : u_prev;
      const double vk = u[cb], vn = u[(cb + 1 < W) ? cb + 1 : cb], vp = u[cb - 1];
      du[col] = wr2 * (uk - un) + wr2 * (uk - up) + wu2 * uk;
      du[cb] = wr2 * (vk - vn) + wr2 * (vk - vp) + wu2 * vk;
    }
    if (col < W) {
      const double uk = u[col];
      const double un = u[(col + 1 < W) ? col + 1 : col];
      const double up = (col > 0) ? u[col - 1] : u_prev;
      du[col] = wr2 * (uk - un) + wr2 * (uk - up) + wu2 * uk;
    }
    CPMPC_GSYNC();
  };

  #pragma unroll 1
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
      for (int id = c.gl; id < nchunk; id += L) {
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
            // Gamma_j <- A Gamma_j for the earlier knots of this chunk, two columns per trip: the loads of a pair are
            // issued ahead of its stores (the compiler cannot move a load above a store into the same array on its
            // own), so the shared-memory latency is paid once per pair
            int j = s * S + i0;
            for (; j + 1 < col; j += 2) {
              const double g0 = Gam[j], g1 = Gam[W + j], g2 = Gam[2 * W + j], g3 = Gam[3 * W + j];
              const double h0 = Gam[j + 1], h1 = Gam[W + j + 1], h2 = Gam[2 * W + j + 1], h3 = Gam[3 * W + j + 1];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const double a = A[4 * r] * g0 + A[4 * r + 1] * g1 + A[4 * r + 2] * g2 + A[4 * r + 3] * g3;
                const double b = A[4 * r] * h0 + A[4 * r + 1] * h1 + A[4 * r + 2] * h2 + A[4 * r + 3] * h3;
                Gam[r * W + j] = a;
                Gam[r * W + j + 1] = b;
              }
            }
            if (j < col) {
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
        #pragma unroll 1
        for (int base = 0; base < nchunk; base += L) {  // uniform trip count: the loop body synchronises
          const int id = base + c.gl;
          const int s = id / cps, cc = id - s * cps;
          const bool mine = id < nchunk && cc < cps - 1;
          double Sx[16], Pr[16];
          if (mine) {
            const double* last = PhiC + 16 * (s * cps + cps - 1);
#pragma unroll
            for (int q = 0; q < 16; ++q) Sx[q] = last[q];
            #pragma unroll 1
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
        #pragma unroll 1
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
        boundary_terms<FX>(k, w, xs, x_cur, set_point, ft, eb);
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
    CPMPC_MARK(1);

    // ---- QP step (qp.cc:679-729) ---------------------------------------------------------------------------------
    double cdx = 0.0;
    bool qp_ok = false;
    bool need_riccati = false;
    if (active) {
      const bool try_condensed = (dbg != nullptr && dbg_qp != 0) ? (dbg_qp == 2) : (k.condensed && !(lambda > 0.0));
      if (try_condensed) {
        qp_ok = condensed_qp<L, FX>(c, w, tri_valid, x_cur, target, cdx);
        // A failed 4x4 multiplier system is a property of the condensed formulation (open-loop sensitivity rows nearly
        // dependent), not necessarily of the QP: restore the gradient and let the Riccati recursion decide.
        if (!qp_ok) fill_gradient();
      }
      need_riccati = !qp_ok && !(dbg != nullptr && dbg_qp == 2);
      if (need_riccati) tri_valid = false;  // the Riccati stage matrix shares the scratch of the tridiagonal factor
    }
    CPMPC_MARK(2);
    if (need_riccati) qp_ok = riccati_qp<L, FX>(c, w, lambda, x_cur, target, cdx);
    CPMPC_MARK(3);
    if (active) {
      if (dbg != nullptr) {
        // debug view: the blocks of this linearisation and the QP step, then stop (the candidates would overwrite Gam)
        CPMPC_GSYNC();
        double* o = dbg;
        #pragma unroll 1
        for (int q = c.gl; q < 16 * nseg; q += L) o[q] = Phi[q];
        o += 16 * nseg;
        #pragma unroll 1
        for (int q = c.gl; q < 4 * W; q += L) o[q] = Gam[q];
        o += 4 * W;
        #pragma unroll 1
        for (int q = c.gl; q < 4 * nseg; q += L) o[q] = ev[q];
        o += 4 * nseg;
        #pragma unroll 1
        for (int q = c.gl; q < 4 * NS; q += L) o[q] = dxs[q];
        o += 4 * NS;
        #pragma unroll 1
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
        #pragma unroll 1
        for (int col = c.gl; col < W; col += L) {
          const double a = wu2 + lambda + wr2 * ((col + 1 < W) ? 2.0 : 1.0);
          const double v = du[col];
          quad += a * v * v;
          if (col + 1 < W) quad -= 2.0 * wr2 * v * du[col + 1];
        }
        #pragma unroll 1
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
          #pragma unroll 1
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

    CPMPC_MARK(4);
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
    #pragma unroll 1
    for (int ls = 0; ls <= k.max_ls; ++ls) {
      if (__all_sync(0xffffffffu, ls_done)) break;
      if (!ls_done && ls > 0) {
        // polynomial step-length selection (nonlinear.cc:418-443,531-605)
        double new_alpha = 0.0;
        bool valid;
        // The quotients and the square root are the straight-line forms of fastmath.cuh (1 ulp; the IEEE division and
        // square root of the CUDA math library carry slow-path branches, and every lane of the warp walks through this
        // code in every trial: 2.7 % of the kernel's samples).  A zero or non-finite denominator yields a non-finite
        // step length as it does with IEEE arithmetic, which the test below rejects.
        if (ls == 1) {
          const double numerator = phi_prev - dphi0 * a_prev - phi0;
          valid = !(dphi0 > 0.0 || numerator <= 0.0);
          new_alpha = (-dphi0 * a_prev * a_prev) * fm::rcp(2.0 * numerator);
        } else {
          const double a0 = a_prev2, a1 = a_prev;
          const double m00 = a0 * a0 * a0, m01 = a0 * a0, m10 = a1 * a1 * a1, m11 = a1 * a1;
          const double r0 = phi_prev2 - phi0 - dphi0 * a0;
          const double r1 = phi_prev - phi0 - dphi0 * a1;
          const double inv_det = fm::rcp(m00 * m11 - m10 * m01);
          const double ca = (m11 * inv_det) * r0 + (-m01 * inv_det) * r1;
          const double cb = (-m10 * inv_det) * r0 + (m00 * inv_det) * r1;
          const double arg_sqrt = cb * cb - 3.0 * ca * dphi0;
          valid = !(ca == 0.0 || arg_sqrt < -1.0e-12);
          double root, inv_root;
          fm::sqrt_rsqrt(arg_sqrt, root, inv_root);
          (void)inv_root;
          new_alpha = (-cb + ((arg_sqrt > 0.0) ? root : 0.0)) * fm::rcp(3.0 * ca);
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
        {
          // two columns per trip, loads ahead of the stores (uc aliases neither u nor du, but the compiler cannot know)
          int col = c.gl;
          #pragma unroll 1
          for (; col + L < W; col += 2 * L) {
            const double ua = u[col], ub = u[col + L], da = du[col], db = du[col + L];
            uc[col] = clampd(__dadd_rn(ua, __dmul_rn(da, alpha)), -k.u_clamp, k.u_clamp);
            uc[col + L] = clampd(__dadd_rn(ub, __dmul_rn(db, alpha)), -k.u_clamp, k.u_clamp);
          }
          if (col < W) uc[col] = clampd(__dadd_rn(u[col], __dmul_rn(du[col], alpha)), -k.u_clamp, k.u_clamp);
        }
        CPMPC_GSYNC();
        CPMPC_PARK();
        evaluate_errors<L, LEAN, FX>(c, w, xc, uc, xkc, x_cur, set_point, u_prev, step_f, step_eq);
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

    CPMPC_MARK(5);
    // ---- lambda update and exit tests (nonlinear.cc:296-343) -------------------------------------------------------
    if (active) {
      bool exit_now = false;
      if (step_result == SR_SUCCESS) {
        #pragma unroll 1
        for (int q = c.gl; q < 4 * NS; q += L) xs[q] = xc[q];
        #pragma unroll 1
        for (int col = c.gl; col < W; col += L) u[col] = uc[col];
        if (cps > 1)
          #pragma unroll 1
          for (int q = c.gl; q < 4 * nchunk; q += L) xk[q] = xkc[q];
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
    CPMPC_MARK(6);
  }
#undef CPMPC_MARK
  res.opt_state = opt_state;
  res.f = cur_f;
  res.eq = cur_eq;
  res.penalty = penalty;
  res.lambda = lambda;
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel: persistent grid-stride over instances; each warp carries 32/L instances at a time.
// ---------------------------------------------------------------------------------------------------------------------
// SIM: the plant step of the fused closed loop is in this kernel (io.sim && !io.split decides at run time); the lean
// instantiation without it serves every plain Step and the split closed loop.
// FX: the problem has the default shape (FX_W, FX_S), known at compile time.
template <int L, bool LEAN = false, bool SIM = true, bool FX = false>
__global__ void __launch_bounds__(256, 1) cpmpc_solve_kernel(const KernelCfg cfg, const KernelIO io) {
  const bool sim_here = SIM ? (io.sim != 0 && !io.split) : false;
  const int cps_guess = LEAN ? 1 : cfg.cps;
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
  const int W = FX ? FX_W : cfg.W, S = FX ? FX_S : cfg.S, NS = FX ? FX_NS : cfg.NS, N = FX ? FX_N : cfg.N;
  double* xs = c.sm + cfg.off_xs;
  double* u = c.sm + cfg.off_u;

  // Work queue: a warp takes G consecutive instances at a time, so warps that draw cheap instances (few line-search
  // trials) go on instead of idling until the slowest warp of a static round is done.  Warp-wide votes stay
  // convergent because all groups of a warp enter and leave a batch together.
  const long long t_begin = clock64();
  unsigned long long last_base = ~0ull;
  struct WarpPhases {
    long long ph[N_PHASES + 1];
  } wph;
  asm volatile("" : : "l"(&wph) : "memory");
#pragma unroll
  for (int i = 0; i < N_PHASES; ++i) wph.ph[i] = 0;
  wph.ph[N_PHASES] = t_begin;
  #pragma unroll 1
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(io.queue, (unsigned long long)G);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned long long)io.B) break;
    last_base = base;
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
    long long* ph = wph.ph;
    asm volatile("" : : "l"(&kc) : "memory");
    // FX instantiations are launched without per-instance weights only (cpmpc_capi.cu Launch): the weights are the
    // handle's, i.e. kernel parameters, and a local that never escapes lets the compiler read them from the constant
    // bank where they are used instead of keeping a per-thread copy in (parked) local memory
    Weights w_uniform;
    if (FX) {
      w_uniform.w_u = (cfg.w_u > 0.0) ? cfg.w_u : 0.0;
      w_uniform.w_r = (cfg.w_r > 0.0) ? cfg.w_r : 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) w_uniform.wt[i] = cfg.wt[i];
    }
    Weights& w = FX ? w_uniform : kc.w;
    double(&x_cur)[4] = kc.x_cur;
    double& set_point = kc.set_point;
    x_cur[0] = x_cur[1] = x_cur[2] = x_cur[3] = 0.0;
    set_point = 0.0;
    if (have) {
      if (c.gl == 0) store_dyn_const(c.sm + cfg.off_dyn, make_dyn_const(io.dyn + inst * CPMPC_DYN_STRIDE));
      if (FX) {
        // (uniform weights, above)
      } else if (io.weights != nullptr) {
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
      if (!FX) {
        if (!(w.w_u > 0.0)) w.w_u = 0.0;
        if (!(w.w_r > 0.0)) w.w_r = 0.0;
      }
      const double* ps = (io.sim ? io.states_io : io.states) + inst * CPMPC_STATE_STRIDE;
#pragma unroll
      for (int i = 0; i < 4; ++i) x_cur[i] = ps[i];
      x_cur[1] = wrap_far_angle(x_cur[1]);
      set_point = io.set_points[inst];
    } else {
      if (c.gl == 0) store_dyn_const(c.sm + cfg.off_dyn, dummy_dyn_const());
      if (!FX) {
        w.w_u = w.w_r = 1.0;
        w.wt[0] = w.wt[1] = w.wt[2] = w.wt[3] = 1.0;
      }
    }

    CPMPC_GSYNC();

    long long& solves = kc.solves;
    long long& sum_iter = kc.sum_iter;
    long long& sum_ls = kc.sum_ls;
    long long& sum_bad = kc.sum_bad;
    solves = sum_iter = sum_ls = sum_bad = 0;
    #pragma unroll 1
    for (long long step = 0; step < io.n_steps; ++step) {
      // ---- BuildProblem + initial guess (optimization.cc:44-71) ---------------------------------------------------
      double& u_prev = kc.u_prev;
      u_prev = 0.0;
      if (have) {
        double* prev = io.prev + inst * N;
        const bool warm = io.has_prev[inst] != 0;
        if (warm) u_prev = prev[4 * NS];  // last step's applied control (optimization.cc:289-291)
        #pragma unroll 1
        for (int q = c.gl; q < W; q += L) {
          u[q] = warm ? prev[4 * NS + ((q + 1 < W) ? q + 1 : q)] : io.u_guess[q];
        }
        CPMPC_GSYNC();
        if (io.guess != nullptr) {
          // split launch: cpmpc_guess_kernel has integrated the shooting states already
          const double* gx = io.guess + inst * (4 * NS);
          #pragma unroll 1
          for (int q = c.gl; q < 4 * NS; q += L) xs[q] = gx[q];
        } else if (c.gl == 0) {
          // FillInitialGuess: all shooting states re-integrated from x(0) (optimization.cc:333-351); the states at
          // the chunk boundaries inside a segment are kept for the first linearisation
          DynConst Pl;
          load_dyn_const(Pl, c.sm + cfg.off_dyn);
          double x[4] = {x_cur[0], x_cur[1], x_cur[2], x_cur[3]};
          double* xk = c.sm + cfg.off_xk;
#pragma unroll
          for (int i = 0; i < 4; ++i) xs[i] = x[i];
          #pragma unroll 1
          for (int s = 1; s < NS; ++s) {
            int cc = 1;
            int nb = (cps_guess > 1) ? (S / cps_guess) : S;
            #pragma unroll 1
            for (int i = 0; i < S; ++i) {
              if (i == nb) {
                double* dst = xk + 4 * ((s - 1) * cps_guess + cc);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = x[q];
                ++cc;
                nb = (cc < cps_guess) ? (cc * S) / cps_guess : S;
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
      double* ilog = (!LEAN && io.iter_log != nullptr && have)
                         ? io.iter_log + (size_t)inst * cfg.max_iter * CPMPC_ITERLOG_STRIDE
                         : nullptr;
      double* dbg = (!LEAN && io.dbg != nullptr && have) ? io.dbg + (size_t)inst * io.dbg_stride : nullptr;
      if (!LEAN) {
        const long long now = clock64();
        ph[0] += now - ph[N_PHASES];
        ph[N_PHASES] = now;
      }
      solve_group<L, LEAN, FX>(c, have, w, x_cur, set_point, u_prev, ilog, res, ph, dbg, io.dbg_lambda, io.dbg_qp);

      if (have) {
        // ---- outputs (optimization.cc:83-96) ------------------------------------------------------------------
        // A rejected instance keeps its previous warm start: the reference throws inside Solve()
        // (nonlinear.cc:356-357) before previous_solution_ is assigned (optimization.cc:84-85).  The debug view
        // never touches the warm start either.
        const bool keep = res.termination != TS_REJECTED && (LEAN || io.dbg == nullptr);
        double* prev = io.prev + inst * N;
        if (keep)
          #pragma unroll 1
          for (int q = c.gl; q < 4 * NS; q += L) prev[q] = xs[q];
        #pragma unroll 1
        for (int q = c.gl; q < W; q += L) {
          if (keep) prev[4 * NS + q] = u[q];
          if (io.u_out != nullptr) io.u_out[inst * W + q] = u[q];
        }
        if (c.gl == 0) {
          if (keep) io.has_prev[inst] = 1;
          if (io.pred_out != nullptr && !io.split) {
            // ComputePredictedStates (optimization.cc:353-371)
            DynConst Pl;
            load_dyn_const(Pl, c.sm + cfg.off_dyn);
            double x[4] = {x_cur[0], x_cur[1], x_cur[2], x_cur[3]};
            double* po = io.pred_out + (size_t)inst * W * 4;
            #pragma unroll 1
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
          if (sim_here) {
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
        if (sim_here) {
          const int src = (c.gmask == 0xffffffffu) ? 0 : (__ffs(c.gmask) - 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) x_cur[i] = __shfl_sync(c.gmask, x_cur[i], src);
        }
        CPMPC_GSYNC();
      }
      if (!LEAN) {
        const long long now = clock64();
        ph[7] += now - ph[N_PHASES];
        ph[N_PHASES] = now;
      }
    }
    if (have && c.gl == 0) {
      if (sim_here) {
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
      unsigned long long* rec = io.warp_cycles + (size_t)WARP_REC * ((size_t)blockIdx.x * WARPS_PER_BLOCK + warp);
      rec[0] = (unsigned long long)(clock64() - t_begin);
      rec[1] = last_base;
#pragma unroll
      for (int i = 0; i < N_PHASES; ++i) rec[2 + i] = (unsigned long long)wph.ph[i];
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

// The plant constants exactly as a lane of cpmpc_solve_kernel sees them: every one a finished double that went through
// shared memory.  Without the barrier the compiler may contract a multiplication inside make_dyn_const with an addition
// of the dynamics into one FMA, and the rollouts of the split launch would round differently from the fused kernel's.
__device__ __forceinline__ DynConst make_dyn_const_opaque(const double* __restrict__ p) {
  double tmp[DYN_CONST_DOUBLES];
  store_dyn_const(tmp, make_dyn_const(p));
#pragma unroll
  for (int i = 0; i < DYN_CONST_DOUBLES; ++i) asm volatile("" : "+d"(tmp[i]));
  DynConst c;
  load_dyn_const(c, tmp);
  return c;
}

// ---------------------------------------------------------------------------------------------------------------------
// Split launch, kernel 1 of 3: FillInitialGuess (optimization.cc:333-351) for every instance of the batch, one THREAD
// per instance.  Inside cpmpc_solve_kernel this 40-knot serial rollout occupies one lane of each group (a quarter of a
// warp at the defaults, an eighth at W = 160) while the others wait; here every lane integrates its own plant.  Same
// controls (warm start shifted by one knot, optimization.cc:57-61, or the cold-start guess), same rk4_step, same order:
// the shooting states are bit-identical to the fused path's (profiles/ab_compare.py).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 4) cpmpc_guess_kernel(const KernelCfg cfg, const KernelIO io, double* __restrict__ guess) {
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= io.B) return;
  const int W = cfg.W, S = cfg.S, NS = cfg.NS, N = cfg.N;
  const DynConst P = make_dyn_const_opaque(io.dyn + inst * CPMPC_DYN_STRIDE);
  const double* ps = (io.sim ? io.states_io : io.states) + inst * CPMPC_STATE_STRIDE;
  double x[4] = {ps[0], wrap_far_angle(ps[1]), ps[2], ps[3]};
  const bool warm = io.has_prev[inst] != 0;
  const double* pu = warm ? io.prev + inst * N + 4 * NS : io.u_guess;
  double* gx = guess + inst * (4 * NS);
#pragma unroll
  for (int i = 0; i < 4; ++i) gx[i] = x[i];
  #pragma unroll 1
  for (int s = 1; s < NS; ++s) {
    #pragma unroll 1
    for (int i = 0; i < S; ++i) {
      const int q = (s - 1) * S + i;
      const double uq = warm ? pu[(q + 1 < W) ? q + 1 : q] : pu[q];
      rk4_step<false>(P, x, uq, cfg.dt);
      x[1] = mod_pi(x[1]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) gx[4 * s + i] = x[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Split launch, kernel 3 of 3: ComputePredictedStates (optimization.cc:353-371) and, in the closed loop, the plant's
// Simulator::Step(control_dt, u[0]) with 1 ms sub-steps (simulator.cc:11-36), one THREAD per instance, from the controls
// the solve kernel left in u[B, W].
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 4) cpmpc_post_kernel(const KernelCfg cfg, const KernelIO io, const double* __restrict__ u) {
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= io.B) return;
  const int W = cfg.W;
  const DynConst P = make_dyn_const_opaque(io.dyn + inst * CPMPC_DYN_STRIDE);
  const double* ps = (io.sim ? io.states_io : io.states) + inst * CPMPC_STATE_STRIDE;
  const double x0[4] = {ps[0], wrap_far_angle(ps[1]), ps[2], ps[3]};
  const double* pu = u + inst * W;
  if (io.pred_out != nullptr) {
    double x[4] = {x0[0], x0[1], x0[2], x0[3]};
    double* po = io.pred_out + (size_t)inst * W * 4;
    #pragma unroll 1
    for (int q = 0; q < W; ++q) {
      rk4_step<false>(P, x, pu[q], cfg.dt);
      x[1] = mod_pi(x[1]);
      po[4 * q] = x[0];
      po[4 * q + 1] = x[1];
      po[4 * q + 2] = x[2];
      po[4 * q + 3] = x[3];
    }
  }
  if (io.sim) {
    const double u0 = pu[0];
    if (io.controls_out != nullptr) io.controls_out[inst * io.n_steps_total + io.step_index] = u0;
    double dt = cfg.dt;
    double x[4] = {x0[0], x0[1], x0[2], x0[3]};
    if (isfinite(u0)) {
      while (dt > 0.0) {
        rk4_step<false>(P, x, u0, fmin(dt, 0.001));
        x[1] = mod_pi(x[1]);
        dt -= 0.001;
      }
    }
    double* pw = io.states_io + inst * CPMPC_STATE_STRIDE;
#pragma unroll
    for (int i = 0; i < 4; ++i) pw[i] = x[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Batched plant simulator: Simulator::Step (optimization/simulator.cc:11-36), one thread per plant.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void cpmpc_sim_kernel(const long long B, double* __restrict__ states, const double* __restrict__ dyn,
                                 const double* __restrict__ u, const double dt_total,
                                 const double* __restrict__ f_base, const double* __restrict__ f_mass) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const DynConst P = make_dyn_const(dyn + i * CPMPC_DYN_STRIDE);
  double x[4] = {states[4 * i], wrap_far_angle(states[4 * i + 1]), states[4 * i + 2], states[4 * i + 3]};
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
// The elementary functions of fastmath.cuh as the device evaluates them (cpmpc_debug_fastmath).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void cpmpc_fastmath_kernel(const long long n, const double* __restrict__ x, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double s, c, root, inv_root;
  fm::sincos_bf(v, s, c);
  fm::sqrt_rsqrt(fabs(v), root, inv_root);
  out[6 * i] = s;
  out[6 * i + 1] = c;
  out[6 * i + 2] = fm::tanh_bf(v);
  out[6 * i + 3] = fm::rcp(v);
  out[6 * i + 4] = root;
  out[6 * i + 5] = inv_root;
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
