// ORACLE — TEST INFRASTRUCTURE ONLY.  Second glue file of oracle/_ref/libref.so: entry points that look INSIDE the
// reference's solver, for element-wise parity checks of the device path ("what's missing" 1-3):
//
//   oracle_linearize_qp        the QP of ONE linearisation of the reference's own problem — G, c, A_eq, b_eq exactly as
//                              ConstrainedNonlinearLeastSquares::LinearizeAndFillQP (nonlinear.cc:170-214) fills them
//                              from Optimization::BuildProblem (optimization.cc:194-301) at the initial guess of
//                              Optimization::Step (optimization.cc:46-71) — and the step QPNullSpaceSolver::Solve
//                              (qp.cc:679-729) returns for it.  The shooting blocks Phi / Gamma / e of
//                              CreateDynamicalConstraint (optimization.cc:99-160) are rows of A_eq / b_eq.
//   oracle_closed_loop_trace   the closed loop of optimization_test.cc:38-61 with a per-tick record of what every Step
//                              was given (plant state, warm start) and what it returned, so that the device path can be
//                              "teacher-forced": fed the reference's inputs tick by tick and compared solve by solve.
//   oracle_solver_summary      mini_opt::NLSSolverOutputs::ToString (structs.cc:159-268) on a solver log handed in as
//                              plain arrays (the layout of cpmpc's iteration log), for byte-for-byte comparison with the
//                              summaries of the C++ facade and the pypendulum mirror.
//
// BuildProblem, FillInitialGuess and LinearizeAndFillQP are private members of the reference's classes.  This file reads
// them through the usual test-only access trick (`private` re-defined for the two headers below; the class layouts
// are unchanged, every other translation unit of libref.so is compiled unmodified).  Nothing of the reference is
// copied: the file only calls it.
#include <omp.h>

#include <cmath>
#include <cstring>
#include <exception>
#include <limits>
#include <memory>
#include <string>
#include <vector>

// standard headers first, so that the access trick below cannot touch them
#include <algorithm>
#include <functional>
#include <optional>
#include <ostream>
#include <variant>

#include <fmt/ostream.h>
#include <Eigen/Core>
#include <mini_opt/assertions.hpp>
#include <mini_opt/qp.hpp>
#include <mini_opt/residual.hpp>
#include <mini_opt/structs.hpp>

#define private public
#include <mini_opt/nonlinear.hpp>
#include "optimization.hpp"
#undef private

#include "simulator.hpp"

#include "oracle_capi.h"

namespace {

using pendulum::Optimization;
using pendulum::OptimizationOutputs;
using pendulum::OptimizationParams;
using pendulum::SingleCartPoleParams;
using pendulum::SingleCartPoleState;

SingleCartPoleParams DynFromArray(const double* d) { return {d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8]}; }

OptimizationParams OptFromArray(const double* o) {
  OptimizationParams p{};
  p.control_dt = o[0];
  p.window_length = static_cast<std::size_t>(o[1]);
  p.state_spacing = static_cast<std::size_t>(o[2]);
  p.max_iterations = static_cast<std::size_t>(o[3]);
  p.relative_exit_tol = o[4];
  p.absolute_first_derivative_tol = o[5];
  p.equality_penalty_initial = o[6];
  p.u_guess_sinusoid_amplitude = o[7];
  p.u_cost_weight = o[8];
  p.u_derivative_cost_weight = o[9];
  p.b_x_final_cost_weight = o[10];
  p.th_final_cost_weight = o[11];
  p.b_x_dot_final_cost_weight = o[12];
  p.th_dot_final_cost_weight = o[13];
  return p;
}

// The guess Optimization::Step hands to the solver (optimization.cc:46-71), rebuilt here because Step does not expose
// it: previous solution with x(0) overwritten and the controls shifted left by one, or the sinusoid; shooting states
// re-integrated by the reference's own FillInitialGuess.
Eigen::VectorXd InitialGuess(Optimization& o, const OptimizationParams& p, const SingleCartPoleState& state,
                             const SingleCartPoleParams& dyn) {
  const std::size_t NS = p.NumStates();
  const std::size_t W = p.window_length;
  const int N = o.problem_.dimension;
  Eigen::VectorXd guess = Eigen::VectorXd::Zero(N);
  if (o.previous_solution_.rows() > 0) {
    guess = o.previous_solution_;
    for (std::size_t k = 0; k + 1 < W; ++k) guess[4 * NS + k] = guess[4 * NS + k + 1];
  } else {
    for (std::size_t k = 0; k < W; ++k) {
      guess[4 * NS + k] = p.u_guess_sinusoid_amplitude *
                          std::sin(static_cast<double>(k) / static_cast<double>(W) * 2 * M_PI);
    }
  }
  guess[0] = state.b_x;
  guess[1] = state.th_1;
  guess[2] = state.b_x_dot;
  guess[3] = state.th_1_dot;
  o.FillInitialGuess(guess, dyn, NS);
  return guess;
}

void FillStats(const OptimizationOutputs& out, double initial_penalty, double* stats) {
  const mini_opt::NLSSolverOutputs& so = out.solver_outputs;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  double f = nan, eq = nan, penalty = initial_penalty, lambda = 0.0;
  if (!so.iterations.empty()) {
    const mini_opt::NLSIteration& last = so.iterations.back();
    if (last.step_result == mini_opt::StepSizeSelectionResult::SUCCESS && !last.line_search_steps.empty()) {
      f = last.line_search_steps.back().errors.f;
      eq = last.line_search_steps.back().errors.equality;
    } else {
      f = last.errors_initial.f;
      eq = last.errors_initial.equality;
    }
    penalty = last.penalty;
    lambda = last.lambda;
  }
  stats[0] = static_cast<double>(static_cast<int>(so.termination_state));
  stats[1] = static_cast<double>(so.iterations.size());
  stats[2] = static_cast<double>(so.NumLineSearchSteps());
  stats[3] = f;
  stats[4] = eq;
  stats[5] = penalty;
  stats[6] = lambda;
  stats[7] = static_cast<double>(so.NumFailedLineSearches());
}

}  // namespace

extern "C" {

// Dimensions of the QP for `opt`: N variables, K equality rows.
void oracle_qp_dims(const double* opt, int* N_out, int* K_out) {
  const OptimizationParams p = OptFromArray(opt);
  const int NS = static_cast<int>(p.NumStates());
  int n_te = 0;
  for (const double w : {p.b_x_final_cost_weight, p.th_final_cost_weight, p.b_x_dot_final_cost_weight,
                         p.th_dot_final_cost_weight}) {
    n_te += (w < 0.0) ? 1 : 0;
  }
  if (N_out) *N_out = 4 * NS + static_cast<int>(p.window_length);
  if (K_out) *K_out = 4 * (NS - 1) + 4 + n_te;
}

// One linearisation of the reference's problem at the initial guess of Step(), with LM damping `lambda`, and the
// null-space QP step.  prev: NULL (cold start) or N doubles (warm start, as SetPreviousSolution).
// Outputs (any may be NULL): guess[N], G[N*N] (row-major; only the lower triangle is meaningful), c[N], A[K*N], b[K],
// dx[N], errors[2] = {f, equality} at the guess.  Returns the QPNullSpaceTerminationState (0 = SUCCESS), -1 on throw.
int oracle_linearize_qp(const double* opt, const double* state, const double* dyn, double set_point,
                        const double* prev, double lambda, double* guess_out, double* G_out, double* c_out,
                        double* A_out, double* b_out, double* dx_out, double* errors_out) {
  try {
    const OptimizationParams p = OptFromArray(opt);
    Optimization o(p);
    const int N = 4 * static_cast<int>(p.NumStates()) + static_cast<int>(p.window_length);
    if (prev != nullptr) o.SetPreviousSolution(std::vector<double>(prev, prev + N));
    const SingleCartPoleState st{state[0], state[1], state[2], state[3]};
    const SingleCartPoleParams d = DynFromArray(dyn);
    o.BuildProblem(st, d, set_point);
    const Eigen::VectorXd guess = InitialGuess(o, p, st, d);
    int K = 0;
    for (const mini_opt::Residual& eq : o.problem_.equality_constraints) K += eq.Dimension();
    mini_opt::QP qp;
    qp.G.resize(N, N);
    qp.c.resize(N);
    qp.A_eq.resize(K, N);
    qp.b_eq.resize(K);
    const mini_opt::Errors err =
        mini_opt::ConstrainedNonlinearLeastSquares::LinearizeAndFillQP(guess, lambda, o.problem_, &qp);
    mini_opt::QPNullSpaceSolver solver;
    const mini_opt::QPNullSpaceTerminationState t = solver.Solve(qp);
    for (int i = 0; i < N; ++i) {
      if (guess_out) guess_out[i] = guess[i];
      if (c_out) c_out[i] = qp.c[i];
      if (G_out) {
        for (int j = 0; j < N; ++j) G_out[i * N + j] = qp.G(i, j);
      }
      if (dx_out) {
        dx_out[i] = (t == mini_opt::QPNullSpaceTerminationState::SUCCESS)
                        ? solver.variables()[i]
                        : std::numeric_limits<double>::quiet_NaN();
      }
    }
    for (int r = 0; r < K; ++r) {
      if (b_out) b_out[r] = qp.b_eq[r];
      if (A_out) {
        for (int j = 0; j < N; ++j) A_out[r * N + j] = qp.A_eq(r, j);
      }
    }
    if (errors_out) {
      errors_out[0] = err.f;
      errors_out[1] = err.equality;
    }
    return static_cast<int>(t);
  } catch (const std::exception&) {
    return -1;
  }
}

// Closed loop of optimization_test.cc:38-61 for B plants with a full per-tick trace (T = n_steps):
//   state_trace [B*T*4]  plant state handed to Step at tick t
//   prev_trace  [B*T*N]  OptimizationOutputs::previous_solution of tick t = the warm start that Step used (zeros at t=0)
//   u_trace     [B*T*W]  OptimizationOutputs::u
//   stats_trace [B*T*8]  as oracle_batch_step
// Any trace may be NULL.  Returns threads used.
int oracle_closed_loop_trace(const double* opt, int64_t B, int64_t n_steps, const double* states0, const double* dyn,
                             const double* set_points, double* state_trace, double* prev_trace, double* u_trace,
                             double* stats_trace, int n_threads) {
  const OptimizationParams p = OptFromArray(opt);
  const std::size_t W = p.window_length;
  const std::size_t N = p.NumStates() * 4 + W;
  const int threads = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int64_t i = 0; i < B; ++i) {
    const SingleCartPoleParams d = DynFromArray(dyn + 9 * i);
    Optimization o(p);
    pendulum::Simulator sim;
    const double* s = states0 + 4 * i;
    sim.SetState({s[0], s[1], s[2], s[3]});
    const pendulum::Vector2 zero{0.0, 0.0};
    for (int64_t t = 0; t < n_steps; ++t) {
      const int64_t rec = i * n_steps + t;
      const SingleCartPoleState x = sim.GetState();
      if (state_trace) {
        state_trace[4 * rec + 0] = x.b_x;
        state_trace[4 * rec + 1] = x.th_1;
        state_trace[4 * rec + 2] = x.b_x_dot;
        state_trace[4 * rec + 3] = x.th_1_dot;
      }
      double u0 = 0.0;
      try {
        const OptimizationOutputs out = o.Step(x, d, set_points[i]);
        if (prev_trace) {
          double* dst = prev_trace + rec * N;
          if (out.previous_solution.size() == N) {
            std::memcpy(dst, out.previous_solution.data(), sizeof(double) * N);
          } else {
            std::memset(dst, 0, sizeof(double) * N);
          }
        }
        if (u_trace) std::memcpy(u_trace + rec * W, out.u.data(), sizeof(double) * W);
        if (stats_trace) FillStats(out, p.equality_penalty_initial, stats_trace + rec * ORACLE_STATS_STRIDE);
        u0 = out.u.front();
      } catch (const std::exception&) {
        if (stats_trace) {
          for (int k = 0; k < ORACLE_STATS_STRIDE; ++k) stats_trace[rec * ORACLE_STATS_STRIDE + k] = -1.0;
        }
      }
      sim.Step(d, p.control_dt, u0, zero, zero);
    }
  }
  return threads;
}

// mini_opt::NLSSolverOutputs::ToString(use_color = false, include_qp = false) (structs.cc:256-266) for a solver log
// given as arrays: `iter_log` holds n_iterations records of ORACLE_ITERLOG_STRIDE doubles (optimizer_state, lambda,
// f(0), c(0), d_f, d_equality, penalty, step_result, n_steps, then 6 x (alpha, f, equality)).  Writes at most
// `capacity` bytes (NUL-terminated when it fits) and returns the length of the full string.
int64_t oracle_solver_summary(int termination_state, int n_iterations, const double* iter_log, char* out,
                              int64_t capacity) {
  std::vector<mini_opt::NLSIteration> iterations;
  iterations.reserve(n_iterations > 0 ? n_iterations : 0);
  for (int i = 0; i < n_iterations; ++i) {
    const double* r = iter_log + static_cast<std::size_t>(i) * ORACLE_ITERLOG_STRIDE;
    std::vector<mini_opt::LineSearchStep> steps;
    const int n_steps = static_cast<int>(r[8]);
    for (int s = 0; s < n_steps && s < 6; ++s) {
      mini_opt::Errors e{};
      e.f = r[10 + 3 * s];
      e.equality = r[11 + 3 * s];
      steps.push_back(mini_opt::LineSearchStep{r[9 + 3 * s], e});
    }
    mini_opt::Errors e0{};
    e0.f = r[2];
    e0.equality = r[3];
    mini_opt::DirectionalDerivatives dd{};
    dd.d_f = r[4];
    dd.d_equality = r[5];
    iterations.emplace_back(i, static_cast<mini_opt::OptimizerState>(static_cast<int>(r[0])), r[1], e0,
                            mini_opt::QPNullSpaceTerminationState::SUCCESS, std::optional<mini_opt::QPEigenvalues>{}, dd,
                            r[6], static_cast<mini_opt::StepSizeSelectionResult>(static_cast<int>(r[7])),
                            std::move(steps));
  }
  const mini_opt::NLSSolverOutputs outputs(static_cast<mini_opt::NLSTerminationState>(termination_state),
                                           std::move(iterations));
  const std::string text = outputs.ToString(false, false);
  if (out != nullptr && capacity > 0) {
    const std::size_t n = std::min<std::size_t>(text.size(), static_cast<std::size_t>(capacity) - 1);
    std::memcpy(out, text.data(), n);
    out[n] = '\0';
  }
  return static_cast<int64_t>(text.size());
}

}  // extern "C"
