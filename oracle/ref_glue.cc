// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the REFERENCE's own translation units
// (oracle/_ref/libref.so).  Built only by `make -C oracle ref` when /root/reference is present; loaded only by
// tests/, tests/golden/make_golden_ref.py and bench.py's CPU legs.  The product never links or loads it.
//
// This file is original glue: it includes the reference's headers where they lie and calls
//   pendulum::Optimization        (optimization/optimization.hpp:73-108)
//   pendulum::Simulator           (optimization/simulator.hpp:10-29)
//   gen::single_pendulum_dynamics (optimization/single_pendulum_dynamics.hpp:13-186)
//   pendulum::runge_kutta_4th_order[_no_jacobians], pendulum::mod_pi (optimization/integration.hpp:13-73)
//   mini_opt::QPNullSpaceSolver   (dependencies/mini_opt/include/mini_opt/qp.hpp:303-319)
// through the same C signatures as oracle_capi.h, so oracle_py can bind either library with one set of
// prototypes.  Eigen is supplied by oracle/eigen_shim (see its header: interface shim, not Eigen).
//
// Differences from liboracle.so that follow from the reference's API (not from its arithmetic):
//  * final (f, equality) are not returned by Optimization::Step; they are read from the solver log: the last
//    logged iteration's accepted line-search step when it succeeded, otherwise its errors_initial (the
//    variables did not move).  Both are the values EvaluateNonlinearErrors(variables()) would return.
//  * the solution vector of the current call is private (previous_solution_ is returned one call late,
//    optimization.cc:84,94), so sol_out carries NaN in its shooting-state part and u in its control part.
//  * the polynomial line-search helpers are private statics (nonlinear.hpp:206-217): not exported.
#include <omp.h>

#include <cmath>
#include <cstring>
#include <exception>
#include <limits>
#include <vector>

#include "integration.hpp"
#include "optimization.hpp"
#include "simulator.hpp"
#include "single_pendulum_dynamics.hpp"

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

void FillStats(const OptimizationOutputs& out, double initial_penalty, double* stats) {
  if (!stats) return;
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

void FillIterLog(const OptimizationOutputs& out, std::size_t max_iterations, double* log) {
  if (!log) return;
  std::memset(log, 0, sizeof(double) * max_iterations * ORACLE_ITERLOG_STRIDE);
  const auto& its = out.solver_outputs.iterations;
  for (std::size_t i = 0; i < its.size() && i < max_iterations; ++i) {
    double* r = log + i * ORACLE_ITERLOG_STRIDE;
    r[0] = static_cast<double>(static_cast<int>(its[i].optimizer_state));
    r[1] = its[i].lambda;
    r[2] = its[i].errors_initial.f;
    r[3] = its[i].errors_initial.equality;
    r[4] = its[i].directional_derivatives.d_f;
    r[5] = its[i].directional_derivatives.d_equality;
    r[6] = its[i].penalty;
    r[7] = static_cast<double>(static_cast<int>(its[i].step_result));
    r[8] = static_cast<double>(its[i].line_search_steps.size());
    for (std::size_t s = 0; s < its[i].line_search_steps.size() && s < 6; ++s) {
      r[9 + 3 * s] = its[i].line_search_steps[s].alpha;
      r[10 + 3 * s] = its[i].line_search_steps[s].errors.f;
      r[11 + 3 * s] = its[i].line_search_steps[s].errors.equality;
    }
  }
}

void WriteOutputs(const OptimizationOutputs& out, std::size_t W, double* u_out, double* pred_out) {
  if (u_out) std::memcpy(u_out, out.u.data(), sizeof(double) * W);
  if (pred_out) {
    for (std::size_t k = 0; k < W; ++k) {
      pred_out[4 * k + 0] = out.predicted_states[k].b_x;
      pred_out[4 * k + 1] = out.predicted_states[k].th_1;
      pred_out[4 * k + 2] = out.predicted_states[k].b_x_dot;
      pred_out[4 * k + 3] = out.predicted_states[k].th_1_dot;
    }
  }
}

struct OptHandle {
  OptimizationParams params;
  Optimization opt;
  std::vector<double> last_solution;  // what the next Step() will report as previous_solution
  explicit OptHandle(const OptimizationParams& p) : params(p), opt(p) {}
};

Eigen::Vector4d Vec4(const double* x) { return Eigen::Vector4d(x[0], x[1], x[2], x[3]); }

}  // namespace

extern "C" {

const char* oracle_ref_info(void) {
  return "reference translation units (optimization.cc, simulator.cc, mini_opt nonlinear.cc/qp.cc/residual.cc/"
         "structs.cc, single_pendulum_dynamics.hpp) compiled unmodified against oracle/eigen_shim";
}

void oracle_dynamics(const double* dyn, const double* x, double u, const double* f_base, const double* f_mass,
                     double* x_dot, double* J_x, double* J_u) {
  const Eigen::Vector4d xv = Vec4(x);
  const Eigen::Vector2d fb(f_base[0], f_base[1]);
  const Eigen::Vector2d fm(f_mass[0], f_mass[1]);
  Eigen::Vector4d xd;
  if (J_x && J_u) {
    Eigen::Matrix<double, 4, 4> Jx;
    Eigen::Matrix<double, 4, 1> Ju;
    xd = gen::single_pendulum_dynamics(DynFromArray(dyn), xv, u, fb, fm, Jx, Ju);
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < 4; ++j) J_x[4 * i + j] = Jx(i, j);
      J_u[i] = Ju[i];
    }
  } else {
    xd = gen::single_pendulum_dynamics(DynFromArray(dyn), xv, u, fb, fm, nullptr, nullptr);
  }
  for (int i = 0; i < 4; ++i) x_dot[i] = xd[i];
}

double oracle_mod_pi(double angle) { return pendulum::mod_pi(angle); }

void oracle_rk4(const double* dyn, const double* x, double u, double h, const double* f_base, const double* f_mass,
                double* x_new) {
  const SingleCartPoleParams params = DynFromArray(dyn);
  const Eigen::Vector2d fb(f_base[0], f_base[1]);
  const Eigen::Vector2d fm(f_mass[0], f_mass[1]);
  const Eigen::Vector4d out = pendulum::runge_kutta_4th_order_no_jacobians<4>(
      Vec4(x), u, h, [&](const Eigen::Matrix<double, 4, 1>& x_updated, const double u_in) {
        return gen::single_pendulum_dynamics(params, x_updated, u_in, fb, fm, nullptr, nullptr);
      });
  for (int i = 0; i < 4; ++i) x_new[i] = out[i];
}

void oracle_rk4_jac(const double* dyn, const double* x, double u, double h, double* x_new, double* D_x,
                    double* D_u) {
  const SingleCartPoleParams params = DynFromArray(dyn);
  const auto [xn, Dx, Du] = pendulum::runge_kutta_4th_order<4>(
      Vec4(x), u, h, [&params](const auto& x_updated, const double u_in, auto& x_dot_D_x, auto& x_dot_D_u) {
        const Eigen::Vector2d zero = Eigen::Vector2d::Zero();
        return gen::single_pendulum_dynamics(params, x_updated, u_in, zero, zero, x_dot_D_x, x_dot_D_u);
      });
  for (int i = 0; i < 4; ++i) {
    x_new[i] = xn[i];
    D_u[i] = Du[i];
    for (int j = 0; j < 4; ++j) D_x[4 * i + j] = Dx(i, j);
  }
}

void oracle_sim_step(const double* dyn, double* state, double dt, double u, const double* f_base,
                     const double* f_mass) {
  pendulum::Simulator sim;
  sim.SetState({state[0], state[1], state[2], state[3]});
  sim.Step(DynFromArray(dyn), dt, u, pendulum::Vector2{f_base[0], f_base[1]},
           pendulum::Vector2{f_mass[0], f_mass[1]});
  const SingleCartPoleState s = sim.GetState();
  state[0] = s.b_x;
  state[1] = s.th_1;
  state[2] = s.b_x_dot;
  state[3] = s.th_1_dot;
}

int oracle_qp_nullspace(int N, int K, const double* G, const double* c, const double* A, const double* b,
                        double* x_out) {
  mini_opt::QP qp(N);
  qp.A_eq.resize(K, N);
  qp.b_eq.resize(K);
  for (int i = 0; i < N; ++i) {
    qp.c[i] = c[i];
    for (int j = 0; j < N; ++j) qp.G(i, j) = G[i * N + j];
  }
  for (int i = 0; i < K; ++i) {
    qp.b_eq[i] = b[i];
    for (int j = 0; j < N; ++j) qp.A_eq(i, j) = A[i * N + j];
  }
  mini_opt::QPNullSpaceSolver solver;
  const mini_opt::QPNullSpaceTerminationState t = solver.Solve(qp);
  if (t == mini_opt::QPNullSpaceTerminationState::SUCCESS) {
    for (int i = 0; i < N; ++i) x_out[i] = solver.variables()[i];
  }
  return static_cast<int>(t);
}

void* oracle_opt_create(const double* opt) {
  try {
    return new OptHandle(OptFromArray(opt));
  } catch (const std::exception&) {
    return nullptr;
  }
}

void oracle_opt_destroy(void* handle) { delete static_cast<OptHandle*>(handle); }

void oracle_opt_reset(void* handle) { static_cast<OptHandle*>(handle)->opt.Reset(); }

void oracle_opt_set_previous_solution(void* handle, const double* sol, int64_t n) {
  static_cast<OptHandle*>(handle)->opt.SetPreviousSolution(std::vector<double>(sol, sol + n));
}

int64_t oracle_opt_get_previous_solution(void*, double*, int64_t) {
  return 0;  // private in the reference (optimization.hpp:107); see the header comment
}

int oracle_opt_step(void* handle, const double* state, const double* dyn, double set_point, double* u_out,
                    double* pred_out, double* stats, double* iter_log) {
  OptHandle* h = static_cast<OptHandle*>(handle);
  try {
    const OptimizationOutputs out =
        h->opt.Step({state[0], state[1], state[2], state[3]}, DynFromArray(dyn), set_point);
    WriteOutputs(out, h->params.window_length, u_out, pred_out);
    FillStats(out, h->params.equality_penalty_initial, stats);
    FillIterLog(out, h->params.max_iterations, iter_log);
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

int oracle_batch_step(const double* opt, int64_t B, const double* states, const double* dyn,
                      const double* set_points, const double* weights, const double* prev,
                      const uint8_t* has_prev, double* u_out, double* pred_out, double* sol_out, double* stats,
                      double* iter_log, int n_threads) {
  const OptimizationParams base = OptFromArray(opt);
  const std::size_t W = base.window_length;
  const std::size_t N = base.NumStates() * 4 + W;
  const int threads = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads)
  for (int64_t i = 0; i < B; ++i) {
    OptimizationParams p = base;
    if (weights) {
      const double* w = weights + i * ORACLE_WEIGHTS_STRIDE;
      p.u_cost_weight = w[0];
      p.u_derivative_cost_weight = w[1];
      p.b_x_final_cost_weight = w[2];
      p.th_final_cost_weight = w[3];
      p.b_x_dot_final_cost_weight = w[4];
      p.th_dot_final_cost_weight = w[5];
    }
    try {
      Optimization o(p);
      if (prev && has_prev && has_prev[i]) {
        o.SetPreviousSolution(std::vector<double>(prev + i * N, prev + (i + 1) * N));
      }
      const double* s = states + 4 * i;
      const OptimizationOutputs out = o.Step({s[0], s[1], s[2], s[3]}, DynFromArray(dyn + 9 * i), set_points[i]);
      WriteOutputs(out, W, u_out ? u_out + i * W : nullptr, pred_out ? pred_out + i * W * 4 : nullptr);
      if (sol_out) {
        double* sol = sol_out + i * N;
        for (std::size_t k = 0; k < N - W; ++k) sol[k] = std::numeric_limits<double>::quiet_NaN();
        std::memcpy(sol + (N - W), out.u.data(), sizeof(double) * W);
      }
      FillStats(out, p.equality_penalty_initial, stats ? stats + i * ORACLE_STATS_STRIDE : nullptr);
      FillIterLog(out, p.max_iterations, iter_log ? iter_log + i * p.max_iterations * ORACLE_ITERLOG_STRIDE : nullptr);
    } catch (const std::exception&) {
      if (stats) {
        for (int k = 0; k < ORACLE_STATS_STRIDE; ++k) stats[i * ORACLE_STATS_STRIDE + k] = -1.0;
      }
    }
  }
  return threads;
}

int oracle_closed_loop(const double* opt, int64_t B, int64_t n_steps, double* states, const double* dyn,
                       const double* set_points, double* controls_out, double* last_pred_terminal_out,
                       int64_t* counts_out, int64_t* term_hist_out, int n_threads) {
  const OptimizationParams p = OptFromArray(opt);
  const int threads = n_threads > 0 ? n_threads : omp_get_max_threads();
  int64_t solves = 0, iters = 0, ls = 0, bad = 0;
  int64_t hist[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : solves, iters, ls, bad, hist[:7])
  for (int64_t i = 0; i < B; ++i) {
    const SingleCartPoleParams d = DynFromArray(dyn + 9 * i);
    Optimization o(p);
    pendulum::Simulator sim;
    double* s = states + 4 * i;
    sim.SetState({s[0], s[1], s[2], s[3]});
    const pendulum::Vector2 zero{0.0, 0.0};
    for (int64_t t = 0; t < n_steps; ++t) {
      const OptimizationOutputs out = o.Step(sim.GetState(), d, set_points[i]);
      ++solves;
      iters += static_cast<int64_t>(out.solver_outputs.iterations.size());
      ls += static_cast<int64_t>(out.solver_outputs.NumLineSearchSteps());
      const int term = static_cast<int>(out.solver_outputs.termination_state);
      hist[term] += 1;
      if (out.solver_outputs.termination_state == mini_opt::NLSTerminationState::QP_INDEFINITE ||
          out.solver_outputs.termination_state == mini_opt::NLSTerminationState::MAX_LAMBDA) {
        ++bad;
      }
      if (controls_out) controls_out[i * n_steps + t] = out.u.front();
      if (last_pred_terminal_out && t + 1 == n_steps) {
        const SingleCartPoleState& ts = out.predicted_states.back();
        last_pred_terminal_out[4 * i + 0] = ts.b_x;
        last_pred_terminal_out[4 * i + 1] = ts.th_1;
        last_pred_terminal_out[4 * i + 2] = ts.b_x_dot;
        last_pred_terminal_out[4 * i + 3] = ts.th_1_dot;
      }
      sim.Step(d, p.control_dt, out.u.front(), zero, zero);
    }
    const SingleCartPoleState f = sim.GetState();
    s[0] = f.b_x;
    s[1] = f.th_1;
    s[2] = f.b_x_dot;
    s[3] = f.th_1_dot;
  }
  if (counts_out) {
    counts_out[0] += solves;
    counts_out[1] += iters;
    counts_out[2] += ls;
    counts_out[3] += bad;
  }
  if (term_hist_out) {
    for (int k = 0; k < 7; ++k) term_hist_out[k] += hist[k];
  }
  return threads;
}

}  // extern "C"
