// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_capi.h).
#include "oracle_capi.h"

#include <omp.h>

#include <cstring>
#include <exception>

#include "cartpole.hpp"
#include "double_cartpole.hpp"
#include "double_pendulum_gen.hpp"

using namespace oracle;

namespace {

SingleCartPoleParams DynFromArray(const double* d) { return {d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8]}; }

OptimizationParams OptFromArray(const double* o) {
  OptimizationParams p;
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

template <class Outputs>
void FillStats(const Outputs& out, double* stats) {
  if (!stats) return;
  const auto& so = out.solver_outputs;
  stats[0] = static_cast<double>(static_cast<int>(so.termination_state));
  stats[1] = static_cast<double>(so.iterations.size());
  stats[2] = static_cast<double>(so.NumLineSearchSteps());
  stats[3] = out.final_errors.f;
  stats[4] = out.final_errors.equality;
  stats[5] = out.final_penalty;
  stats[6] = so.iterations.empty() ? 0.0 : so.iterations.back().lambda;
  stats[7] = static_cast<double>(so.NumFailedLineSearches());
}

template <class Outputs>
void FillIterLog(const Outputs& out, std::size_t max_iterations, double* log) {
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
  explicit OptHandle(const OptimizationParams& p) : params(p), opt(p) {}
};

}  // namespace

extern "C" {

void oracle_dynamics(const double* dyn, const double* x, double u, const double* f_base, const double* f_mass,
                     double* x_dot, double* J_x, double* J_u) {
  SinglePendulumDynamics(DynFromArray(dyn), x, u, f_base, f_mass, x_dot, J_x, J_u);
}

double oracle_mod_pi(double angle) { return ModPi(angle); }

void oracle_rk4(const double* dyn, const double* x, double u, double h, const double* f_base, const double* f_mass,
                double* x_new) {
  RungeKutta4NoJacobians(DynFromArray(dyn), x, u, h, f_base, f_mass, x_new);
}

void oracle_rk4_jac(const double* dyn, const double* x, double u, double h, double* x_new, double* D_x,
                    double* D_u) {
  RungeKutta4(DynFromArray(dyn), x, u, h, x_new, D_x, D_u);
}

void oracle_sim_step(const double* dyn, double* state, double dt, double u, const double* f_base,
                     const double* f_mass) {
  Simulator sim;
  sim.SetState({state[0], state[1], state[2], state[3]});
  sim.Step(DynFromArray(dyn), dt, u, f_base, f_mass);
  const SingleCartPoleState s = sim.GetState();
  state[0] = s.b_x;
  state[1] = s.th_1;
  state[2] = s.b_x_dot;
  state[3] = s.th_1_dot;
}

int oracle_qp_nullspace(int N, int K, const double* G, const double* c, const double* A, const double* b,
                        double* x_out) {
  QP qp;
  qp.G.Resize(N, N);
  qp.A_eq.Resize(K, N);
  qp.c.assign(c, c + N);
  qp.b_eq.assign(b, b + K);
  std::memcpy(qp.G.a.data(), G, sizeof(double) * N * N);
  std::memcpy(qp.A_eq.a.data(), A, sizeof(double) * K * N);
  QPNullSpaceSolver solver;
  const QPNullSpaceTerminationState t = solver.Solve(qp);
  if (t == QPNullSpaceTerminationState::SUCCESS) std::memcpy(x_out, solver.variables().data(), sizeof(double) * N);
  return static_cast<int>(t);
}

int oracle_quadratic_approx_minimum(double phi_0, double phi_prime_0, double alpha_0, double phi_alpha_0,
                                    double* out) {
  const auto v = ConstrainedNonlinearLeastSquares::QuadraticApproxMinimum(phi_0, phi_prime_0, alpha_0, phi_alpha_0);
  if (v) *out = *v;
  return v.has_value() ? 1 : 0;
}

void oracle_cubic_approx_coeffs(double phi_0, double phi_prime_0, double alpha_0, double phi_alpha_0,
                                double alpha_1, double phi_alpha_1, double* ab_out) {
  ConstrainedNonlinearLeastSquares::CubicApproxCoeffs(phi_0, phi_prime_0, alpha_0, phi_alpha_0, alpha_1,
                                                      phi_alpha_1, ab_out);
}

int oracle_cubic_approx_minimum(double phi_prime_0, const double* ab, double* out) {
  const auto v = ConstrainedNonlinearLeastSquares::CubicApproxMinimum(phi_prime_0, ab);
  if (v) *out = *v;
  return v.has_value() ? 1 : 0;
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

int64_t oracle_opt_get_previous_solution(void* handle, double* out, int64_t capacity) {
  const auto& p = static_cast<OptHandle*>(handle)->opt.previous_solution();
  const int64_t n = static_cast<int64_t>(p.size());
  if (n > capacity) return -n;
  std::memcpy(out, p.data(), sizeof(double) * n);
  return n;
}

int oracle_opt_step(void* handle, const double* state, const double* dyn, double set_point, double* u_out,
                    double* pred_out, double* stats, double* iter_log) {
  OptHandle* h = static_cast<OptHandle*>(handle);
  try {
    const OptimizationOutputs out =
        h->opt.Step({state[0], state[1], state[2], state[3]}, DynFromArray(dyn), set_point);
    WriteOutputs(out, h->params.window_length, u_out, pred_out);
    FillStats(out, stats);
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
      if (sol_out) std::memcpy(sol_out + i * N, o.previous_solution().data(), sizeof(double) * N);
      FillStats(out, stats ? stats + i * ORACLE_STATS_STRIDE : nullptr);
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
    Simulator sim;
    double* s = states + 4 * i;
    sim.SetState({s[0], s[1], s[2], s[3]});
    const double zero[2] = {0.0, 0.0};
    for (int64_t t = 0; t < n_steps; ++t) {
      const OptimizationOutputs out = o.Step(sim.GetState(), d, set_points[i]);
      ++solves;
      iters += static_cast<int64_t>(out.solver_outputs.iterations.size());
      ls += static_cast<int64_t>(out.solver_outputs.NumLineSearchSteps());
      const int term = static_cast<int>(out.solver_outputs.termination_state);
      hist[term] += 1;
      if (out.solver_outputs.termination_state == NLSTerminationState::QP_INDEFINITE ||
          out.solver_outputs.termination_state == NLSTerminationState::MAX_LAMBDA) {
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

// ---- double pendulum on a cart (double_cartpole.hpp) ---------------------------------------------------------------

void oracle_dp_dynamics(const double* dyn, const double* x, double u, double* x_dot, double* J_x, double* J_u) {
  DoublePendulumDynamics(dyn, x, u, x_dot, J_x, J_u);
}

void oracle_dp_rk4(const double* dyn, const double* x, double u, double h, double* x_new) {
  DoubleRungeKutta4NoJacobians({dyn[0], dyn[1], dyn[2], dyn[3], dyn[4], dyn[5]}, x, u, h, x_new);
}

void oracle_dp_rk4_jac(const double* dyn, const double* x, double u, double h, double* x_new, double* D_x,
                       double* D_u) {
  DoubleRungeKutta4({dyn[0], dyn[1], dyn[2], dyn[3], dyn[4], dyn[5]}, x, u, h, x_new, D_x, D_u);
}

int oracle_dp_batch_step(const double* opt, int64_t B, const double* states, const double* dyn,
                         const double* set_points, const double* weights, const double* prev,
                         const uint8_t* has_prev, double* u_out, double* pred_out, double* sol_out, double* stats,
                         double* iter_log, int n_threads) {
  DoubleOptimizationParams base;
  base.control_dt = opt[0];
  base.window_length = static_cast<std::size_t>(opt[1]);
  base.state_spacing = static_cast<std::size_t>(opt[2]);
  base.max_iterations = static_cast<std::size_t>(opt[3]);
  base.relative_exit_tol = opt[4];
  base.absolute_first_derivative_tol = opt[5];
  base.equality_penalty_initial = opt[6];
  base.u_guess_sinusoid_amplitude = opt[7];
  base.u_cost_weight = opt[8];
  base.u_derivative_cost_weight = opt[9];
  for (int i = 0; i < 6; ++i) base.final_cost_weight[i] = opt[10 + i];
  const std::size_t W = base.window_length;
  const std::size_t N = base.NumStates() * 6 + W;
  const int threads = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads)
  for (int64_t i = 0; i < B; ++i) {
    DoubleOptimizationParams p = base;
    if (weights) {
      const double* w = weights + i * ORACLE_DP_WEIGHTS_STRIDE;
      p.u_cost_weight = w[0];
      p.u_derivative_cost_weight = w[1];
      for (int k = 0; k < 6; ++k) p.final_cost_weight[k] = w[2 + k];
    }
    try {
      DoubleOptimization o(p);
      if (prev && has_prev && has_prev[i]) {
        o.SetPreviousSolution(std::vector<double>(prev + i * N, prev + (i + 1) * N));
      }
      const double* d = dyn + 6 * i;
      const DoubleOptimizationOutputs out =
          o.Step(states + 6 * i, {d[0], d[1], d[2], d[3], d[4], d[5]}, set_points[i]);
      if (u_out) std::memcpy(u_out + i * W, out.u.data(), sizeof(double) * W);
      if (pred_out) std::memcpy(pred_out + i * W * 6, out.predicted_states.data(), sizeof(double) * W * 6);
      if (sol_out) std::memcpy(sol_out + i * N, o.previous_solution().data(), sizeof(double) * N);
      FillStats(out, stats ? stats + i * ORACLE_STATS_STRIDE : nullptr);
      FillIterLog(out, p.max_iterations, iter_log ? iter_log + i * p.max_iterations * ORACLE_ITERLOG_STRIDE : nullptr);
    } catch (const std::exception&) {
      if (stats) {
        for (int k = 0; k < ORACLE_STATS_STRIDE; ++k) stats[i * ORACLE_STATS_STRIDE + k] = -1.0;
      }
    }
  }
  return threads;
}

}  // extern "C"
