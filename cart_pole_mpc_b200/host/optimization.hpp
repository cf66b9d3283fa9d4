// C++ host facade over the C ABI (include/cpmpc.h): keeps the reference's controller API
// (optimization/optimization.hpp:12-108, optimization/simulator.hpp:10-29) so that wrapper/wrapper.cc,
// optimization_test.cc and any other caller of pendulum::Optimization / pendulum::Simulator compile against it, and
// adds pendulum::BatchedOptimization for B independent controllers on one GPU.
//
// Header-only; link with -lcpmpc.  Errors: precondition failures and CUDA errors are re-thrown as std::exception
// subclasses (the reference throws assert::default_error : std::exception, mini_opt/assertions.hpp:49-73); numerical
// outcomes stay in-band in the per-instance statistics.
#pragma once
#include <charconv>
#include <cstddef>
#include <optional>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cpmpc.h"
#include "structs.hpp"

namespace mini_opt {

// mini_opt/structs.hpp:232-248 (declaration order == numeric value used by the C ABI)
enum class NLSTerminationState : int {
  MAX_ITERATIONS = 0,
  SATISFIED_ABSOLUTE_TOL,
  SATISFIED_RELATIVE_TOL,
  SATISFIED_FIRST_ORDER_TOL,
  MAX_LAMBDA,
  QP_INDEFINITE,
  USER_CALLBACK,
};

// mini_opt/structs.hpp:215-228
enum class StepSizeSelectionResult : int {
  SUCCESS = 0,
  MAX_ITERATIONS,
  FIRST_ORDER_SATISFIED,
  POSITIVE_DERIVATIVE,
  FAILURE_NON_FINITE_COST,
  FAILURE_INVALID_ALPHA,
};

// mini_opt/structs.hpp:159-164
enum class OptimizerState : int { NOMINAL = 0, ATTEMPTING_RESTORE_LM = 1 };

// mini_opt/structs.hpp:147-155.  The equality-only problem of Optimization always takes the null-space QP
// (nonlinear.cc:83-89); NOT_POSITIVE_DEFINITE ends the solve before the iteration is logged (nonlinear.cc:103-105),
// so every logged iteration carries SUCCESS.
enum class QPNullSpaceTerminationState : int { SUCCESS = 0, NOT_POSITIVE_DEFINITE };

// mini_opt/structs.hpp:262-274: only filled when Params::log_qp_eigenvalues is set, which Optimization never does
// (optimization.cc:73-79); kept so that code reading NLSIteration::qp_eigenvalues compiles.
struct QPEigenvalues {
  double min;
  double max;
  double abs_min;
};

// Names as printed by the reference's ostream operators (mini_opt/source/structs.cc:17-120).
inline const char* ToName(NLSTerminationState s) {
  static const char* const names[] = {"MAX_ITERATIONS",  "SATISFIED_ABSOLUTE_TOL", "SATISFIED_RELATIVE_TOL",
                                      "SATISFIED_FIRST_ORDER_TOL", "MAX_LAMBDA", "QP_INDEFINITE", "USER_CALLBACK"};
  const int i = static_cast<int>(s);
  return (i >= 0 && i < 7) ? names[i] : "REJECTED_NON_FINITE";
}
inline const char* ToName(StepSizeSelectionResult r) {
  static const char* const names[] = {"SUCCESS", "MAX_ITERATIONS", "FIRST_ORDER_SATISFIED", "POSITIVE_DERIVATIVE",
                                      "FAILURE_NON_FINITE_COST", "FAILURE_INVALID_ALPHA"};
  const int i = static_cast<int>(r);
  return (i >= 0 && i < 6) ? names[i] : "?";
}
inline const char* ToName(OptimizerState s) {
  return s == OptimizerState::NOMINAL ? "NOMINAL" : "ATTEMPTING_RESTORE_LM";
}

namespace detail {
// fmt's "{}" for a double: the shortest decimal string that round-trips
inline std::string Shortest(double v) {
  char buf[40];
  const auto r = std::to_chars(buf, buf + sizeof(buf), v);
  return std::string(buf, r.ptr);
}
inline std::string Printf(const char* format, double v) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), format, v);
  return buf;
}
}  // namespace detail

// mini_opt/structs.hpp:169-212
struct Errors {
  double f{0};
  double equality{0};
  constexpr double Total(double penalty) const noexcept { return f + penalty * equality; }
};
struct DirectionalDerivatives {
  double d_f;
  double d_equality;
};
struct LineSearchStep {
  double alpha;
  Errors errors;
};

// mini_opt/structs.hpp:277-322 (the interior-point members do not exist on this path)
struct NLSIteration {
  int iteration;
  OptimizerState optimizer_state;
  double lambda;
  Errors errors_initial;
  QPNullSpaceTerminationState qp_outputs{QPNullSpaceTerminationState::SUCCESS};
  std::optional<QPEigenvalues> qp_eigenvalues{};
  DirectionalDerivatives directional_derivatives;
  double penalty;
  StepSizeSelectionResult step_result;
  std::vector<LineSearchStep> line_search_steps;

  // Same lines as NLSIteration::ToString (mini_opt/source/structs.cc:159-218) without colour codes; the QP line is
  // always the null-space solver's SUCCESS (an indefinite QP ends the solve before the iteration is logged).
  std::string ToString() const {
    using detail::Printf;
    using detail::Shortest;
    std::string r = "Iteration # " + std::to_string(iteration) + ", state = " + ToName(optimizer_state) +
                    ", lambda = " + Shortest(lambda) + "\n";
    r += "  f(0): " + Printf("%.16e", errors_initial.f) + ", c(0): " + Printf("%.16e", errors_initial.equality) +
         ", total: " + Printf("%.16e", errors_initial.Total(penalty)) + "\n";
    r += "  penalty = " + Printf("%.16f", penalty) + "\n";
    r += "  QP: SUCCESS\n";
    r += "  df/dalpha = " + Shortest(directional_derivatives.d_f) + ", dc/dalpha = " +
         Shortest(directional_derivatives.d_equality) + "\n";
    r += std::string("  Search result: ") + ToName(step_result) + "\n";
    std::size_t i = 0;
    for (const LineSearchStep& step : line_search_steps) {
      const std::string n = std::to_string(i++);
      r += "  f(" + n + "): " + Printf("%.16e", step.errors.f) + ", c(" + n + "): " +
           Printf("%.16e", step.errors.equality) + ", total: " + Printf("%.16e", step.errors.Total(penalty)) +
           ", alpha = " + Printf("%.10f", step.alpha) + "\n";
    }
    return r;
  }
};

// mini_opt/structs.hpp:329-347
struct NLSSolverOutputs {
  NLSTerminationState termination_state{NLSTerminationState::MAX_ITERATIONS};
  std::vector<NLSIteration> iterations;
  Errors final_errors;       // Errors at the returned variables (extra)
  double final_penalty{0};   // extra
  // mini_opt/source/structs.cc:220-231: one QP solve per iteration on the null-space path
  std::size_t NumQPIterations() const noexcept { return iterations.size(); }
  std::size_t NumLineSearchSteps() const noexcept {
    std::size_t n = 0;
    for (const auto& it : iterations) n += it.line_search_steps.size();
    return n;
  }
  std::size_t NumFailedLineSearches() const noexcept {
    std::size_t n = 0;
    for (const auto& it : iterations) {
      n += (it.step_result != StepSizeSelectionResult::SUCCESS &&
            it.step_result != StepSizeSelectionResult::FIRST_ORDER_SATISFIED);
    }
    return n;
  }
  // mini_opt/source/structs.cc:256-266
  std::string ToString() const {
    std::string r;
    for (const auto& it : iterations) r += it.ToString();
    r += std::string("  termination = ") + ToName(termination_state) + "\n";
    return r;
  }
};

}  // namespace mini_opt

namespace pendulum {

// optimization/optimization.hpp:12-53, field for field and with the same defaults.
struct OptimizationParams {
  double control_dt{0.01};
  std::size_t window_length{40};
  std::size_t state_spacing{10};
  std::size_t max_iterations{8};
  double relative_exit_tol{1.0e-5};
  double absolute_first_derivative_tol{1.0e-6};
  double equality_penalty_initial{1.0};
  double u_guess_sinusoid_amplitude{10.0};
  double u_cost_weight{0.1};
  double u_derivative_cost_weight{0.1};
  double b_x_final_cost_weight{150.0};
  double th_final_cost_weight{-1.0};
  double b_x_dot_final_cost_weight{-1.0};
  double th_dot_final_cost_weight{-1.0};
  constexpr std::size_t NumStates() const noexcept { return window_length / state_spacing + 1; }
};

// optimization/optimization.hpp:55-70
struct OptimizationOutputs {
  SingleCartPoleState initial_state;
  std::vector<double> previous_solution;
  mini_opt::NLSSolverOutputs solver_outputs;
  std::vector<double> u;
  std::vector<SingleCartPoleState> predicted_states;
};

namespace detail {

struct CpmpcError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void Check(int rc) {
  if (rc == CPMPC_OK) return;
  const std::string msg = cpmpc_last_error();
  if (rc == CPMPC_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  throw CpmpcError("cpmpc error " + std::to_string(rc) + ": " + msg);
}

inline cpmpc_params ToC(const OptimizationParams& p) {
  cpmpc_params c;
  c.control_dt = p.control_dt;
  c.window_length = static_cast<int64_t>(p.window_length);
  c.state_spacing = static_cast<int64_t>(p.state_spacing);
  c.max_iterations = static_cast<int64_t>(p.max_iterations);
  c.relative_exit_tol = p.relative_exit_tol;
  c.absolute_first_derivative_tol = p.absolute_first_derivative_tol;
  c.equality_penalty_initial = p.equality_penalty_initial;
  c.u_guess_sinusoid_amplitude = p.u_guess_sinusoid_amplitude;
  c.u_cost_weight = p.u_cost_weight;
  c.u_derivative_cost_weight = p.u_derivative_cost_weight;
  c.b_x_final_cost_weight = p.b_x_final_cost_weight;
  c.th_final_cost_weight = p.th_final_cost_weight;
  c.b_x_dot_final_cost_weight = p.b_x_dot_final_cost_weight;
  c.th_dot_final_cost_weight = p.th_dot_final_cost_weight;
  return c;
}

}  // namespace detail

// B independent MPC controllers on one GPU.  Buffers are host memory; see include/cpmpc.h for the device-pointer
// entry points used by callers that keep their data on the GPU.
class BatchedOptimization {
 public:
  // `flags`: 0 (throughput mapping), CPMPC_FLAG_LOW_LATENCY or CPMPC_FLAG_MIN_LATENCY (include/cpmpc.h)
  BatchedOptimization(const OptimizationParams& params, std::size_t batch, int device = 0, std::uint32_t flags = 0)
      : params_(params), batch_(batch) {
    const cpmpc_params c = detail::ToC(params);
    detail::Check(cpmpc_create_ex(&c, static_cast<int64_t>(batch), device, flags, &handle_));
  }
  ~BatchedOptimization() { cpmpc_destroy(handle_); }
  BatchedOptimization(const BatchedOptimization&) = delete;
  BatchedOptimization& operator=(const BatchedOptimization&) = delete;

  std::size_t batch() const noexcept { return batch_; }
  std::size_t solution_dim() const noexcept { return static_cast<std::size_t>(cpmpc_solution_dim(handle_)); }
  const OptimizationParams& params() const noexcept { return params_; }
  cpmpc_handle handle() const noexcept { return handle_; }

  // Optimization::Step for every instance.  u_out: [B*W]; predicted_out: [B*W] states or nullptr; stats: [B] or nullptr.
  void Step(const SingleCartPoleState* states, const SingleCartPoleParams* dynamics_params,
            const double* b_x_set_points, double* u_out, SingleCartPoleState* predicted_out = nullptr,
            cpmpc_stats* stats_out = nullptr, const double* weights = nullptr) {
    detail::Check(cpmpc_step(handle_, reinterpret_cast<const double*>(states),
                             reinterpret_cast<const double*>(dynamics_params), b_x_set_points, weights, u_out,
                             reinterpret_cast<double*>(predicted_out), stats_out));
  }
  void Reset(const std::uint8_t* mask = nullptr) { detail::Check(cpmpc_reset(handle_, mask)); }
  void SetPreviousSolution(const double* solutions, const std::uint8_t* mask = nullptr) {
    detail::Check(cpmpc_set_previous_solution(handle_, solutions, mask));
  }
  void GetPreviousSolution(double* solutions_out, std::uint8_t* has_prev_out = nullptr) const {
    detail::Check(cpmpc_get_previous_solution(handle_, solutions_out, has_prev_out));
  }
  void EnableIterationLog(bool enable) { detail::Check(cpmpc_enable_iteration_log(handle_, enable ? 1 : 0)); }
  void GetIterationLog(double* log_out) const { detail::Check(cpmpc_get_iteration_log(handle_, log_out)); }
  // n_steps x { Step; Simulator::Step(control_dt, u[0]) } on the device; states updated in place.
  void ClosedLoop(std::size_t n_steps, SingleCartPoleState* states, const SingleCartPoleParams* dynamics_params,
                  const double* b_x_set_points, double* controls_out = nullptr, std::int64_t* stats_sum_out = nullptr,
                  const double* weights = nullptr) {
    detail::Check(cpmpc_closed_loop(handle_, static_cast<int64_t>(n_steps), reinterpret_cast<double*>(states),
                                    reinterpret_cast<const double*>(dynamics_params), b_x_set_points, weights,
                                    controls_out, stats_sum_out));
  }

 private:
  OptimizationParams params_;
  std::size_t batch_;
  cpmpc_handle handle_{nullptr};
};

// Drop-in for pendulum::Optimization (optimization/optimization.hpp:73-108): a GPU batch of one.
class Optimization {
 public:
  explicit Optimization(const OptimizationParams& params) : impl_(params, 1) { impl_.EnableIterationLog(true); }

  [[nodiscard]] OptimizationOutputs Step(const SingleCartPoleState& current_state,
                                         const SingleCartPoleParams& dynamics_params, const double b_x_set_point) {
    const OptimizationParams& p = impl_.params();
    OptimizationOutputs out;
    out.initial_state = current_state;
    // the solution of the call BEFORE this one (optimization.cc:84,94); empty on the first call
    std::vector<double> prev(impl_.solution_dim());
    std::uint8_t has_prev = 0;
    impl_.GetPreviousSolution(prev.data(), &has_prev);
    if (has_prev) out.previous_solution = std::move(prev);
    out.u.resize(p.window_length);
    out.predicted_states.resize(p.window_length);
    cpmpc_stats st;
    impl_.Step(&current_state, &dynamics_params, &b_x_set_point, out.u.data(), out.predicted_states.data(), &st);
    std::vector<double> log(p.max_iterations * CPMPC_ITERLOG_STRIDE);
    impl_.GetIterationLog(log.data());
    auto& so = out.solver_outputs;
    so.termination_state = static_cast<mini_opt::NLSTerminationState>(st.termination_state < 0 ? 0 : st.termination_state);
    so.final_errors = {st.f, st.equality};
    so.final_penalty = st.penalty;
    if (st.termination_state < 0) {
      // the reference throws when the cost at the linearisation point is not finite (nonlinear.cc:356-357)
      throw std::domain_error("non-finite cost at the linearization point");
    }
    for (int i = 0; i < st.n_iterations; ++i) {
      const double* r = log.data() + static_cast<std::size_t>(i) * CPMPC_ITERLOG_STRIDE;
      mini_opt::NLSIteration it;
      it.iteration = i;
      it.optimizer_state = static_cast<mini_opt::OptimizerState>(static_cast<int>(r[0]));
      it.lambda = r[1];
      it.errors_initial = {r[2], r[3]};
      it.directional_derivatives = {r[4], r[5]};
      it.penalty = r[6];
      it.step_result = static_cast<mini_opt::StepSizeSelectionResult>(static_cast<int>(r[7]));
      const int n_steps = static_cast<int>(r[8]);
      for (int s = 0; s < n_steps && s < 6; ++s) {
        it.line_search_steps.push_back({r[9 + 3 * s], {r[10 + 3 * s], r[11 + 3 * s]}});
      }
      so.iterations.push_back(std::move(it));
    }
    return out;
  }

  void Reset() { impl_.Reset(); }

  void SetPreviousSolution(const std::vector<double>& guess) {
    if (guess.empty()) {
      impl_.Reset();  // resize(0) in the reference == no warm start
      return;
    }
    if (guess.size() != impl_.solution_dim()) throw std::invalid_argument("guess has the wrong dimension");
    impl_.SetPreviousSolution(guess.data());
  }

 private:
  BatchedOptimization impl_;
};

// Drop-in for pendulum::Simulator (optimization/simulator.hpp:10-29); the integration runs on the GPU.
class Simulator {
 public:
  void Step(const SingleCartPoleParams& params, double dt, double u, const Vector2& f_base, const Vector2& f_mass) {
    const double fb[2] = {f_base.x, f_base.y};
    const double fm[2] = {f_mass.x, f_mass.y};
    detail::Check(cpmpc_sim_step(device_, 1, reinterpret_cast<double*>(&state_),
                                 reinterpret_cast<const double*>(&params), &u, dt, fb, fm));
  }
  SingleCartPoleState GetState() const noexcept { return state_; }
  void SetState(const SingleCartPoleState& state) noexcept { state_ = state; }

 private:
  int device_{0};
  SingleCartPoleState state_{0.0, -1.5707963267948966, 0.0, 0.0};
};

}  // namespace pendulum
