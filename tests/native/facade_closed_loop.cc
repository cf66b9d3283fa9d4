// The reference's end-to-end test (optimization/optimization_test.cc:12-77) written against the C++ host facade:
// only the include lines and the namespace of NLSTerminationState's header differ from the original.
// Build: g++ -std=c++17 facade_closed_loop.cc -L<pkg> -lcpmpc ; exit code 0 == every assertion of the reference holds.
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../cart_pole_mpc_b200/host/optimization.hpp"

#define ASSERT_NEAR(expected, actual, tol)                                                              \
  do {                                                                                                  \
    if (!(std::abs((expected) - (actual)) <= (tol))) {                                                  \
      std::printf("ASSERT_NEAR failed at line %d: %.17g vs %.17g\n", __LINE__, (double)(expected),      \
                  (double)(actual));                                                                    \
      return 1;                                                                                         \
    }                                                                                                   \
  } while (0)
#define ASSERT_NE(a, b)                                        \
  do {                                                         \
    if ((a) == (b)) {                                          \
      std::printf("ASSERT_NE failed at line %d\n", __LINE__); \
      return 1;                                                \
    }                                                          \
  } while (0)

using namespace pendulum;

int main() {
  constexpr std::size_t num_steps = 200;

  OptimizationParams optimization_params{};
  optimization_params.control_dt = 0.01;
  optimization_params.window_length = 40;
  optimization_params.state_spacing = 5;
  optimization_params.max_iterations = 10;

  constexpr SingleCartPoleParams dynamics_params{1.0, 0.1, 0.25, 9.81, 0.03, 0.1, 0.13, 0.8, 100.0};
  constexpr SingleCartPoleState x0{0.0, -M_PI / 2, 0.0, 0.0};

  std::vector<SingleCartPoleState> states{};
  states.push_back(x0);
  std::vector<double> controls{};

  Simulator sim{};
  sim.SetState(x0);

  // precondition failures surface as std::exception, like the reference's F_ASSERTs
  try {
    OptimizationParams bad = optimization_params;
    bad.state_spacing = 7;
    Optimization should_throw{bad};
    std::printf("invalid params were accepted\n");
    return 1;
  } catch (const std::exception&) {
  }

  Optimization optimization{optimization_params};
  std::size_t total_iterations = 0;
  for (std::size_t t = 0; t < num_steps; ++t) {
    const OptimizationOutputs outputs = optimization.Step(sim.GetState(), dynamics_params, 0.0);
    ASSERT_NE(mini_opt::NLSTerminationState::QP_INDEFINITE, outputs.solver_outputs.termination_state);
    ASSERT_NE(mini_opt::NLSTerminationState::MAX_LAMBDA, outputs.solver_outputs.termination_state);
    if (t == 0 && !outputs.previous_solution.empty()) {
      std::printf("previous_solution must be empty on the first call\n");
      return 1;
    }
    if (t == 1 && outputs.previous_solution.size() != 4 * 9 + 40) {
      std::printf("previous_solution has the wrong size\n");
      return 1;
    }
    if (t == 0) {
      // NLSSolverOutputs::ToString (mini_opt/source/structs.cc:159-266): one block per iteration + the termination line
      const std::string summary = outputs.solver_outputs.ToString();
      std::size_t blocks = 0;
      for (std::size_t pos = summary.find("Iteration # "); pos != std::string::npos;
           pos = summary.find("Iteration # ", pos + 1)) {
        ++blocks;
      }
      if (blocks != outputs.solver_outputs.iterations.size() ||
          summary.find("Iteration # 0, state = NOMINAL, lambda = 0\n") != 0 ||
          summary.find("  QP: SUCCESS\n") == std::string::npos ||
          summary.find("  termination = ") == std::string::npos) {
        std::printf("unexpected solver summary:\n%s", summary.c_str());
        return 1;
      }
      std::printf("SUMMARY_BEGIN\n%sSUMMARY_END\n", summary.c_str());
    }
    total_iterations += outputs.solver_outputs.iterations.size();
    const SingleCartPoleState& terminal_state = outputs.predicted_states.back();
    if (t > 20) {
      ASSERT_NEAR(0.0, terminal_state.b_x_dot, 1.0e-4);
      ASSERT_NEAR(0.0, terminal_state.th_1_dot, 1.0e-4);
      ASSERT_NEAR(M_PI / 2, terminal_state.th_1, 1.0e-4);
    }
    states.push_back(sim.GetState());
    controls.push_back(outputs.u.front());
    sim.Step(dynamics_params, optimization_params.control_dt, outputs.u.front(), {0, 0}, {0, 0});
  }
  const SingleCartPoleState& terminal_state = states.back();
  ASSERT_NEAR(0.0, terminal_state.b_x_dot, 1.0e-4);
  ASSERT_NEAR(0.0, terminal_state.th_1_dot, 1.0e-3);
  ASSERT_NEAR(M_PI / 2, terminal_state.th_1, 1.0e-4);
  std::printf("FACADE_CLOSED_LOOP_OK mean SQP iterations %.2f\n", total_iterations / 200.0);
  return 0;
}
