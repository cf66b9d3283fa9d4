// ORACLE — TEST INFRASTRUCTURE ONLY (see sqp.hpp header).  Pinned against the reference's compiled sources
// (oracle/_ref/libref.so): dynamics, RK4, RK4 Jacobians and the simulator are BIT-EQUAL to
// single_pendulum_dynamics.hpp / integration.hpp / simulator.cc (tests/test_reference_pinning.py).
//
// CPU restatement of the cart-pole side of the reference hot path:
//   dynamics + Jacobians     optimization/single_pendulum_dynamics.hpp:13-186
//   RK4 (+Jacobian chain)    optimization/integration.hpp:13-73
//   MPC controller           optimization/optimization.cc:13-371, optimization/optimization.hpp:12-108
//   plant simulator          optimization/simulator.cc:11-36
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

#include "sqp.hpp"

namespace oracle {

// optimization/structs.hpp:8-26 (field order preserved)
struct SingleCartPoleParams {
  double m_b, m_1, l_1, g, mu_b, v_mu_b, c_d_1, x_s, k_s;
};

// optimization/structs.hpp:44-49
struct SingleCartPoleState {
  double b_x, th_1, b_x_dot, th_1_dot;
};

// x_dot = f(x, u); J_x (row-major 4x4) and J_u (4) are optional (nullptr to skip).
void SinglePendulumDynamics(const SingleCartPoleParams& params, const double x[4], double u,
                            const double f_base[2], const double f_mass[2], double x_dot[4], double* J_x,
                            double* J_u);

// integration.hpp:65-73
double ModPi(double angle);

// integration.hpp:53-62 with the optimiser's zero external forces unless given.
void RungeKutta4NoJacobians(const SingleCartPoleParams& params, const double x[4], double u, double h,
                            const double f_base[2], const double f_mass[2], double x_new[4]);

// integration.hpp:13-49.  x_new_D_x is row-major 4x4, x_new_D_u has 4 entries.
void RungeKutta4(const SingleCartPoleParams& params, const double x[4], double u, double h, double x_new[4],
                 double x_new_D_x[16], double x_new_D_u[4]);

// optimization/optimization.hpp:12-53 (defaults identical)
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
  std::size_t NumStates() const { return window_length / state_spacing + 1; }
};

// optimization/optimization.hpp:55-70
struct OptimizationOutputs {
  SingleCartPoleState initial_state;
  std::vector<double> previous_solution;
  NLSSolverOutputs solver_outputs{NLSTerminationState::MAX_ITERATIONS, {}};
  std::vector<double> u;
  std::vector<SingleCartPoleState> predicted_states;
  // Extra (not in the reference struct): errors at the final variables, evaluated with
  // EvaluateNonlinearErrors (nonlinear.cc:279-293), and the final penalty — the "final cost" the parity
  // tests compare.
  Errors final_errors;
  double final_penalty{0.0};
};

// optimization/optimization.hpp:73-108
class Optimization {
 public:
  explicit Optimization(const OptimizationParams& params);
  OptimizationOutputs Step(const SingleCartPoleState& current_state, const SingleCartPoleParams& dynamics_params,
                           double b_x_set_point);
  void Reset() { previous_solution_.clear(); }
  void SetPreviousSolution(const std::vector<double>& guess) { previous_solution_ = guess; }
  const std::vector<double>& previous_solution() const { return previous_solution_; }

 private:
  void BuildProblem(const SingleCartPoleState& current_state, const SingleCartPoleParams& dynamics_params,
                    double b_x_set_point);
  void FillInitialGuess(Vec& guess, const SingleCartPoleParams& dynamics_params, std::size_t num_states) const;
  std::vector<SingleCartPoleState> ComputePredictedStates(const double* u_out,
                                                          const SingleCartPoleParams& dynamics_params,
                                                          const SingleCartPoleState& x_current) const;

  OptimizationParams params_;
  Problem problem_;
  std::unique_ptr<ConstrainedNonlinearLeastSquares> solver_;
  Vec previous_solution_;
};

// optimization/simulator.hpp:10-29
class Simulator {
 public:
  void Step(const SingleCartPoleParams& params, double dt, double u, const double f_base[2],
            const double f_mass[2]);
  SingleCartPoleState GetState() const { return {state_[0], state_[1], state_[2], state_[3]}; }
  void SetState(const SingleCartPoleState& s) {
    state_[0] = s.b_x;
    state_[1] = s.th_1;
    state_[2] = s.b_x_dot;
    state_[3] = s.th_1_dot;
  }

 private:
  void SubStep(const SingleCartPoleParams& params, double dt, double u, const double f_base[2],
               const double f_mass[2]);
  double state_[4] = {0.0, -1.5707963267948966, 0.0, 0.0};
};

}  // namespace oracle
