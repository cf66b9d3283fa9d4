// ORACLE — TEST INFRASTRUCTURE ONLY (see sqp.hpp header).  PARITY UNPINNED: the reference holds no implementation of
// the double-pendulum controller at all (symbolic/dynamics_double.py is "work in progress", no generated dynamics, no
// optimiser; the optimiser carries a TODO for it, optimization/optimization.cc:197-199).
//
// This is the nx = 6 generalisation of the reference's single-pendulum controller, restated on top of the same
// mini_opt restatement (sqp.hpp): the same multiple-shooting layout [x(0) .. x(NS-1) | u(0) .. u(W-1)]
// (optimization.cc:24-37) with x = [b_x, th_1, th_2, b_x', th_1', th_2'] (dynamics_double.py:118-129), the same costs
// (optimization.cc:270-301), the same terminal rule (negative weight = hard equality, optimization.cc:236-267) with
// both angles driven to +pi/2 (upright) and all rates to zero, the same retraction (optimization.cc:309-329) wrapping
// both angles, the same RK4 (integration.hpp:13-62).
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

#include "sqp.hpp"

namespace oracle {

constexpr int kDpNx = 6;

// dynamics_double.py:13-22 (field order preserved)
struct DoubleCartPoleParams {
  double m_b, m_1, m_2, l_1, l_2, g;
};

struct DoubleOptimizationParams {
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
  // terminal weights in state order; negative = hard equality
  double final_cost_weight[kDpNx]{150.0, -1.0, -1.0, -1.0, -1.0, -1.0};
  std::size_t NumStates() const { return window_length / state_spacing + 1; }
};

struct DoubleOptimizationOutputs {
  std::vector<double> previous_solution;
  NLSSolverOutputs solver_outputs{NLSTerminationState::MAX_ITERATIONS, {}};
  std::vector<double> u;
  std::vector<double> predicted_states;  // W x 6
  Errors final_errors;
  double final_penalty{0.0};
};

void DoubleRungeKutta4NoJacobians(const DoubleCartPoleParams& p, const double x[6], double u, double h,
                                  double x_new[6]);
void DoubleRungeKutta4(const DoubleCartPoleParams& p, const double x[6], double u, double h, double x_new[6],
                       double x_new_D_x[36], double x_new_D_u[6]);

class DoubleOptimization {
 public:
  explicit DoubleOptimization(const DoubleOptimizationParams& params);
  DoubleOptimizationOutputs Step(const double state[6], const DoubleCartPoleParams& dynamics_params,
                                 double b_x_set_point);
  void Reset() { previous_solution_.clear(); }
  void SetPreviousSolution(const std::vector<double>& guess) { previous_solution_ = guess; }
  const std::vector<double>& previous_solution() const { return previous_solution_; }

 private:
  void BuildProblem(const double state[6], const DoubleCartPoleParams& dynamics_params, double b_x_set_point);

  DoubleOptimizationParams params_;
  Problem problem_;
  std::unique_ptr<ConstrainedNonlinearLeastSquares> solver_;
  Vec previous_solution_;
};

}  // namespace oracle
