// ORACLE — TEST INFRASTRUCTURE ONLY (see linalg.hpp header).  PARITY PINNED against the reference's own sources:
// the reference's tests hold no golden control/cost/iteration vectors and Eigen is an empty submodule, so the
// reference's translation units are compiled unmodified against an Eigen-interface shim (oracle/eigen_shim,
// `make -C oracle ref` -> oracle/_ref/libref.so) and this restatement is checked against that library
// (tests/test_reference_pinning.py: physics bit-equal, 100 % iteration-count match, >= 99 % of config-2 solves
// within 1e-8 relative final cost) and against the reference's known-answer tests (tests/test_oracle_kats.py).
// What stays unpinned is Eigen's own blocked summation order inside QR / GEMM / LLT (the shim restates the
// published algorithms, see linalg.hpp); the QP step is unique, so this is a round-off-level statement.
//
// CPU restatement of the part of gareth-cross/mini_opt that the cart-pole MPC reaches:
//   Residual scatter semantics     dependencies/mini_opt/include/mini_opt/residual.hpp:174-250
//   QPNullSpaceSolver::Solve       dependencies/mini_opt/source/qp.cc:679-729
//   ConstrainedNonlinearLeastSquares (SQP + L1 merit line search + LM state machine)
//                                  dependencies/mini_opt/source/nonlinear.cc:75-605
//   result/stat types              dependencies/mini_opt/include/mini_opt/structs.hpp:147-347
//   default solver parameters      dependencies/mini_opt/include/mini_opt/nonlinear.hpp:64-124
// Only the branches an equality-only problem can reach are restated (no interior-point QP, no inequality
// constraints, no user-exit callback, no eigenvalue logging).
#pragma once
#include <functional>
#include <optional>
#include <vector>

#include "linalg.hpp"

namespace oracle {

// mini_opt/structs.hpp:169-186
struct Errors {
  double f{0.0};
  double equality{0.0};
  double Total(double penalty) const { return f + penalty * equality; }
  double LInfinity() const { return std::max(f, equality); }
  bool ContainsInvalidValues() const { return !std::isfinite(f) || !std::isfinite(equality); }
};

// mini_opt/structs.hpp:189-203
struct DirectionalDerivatives {
  double d_f{0.0};
  double d_equality{0.0};
  double Total(double penalty) const { return d_f + penalty * d_equality; }
  double LInfinity() const { return std::max(std::abs(d_f), std::abs(d_equality)); }
};

// mini_opt/structs.hpp:206-212
struct LineSearchStep {
  double alpha;
  Errors errors;
};

// Numeric values follow the declaration order of the reference enums (mini_opt/structs.hpp:215-248,159-164).
enum class StepSizeSelectionResult : int {
  SUCCESS = 0,
  MAX_ITERATIONS = 1,
  FIRST_ORDER_SATISFIED = 2,
  POSITIVE_DERIVATIVE = 3,
  FAILURE_NON_FINITE_COST = 4,
  FAILURE_INVALID_ALPHA = 5,
};
enum class NLSTerminationState : int {
  MAX_ITERATIONS = 0,
  SATISFIED_ABSOLUTE_TOL = 1,
  SATISFIED_RELATIVE_TOL = 2,
  SATISFIED_FIRST_ORDER_TOL = 3,
  MAX_LAMBDA = 4,
  QP_INDEFINITE = 5,
  USER_CALLBACK = 6,
};
enum class OptimizerState : int { NOMINAL = 0, ATTEMPTING_RESTORE_LM = 1 };
enum class QPNullSpaceTerminationState : int { SUCCESS = 0, NOT_POSITIVE_DEFINITE = 1, RANK_DEFICIENT = 2 };
enum class LineSearchStrategy : int { ARMIJO_BACKTRACK = 0, POLYNOMIAL_APPROXIMATION = 1 };

// A residual over a sparse subset of the variables.  `func(params, r, J)` receives the gathered parameters,
// writes `dim` residuals and, when J != nullptr, the row-major dim x index.size() Jacobian.
struct Residual {
  std::vector<int> index;
  int dim{0};
  std::function<void(const double* params, double* r, double* J)> func;

  int Dimension() const { return dim; }
  void ErrorVector(const Vec& vars, double* r_out) const;
  // G(lower) += J^T J, c += J^T r, returns 0.5 |r|^2   (residual.hpp:185-226)
  double UpdateHessian(const Vec& vars, Mat* H, Vec* b) const;
  // Writes rows [row, row+dim) of A and b                (residual.hpp:229-250)
  void UpdateJacobian(const Vec& vars, Mat* A, Vec* b, int row) const;
};

// nonlinear.hpp:33-58 (inequality constraints omitted: never present on this path)
struct Problem {
  int dimension{0};
  std::vector<Residual> costs;
  std::vector<Residual> equality_constraints;
  void clear() {
    costs.clear();
    equality_constraints.clear();
    dimension = 0;
  }
};

// qp.hpp:28-60 restricted to the equality-constrained form.
struct QP {
  Mat G;
  Vec c;
  Mat A_eq;
  Vec b_eq;
};

// qp.cc:679-729
class QPNullSpaceSolver {
 public:
  QPNullSpaceTerminationState Solve(const QP& p);
  const Vec& variables() const { return x_; }

 private:
  Vec x_;
};

struct NLSIteration {
  int iteration;
  OptimizerState optimizer_state;
  double lambda;
  Errors errors_initial;
  QPNullSpaceTerminationState qp_term;
  DirectionalDerivatives directional_derivatives;
  double penalty;
  StepSizeSelectionResult step_result;
  std::vector<LineSearchStep> line_search_steps;
};

struct NLSSolverOutputs {
  NLSTerminationState termination_state;
  std::vector<NLSIteration> iterations;
  std::size_t NumLineSearchSteps() const {
    std::size_t n = 0;
    for (const auto& it : iterations) n += it.line_search_steps.size();
    return n;
  }
  // structs.cc: number of iterations whose step selection was not SUCCESS/FIRST_ORDER_SATISFIED
  std::size_t NumFailedLineSearches() const {
    std::size_t n = 0;
    for (const auto& it : iterations) {
      n += (it.step_result == StepSizeSelectionResult::MAX_ITERATIONS ||
            it.step_result == StepSizeSelectionResult::POSITIVE_DERIVATIVE ||
            it.step_result == StepSizeSelectionResult::FAILURE_NON_FINITE_COST ||
            it.step_result == StepSizeSelectionResult::FAILURE_INVALID_ALPHA)
               ? 1
               : 0;
    }
    return n;
  }
};

class ConstrainedNonlinearLeastSquares {
 public:
  // nonlinear.hpp:64-124 (defaults identical)
  struct Params {
    int max_iterations{10};
    int max_qp_iterations{10};
    double termination_kkt_tolerance{1.0e-6};
    double absolute_exit_tol{1.0e-12};
    double relative_exit_tol{1.0e-5};
    double absolute_first_derivative_tol{1.0e-6};
    int max_line_search_iterations{2};
    LineSearchStrategy line_search_strategy{LineSearchStrategy::POLYNOMIAL_APPROXIMATION};
    double armijo_search_tau{0.8};
    double equality_penalty_initial{1.0};
    double equality_penalty_scale_factor{1.01};
    double equality_penalty_rho{0.1};
    double lambda_initial{0.0};
    double lambda_failure_init{1.0e-2};
    double lambda_decrease_on_success{0.1};
    double lambda_decrease_on_restore{0.8};
    double max_lambda{1.0};
    double min_lambda{0.0};
  };

  using Retraction = std::function<void(Vec& x, const Vec& dx, double alpha)>;

  explicit ConstrainedNonlinearLeastSquares(const Problem* problem, Retraction retraction = nullptr);

  NLSSolverOutputs Solve(const Params& params, const Vec& variables);

  const Vec& variables() const { return variables_; }

  Errors EvaluateNonlinearErrors(const Vec& vars) const;

  // Static pieces that the reference unit-tests directly (nonlinear_test.cc:109-250).
  static Errors LinearizeAndFillQP(const Vec& variables, double lambda, const Problem& problem, QP* qp);
  static DirectionalDerivatives ComputeQPCostDerivative(const QP& qp, const Vec& dx);
  static double SelectPenalty(const QP& qp, const Vec& dx, double rho);
  static std::optional<double> QuadraticApproxMinimum(double phi_0, double phi_prime_0, double alpha_0,
                                                      double phi_alpha_0);
  static void CubicApproxCoeffs(double phi_0, double phi_prime_0, double alpha_0, double phi_alpha_0,
                                double alpha_1, double phi_alpha_1, double ab_out[2]);
  static std::optional<double> CubicApproxMinimum(double phi_prime_0, const double ab[2]);

 private:
  void RetractCandidateVars(double alpha);
  StepSizeSelectionResult SelectStepSize(int max_iterations, double abs_first_derivative_tol,
                                         const Errors& errors_pre, const DirectionalDerivatives& derivatives,
                                         double penalty, double armijo_c1, LineSearchStrategy strategy,
                                         double backtrack_search_tau);
  std::optional<double> ComputeAlphaPolynomialApproximation(int iteration, const Errors& errors_pre,
                                                            const DirectionalDerivatives& derivatives,
                                                            double penalty) const;
  std::optional<NLSTerminationState> UpdateLambdaAndCheckExitConditions(const Params& params,
                                                                        StepSizeSelectionResult step_result,
                                                                        const Errors& initial_errors,
                                                                        double penalty, double& lambda);

  const Problem* p_;
  Retraction custom_retraction_;
  QP qp_;
  QPNullSpaceSolver solver_;
  Vec variables_;
  Vec candidate_vars_;
  Vec dx_;
  std::vector<LineSearchStep> steps_;
  OptimizerState state_{OptimizerState::NOMINAL};
};

}  // namespace oracle
