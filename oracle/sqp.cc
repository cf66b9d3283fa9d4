// ORACLE — TEST INFRASTRUCTURE ONLY (see sqp.hpp header for scope and citations).
#include "sqp.hpp"

#include <stdexcept>

namespace oracle {

// ---------------------------------------------------------------------------------------------------------------------
// Residual: gather -> evaluate -> scatter   (residual.hpp:146-250)
// ---------------------------------------------------------------------------------------------------------------------

static void Gather(const Vec& vars, const std::vector<int>& index, Vec* out) {
  out->resize(index.size());
  for (std::size_t local = 0; local < index.size(); ++local) {
    const int i = index[local];
    if (i < 0 || i >= static_cast<int>(vars.size())) throw std::out_of_range("Index exceeds the number of params");
    (*out)[local] = vars[i];
  }
}

void Residual::ErrorVector(const Vec& vars, double* r_out) const {
  Vec p;
  Gather(vars, index, &p);
  func(p.data(), r_out, nullptr);
}

double Residual::UpdateHessian(const Vec& vars, Mat* const H, Vec* const b) const {
  Vec p;
  Gather(vars, index, &p);
  const int P = static_cast<int>(index.size());
  Vec r(dim, 0.0), J(static_cast<std::size_t>(dim) * P, 0.0);
  func(p.data(), r.data(), J.data());

  // lower-triangular scatter, in the reference's loop order (residual.hpp:204-224)
  for (int row_local = 0; row_local < P; ++row_local) {
    const int row_global = index[row_local];
    for (int col_local = 0; col_local <= row_local; ++col_local) {
      const int col_global = index[col_local];
      double JtJ = 0.0;
      for (int k = 0; k < dim; ++k) JtJ += J[k * P + row_local] * J[k * P + col_local];
      if (col_global <= row_global) {
        (*H)(row_global, col_global) += JtJ;
      } else {
        (*H)(col_global, row_global) += JtJ;
      }
    }
    double Jtr = 0.0;
    for (int k = 0; k < dim; ++k) Jtr += J[k * P + row_local] * r[k];
    (*b)[row_global] += Jtr;
  }
  double sq = 0.0;
  for (int k = 0; k < dim; ++k) sq += r[k] * r[k];
  return 0.5 * sq;
}

void Residual::UpdateJacobian(const Vec& vars, Mat* const A, Vec* const b, const int row) const {
  Vec p;
  Gather(vars, index, &p);
  const int P = static_cast<int>(index.size());
  Vec J(static_cast<std::size_t>(dim) * P, 0.0);
  func(p.data(), b->data() + row, J.data());
  for (int col_local = 0; col_local < P; ++col_local) {
    const int col_global = index[col_local];
    for (int k = 0; k < dim; ++k) (*A)(row + k, col_global) = J[k * P + col_local];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Null-space QP   (qp.cc:679-729)
// ---------------------------------------------------------------------------------------------------------------------

QPNullSpaceTerminationState QPNullSpaceSolver::Solve(const QP& p) {
  const int K = p.A_eq.rows;  // num equality constraints
  const int N = p.A_eq.cols;  // num params
  if (K <= 0) throw std::invalid_argument("Problem must have at least one equality constraint");

  // QR of A_eq^T with column pivoting: A_eq^T P = Q R
  Mat At(N, K);
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < N; ++j) At(j, i) = p.A_eq(i, j);
  ColPivQR qr;
  qr.Compute(At);
  const Mat Q = qr.MatrixQ();
  const int rank = qr.Rank();
  if (rank < K) {
    // The reference mixes `K` (R.topRows(K)) and `rank` (Q.leftCols(rank)) here, which is only meaningful at
    // full row rank (qp.cc:696-706).  A rank-deficient A_eq is reported as its own failure class.
    x_.assign(N, 0.0);
    return QPNullSpaceTerminationState::RANK_DEFICIENT;
  }
  const int Z = N - rank;

  // permuted_rhs = P^T * (-b)
  Vec rhs(K);
  for (int j = 0; j < K; ++j) rhs[j] = -p.b_eq[qr.perm[j]];

  // solve R1^T w = rhs (R1^T lower triangular), then u = Q1 w
  Vec w(K);
  for (int i = 0; i < K; ++i) {
    double s = rhs[i];
    for (int j = 0; j < i; ++j) s -= qr.qr(j, i) * w[j];
    w[i] = s / qr.qr(i, i);
  }
  Vec u(N, 0.0);
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    for (int j = 0; j < rank; ++j) s += Q(i, j) * w[j];
    u[i] = s;
  }

  // G_reduced = Q2^T sym(G) Q2
  // Both products are swept so that the innermost loop runs along a row of the row-major operands (the compiler
  // vectorises it); every entry still receives its terms in increasing j (resp. i) starting from 0.0, i.e. exactly
  // the sums of the plain dot-product form.
  Mat GQ2(N, Z);
  for (int i = 0; i < N; ++i) {
    double* out = &GQ2(i, 0);
    for (int j = 0; j < N; ++j) {
      const double g = (j <= i) ? p.G(i, j) : p.G(j, i);
      const double* qj = Q.a.data() + static_cast<std::size_t>(j) * Q.cols + rank;
      for (int z = 0; z < Z; ++z) out[z] += g * qj[z];
    }
  }
  Mat Gr(Z, Z);
  for (int a = 0; a < Z; ++a) {
    double* out = &Gr(a, 0);
    for (int i = 0; i < N; ++i) {
      const double q = Q(i, rank + a);
      const double* gi = &GQ2(i, 0);
      for (int b = 0; b < Z; ++b) out[b] += q * gi[b];
    }
  }

  // Cholesky of the lower triangle; failure => not positive definite
  if (CholeskyLowerInPlace(Gr) >= 0) {
    return QPNullSpaceTerminationState::NOT_POSITIVE_DEFINITE;
  }

  // y = -(Q2^T (c + G u))
  const Vec Gu = SymLowerTimes(p.G, u);
  Vec y(Z);
  for (int z = 0; z < Z; ++z) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += Q(i, rank + z) * (p.c[i] + Gu[i]);
    y[z] = -s;
  }
  CholeskySolveInPlace(Gr, y);

  // x = u + Q2 y
  x_.resize(N);
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    for (int z = 0; z < Z; ++z) s += Q(i, rank + z) * y[z];
    x_[i] = u[i] + s;
  }
  return QPNullSpaceTerminationState::SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------------
// SQP outer loop   (nonlinear.cc)
// ---------------------------------------------------------------------------------------------------------------------

ConstrainedNonlinearLeastSquares::ConstrainedNonlinearLeastSquares(const Problem* const problem,
                                                                   Retraction retraction)
    : p_(problem), custom_retraction_(std::move(retraction)) {
  if (p_ == nullptr || p_->dimension <= 0) throw std::invalid_argument("Need at least one variable");
  // nonlinear.cc:19-40 — sizes are fixed at construction from the problem as it is *now*.
  qp_.G.Resize(p_->dimension, p_->dimension);
  qp_.c.assign(p_->dimension, 0.0);
  int total_eq_size = 0;
  for (const Residual& constraint : p_->equality_constraints) total_eq_size += constraint.Dimension();
  qp_.A_eq.Resize(total_eq_size, p_->dimension);
  qp_.b_eq.assign(total_eq_size, 0.0);
  variables_.assign(p_->dimension, 0.0);
  candidate_vars_.assign(p_->dimension, 0.0);
  dx_.assign(p_->dimension, 0.0);
}

static void CheckParams(const ConstrainedNonlinearLeastSquares::Params& params) {
  // nonlinear.cc:49-73
  const bool ok = params.max_iterations >= 0 && params.max_qp_iterations >= 1 &&
                  params.termination_kkt_tolerance > 0 && params.absolute_exit_tol > 0 &&
                  params.max_line_search_iterations >= 0 && params.relative_exit_tol >= 0 &&
                  params.relative_exit_tol <= 1 && params.absolute_first_derivative_tol >= 0 &&
                  params.armijo_search_tau > 0 && params.armijo_search_tau < 1 &&
                  params.equality_penalty_initial >= 0 && params.equality_penalty_scale_factor >= 1.0 &&
                  params.equality_penalty_rho >= 0 && params.equality_penalty_rho < 1 && params.max_lambda >= 0 &&
                  params.min_lambda <= params.max_lambda && params.lambda_initial >= params.min_lambda &&
                  params.lambda_initial <= params.max_lambda && params.lambda_failure_init >= 0 &&
                  params.lambda_decrease_on_success >= 0 && params.lambda_decrease_on_success < 1.0 &&
                  params.lambda_decrease_on_restore >= 0 && params.lambda_decrease_on_restore < 1.0;
  if (!ok) throw std::invalid_argument("invalid ConstrainedNonlinearLeastSquares::Params");
}

NLSSolverOutputs ConstrainedNonlinearLeastSquares::Solve(const Params& params, const Vec& variables) {
  CheckParams(params);
  if (p_->equality_constraints.empty()) {
    throw std::invalid_argument("oracle restates only the equality-constrained (null-space) path");
  }
  variables_ = variables;
  state_ = OptimizerState::NOMINAL;

  double lambda{params.lambda_initial};
  double penalty{params.equality_penalty_initial};
  std::vector<NLSIteration> iterations;
  iterations.reserve(10);
  for (int iter = 0; iter < params.max_iterations; ++iter) {
    const Errors errors_pre = LinearizeAndFillQP(variables_, lambda, *p_, &qp_);

    // ComputeStepDirection (nonlinear.cc:216-259)
    const QPNullSpaceTerminationState qp_term = solver_.Solve(qp_);
    if (qp_term == QPNullSpaceTerminationState::SUCCESS) {
      dx_ = solver_.variables();
    } else {
      std::fill(dx_.begin(), dx_.end(), 0.0);
    }
    if (qp_term != QPNullSpaceTerminationState::SUCCESS) {
      // NOT_POSITIVE_DEFINITE in the reference; the oracle folds RANK_DEFICIENT into the same exit.
      return {NLSTerminationState::QP_INDEFINITE, std::move(iterations)};
    }

    const DirectionalDerivatives directional_derivative = ComputeQPCostDerivative(qp_, dx_);

    // Raise the penalty if needed (nonlinear.cc:112-118)
    const double new_penalty = SelectPenalty(qp_, dx_, params.equality_penalty_rho);
    if (new_penalty > penalty) {
      penalty = new_penalty * params.equality_penalty_scale_factor;
    }

    constexpr double armijo_c1 = 1.0e-4;
    const StepSizeSelectionResult step_result = SelectStepSize(
        params.max_line_search_iterations, params.absolute_first_derivative_tol, errors_pre,
        directional_derivative, penalty, armijo_c1, params.line_search_strategy, params.armijo_search_tau);

    const double old_lambda = lambda;
    const std::optional<NLSTerminationState> maybe_exit =
        UpdateLambdaAndCheckExitConditions(params, step_result, errors_pre, penalty, lambda);

    iterations.push_back(NLSIteration{iter, state_, old_lambda, errors_pre, qp_term, directional_derivative,
                                      penalty, step_result, std::move(steps_)});
    steps_.clear();

    if (maybe_exit.has_value()) {
      return {*maybe_exit, std::move(iterations)};
    }
  }
  return {NLSTerminationState::MAX_ITERATIONS, std::move(iterations)};
}

void ConstrainedNonlinearLeastSquares::RetractCandidateVars(const double alpha) {
  candidate_vars_ = variables_;
  if (custom_retraction_) {
    custom_retraction_(candidate_vars_, dx_, alpha);
  } else {
    for (std::size_t i = 0; i < candidate_vars_.size(); ++i) candidate_vars_[i] += dx_[i] * alpha;
  }
}

Errors ConstrainedNonlinearLeastSquares::LinearizeAndFillQP(const Vec& variables, const double lambda,
                                                            const Problem& problem, QP* const qp) {
  // nonlinear.cc:170-214
  Errors output_errors{};
  qp->G.SetZero();
  std::fill(qp->c.begin(), qp->c.end(), 0.0);
  for (const Residual& cost : problem.costs) {
    output_errors.f += cost.UpdateHessian(variables, &qp->G, &qp->c);
  }
  if (lambda > 0) {
    for (int i = 0; i < qp->G.rows; ++i) qp->G(i, i) += lambda;
  }
  qp->A_eq.SetZero();
  std::fill(qp->b_eq.begin(), qp->b_eq.end(), 0.0);
  int row = 0;
  for (const Residual& eq : problem.equality_constraints) {
    const int dim = eq.Dimension();
    if (row + dim > qp->A_eq.rows) throw std::out_of_range("equality rows exceed A_eq");
    eq.UpdateJacobian(variables, &qp->A_eq, &qp->b_eq, row);
    output_errors.equality += L1Norm(qp->b_eq.data() + row, dim);
    row += dim;
  }
  return output_errors;
}

Errors ConstrainedNonlinearLeastSquares::EvaluateNonlinearErrors(const Vec& vars) const {
  // nonlinear.cc:279-293
  Errors output_errors{};
  double buffer[16];
  for (const Residual& cost : p_->costs) {
    cost.ErrorVector(vars, buffer);
    double sq = 0.0;
    for (int k = 0; k < cost.Dimension(); ++k) sq += buffer[k] * buffer[k];
    output_errors.f += 0.5 * sq;
  }
  for (const Residual& eq : p_->equality_constraints) {
    eq.ErrorVector(vars, buffer);
    output_errors.equality += L1Norm(buffer, eq.Dimension());
  }
  return output_errors;
}

std::optional<NLSTerminationState> ConstrainedNonlinearLeastSquares::UpdateLambdaAndCheckExitConditions(
    const Params& params, const StepSizeSelectionResult step_result, const Errors& initial_errors,
    const double penalty, double& lambda) {
  // nonlinear.cc:296-343
  if (step_result == StepSizeSelectionResult::SUCCESS) {
    variables_.swap(candidate_vars_);
    if (state_ == OptimizerState::ATTEMPTING_RESTORE_LM) {
      lambda = std::max(lambda * params.lambda_decrease_on_restore, params.min_lambda);
    } else {
      lambda = std::max(lambda * params.lambda_decrease_on_success, params.min_lambda);
    }
    state_ = OptimizerState::NOMINAL;

    const LineSearchStep& final_step = steps_.back();
    if (final_step.errors.LInfinity() < params.absolute_exit_tol) {
      return NLSTerminationState::SATISFIED_ABSOLUTE_TOL;
    } else if (final_step.errors.Total(penalty) > initial_errors.Total(penalty) * (1 - params.relative_exit_tol)) {
      return NLSTerminationState::SATISFIED_RELATIVE_TOL;
    }
  } else if (step_result == StepSizeSelectionResult::FIRST_ORDER_SATISFIED) {
    return NLSTerminationState::SATISFIED_FIRST_ORDER_TOL;
  } else if (step_result == StepSizeSelectionResult::MAX_ITERATIONS ||
             step_result == StepSizeSelectionResult::POSITIVE_DERIVATIVE) {
    if (state_ == OptimizerState::NOMINAL) {
      lambda = std::max(params.lambda_failure_init, lambda * 10.0);
      state_ = OptimizerState::ATTEMPTING_RESTORE_LM;
    } else {
      lambda *= 10.0;
    }
    if (lambda > params.max_lambda) {
      return NLSTerminationState::MAX_LAMBDA;
    }
  }
  // FAILURE_* results fall through: no update of variables, lambda or state.
  return std::nullopt;
}

StepSizeSelectionResult ConstrainedNonlinearLeastSquares::SelectStepSize(
    const int max_iterations, const double abs_first_derivative_tol, const Errors& errors_pre,
    const DirectionalDerivatives& derivatives, const double penalty, const double armijo_c1,
    const LineSearchStrategy strategy, const double backtrack_search_tau) {
  // nonlinear.cc:351-416
  if (!(penalty > 0.0)) throw std::domain_error("penalty must be positive");
  if (errors_pre.ContainsInvalidValues()) throw std::domain_error("non-finite cost at the linearization point");
  steps_.clear();
  steps_.reserve(5);

  const double directional_derivative = derivatives.Total(penalty);

  double alpha{1};
  for (int iter = 0; iter <= max_iterations; ++iter) {
    if (strategy == LineSearchStrategy::POLYNOMIAL_APPROXIMATION) {
      if (iter > 0) {
        const std::optional<double> new_alpha =
            ComputeAlphaPolynomialApproximation(iter, errors_pre, derivatives, penalty);
        if (!new_alpha.has_value() || !std::isfinite(*new_alpha) || *new_alpha <= 0.0 || *new_alpha >= alpha) {
          return StepSizeSelectionResult::FAILURE_INVALID_ALPHA;
        } else {
          alpha = new_alpha.value();
        }
      }
    } else {
      if (iter > 0) alpha = alpha * backtrack_search_tau;
    }

    RetractCandidateVars(alpha);

    const Errors errors_step = EvaluateNonlinearErrors(candidate_vars_);
    steps_.push_back(LineSearchStep{alpha, errors_step});

    if (errors_step.ContainsInvalidValues()) {
      return StepSizeSelectionResult::FAILURE_NON_FINITE_COST;
    }
    if (derivatives.LInfinity() < abs_first_derivative_tol) {
      return StepSizeSelectionResult::FIRST_ORDER_SATISFIED;
    } else if (directional_derivative > 0) {
      return StepSizeSelectionResult::POSITIVE_DERIVATIVE;
    }
    if (errors_step.Total(penalty) <= errors_pre.Total(penalty) + directional_derivative * alpha * armijo_c1) {
      return StepSizeSelectionResult::SUCCESS;
    }
  }
  return StepSizeSelectionResult::MAX_ITERATIONS;
}

std::optional<double> ConstrainedNonlinearLeastSquares::ComputeAlphaPolynomialApproximation(
    const int iteration, const Errors& errors_pre, const DirectionalDerivatives& derivatives,
    const double penalty) const {
  // nonlinear.cc:418-443
  if (iteration == 1) {
    const LineSearchStep& prev_step = steps_.back();
    return QuadraticApproxMinimum(errors_pre.Total(penalty), derivatives.Total(penalty), prev_step.alpha,
                                  prev_step.errors.Total(penalty));
  }
  const LineSearchStep& second_last_step = steps_[steps_.size() - 2];
  const LineSearchStep& last_step = steps_.back();
  double ab[2];
  CubicApproxCoeffs(errors_pre.Total(penalty), derivatives.Total(penalty), second_last_step.alpha,
                    second_last_step.errors.Total(penalty), last_step.alpha, last_step.errors.Total(penalty), ab);
  return CubicApproxMinimum(derivatives.Total(penalty), ab);
}

static double Sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

DirectionalDerivatives ConstrainedNonlinearLeastSquares::ComputeQPCostDerivative(const QP& qp, const Vec& dx) {
  // nonlinear.cc:457-482
  DirectionalDerivatives out{};
  out.d_f = Dot(qp.c, dx);
  for (int i = 0; i < qp.A_eq.rows; ++i) {
    double row_dot = 0.0;
    for (int j = 0; j < qp.A_eq.cols; ++j) row_dot += qp.A_eq(i, j) * dx[j];
    out.d_equality += Sign(qp.b_eq[i]) * row_dot;
  }
  return out;
}

double ConstrainedNonlinearLeastSquares::SelectPenalty(const QP& qp, const Vec& dx, const double rho) {
  // nonlinear.cc:484-502, null-space branch (no Lagrange multipliers available): inequality (18.36).
  const double l1_eq = std::max(L1Norm(qp.b_eq.data(), static_cast<int>(qp.b_eq.size())),
                                std::numeric_limits<double>::epsilon());
  const Vec Gdx = SymLowerTimes(qp.G, dx);
  const double quadratic_cost_approx = Dot(qp.c, dx) + 0.5 * std::max(0.0, Dot(dx, Gdx));
  return quadratic_cost_approx / ((1 - rho) * l1_eq);
}

std::optional<double> ConstrainedNonlinearLeastSquares::QuadraticApproxMinimum(const double phi_0,
                                                                              const double phi_prime_0,
                                                                              const double alpha_0,
                                                                              const double phi_alpha_0) {
  // nonlinear.cc:531-538
  const double numerator = phi_alpha_0 - phi_prime_0 * alpha_0 - phi_0;
  if (phi_prime_0 > 0 || numerator <= 0) return std::nullopt;
  return -phi_prime_0 * alpha_0 * alpha_0 / (2.0 * numerator);
}

void ConstrainedNonlinearLeastSquares::CubicApproxCoeffs(const double phi_0, const double phi_prime_0,
                                                         const double alpha_0, const double phi_alpha_0,
                                                         const double alpha_1, const double phi_alpha_1,
                                                         double ab_out[2]) {
  // nonlinear.cc:563-578.  Eigen's Matrix2d::inverse() is the adjugate scaled by 1/det.
  if (!(alpha_1 > 0) || !(alpha_0 > alpha_1)) throw std::domain_error("cubic fit needs alpha_0 > alpha_1 > 0");
  const double a00 = alpha_0 * alpha_0 * alpha_0, a01 = alpha_0 * alpha_0;
  const double a10 = alpha_1 * alpha_1 * alpha_1, a11 = alpha_1 * alpha_1;
  const double r0 = phi_alpha_0 - phi_0 - phi_prime_0 * alpha_0;
  const double r1 = phi_alpha_1 - phi_0 - phi_prime_0 * alpha_1;
  const double inv_det = 1.0 / (a00 * a11 - a10 * a01);
  const double i00 = a11 * inv_det, i01 = -a01 * inv_det;
  const double i10 = -a10 * inv_det, i11 = a00 * inv_det;
  ab_out[0] = i00 * r0 + i01 * r1;
  ab_out[1] = i10 * r0 + i11 * r1;
}

std::optional<double> ConstrainedNonlinearLeastSquares::CubicApproxMinimum(const double phi_prime_0,
                                                                          const double ab[2]) {
  // nonlinear.cc:594-605
  const double a = ab[0];
  const double b = ab[1];
  const double arg_sqrt = b * b - 3 * a * phi_prime_0;
  constexpr double kNegativeTol = -1.0e-12;
  if (a == 0.0 || arg_sqrt < kNegativeTol) return std::nullopt;
  const double numerator = -b + std::sqrt(std::max(arg_sqrt, 0.));
  return numerator / (3 * a);
}

}  // namespace oracle
