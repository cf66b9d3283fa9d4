// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Known-answer / behavioural checks that pin the oracle to the reference's OWN tests.  Each case restates one
// gtest case of the reference (file:line given at the case) against the oracle's classes, with the reference's
// tolerances.  Usage: `oracle_kat` runs every case, `oracle_kat <name>` runs one.  Exit code = number of failures.
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "cartpole.hpp"

using namespace oracle;

namespace {

int g_failures = 0;
const char* g_case = "";

#define CHECK_TRUE(cond)                                                              \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      std::printf("  [%s] CHECK FAILED %s:%d: %s\n", g_case, __FILE__, __LINE__, #cond); \
      ++g_failures;                                                                   \
      return;                                                                         \
    }                                                                                 \
  } while (0)

#define CHECK_NEAR(expected, actual, tol)                                                                  \
  do {                                                                                                     \
    const double e_ = (expected), a_ = (actual);                                                           \
    if (!(std::abs(e_ - a_) <= (tol))) {                                                                   \
      std::printf("  [%s] CHECK_NEAR FAILED %s:%d: expected %.17g got %.17g (tol %g)\n", g_case, __FILE__, \
                  __LINE__, e_, a_, static_cast<double>(tol));                                             \
      ++g_failures;                                                                                        \
      return;                                                                                              \
    }                                                                                                      \
  } while (0)

constexpr double kPico = 1.0e-12, kNano = 1.0e-9, kMicro = 1.0e-6;

// 6th-order central difference (geometry_utils numerical_derivative.hpp:29-38; integration_test.cc:10-20)
template <typename F>
double NumericalDerivative(double h, F func) {
  const double c1 = func(h) - func(-h);
  const double c2 = func(2 * h) - func(-2 * h);
  const double c3 = func(3 * h) - func(-3 * h);
  return (c1 * 45 - c2 * 9 + c3) / (60 * h);
}

// A residual r = A x - b over `index` with constant Jacobian A (dim x P row-major).
Residual LinearResidual(std::vector<int> index, int dim, std::vector<double> A, std::vector<double> b) {
  Residual r;
  const int P = static_cast<int>(index.size());
  r.index = std::move(index);
  r.dim = dim;
  r.func = [A, b, dim, P](const double* x, double* out, double* J) {
    for (int i = 0; i < dim; ++i) {
      double s = 0.0;
      for (int j = 0; j < P; ++j) s += A[i * P + j] * x[j];
      out[i] = s - b[i];
    }
    if (J) std::memcpy(J, A.data(), sizeof(double) * dim * P);
  };
  return r;
}

// dependencies/mini_opt/test/qp_test.cc:576-630
void TestNullSpaceSolver1() {
  Residual cost = LinearResidual({0, 1}, 2, {-2.0, 1.4, 2.2, -3.5}, {-0.8, 1.3});
  constexpr double x1_pinned_value = 0.3;
  Residual eq = LinearResidual({1}, 1, {1.0}, {x1_pinned_value});

  const Vec initial_values(2, 0.0);
  QP qp;
  qp.G.Resize(2, 2);
  qp.c.assign(2, 0.0);
  qp.A_eq.Resize(1, 2);
  qp.b_eq.assign(1, 0.0);
  cost.UpdateHessian(initial_values, &qp.G, &qp.c);
  eq.UpdateJacobian(initial_values, &qp.A_eq, &qp.b_eq, 0);

  QPNullSpaceSolver solver;
  CHECK_TRUE(solver.Solve(qp) == QPNullSpaceTerminationState::SUCCESS);
  const Vec x = solver.variables();
  CHECK_NEAR(x1_pinned_value, x[1], 1.0e-12);
  CHECK_NEAR(0.860859728506787, x[0], 1.0e-12);

  auto cost_at = [](double x0, double x1) {
    const double r0 = -2.0 * x0 + 1.4 * x1 + 0.8, r1 = 2.2 * x0 - 3.5 * x1 - 1.3;
    return r0 * r0 + r1 * r1;
  };
  CHECK_TRUE(cost_at(x[0] + 1e-5, x[1]) > cost_at(x[0], x[1]));
  CHECK_TRUE(cost_at(x[0] - 1e-5, x[1]) > cost_at(x[0], x[1]));
}

// dependencies/mini_opt/test/qp_test.cc:633-707
void TestNullSpaceSolver2() {
  Residual cost_1 = LinearResidual({0, 1}, 2, {1.7, -0.2, 2.3, 1.2}, {5.4, -3.4});
  Residual cost_2 = LinearResidual({1, 2}, 2, {-5.0, 3.3, 9.1, 1.9}, {-3.3, 4.4});
  Residual cost_3 = LinearResidual({0, 3}, 2, {0.2, -0.5, 1.1, -3.1}, {0.5, 0.0});
  Residual eq = LinearResidual({1, 3}, 2, {1.1, -1.1, 0.3, 0.6}, {1.3, -4.0});

  const Vec initial_values(4, 0.0);
  QP qp;
  qp.G.Resize(4, 4);
  qp.c.assign(4, 0.0);
  qp.A_eq.Resize(2, 4);
  qp.b_eq.assign(2, 0.0);
  cost_1.UpdateHessian(initial_values, &qp.G, &qp.c);
  cost_2.UpdateHessian(initial_values, &qp.G, &qp.c);
  cost_3.UpdateHessian(initial_values, &qp.G, &qp.c);
  eq.UpdateJacobian(initial_values, &qp.A_eq, &qp.b_eq, 0);

  QPNullSpaceSolver solver;
  CHECK_TRUE(solver.Solve(qp) == QPNullSpaceTerminationState::SUCCESS);
  const Vec x = solver.variables();
  CHECK_NEAR(-3.656565656565657, x[1], 1.0e-15);
  CHECK_NEAR(-4.838383838383838, x[3], 1.0e-15);
  CHECK_NEAR(-0.707724112814252, x[0], 1.0e-14);
  CHECK_NEAR(0.0247370254266801, x[2], 1.0e-14);
}

// dependencies/mini_opt/test/nonlinear_test.cc:109-183
void TestComputeQPCostDerivative() {
  Residual cost;
  cost.index = {0, 2, 1, 4, 3};
  cost.dim = 3;
  cost.func = [](const double* p, double* r, double* J) {
    r[0] = p[0] * p[1] - p[3];
    r[1] = std::sin(p[1]) * std::cos(p[2]);
    r[2] = p[3] * p[0] - p[2];
    if (J) {
      std::fill(J, J + 15, 0.0);
      J[0 * 5 + 0] = p[1];
      J[0 * 5 + 1] = p[0];
      J[0 * 5 + 3] = -1;
      J[1 * 5 + 1] = std::cos(p[1]) * std::cos(p[2]);
      J[1 * 5 + 2] = -std::sin(p[2]) * std::sin(p[1]);
      J[2 * 5 + 0] = p[3];
      J[2 * 5 + 2] = -1;
      J[2 * 5 + 3] = p[0];
    }
  };
  Residual eq;
  eq.index = {3, 4, 0};
  eq.dim = 2;
  eq.func = [](const double* p, double* r, double* J) {
    r[0] = p[0] - p[1];
    r[1] = -p[2] * p[2];
    if (J) {
      std::fill(J, J + 6, 0.0);
      J[0] = 1;
      J[1] = -1;
      J[1 * 3 + 2] = -2 * p[2];
    }
  };

  const Vec params = {-0.5, 0.2, 0.3, -0.8, 1.2};
  const Vec dx = {0.1, 0.25, -0.87, 1.1, -0.02};
  constexpr double penalty = 0.334;

  const double J_numerical = NumericalDerivative(0.01, [&](double alpha) {
    Vec v(5);
    for (int i = 0; i < 5; ++i) v[i] = params[i] + dx[i] * alpha;
    double c_out[3], e_out[2];
    cost.ErrorVector(v, c_out);
    eq.ErrorVector(v, e_out);
    return 0.5 * (c_out[0] * c_out[0] + c_out[1] * c_out[1] + c_out[2] * c_out[2]) +
           penalty * (std::abs(e_out[0]) + std::abs(e_out[1]));
  });

  Problem problem;
  problem.dimension = 5;
  problem.costs.push_back(cost);
  problem.equality_constraints.push_back(eq);
  QP qp;
  qp.G.Resize(5, 5);
  qp.c.assign(5, 0.0);
  qp.A_eq.Resize(2, 5);
  qp.b_eq.assign(2, 0.0);
  ConstrainedNonlinearLeastSquares::LinearizeAndFillQP(params, 0.0, problem, &qp);
  const DirectionalDerivatives J_analytical = ConstrainedNonlinearLeastSquares::ComputeQPCostDerivative(qp, dx);
  CHECK_NEAR(J_numerical, J_analytical.Total(penalty), kPico);
}

// dependencies/mini_opt/test/nonlinear_test.cc:185-207
void TestQuadraticApproxMinimum() {
  constexpr double alpha_0 = 0.8, phi_0 = 2.0, phi_prime_0 = -1.2, phi_alpha_0 = 2.2;
  const auto solution =
      ConstrainedNonlinearLeastSquares::QuadraticApproxMinimum(phi_0, phi_prime_0, alpha_0, phi_alpha_0);
  CHECK_TRUE(solution.has_value());
  const double a = (phi_alpha_0 - phi_0 - alpha_0 * phi_prime_0) / std::pow(alpha_0, 2);
  const double b = phi_prime_0;
  CHECK_NEAR(-b / (2 * a), solution.value(), kPico);
  CHECK_TRUE(!ConstrainedNonlinearLeastSquares::QuadraticApproxMinimum(phi_0, 1.3, alpha_0, phi_alpha_0));
  CHECK_TRUE(!ConstrainedNonlinearLeastSquares::QuadraticApproxMinimum(phi_0, 1.3, alpha_0,
                                                                       (phi_prime_0 * alpha_0 - phi_0) * 1.01));
}

// dependencies/mini_opt/test/nonlinear_test.cc:210-250
void TestCubicApproxCoeffs() {
  constexpr double alpha_0 = 0.8, alpha_1 = 0.4, phi_0 = 1.44, phi_prime_0 = -1.23;
  constexpr double phi_alpha_0 = 2.2, phi_alpha_1 = 1.6;
  double ab[2];
  ConstrainedNonlinearLeastSquares::CubicApproxCoeffs(phi_0, phi_prime_0, alpha_0, phi_alpha_0, alpha_1,
                                                      phi_alpha_1, ab);
  CHECK_NEAR(ab[0] * std::pow(alpha_0, 3) + ab[1] * std::pow(alpha_0, 2) + alpha_0 * phi_prime_0 + phi_0,
             phi_alpha_0, kPico);
  CHECK_NEAR(ab[0] * std::pow(alpha_1, 3) + ab[1] * std::pow(alpha_1, 2) + alpha_1 * phi_prime_0 + phi_0,
             phi_alpha_1, kPico);
  const auto min_alpha = ConstrainedNonlinearLeastSquares::CubicApproxMinimum(phi_prime_0, ab);
  CHECK_TRUE(min_alpha.has_value());
  CHECK_NEAR(0.0, 3 * ab[0] * std::pow(*min_alpha, 2) + 2 * ab[1] * *min_alpha + phi_prime_0, kPico);
  CHECK_TRUE(6 * ab[0] * *min_alpha + 2 * ab[1] > 0);
  const double bad1[2] = {0.0, ab[1]};
  CHECK_TRUE(!ConstrainedNonlinearLeastSquares::CubicApproxMinimum(phi_prime_0, bad1));
  const double bad2[2] = {(ab[1] * ab[1] / (3 * phi_prime_0)) * 1.01, ab[1]};
  CHECK_TRUE(!ConstrainedNonlinearLeastSquares::CubicApproxMinimum(phi_prime_0, bad2));
}

// dependencies/mini_opt/test/nonlinear_test.cc:745-825 — null-space QP + polynomial line search + lambda>0
void TestSphereWithNonlinearEqualityConstraints() {
  constexpr int N = 6;
  Problem problem;
  problem.dimension = N;
  {
    Residual sphere;
    sphere.index.resize(N);
    std::iota(sphere.index.begin(), sphere.index.end(), 0);
    sphere.dim = N;
    sphere.func = [](const double* x, double* r, double* J) {
      for (int i = 0; i < N; ++i) r[i] = x[i];
      if (J) {
        std::fill(J, J + N * N, 0.0);
        for (int i = 0; i < N; ++i) J[i * N + i] = 1.0;
      }
    };
    problem.costs.push_back(sphere);
  }
  auto product = [](int i0, int i1, double v) {
    Residual r;
    r.index = {i0, i1};
    r.dim = 1;
    r.func = [v](const double* x, double* out, double* J) {
      if (J) {
        J[0] = x[1];
        J[1] = x[0];
      }
      out[0] = x[0] * x[1] - v;
    };
    return r;
  };
  problem.equality_constraints.push_back(product(0, 1, 4.0));
  problem.equality_constraints.push_back(product(2, 3, 9.0));

  ConstrainedNonlinearLeastSquares nls(&problem);
  ConstrainedNonlinearLeastSquares::Params p{};
  p.max_iterations = 100;
  p.max_qp_iterations = 1;
  p.relative_exit_tol = kPico;
  p.absolute_first_derivative_tol = kNano;
  p.termination_kkt_tolerance = kMicro;
  p.lambda_initial = 0.001;

  std::default_random_engine engine{7};
  std::uniform_real_distribution<double> dist{-30.0, 30.0};
  std::vector<Vec> guesses;
  for (int i = 0; i < 100; ++i) {
    Vec g(N);
    for (int j = 0; j < N; ++j) g[j] = dist(engine);
    guesses.push_back(g);
  }
  std::vector<Vec> solutions;
  for (double x0 : {-2.0, 2.0})
    for (double x2 : {-3.0, 3.0}) solutions.push_back({x0, x0, x2, x2, 0.0, 0.0});

  std::size_t total_iters = 0;
  for (const Vec& guess : guesses) {
    const NLSSolverOutputs outputs = nls.Solve(p, guess);
    total_iters += outputs.iterations.size();
    const bool satisfied = outputs.termination_state == NLSTerminationState::SATISFIED_ABSOLUTE_TOL ||
                           outputs.termination_state == NLSTerminationState::SATISFIED_RELATIVE_TOL ||
                           outputs.termination_state == NLSTerminationState::SATISFIED_FIRST_ORDER_TOL;
    CHECK_TRUE(satisfied);
    double best = 1e300;
    const Vec* best_sol = nullptr;
    for (const Vec& s : solutions) {
      double d = 0.0;
      for (int i = 0; i < N; ++i) d += (s[i] - nls.variables()[i]) * (s[i] - nls.variables()[i]);
      if (d < best) {
        best = d;
        best_sol = &s;
      }
    }
    for (int i = 0; i < N; ++i) CHECK_NEAR((*best_sol)[i], nls.variables()[i], 5.0e-5);
    CHECK_TRUE(outputs.NumFailedLineSearches() == 0);
  }
  std::printf("  [%s] mean SQP iterations over 100 starts: %.2f\n", g_case, total_iters / 100.0);
}

// optimization/integration_test.cc:45-80
void TestIntegrationDerivatives() {
  const double x[4] = {1.2, 0.7, 0.4, -0.15};
  constexpr double u = 0.1, dt = 0.01;
  const SingleCartPoleParams params{1.0, 0.1, 0.25, 9.81, 0.0, 0.1, 0.0, 0.8, 10.0};
  const double zero[2] = {0.0, 0.0};
  double x_new[4], D_x[16], D_u[4];
  RungeKutta4(params, x, u, dt, x_new, D_x, D_u);

  double sq_x = 0.0, sq_u = 0.0;
  for (int j = 0; j < 4; ++j) {
    for (int i = 0; i < 4; ++i) {
      const double num = NumericalDerivative(0.01, [&](double d) {
        double xp[4] = {x[0], x[1], x[2], x[3]};
        xp[j] += d;
        double out[4];
        RungeKutta4NoJacobians(params, xp, u, dt, zero, zero, out);
        return out[i];
      });
      sq_x += (D_x[i * 4 + j] - num) * (D_x[i * 4 + j] - num);
    }
  }
  for (int i = 0; i < 4; ++i) {
    const double num = NumericalDerivative(0.01, [&](double d) {
      double out[4];
      RungeKutta4NoJacobians(params, x, u + d, dt, zero, zero, out);
      return out[i];
    });
    sq_u += (D_u[i] - num) * (D_u[i] - num);
  }
  CHECK_TRUE(std::sqrt(sq_x) < 1.0e-12);
  CHECK_TRUE(std::sqrt(sq_u) < 1.0e-12);
  // the value part of the Jacobian variant must equal the plain integrator
  double x_plain[4];
  RungeKutta4NoJacobians(params, x, u, dt, zero, zero, x_plain);
  for (int i = 0; i < 4; ++i) CHECK_TRUE(x_plain[i] == x_new[i]);
}

// Same check as above but with friction, drag and an engaged bumper spring so every Jacobian branch of
// single_pendulum_dynamics.hpp:104-178 is exercised (stronger than the reference's case; looser FD tolerance).
void TestIntegrationDerivativesAllTerms() {
  const double x[4] = {0.85, -0.9, 0.6, 1.3};
  constexpr double u = -2.5, dt = 0.01;
  const SingleCartPoleParams params{1.3, 0.2, 0.3, 9.81, 0.05, 0.1, 0.02, 0.8, 100.0};
  const double zero[2] = {0.0, 0.0};
  double x_new[4], D_x[16], D_u[4];
  RungeKutta4(params, x, u, dt, x_new, D_x, D_u);
  for (int j = 0; j < 4; ++j) {
    for (int i = 0; i < 4; ++i) {
      const double num = NumericalDerivative(1e-3, [&](double d) {
        double xp[4] = {x[0], x[1], x[2], x[3]};
        xp[j] += d;
        double out[4];
        RungeKutta4NoJacobians(params, xp, u, dt, zero, zero, out);
        return out[i];
      });
      CHECK_NEAR(num, D_x[i * 4 + j], 1.0e-9);
    }
  }
  for (int i = 0; i < 4; ++i) {
    const double num = NumericalDerivative(1e-3, [&](double d) {
      double out[4];
      RungeKutta4NoJacobians(params, x, u + d, dt, zero, zero, out);
      return out[i];
    });
    CHECK_NEAR(num, D_u[i], 1.0e-9);
  }
}

// optimization/integration_test.cc:82-102
void TestFrictionDissipation() {
  const SingleCartPoleParams params{1.0, 0.5, 0.4, 9.81, 0.5, 0.1, 0.0, 0.0, 0.0};
  const double zero[2] = {0.0, 0.0};
  double x[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < 20000; ++i) {
    double xn[4];
    RungeKutta4NoJacobians(params, x, 0.0, 0.01, zero, zero, xn);
    std::memcpy(x, xn, sizeof(x));
  }
  CHECK_NEAR(0.0, x[2], 1.0e-6);
  CHECK_NEAR(0.0, x[3], 1.0e-4);
}

// optimization/integration_test.cc:104-125
void TestDragDissipation() {
  const SingleCartPoleParams params{0.8, 0.1, 0.4, 9.81, 0.01, 0.1, 5.0, 0.0, 0.0};
  const double zero[2] = {0.0, 0.0};
  double x[4] = {0.0, -M_PI, 0.0, 0.0};
  for (int i = 0; i < 10000; ++i) {
    double xn[4];
    RungeKutta4NoJacobians(params, x, 0.0, 0.01, zero, zero, xn);
    std::memcpy(x, xn, sizeof(x));
  }
  CHECK_NEAR(0.0, x[2], 1.0e-6);
  CHECK_NEAR(0.0, x[3], 3.0e-5);
}

// optimization/integration_test.cc:127-175
void TestExternalForceSymmetry() {
  const SingleCartPoleParams params{1.0, 0.1, 0.25, 9.81, 0.1, 0.1, 0.0, 0.0, 0.0};
  const double zero[2] = {0.0, 0.0};
  constexpr int num_iterations = 3000, force_duration = 500;
  constexpr double force_magnitude = 5.0;
  double xl[4] = {0.0, -M_PI / 2, 0.0, 0.0}, xr[4] = {0.0, -M_PI / 2, 0.0, 0.0};
  for (int i = 0; i < num_iterations; ++i) {
    const double fl[2] = {i < force_duration ? force_magnitude : 0.0, 0.0};
    const double fr[2] = {i < force_duration ? -force_magnitude : 0.0, 0.0};
    double xn[4];
    RungeKutta4NoJacobians(params, xl, 0.0, 0.001, fl, zero, xn);
    std::memcpy(xl, xn, sizeof(xl));
    RungeKutta4NoJacobians(params, xr, 0.0, 0.001, fr, zero, xn);
    std::memcpy(xr, xn, sizeof(xr));
  }
  CHECK_TRUE(xl[0] > 0);
  CHECK_TRUE(xr[0] < 0);
  CHECK_NEAR(xl[0], -xr[0], 1.0e-12);
  CHECK_NEAR(xl[2], -xr[2], 1.0e-12);
  CHECK_NEAR(-M_PI / 2 - xl[1], xr[1] - -M_PI / 2, 1.0e-12);
  CHECK_NEAR(xl[3], -xr[3], 1.0e-12);
}

// optimization/optimization_test.cc:12-77
void TestCartPoleMultipleShootingClosedLoop() {
  constexpr std::size_t num_steps = 200;
  OptimizationParams op{};
  op.control_dt = 0.01;
  op.window_length = 40;
  op.state_spacing = 5;
  op.max_iterations = 10;
  const SingleCartPoleParams dyn{1.0, 0.1, 0.25, 9.81, 0.03, 0.1, 0.13, 0.8, 100.0};
  const SingleCartPoleState x0{0.0, -M_PI / 2, 0.0, 0.0};
  const double zero[2] = {0.0, 0.0};

  Simulator sim;
  sim.SetState(x0);
  Optimization optimization{op};
  std::size_t iters = 0, ls = 0;
  int hist[7] = {0, 0, 0, 0, 0, 0, 0};
  for (std::size_t t = 0; t < num_steps; ++t) {
    const OptimizationOutputs outputs = optimization.Step(sim.GetState(), dyn, 0.0);
    CHECK_TRUE(outputs.solver_outputs.termination_state != NLSTerminationState::QP_INDEFINITE);
    CHECK_TRUE(outputs.solver_outputs.termination_state != NLSTerminationState::MAX_LAMBDA);
    iters += outputs.solver_outputs.iterations.size();
    ls += outputs.solver_outputs.NumLineSearchSteps();
    hist[static_cast<int>(outputs.solver_outputs.termination_state)]++;
    const SingleCartPoleState& terminal = outputs.predicted_states.back();
    if (t > 20) {
      CHECK_NEAR(0.0, terminal.b_x_dot, 1.0e-4);
      CHECK_NEAR(0.0, terminal.th_1_dot, 1.0e-4);
      CHECK_NEAR(M_PI / 2, terminal.th_1, 1.0e-4);
    }
    sim.Step(dyn, op.control_dt, outputs.u.front(), zero, zero);
  }
  // NOTE: the reference pushes sim.GetState() *before* the last sim.Step, so `states.back()` is the state at the
  // start of step 199 (optimization_test.cc:57-60).  Check both that one and the truly final one.
  const SingleCartPoleState final_state = sim.GetState();
  CHECK_NEAR(0.0, final_state.b_x_dot, 1.0e-4);
  CHECK_NEAR(0.0, final_state.th_1_dot, 1.0e-3);
  CHECK_NEAR(M_PI / 2, final_state.th_1, 1.0e-4);
  std::printf("  [%s] mean SQP iterations %.2f, merit evals %.2f; term hist: max_iter %d abs %d rel %d first %d\n",
              g_case, iters / 200.0, ls / 200.0, hist[0], hist[1], hist[2], hist[3]);
}

struct Case {
  const char* name;
  void (*fn)();
};
const Case kCases[] = {
    {"null_space_solver_1", TestNullSpaceSolver1},
    {"null_space_solver_2", TestNullSpaceSolver2},
    {"qp_cost_derivative", TestComputeQPCostDerivative},
    {"quadratic_approx_minimum", TestQuadraticApproxMinimum},
    {"cubic_approx_coeffs", TestCubicApproxCoeffs},
    {"sphere_nonlinear_equality", TestSphereWithNonlinearEqualityConstraints},
    {"integration_derivatives", TestIntegrationDerivatives},
    {"integration_derivatives_all_terms", TestIntegrationDerivativesAllTerms},
    {"friction_dissipation", TestFrictionDissipation},
    {"drag_dissipation", TestDragDissipation},
    {"external_force_symmetry", TestExternalForceSymmetry},
    {"cartpole_closed_loop", TestCartPoleMultipleShootingClosedLoop},
};

}  // namespace

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--list") {
    for (const Case& c : kCases) std::printf("%s\n", c.name);
    return 0;
  }
  int ran = 0;
  for (const Case& c : kCases) {
    if (argc > 1 && std::string(argv[1]) != c.name) continue;
    g_case = c.name;
    const int before = g_failures;
    c.fn();
    std::printf("%s %s\n", g_failures == before ? "PASS" : "FAIL", c.name);
    ++ran;
  }
  if (ran == 0) {
    std::printf("no such case\n");
    return 99;
  }
  return g_failures;
}
