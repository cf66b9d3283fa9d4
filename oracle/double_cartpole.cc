// ORACLE — TEST INFRASTRUCTURE ONLY (see double_cartpole.hpp).
#include "double_cartpole.hpp"

#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "cartpole.hpp"  // ModPi
#include "double_pendulum_gen.hpp"

namespace oracle {

namespace {

constexpr int NX = kDpNx;
constexpr bool kIsAngle[NX] = {false, true, true, false, false, false};

void Dyn(const DoubleCartPoleParams& p, const double x[NX], const double u, double x_dot[NX], double* J_x,
         double* J_u) {
  const double pp[6] = {p.m_b, p.m_1, p.m_2, p.l_1, p.l_2, p.g};
  DoublePendulumDynamics(pp, x, u, x_dot, J_x, J_u);
}

void MulNN(const double* a, const double* b, double* out) {
  for (int i = 0; i < NX; ++i)
    for (int j = 0; j < NX; ++j) {
      double s = 0.0;
      for (int k = 0; k < NX; ++k) s += a[i * NX + k] * b[k * NX + j];
      out[i * NX + j] = s;
    }
}
void MulN1(const double* a, const double* v, double* out) {
  for (int i = 0; i < NX; ++i) {
    double s = 0.0;
    for (int k = 0; k < NX; ++k) s += a[i * NX + k] * v[k];
    out[i] = s;
  }
}

// [x(0) .. x(NS-1) | u(0) .. u(W-1)]   (optimization.cc:24-37 with 6 states per node)
int StateIndex(const int comp, const std::size_t node) { return static_cast<int>(node) * NX + comp; }
int ControlIndex(const std::size_t k, const std::size_t num_states) {
  return static_cast<int>(num_states) * NX + static_cast<int>(k);
}

Residual ScalarResidual(const int index, const bool wrap, const double desired, const double weight) {
  Residual r;
  r.index = {index};
  r.dim = 1;
  r.func = [desired, weight, wrap](const double* vars, double* out, double* J) {
    if (J) J[0] = weight;
    out[0] = wrap ? ModPi(vars[0] - desired) * weight : (vars[0] - desired) * weight;
  };
  return r;
}

// Shooting-segment defect and its 6 x (12+S) Jacobian (optimization.cc:99-160 with 6 states)
Residual DynamicalConstraint(std::vector<int> indices, const DoubleCartPoleParams params,
                             const std::size_t state_spacing, const double dt) {
  Residual r;
  r.index = std::move(indices);
  r.dim = NX;
  r.func = [params, state_spacing, dt](const double* vars, double* out, double* J) {
    const double* x_k = vars;
    const double* x_kpl = vars + NX;
    const double* u_k = vars + 2 * NX;
    const int P = 2 * NX + static_cast<int>(state_spacing);
    double x[NX];
    for (int j = 0; j < NX; ++j) x[j] = x_k[j];
    std::vector<double> step_D_x, step_D_u;
    if (J) {
      step_D_x.resize(NX * NX * state_spacing);
      step_D_u.resize(NX * state_spacing);
    }
    for (std::size_t i = 0; i < state_spacing; ++i) {
      double x_next[NX];
      if (J) {
        DoubleRungeKutta4(params, x, u_k[i], dt, x_next, &step_D_x[NX * NX * i], &step_D_u[NX * i]);
      } else {
        DoubleRungeKutta4NoJacobians(params, x, u_k[i], dt, x_next);
      }
      for (int j = 0; j < NX; ++j) x[j] = x_next[j];
    }
    for (int j = 0; j < NX; ++j)
      if (kIsAngle[j]) x[j] = ModPi(x[j]);
    if (J) {
      for (int i = 0; i < NX * P; ++i) J[i] = 0.0;
      double chain[NX * NX];
      for (int i = 0; i < NX * NX; ++i) chain[i] = (i % (NX + 1) == 0) ? 1.0 : 0.0;
      for (int i = static_cast<int>(state_spacing) - 1; i >= 0; --i) {
        double col[NX];
        MulN1(chain, &step_D_u[NX * i], col);
        for (int rr = 0; rr < NX; ++rr) J[rr * P + 2 * NX + i] = col[rr];
        double next[NX * NX];
        MulNN(chain, &step_D_x[NX * NX * i], next);
        for (int j = 0; j < NX * NX; ++j) chain[j] = next[j];
      }
      for (int rr = 0; rr < NX; ++rr)
        for (int cc = 0; cc < NX; ++cc) {
          J[rr * P + cc] = chain[rr * NX + cc];
          J[rr * P + NX + cc] = (rr == cc) ? -1.0 : 0.0;
        }
    }
    for (int j = 0; j < NX; ++j) {
      out[j] = x[j] - x_kpl[j];
      if (kIsAngle[j]) out[j] = ModPi(out[j]);
    }
  };
  return r;
}

}  // namespace

void DoubleRungeKutta4NoJacobians(const DoubleCartPoleParams& p, const double x[6], const double u, const double h,
                                  double x_new[6]) {
  double k1[NX], k2[NX], k3[NX], k4[NX], arg[NX];
  Dyn(p, x, u, k1, nullptr, nullptr);
  for (int i = 0; i < NX; ++i) arg[i] = x[i] + k1[i] * h / 2.0;
  Dyn(p, arg, u, k2, nullptr, nullptr);
  for (int i = 0; i < NX; ++i) arg[i] = x[i] + k2[i] * h / 2.0;
  Dyn(p, arg, u, k3, nullptr, nullptr);
  for (int i = 0; i < NX; ++i) arg[i] = x[i] + k3[i] * h;
  Dyn(p, arg, u, k4, nullptr, nullptr);
  for (int i = 0; i < NX; ++i) x_new[i] = x[i] + (h / 6.0) * (k1[i] + k2[i] * 2.0 + k3[i] * 2.0 + k4[i]);
}

void DoubleRungeKutta4(const DoubleCartPoleParams& p, const double x[6], const double u, const double h,
                       double x_new[6], double x_new_D_x[36], double x_new_D_u[6]) {
  constexpr int NN = NX * NX;
  double k1[NX], k2[NX], k3[NX], k4[NX], arg[NX];
  double k1_D_x[NN], k2_D_arg[NN], k3_D_arg[NN], k4_D_arg[NN];
  double k1_D_u[NX], k2_D_u_arg[NX], k3_D_u_arg[NX], k4_D_u_arg[NX];
  Dyn(p, x, u, k1, k1_D_x, k1_D_u);
  for (int i = 0; i < NX; ++i) arg[i] = x[i] + k1[i] * h / 2.0;
  Dyn(p, arg, u, k2, k2_D_arg, k2_D_u_arg);
  for (int i = 0; i < NX; ++i) arg[i] = x[i] + k2[i] * h / 2.0;
  Dyn(p, arg, u, k3, k3_D_arg, k3_D_u_arg);
  for (int i = 0; i < NX; ++i) arg[i] = x[i] + k3[i] * h;
  Dyn(p, arg, u, k4, k4_D_arg, k4_D_u_arg);
  for (int i = 0; i < NX; ++i) x_new[i] = x[i] + (h / 6.0) * (k1[i] + k2[i] * 2.0 + k3[i] * 2.0 + k4[i]);

  // chain rule through the stages (integration.hpp:36-46)
  const auto eye = [](int i) { return (i % (NX + 1) == 0) ? 1.0 : 0.0; };
  double inner[NN], k2_D_x[NN], k3_D_x[NN], k4_D_x[NN];
  for (int i = 0; i < NN; ++i) inner[i] = eye(i) + k1_D_x[i] * h / 2.0;
  MulNN(k2_D_arg, inner, k2_D_x);
  for (int i = 0; i < NN; ++i) inner[i] = eye(i) + k2_D_x[i] * h / 2.0;
  MulNN(k3_D_arg, inner, k3_D_x);
  for (int i = 0; i < NN; ++i) inner[i] = eye(i) + k3_D_x[i] * h;
  MulNN(k4_D_arg, inner, k4_D_x);
  double tmp[NX], k2_D_u[NX], k3_D_u[NX], k4_D_u[NX];
  MulN1(k2_D_arg, k1_D_u, tmp);
  for (int i = 0; i < NX; ++i) k2_D_u[i] = tmp[i] * (h / 2.0) + k2_D_u_arg[i];
  MulN1(k3_D_arg, k2_D_u, tmp);
  for (int i = 0; i < NX; ++i) k3_D_u[i] = tmp[i] * (h / 2.0) + k3_D_u_arg[i];
  MulN1(k4_D_arg, k3_D_u, tmp);
  for (int i = 0; i < NX; ++i) k4_D_u[i] = tmp[i] * h + k4_D_u_arg[i];
  for (int i = 0; i < NN; ++i)
    x_new_D_x[i] = eye(i) + (h / 6.0) * (k1_D_x[i] + k2_D_x[i] * 2.0 + k3_D_x[i] * 2.0 + k4_D_x[i]);
  for (int i = 0; i < NX; ++i)
    x_new_D_u[i] = (h / 6.0) * (k1_D_u[i] + k2_D_u[i] * 2.0 + k3_D_u[i] * 2.0 + k4_D_u[i]);
}

DoubleOptimization::DoubleOptimization(const DoubleOptimizationParams& params) : params_(params) {
  // optimization.cc:13-22
  if (!(params.control_dt > 0)) throw std::invalid_argument("control_dt must be > 0");
  if (!(params.window_length >= 1)) throw std::invalid_argument("window_length must be >= 1");
  if (params.state_spacing == 0 || params.window_length % params.state_spacing != 0)
    throw std::invalid_argument("state_spacing must divide into window_length cleanly");
  if (!(params.max_iterations >= 1)) throw std::invalid_argument("max_iterations must be >= 1");
  if (!(params.u_cost_weight >= 0.0)) throw std::invalid_argument("u_cost_weight must be >= 0");
  if (!(params.u_derivative_cost_weight >= 0.0)) throw std::invalid_argument("u_derivative_cost_weight must be >= 0");
}

DoubleOptimizationOutputs DoubleOptimization::Step(const double state[6], const DoubleCartPoleParams& dyn,
                                                   const double b_x_set_point) {
  BuildProblem(state, dyn, b_x_set_point);
  const std::size_t NS = params_.NumStates(), W = params_.window_length, S = params_.state_spacing;
  Vec guess(problem_.dimension, 0.0);
  if (!previous_solution_.empty()) {
    guess = previous_solution_;  // optimization.cc:50-57
    for (int i = 0; i < NX; ++i) guess[StateIndex(i, 0)] = state[i];
    for (std::size_t k = 0; k + 1 < W; ++k) guess[ControlIndex(k, NS)] = guess[ControlIndex(k + 1, NS)];
  } else {
    for (int i = 0; i < NX; ++i) guess[StateIndex(i, 0)] = state[i];  // optimization.cc:59-67
    for (std::size_t k = 0; k < W; ++k)
      guess[ControlIndex(k, NS)] = params_.u_guess_sinusoid_amplitude *
                                   std::sin(static_cast<double>(k) / static_cast<double>(W) * 2 * M_PI);
  }
  // FillInitialGuess (optimization.cc:333-351)
  {
    double x[NX];
    for (int i = 0; i < NX; ++i) x[i] = guess[StateIndex(i, 0)];
    for (std::size_t s = 1; s < NS; ++s) {
      for (std::size_t k = 0; k < S; ++k) {
        double xn[NX];
        DoubleRungeKutta4NoJacobians(dyn, x, guess[ControlIndex((s - 1) * S + k, NS)], params_.control_dt, xn);
        for (int i = 0; i < NX; ++i) x[i] = kIsAngle[i] ? ModPi(xn[i]) : xn[i];
      }
      for (int i = 0; i < NX; ++i) guess[StateIndex(i, s)] = x[i];
    }
  }
  ConstrainedNonlinearLeastSquares::Params p{};  // optimization.cc:73-79
  p.max_iterations = static_cast<int>(params_.max_iterations);
  p.max_line_search_iterations = 5;
  p.absolute_first_derivative_tol = params_.absolute_first_derivative_tol;
  p.equality_penalty_initial = params_.equality_penalty_initial;
  p.relative_exit_tol = params_.relative_exit_tol;
  NLSSolverOutputs outputs = solver_->Solve(p, guess);

  std::vector<double> previous_solution = previous_solution_;
  previous_solution_ = solver_->variables();
  const double* u_out = previous_solution_.data() + (previous_solution_.size() - W);

  DoubleOptimizationOutputs result;
  result.previous_solution = std::move(previous_solution);
  result.final_errors = solver_->EvaluateNonlinearErrors(solver_->variables());
  result.final_penalty = outputs.iterations.empty() ? p.equality_penalty_initial : outputs.iterations.back().penalty;
  result.solver_outputs = std::move(outputs);
  result.u.assign(u_out, u_out + W);
  // ComputePredictedStates (optimization.cc:353-371)
  result.predicted_states.resize(W * NX);
  double x[NX];
  for (int i = 0; i < NX; ++i) x[i] = state[i];
  for (std::size_t k = 0; k < W; ++k) {
    double xn[NX];
    DoubleRungeKutta4NoJacobians(dyn, x, u_out[k], params_.control_dt, xn);
    for (int i = 0; i < NX; ++i) {
      x[i] = kIsAngle[i] ? ModPi(xn[i]) : xn[i];
      result.predicted_states[k * NX + i] = x[i];
    }
  }
  return result;
}

void DoubleOptimization::BuildProblem(const double state[6], const DoubleCartPoleParams& dyn,
                                      const double b_x_set_point) {
  // optimization.cc:194-331 with 6 states per node
  const std::size_t NS = params_.NumStates(), W = params_.window_length, S = params_.state_spacing;
  problem_.clear();
  problem_.dimension = static_cast<int>(NS) * NX + static_cast<int>(W);
  for (std::size_t s = 0; s + 1 < NS; ++s) {
    std::vector<int> indices;
    for (const std::size_t node : {s, s + 1})
      for (int i = 0; i < NX; ++i) indices.push_back(StateIndex(i, node));
    for (std::size_t k = 0; k < S; ++k) indices.push_back(ControlIndex(S * s + k, NS));
    problem_.equality_constraints.push_back(DynamicalConstraint(std::move(indices), dyn, S, params_.control_dt));
  }
  for (int i = 0; i < NX; ++i)
    problem_.equality_constraints.push_back(ScalarResidual(StateIndex(i, 0), kIsAngle[i], state[i], 1.0));
  const double target[NX] = {b_x_set_point, M_PI / 2, M_PI / 2, 0.0, 0.0, 0.0};
  for (int i = 0; i < NX; ++i) {
    const double w = params_.final_cost_weight[i];
    if (w >= 0.0) {
      problem_.costs.push_back(ScalarResidual(StateIndex(i, NS - 1), kIsAngle[i], target[i], w));
    } else {
      problem_.equality_constraints.push_back(ScalarResidual(StateIndex(i, NS - 1), kIsAngle[i], target[i], 1.0));
    }
  }
  if (params_.u_derivative_cost_weight > 0.0) {
    const double weight = params_.u_derivative_cost_weight;
    for (std::size_t k = 0; k + 1 < W; ++k) {
      Residual r;
      r.index = {ControlIndex(k, NS), ControlIndex(k + 1, NS)};
      r.dim = 1;
      r.func = [weight](const double* uu, double* out, double* J) {
        if (J) {
          J[0] = weight;
          J[1] = -weight;
        }
        out[0] = (uu[0] - uu[1]) * weight;
      };
      problem_.costs.push_back(std::move(r));
    }
    const double u_prev = !previous_solution_.empty() ? previous_solution_[ControlIndex(0, NS)] : 0.0;
    problem_.costs.push_back(ScalarResidual(ControlIndex(0, NS), false, u_prev, weight));
  }
  if (params_.u_cost_weight > 0.0)
    for (std::size_t k = 0; k < W; ++k)
      problem_.costs.push_back(ScalarResidual(ControlIndex(k, NS), false, 0.0, params_.u_cost_weight));
  if (solver_) return;
  const auto retraction = [NS, W](Vec& x, const Vec& dx, const double alpha) {  // optimization.cc:309-329
    for (std::size_t i = 0; i < x.size(); ++i) x[i] += dx[i] * alpha;
    for (std::size_t s = 0; s < NS; ++s) {
      for (int i = 0; i < NX; ++i)
        if (kIsAngle[i]) x[StateIndex(i, s)] = ModPi(x[StateIndex(i, s)]);
      x[StateIndex(0, s)] = std::clamp(x[StateIndex(0, s)], -5.0, 5.0);
    }
    for (std::size_t k = 0; k < W; ++k) x[ControlIndex(k, NS)] = std::clamp(x[ControlIndex(k, NS)], -300.0, 300.0);
  };
  solver_ = std::make_unique<ConstrainedNonlinearLeastSquares>(&problem_, retraction);
}

}  // namespace oracle
