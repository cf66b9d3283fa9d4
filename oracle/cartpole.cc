// ORACLE — TEST INFRASTRUCTURE ONLY (see cartpole.hpp header for scope and citations).
#include "cartpole.hpp"

#include <cmath>
#include <stdexcept>

namespace oracle {

// ---------------------------------------------------------------------------------------------------------------------
// Dynamics.  Restates the expression graph of optimization/single_pendulum_dynamics.hpp:33-185 with physical
// names; the association of every product/sum follows the generated code so results agree to the last bit when
// compiled without FMA contraction.
// ---------------------------------------------------------------------------------------------------------------------
void SinglePendulumDynamics(const SingleCartPoleParams& params, const double x[4], const double u,
                            const double f_base[2], const double f_mass[2], double x_dot[4], double* const J_x,
                            double* const J_u) {
  const double bumper = params.x_s;
  const double b_x = x[0];
  // one-sided bumper springs at +-x_s (hpp:33-50)
  const double over_right = b_x + -bumper;
  const bool right_active = 0.0 < over_right;
  const double pen_right = right_active ? over_right : 0.0;
  const double over_left = -(b_x + bumper);
  const bool left_active = 0.0 < over_left;
  const double pen_left = left_active ? over_left : 0.0;
  // smoothed-Coulomb cutoff velocity floored at 1e-6 (hpp:51-57)
  const double v_cut = (1e-06 < params.v_mu_b) ? params.v_mu_b : 1e-06;

  const double th = x[1];
  const double s = std::sin(th);
  const double len = params.l_1;
  const double th_dot = x[3];
  const double neg_s = -s;
  const double lw = th_dot * len;  // tangential speed of the mass about the pivot
  const double c = std::cos(th);
  const double b_x_dot = x[2];
  const double cc = c * c;
  const double vx = b_x_dot + lw * neg_s;  // mass velocity, x component
  const double lw_cc = lw * cc;
  const double speed_sq = lw * lw_cc + vx * vx;  // |mass velocity|^2
  const double speed = std::sqrt(speed_sq);
  const double speed15 = 1.5 * speed;
  const double vx2 = vx * 2.0;
  const double len2 = len * 2.0;
  const double proj = vx * neg_s + lw_cc;
  const bool moving = 0.0 < speed_sq;
  double drag_x, drag_th;
  if (moving) {
    drag_x = vx2 * speed15;
    drag_th = proj * len2 * speed15;
  } else {
    drag_x = 0.0;
    drag_th = 0.0;
  }
  const double m1 = params.m_1;
  const double grav = params.g;
  const double fm_y = f_mass[1];
  const double fm_x = f_mass[0];
  const double m_total = m1 + params.m_b;
  const double inv_v_cut = 1.0 / v_cut;
  const double s_m1 = s * m1;
  const double drag_k = (-0.16666666666666666) * params.c_d_1;
  const double fric_k = (m_total * params.mu_b) * -grav;
  const double fric_tanh = std::tanh(b_x_dot * inv_v_cut);
  const double inv_len = 1.0 / len;
  const double fb_x = f_base[0];
  const double det = m1 * (m_total + s_m1 * neg_s);
  const double m_over_l2 = m_total * (1.0 / (len * len));
  const double q_th = drag_th * drag_k + len * (fm_x * neg_s + c * (fm_y + grav * -m1));
  const double s_m1_over_l = inv_len * s_m1;
  const double q_x = u + fb_x + fm_x + fric_tanh * fric_k + drag_x * drag_k + m1 * lw * (th_dot * c) +
                     params.k_s * (pen_left + -pen_right);
  const double s_over_l = s * inv_len;
  const double inv_det = 1.0 / det;
  const double inv_det_l = inv_det * inv_len;
  const double num_th = q_x * s_m1_over_l + q_th * m_over_l2;
  const double num_x = q_x + q_th * s_over_l;
  const double ju_th = s_m1 * inv_det_l;
  const double ju_x = m1 * inv_det;

  if (J_x != nullptr) {
    const double sw_right = right_active ? 1.0 : 0.0;
    const double sw_left = left_active ? -1.0 : 0.0;
    const double neg_c = -c;
    double d_th_q, d_th_x, d_v_l, d_v_x, d_w_th;
    if (moving) {
      const double inv_speed = 1.0 / speed;
      const double k075 = 0.75;
      const double k15 = 2.0 * k075;
      const double proj_n = k15 * (inv_speed * proj);
      const double vx_rel = vx + s * lw;
      const double speed3 = speed * 3.0;
      d_th_q = -((lw * proj_n + speed15) * (c * len) * (2.0 * vx_rel));
      d_th_x = (inv_speed * vx_rel * (vx2 * k15) + speed3) * lw * neg_c;
      d_v_l = len * (speed3 * neg_s + vx2 * proj_n);
      d_v_x = vx2 * vx2 * (k075 * inv_speed) + speed3;
      d_w_th = len * (proj * proj_n + (cc + s * s) * speed15) * len2;
    } else {
      d_th_q = 0.0;
      d_th_x = 0.0;
      d_v_l = 0.0;
      d_v_x = 0.0;
      d_w_th = 0.0;
    }
    const double dq_th_dth = d_th_q * drag_k + len * (fm_x * neg_c + s * (m1 * grav + -fm_y));
    const double dq_x_dth = lw * s_m1 * -th_dot + d_th_x * drag_k;
    const double m1_c = m1 * c;
    const double dv_l_k = d_v_l * drag_k;
    const double inv_det_sq = 1.0 / (det * det);
    const double dq_x_dw = 2.0 * lw * m1_c + dv_l_k;
    const double dw_th_k = d_w_th * drag_k;
    const double dq_x_dv = inv_v_cut * (1.0 + fric_tanh * -fric_tanh) * fric_k + d_v_x * drag_k;
    const double spring_sw = params.k_s * (sw_left + -sw_right);
    J_x[0] = 0.0;
    J_x[1] = 0.0;
    J_x[2] = 1.0;
    J_x[3] = 0.0;
    J_x[4] = 0.0;
    J_x[5] = 0.0;
    J_x[6] = 0.0;
    J_x[7] = 1.0;
    J_x[8] = ju_x * spring_sw;
    J_x[9] = m1 * (num_x * s_m1 * m1_c * (2.0 * inv_det_sq) +
                   inv_det * (dq_x_dth + inv_len * (s * dq_th_dth + c * q_th)));
    J_x[10] = (dq_x_dv + s_over_l * dv_l_k) * ju_x;
    J_x[11] = (dq_x_dw + s_over_l * dw_th_k) * ju_x;
    J_x[12] = spring_sw * ju_th;
    J_x[13] = inv_det * dq_th_dth * m_over_l2 +
              m1 * (inv_det_sq * num_th * s_m1 * (c * 2.0) + (c * q_x + s * dq_x_dth) * inv_det_l);
    J_x[14] = inv_det * (m_over_l2 * dv_l_k + dq_x_dv * s_m1_over_l);
    J_x[15] = inv_det * (m_over_l2 * dw_th_k + dq_x_dw * s_m1_over_l);
  }
  if (J_u != nullptr) {
    J_u[0] = 0.0;
    J_u[1] = 0.0;
    J_u[2] = ju_x;
    J_u[3] = ju_th;
  }
  x_dot[0] = b_x_dot;
  x_dot[1] = th_dot;
  x_dot[2] = num_x * ju_x;
  x_dot[3] = inv_det * num_th;
}

double ModPi(double angle) {
  constexpr double pi = M_PI;
  constexpr double two_pi = 2 * pi;
  angle = std::fmod(angle, two_pi);
  angle += (angle < 0) * two_pi;
  angle -= (angle > pi) * two_pi;
  return angle;
}

void RungeKutta4NoJacobians(const SingleCartPoleParams& params, const double x[4], const double u, const double h,
                            const double f_base[2], const double f_mass[2], double x_new[4]) {
  double k1[4], k2[4], k3[4], k4[4], arg[4];
  SinglePendulumDynamics(params, x, u, f_base, f_mass, k1, nullptr, nullptr);
  for (int i = 0; i < 4; ++i) arg[i] = x[i] + k1[i] * h / 2.0;
  SinglePendulumDynamics(params, arg, u, f_base, f_mass, k2, nullptr, nullptr);
  for (int i = 0; i < 4; ++i) arg[i] = x[i] + k2[i] * h / 2.0;
  SinglePendulumDynamics(params, arg, u, f_base, f_mass, k3, nullptr, nullptr);
  for (int i = 0; i < 4; ++i) arg[i] = x[i] + k3[i] * h;
  SinglePendulumDynamics(params, arg, u, f_base, f_mass, k4, nullptr, nullptr);
  for (int i = 0; i < 4; ++i) x_new[i] = x[i] + (h / 6.0) * (k1[i] + k2[i] * 2.0 + k3[i] * 2.0 + k4[i]);
}

// out = a * b for row-major 4x4.
static void Mul44(const double a[16], const double b[16], double out[16]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      out[i * 4 + j] = s;
    }
}
static void Mul41(const double a[16], const double v[4], double out[4]) {
  for (int i = 0; i < 4; ++i) {
    double s = 0.0;
    for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * v[k];
    out[i] = s;
  }
}

void RungeKutta4(const SingleCartPoleParams& params, const double x[4], const double u, const double h,
                 double x_new[4], double x_new_D_x[16], double x_new_D_u[4]) {
  const double zero[2] = {0.0, 0.0};
  double k1[4], k2[4], k3[4], k4[4], arg[4];
  double k1_D_x[16], k2_D_arg[16], k3_D_arg[16], k4_D_arg[16];
  double k1_D_u[4], k2_D_u_arg[4], k3_D_u_arg[4], k4_D_u_arg[4];

  SinglePendulumDynamics(params, x, u, zero, zero, k1, k1_D_x, k1_D_u);
  for (int i = 0; i < 4; ++i) arg[i] = x[i] + k1[i] * h / 2.0;
  SinglePendulumDynamics(params, arg, u, zero, zero, k2, k2_D_arg, k2_D_u_arg);
  for (int i = 0; i < 4; ++i) arg[i] = x[i] + k2[i] * h / 2.0;
  SinglePendulumDynamics(params, arg, u, zero, zero, k3, k3_D_arg, k3_D_u_arg);
  for (int i = 0; i < 4; ++i) arg[i] = x[i] + k3[i] * h;
  SinglePendulumDynamics(params, arg, u, zero, zero, k4, k4_D_arg, k4_D_u_arg);
  for (int i = 0; i < 4; ++i) x_new[i] = x[i] + (h / 6.0) * (k1[i] + k2[i] * 2.0 + k3[i] * 2.0 + k4[i]);

  // chain rule through the stages (integration.hpp:36-46)
  double inner[16], k2_D_x[16], k3_D_x[16], k4_D_x[16];
  for (int i = 0; i < 16; ++i) inner[i] = ((i % 5 == 0) ? 1.0 : 0.0) + k1_D_x[i] * h / 2.0;
  Mul44(k2_D_arg, inner, k2_D_x);
  for (int i = 0; i < 16; ++i) inner[i] = ((i % 5 == 0) ? 1.0 : 0.0) + k2_D_x[i] * h / 2.0;
  Mul44(k3_D_arg, inner, k3_D_x);
  for (int i = 0; i < 16; ++i) inner[i] = ((i % 5 == 0) ? 1.0 : 0.0) + k3_D_x[i] * h;
  Mul44(k4_D_arg, inner, k4_D_x);

  double tmp[4], k2_D_u[4], k3_D_u[4], k4_D_u[4];
  Mul41(k2_D_arg, k1_D_u, tmp);
  for (int i = 0; i < 4; ++i) k2_D_u[i] = tmp[i] * (h / 2.0) + k2_D_u_arg[i];
  Mul41(k3_D_arg, k2_D_u, tmp);
  for (int i = 0; i < 4; ++i) k3_D_u[i] = tmp[i] * (h / 2.0) + k3_D_u_arg[i];
  Mul41(k4_D_arg, k3_D_u, tmp);
  for (int i = 0; i < 4; ++i) k4_D_u[i] = tmp[i] * h + k4_D_u_arg[i];

  for (int i = 0; i < 16; ++i) {
    x_new_D_x[i] = ((i % 5 == 0) ? 1.0 : 0.0) +
                   (h / 6.0) * (k1_D_x[i] + k2_D_x[i] * 2.0 + k3_D_x[i] * 2.0 + k4_D_x[i]);
  }
  for (int i = 0; i < 4; ++i) {
    x_new_D_u[i] = (h / 6.0) * (k1_D_u[i] + k2_D_u[i] * 2.0 + k3_D_u[i] * 2.0 + k4_D_u[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Controller   (optimization/optimization.cc)
// ---------------------------------------------------------------------------------------------------------------------

namespace {

enum KeyType : int { B_X = 0, THETA_1 = 1, B_X_DOT = 2, THETA_1_DOT = 3, U = 4 };  // optimization/key.hpp:8-19

// Variable layout [x(0) .. x(NS-1), u(0) .. u(W-1)]   (optimization.cc:24-37)
int MapKey(const KeyType type, const std::size_t index, const std::size_t num_states) {
  if (type == U) return static_cast<int>(num_states) * 4 + static_cast<int>(index);
  return static_cast<int>(index) * 4 + static_cast<int>(type);
}

// Scalar residual (v - desired) * weight, optionally wrapped to (-pi, pi]   (optimization.cc:165-192)
Residual MakeScalarCost(const KeyType key_type, const std::size_t index, const std::size_t num_states,
                        const double desired_value, const double weight) {
  Residual r;
  r.index = {MapKey(key_type, index, num_states)};
  r.dim = 1;
  const bool wrap = key_type == THETA_1;
  r.func = [desired_value, weight, wrap](const double* vars, double* out, double* J) {
    if (J) J[0] = weight;
    out[0] = wrap ? ModPi(vars[0] - desired_value) * weight : (vars[0] - desired_value) * weight;
  };
  return r;
}

// Shooting-segment defect and its 4 x (8+S) Jacobian   (optimization.cc:99-160)
Residual MakeDynamicalConstraint(std::vector<int> indices, const SingleCartPoleParams params,
                                 const std::size_t state_spacing, const double dt) {
  Residual r;
  r.index = std::move(indices);
  r.dim = 4;
  r.func = [params, state_spacing, dt](const double* vars, double* out, double* J) {
    const double* x_k = vars;
    const double* x_kpl = vars + 4;
    const double* u_k = vars + 8;
    const int P = 8 + static_cast<int>(state_spacing);
    const double zero[2] = {0.0, 0.0};

    double x[4] = {x_k[0], x_k[1], x_k[2], x_k[3]};
    std::vector<double> step_D_x, step_D_u;
    if (J) {
      step_D_x.resize(16 * state_spacing);
      step_D_u.resize(4 * state_spacing);
      for (std::size_t i = 0; i < state_spacing; ++i) {
        double x_next[4];
        RungeKutta4(params, x, u_k[i], dt, x_next, &step_D_x[16 * i], &step_D_u[4 * i]);
        for (int j = 0; j < 4; ++j) x[j] = x_next[j];
      }
    } else {
      for (std::size_t i = 0; i < state_spacing; ++i) {
        double x_next[4];
        RungeKutta4NoJacobians(params, x, u_k[i], dt, zero, zero, x_next);
        for (int j = 0; j < 4; ++j) x[j] = x_next[j];
      }
    }
    x[1] = ModPi(x[1]);

    if (J) {
      for (int i = 0; i < 4 * P; ++i) J[i] = 0.0;
      double chain[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
      for (int i = static_cast<int>(state_spacing) - 1; i >= 0; --i) {
        double col[4];
        Mul41(chain, &step_D_u[4 * i], col);
        for (int rr = 0; rr < 4; ++rr) J[rr * P + 8 + i] = col[rr];
        double next[16];
        Mul44(chain, &step_D_x[16 * i], next);
        for (int j = 0; j < 16; ++j) chain[j] = next[j];
      }
      for (int rr = 0; rr < 4; ++rr) {
        for (int cc = 0; cc < 4; ++cc) {
          J[rr * P + cc] = chain[rr * 4 + cc];
          J[rr * P + 4 + cc] = (rr == cc) ? -1.0 : 0.0;
        }
      }
    }
    for (int j = 0; j < 4; ++j) out[j] = x[j] - x_kpl[j];
    out[1] = ModPi(out[1]);
  };
  return r;
}

}  // namespace

Optimization::Optimization(const OptimizationParams& params) : params_(params) {
  // optimization.cc:13-22
  if (!(params.control_dt > 0)) throw std::invalid_argument("control_dt must be > 0");
  if (!(params.window_length >= 1)) throw std::invalid_argument("window_length must be >= 1");
  if (params.state_spacing == 0 || params.window_length % params.state_spacing != 0) {
    throw std::invalid_argument("state_spacing must divide into window_length cleanly");
  }
  if (!(params.max_iterations >= 1)) throw std::invalid_argument("max_iterations must be >= 1");
  if (!(params.u_cost_weight >= 0.0)) throw std::invalid_argument("u_cost_weight must be >= 0");
  if (!(params.u_derivative_cost_weight >= 0.0)) throw std::invalid_argument("u_derivative_cost_weight must be >= 0");
}

OptimizationOutputs Optimization::Step(const SingleCartPoleState& current_state,
                                       const SingleCartPoleParams& dynamics_params, const double b_x_set_point) {
  BuildProblem(current_state, dynamics_params, b_x_set_point);

  Vec guess(problem_.dimension, 0.0);
  const auto num_states = params_.NumStates();
  const double x_cur[4] = {current_state.b_x, current_state.th_1, current_state.b_x_dot, current_state.th_1_dot};
  if (!previous_solution_.empty()) {
    // warm start: previous solution with the controls shifted forward by one (optimization.cc:50-57)
    guess = previous_solution_;
    for (int i = 0; i < 4; ++i) guess[MapKey(static_cast<KeyType>(i), 0, num_states)] = x_cur[i];
    for (std::size_t k = 0; k + 1 < params_.window_length; ++k) {
      guess[MapKey(U, k, num_states)] = guess[MapKey(U, k + 1, num_states)];
    }
  } else {
    // cold start: sinusoidal control guess (optimization.cc:59-67)
    for (int i = 0; i < 4; ++i) guess[MapKey(static_cast<KeyType>(i), 0, num_states)] = x_cur[i];
    for (std::size_t k = 0; k < params_.window_length; ++k) {
      guess[MapKey(U, k, num_states)] =
          params_.u_guess_sinusoid_amplitude *
          std::sin(static_cast<double>(k) / static_cast<double>(params_.window_length) * 2 * M_PI);
    }
  }
  FillInitialGuess(guess, dynamics_params, num_states);

  ConstrainedNonlinearLeastSquares::Params p{};
  p.max_iterations = static_cast<int>(params_.max_iterations);
  p.max_line_search_iterations = 5;
  p.absolute_first_derivative_tol = params_.absolute_first_derivative_tol;
  p.equality_penalty_initial = params_.equality_penalty_initial;
  p.relative_exit_tol = params_.relative_exit_tol;  // the later of the two assignments wins (optimization.cc:75,79)

  NLSSolverOutputs outputs = solver_->Solve(p, guess);

  std::vector<double> previous_solution = previous_solution_;
  previous_solution_ = solver_->variables();

  const double* u_out = previous_solution_.data() + (previous_solution_.size() - params_.window_length);
  std::vector<SingleCartPoleState> predicted = ComputePredictedStates(u_out, dynamics_params, current_state);

  OptimizationOutputs result;
  result.initial_state = current_state;
  result.previous_solution = std::move(previous_solution);
  result.final_errors = solver_->EvaluateNonlinearErrors(solver_->variables());
  result.final_penalty = outputs.iterations.empty() ? p.equality_penalty_initial : outputs.iterations.back().penalty;
  result.solver_outputs = std::move(outputs);
  result.u.assign(u_out, u_out + params_.window_length);
  result.predicted_states = std::move(predicted);
  return result;
}

void Optimization::BuildProblem(const SingleCartPoleState& current_state,
                                const SingleCartPoleParams& dynamics_params, const double b_x_set_point) {
  // optimization.cc:194-331
  const std::size_t num_states = params_.NumStates();
  const std::size_t final_state_index = num_states - 1;

  problem_.clear();
  problem_.dimension = static_cast<int>(num_states) * 4 + static_cast<int>(params_.window_length);

  for (std::size_t s = 0; s + 1 < num_states; ++s) {
    std::vector<int> indices;
    indices.reserve(20);
    for (const std::size_t state : {s, s + 1}) {
      for (int i = 0; i < 4; ++i) indices.push_back(MapKey(static_cast<KeyType>(i), state, num_states));
    }
    for (std::size_t k = 0; k < params_.state_spacing; ++k) {
      indices.push_back(MapKey(U, params_.state_spacing * s + k, num_states));
    }
    problem_.equality_constraints.push_back(
        MakeDynamicalConstraint(std::move(indices), dynamics_params, params_.state_spacing, params_.control_dt));
  }

  const double x_current[4] = {current_state.b_x, current_state.th_1, current_state.b_x_dot,
                               current_state.th_1_dot};
  for (int i = 0; i < 4; ++i) {
    problem_.equality_constraints.push_back(
        MakeScalarCost(static_cast<KeyType>(i), 0, num_states, x_current[i], 1.0));
  }

  // terminal quantities: cost when the weight is >= 0, otherwise a hard equality with weight 1
  const double terminal_weight[4] = {params_.b_x_final_cost_weight, params_.th_final_cost_weight,
                                     params_.b_x_dot_final_cost_weight, params_.th_dot_final_cost_weight};
  const double terminal_target[4] = {b_x_set_point, M_PI / 2, 0.0, 0.0};
  for (int i = 0; i < 4; ++i) {
    if (terminal_weight[i] >= 0.0) {
      problem_.costs.push_back(MakeScalarCost(static_cast<KeyType>(i), final_state_index, num_states,
                                              terminal_target[i], terminal_weight[i]));
    } else {
      problem_.equality_constraints.push_back(
          MakeScalarCost(static_cast<KeyType>(i), final_state_index, num_states, terminal_target[i], 1.0));
    }
  }

  if (params_.u_derivative_cost_weight > 0.0) {
    for (std::size_t k = 0; k + 1 < params_.window_length; ++k) {
      const double weight = params_.u_derivative_cost_weight;
      Residual r;
      r.index = {MapKey(U, k, num_states), MapKey(U, k + 1, num_states)};
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
    const double u_prev = !previous_solution_.empty() ? previous_solution_[MapKey(U, 0, num_states)] : 0.0;
    problem_.costs.push_back(MakeScalarCost(U, 0, num_states, u_prev, params_.u_derivative_cost_weight));
  }

  if (params_.u_cost_weight > 0.0) {
    for (std::size_t k = 0; k < params_.window_length; ++k) {
      problem_.costs.push_back(MakeScalarCost(U, k, num_states, 0.0, params_.u_cost_weight));
    }
  }

  if (solver_) return;

  const std::size_t window_length = params_.window_length;
  const auto retraction = [num_states, window_length](Vec& x, const Vec& dx, const double alpha) {
    for (std::size_t i = 0; i < x.size(); ++i) x[i] += dx[i] * alpha;
    for (std::size_t s = 0; s < num_states; ++s) {
      const int angle_index = MapKey(THETA_1, s, num_states);
      x[angle_index] = ModPi(x[angle_index]);
      const int pos_index = MapKey(B_X, s, num_states);
      x[pos_index] = std::clamp(x[pos_index], -5.0, 5.0);
    }
    for (std::size_t k = 0; k < window_length; ++k) {
      const int index = MapKey(U, k, num_states);
      x[index] = std::clamp(x[index], -300.0, 300.0);
    }
  };
  solver_ = std::make_unique<ConstrainedNonlinearLeastSquares>(&problem_, retraction);
}

void Optimization::FillInitialGuess(Vec& guess, const SingleCartPoleParams& dynamics_params,
                                    const std::size_t num_states) const {
  // optimization.cc:333-351
  const double zero[2] = {0.0, 0.0};
  double x[4];
  for (int i = 0; i < 4; ++i) x[i] = guess[MapKey(static_cast<KeyType>(i), 0, num_states)];
  for (std::size_t s = 1; s < num_states; ++s) {
    for (std::size_t k = 0; k < params_.state_spacing; ++k) {
      double x_next[4];
      RungeKutta4NoJacobians(dynamics_params, x, guess[MapKey(U, (s - 1) * params_.state_spacing + k, num_states)],
                             params_.control_dt, zero, zero, x_next);
      for (int i = 0; i < 4; ++i) x[i] = x_next[i];
      x[1] = ModPi(x[1]);
    }
    for (int i = 0; i < 4; ++i) guess[MapKey(static_cast<KeyType>(i), s, num_states)] = x[i];
  }
}

std::vector<SingleCartPoleState> Optimization::ComputePredictedStates(const double* u_out,
                                                                      const SingleCartPoleParams& dynamics_params,
                                                                      const SingleCartPoleState& x_current) const {
  // optimization.cc:353-371
  const double zero[2] = {0.0, 0.0};
  std::vector<SingleCartPoleState> predicted_states;
  predicted_states.reserve(params_.window_length);
  double x[4] = {x_current.b_x, x_current.th_1, x_current.b_x_dot, x_current.th_1_dot};
  for (std::size_t k = 0; k < params_.window_length; ++k) {
    double x_next[4];
    RungeKutta4NoJacobians(dynamics_params, x, u_out[k], params_.control_dt, zero, zero, x_next);
    for (int i = 0; i < 4; ++i) x[i] = x_next[i];
    x[1] = ModPi(x[1]);
    predicted_states.push_back({x[0], x[1], x[2], x[3]});
  }
  return predicted_states;
}

// ---------------------------------------------------------------------------------------------------------------------
// Simulator   (optimization/simulator.cc:11-36)
// ---------------------------------------------------------------------------------------------------------------------

void Simulator::Step(const SingleCartPoleParams& params, double dt, const double u, const double f_base[2],
                     const double f_mass[2]) {
  if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0");
  if (!std::isfinite(u)) throw std::invalid_argument("u must be finite");
  constexpr double internal_dt = 0.001;
  while (dt > 0.0) {
    SubStep(params, std::min(dt, internal_dt), u, f_base, f_mass);
    dt -= internal_dt;
  }
}

void Simulator::SubStep(const SingleCartPoleParams& params, const double dt, const double u, const double f_base[2],
                        const double f_mass[2]) {
  double x_next[4];
  RungeKutta4NoJacobians(params, state_, u, dt, f_base, f_mass, x_next);
  for (int i = 0; i < 4; ++i) state_[i] = x_next[i];
  state_[1] = ModPi(state_[1]);
}

}  // namespace oracle
