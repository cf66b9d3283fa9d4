// Parameter / state structs of the reference API, kept field for field so that callers of
// optimization/structs.hpp:8-70 compile against this header unchanged (minus the Eigen conversion helpers, which the
// GPU path does not need: states travel as four doubles).
#pragma once
#include <array>

namespace pendulum {

// optimization/structs.hpp:8-41
struct SingleCartPoleParams {
  double m_b;     // Mass of the base (kg).
  double m_1;     // Mass of the object on the pole (kg).
  double l_1;     // Length of the pole (meters).
  double g;       // Gravity (m/s^2).
  double mu_b;    // Friction at the base.
  double v_mu_b;  // Cutoff velocity of the smoothed Coulomb model (m/s).
  double c_d_1;   // Drag coefficient on the pole mass.
  double x_s;     // Position of the bumper springs (meters).
  double k_s;     // Spring constant of the bumper springs (N/m).

  SingleCartPoleParams() noexcept = default;
  constexpr SingleCartPoleParams(double m_b, double m_1, double l_1, double g, double mu_b, double v_mu_b,
                                 double c_d_1, double x_s, double k_s) noexcept
      : m_b(m_b), m_1(m_1), l_1(l_1), g(g), mu_b(mu_b), v_mu_b(v_mu_b), c_d_1(c_d_1), x_s(x_s), k_s(k_s) {}
};
static_assert(sizeof(SingleCartPoleParams) == 9 * sizeof(double), "layout must match CPMPC_DYN_STRIDE");

// optimization/structs.hpp:44-64
struct SingleCartPoleState {
  double b_x;
  double th_1;
  double b_x_dot;
  double th_1_dot;

  SingleCartPoleState() noexcept = default;
  constexpr SingleCartPoleState(double b_x, double th_1, double b_x_dot, double th_1_dot) noexcept
      : b_x(b_x), th_1(th_1), b_x_dot(b_x_dot), th_1_dot(th_1_dot) {}

  // optimization/structs.hpp:60-63 returns an Eigen::Vector4d; the same four numbers in the same order, without Eigen
  constexpr std::array<double, 4> ToVector() const noexcept { return {b_x, th_1, b_x_dot, th_1_dot}; }
};
static_assert(sizeof(SingleCartPoleState) == 4 * sizeof(double), "layout must match CPMPC_STATE_STRIDE");

// optimization/structs.hpp:67-70
struct Vector2 {
  double x;
  double y;
};

}  // namespace pendulum
