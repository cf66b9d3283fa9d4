// Host-side check of cart_pole_mpc_b200/csrc/dynamics.cuh (the device dynamics compiled for the host: same source, the
// reciprocal / rsqrt seeds replaced by float-precision ones) against the oracle's restatement of the generated dynamics
// and of integration.hpp (oracle/cartpole.cc, bit-equal to the compiled reference: tests/test_reference_pinning.py).
//
// The device code evaluates a re-associated form of the equations of motion (parameter-only products folded, the
// factor m_1 cancelled), so agreement is to rounding, not bit for bit: values 1e-13, Jacobian entries 1e-12 relative to
// the largest entry of their row, RK4 steps and their Jacobians 1e-13.  Covers all model terms: friction on both sides
// of its tanh, drag at rest (exactly zero velocity: the selects), both bumpers engaged, external forces.
// Test infrastructure only.  Exit code 1 if a bound is exceeded.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>

#include "../../cart_pole_mpc_b200/csrc/dynamics.cuh"
#include "../../oracle/cartpole.hpp"

namespace {
double rel(double got, double want, double scale) { return std::fabs(got - want) / std::fmax(scale, 1e-3); }
}  // namespace

int main() {
  std::mt19937_64 rng(2024);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  double e_val = 0, e_jac = 0, e_rk = 0, e_rkj = 0, e_force = 0, e_jac_stiff = 0, e_rkj_stiff = 0;
  for (int it = 0; it < 200000; ++it) {
    oracle::SingleCartPoleParams p;
    p.m_b = 0.1 + 1.9 * U(rng);
    p.m_1 = 0.1 + 0.9 * U(rng);
    p.l_1 = 0.05 + 0.45 * U(rng);
    p.g = 9.81;
    p.mu_b = (it % 7 == 0) ? 0.0 : 0.5 * U(rng);
    const bool stiff = it % 11 == 0;  // cut-off velocity at its floor: d tanh / dv = 1e6 (1 - tanh^2) amplifies rounding
    p.v_mu_b = stiff ? 1e-9 : 0.01 + 0.3 * U(rng);
    p.c_d_1 = (it % 5 == 0) ? 0.0 : 0.15 * U(rng);
    p.x_s = 0.8;
    p.k_s = 100.0;
    double x[4] = {-1.2 + 2.4 * U(rng), -7.0 + 14.0 * U(rng), -3.0 + 6.0 * U(rng), -12.0 + 24.0 * U(rng)};
    if (it % 13 == 0) x[2] = x[3] = 0.0;            // at rest: drag selects
    if (it % 17 == 0) x[2] *= 1e-3;                 // inside the friction tanh
    if (it % 19 == 0) x[0] = (it & 1) ? 0.8 : -0.8; // exactly at a bumper
    const double u = -150.0 + 300.0 * U(rng);
    const double pv[9] = {p.m_b, p.m_1, p.l_1, p.g, p.mu_b, p.v_mu_b, p.c_d_1, p.x_s, p.k_s};
    const cpmpc::DynConst P = cpmpc::make_dyn_const(pv);
    const double zero[2] = {0.0, 0.0};

    double xd[4], Jx[16], Ju[4];
    oracle::SinglePendulumDynamics(p, x, u, zero, zero, xd, Jx, Ju);
    double ax, at, fx[4], fth[4], gx, gth;
    cpmpc::dyn_eval<false>(P, x[0], x[1], x[2], x[3], u, 0.0, 0.0, 0.0, ax, at);
    const double vs = std::fmax(std::fabs(xd[2]), std::fabs(xd[3]));
    e_val = std::fmax(e_val, std::fmax(rel(ax, xd[2], vs), rel(at, xd[3], vs)));
    double ax2, at2;
    cpmpc::dyn_eval_jac(P, x[0], x[1], x[2], x[3], u, ax2, at2, fx, fth, gx, gth);
    e_val = std::fmax(e_val, std::fmax(rel(ax2, xd[2], vs), rel(at2, xd[3], vs)));
    double rs2 = std::fabs(Ju[2]), rs3 = std::fabs(Ju[3]);
    for (int j = 0; j < 4; ++j) {
      rs2 = std::fmax(rs2, std::fabs(Jx[8 + j]));
      rs3 = std::fmax(rs3, std::fabs(Jx[12 + j]));
    }
    double& ej = stiff ? e_jac_stiff : e_jac;
    for (int j = 0; j < 4; ++j) {
      ej = std::fmax(ej, rel(fx[j], Jx[8 + j], rs2));
      ej = std::fmax(ej, rel(fth[j], Jx[12 + j], rs3));
    }
    ej = std::fmax(ej, std::fmax(rel(gx, Ju[2], rs2), rel(gth, Ju[3], rs3)));

    // simulator variant with external forces
    const double fb[2] = {-5.0 + 10.0 * U(rng), 0.0}, fmass[2] = {-2.0 + 4.0 * U(rng), -2.0 + 4.0 * U(rng)};
    oracle::SinglePendulumDynamics(p, x, u, fb, fmass, xd, nullptr, nullptr);
    cpmpc::dyn_eval<true>(P, x[0], x[1], x[2], x[3], u, fb[0], fmass[0], fmass[1], ax, at);
    const double vs2 = std::fmax(std::fabs(xd[2]), std::fabs(xd[3]));
    e_force = std::fmax(e_force, std::fmax(rel(ax, xd[2], vs2), rel(at, xd[3], vs2)));

    // one RK4 step with and without Jacobians
    double xn[4], A[16], B[4];
    oracle::RungeKutta4(p, x, u, 0.01, xn, A, B);
    double y[4] = {x[0], x[1], x[2], x[3]};
    cpmpc::rk4_step<false>(P, y, u, 0.01);
    double z[4] = {x[0], x[1], x[2], x[3]}, Ad[16], Bd[4];
    cpmpc::rk4_step_jac(P, z, u, 0.01, Ad, Bd);
    for (int i = 0; i < 4; ++i) {
      const double sc = std::fmax(1.0, std::fabs(xn[i]));
      e_rk = std::fmax(e_rk, std::fmax(rel(y[i], xn[i], sc), rel(z[i], xn[i], sc)));
      double rs = std::fabs(B[i]);
      for (int j = 0; j < 4; ++j) rs = std::fmax(rs, std::fabs(A[4 * i + j]));
      double& er = stiff ? e_rkj_stiff : e_rkj;
      for (int j = 0; j < 4; ++j) er = std::fmax(er, rel(Ad[4 * i + j], A[4 * i + j], rs));
      er = std::fmax(er, rel(Bd[i], B[i], rs));
    }
  }
  std::printf("accelerations %.3e\nacceleration Jacobians %.3e (friction cut-off at its floor: %.3e)\n"
              "with external forces %.3e\nRK4 step %.3e\nRK4 step Jacobians %.3e (%.3e)\n",
              e_val, e_jac, e_jac_stiff, e_force, e_rk, e_rkj, e_rkj_stiff);
  const bool ok = e_val < 1e-13 && e_jac < 1e-12 && e_jac_stiff < 1e-10 && e_force < 1e-13 && e_rk < 1e-13 &&
                  e_rkj < 2e-11 && e_rkj_stiff < 1e-10;
  return ok ? 0 : 1;
}
