/* ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points of the CPU restatement (liboracle.so), bound with ctypes
 * from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
 * library (libcpmpc.so, include/cpmpc.h) never links or loads this.
 *
 * All matrices are row-major double arrays.  `opt[14]` is optimization/optimization.hpp:12-48 flattened in
 * declaration order: control_dt, window_length, state_spacing, max_iterations, relative_exit_tol,
 * absolute_first_derivative_tol, equality_penalty_initial, u_guess_sinusoid_amplitude, u_cost_weight,
 * u_derivative_cost_weight, b_x_final_cost_weight, th_final_cost_weight, b_x_dot_final_cost_weight,
 * th_dot_final_cost_weight.  `dyn[9]` is optimization/structs.hpp:8-26 in declaration order.
 */
#ifndef ORACLE_CAPI_H_
#define ORACLE_CAPI_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_STATS_STRIDE 8      /* term, n_iter, n_ls, f, equality, penalty, last_lambda, n_failed_ls */
#define ORACLE_ITERLOG_STRIDE 27   /* state, lambda, f0, eq0, d_f, d_eq, penalty, step_result, n_steps, 6x(alpha,f,eq) */
#define ORACLE_WEIGHTS_STRIDE 6    /* u, u_derivative, b_x_final, th_final, b_x_dot_final, th_dot_final */

/* single_pendulum_dynamics.hpp:13-186.  J_x (16) / J_u (4) may be NULL. */
void oracle_dynamics(const double* dyn, const double* x, double u, const double* f_base, const double* f_mass,
                     double* x_dot, double* J_x, double* J_u);
/* integration.hpp:65-73 */
double oracle_mod_pi(double angle);
/* integration.hpp:53-62 */
void oracle_rk4(const double* dyn, const double* x, double u, double h, const double* f_base,
                const double* f_mass, double* x_new);
/* integration.hpp:13-49 */
void oracle_rk4_jac(const double* dyn, const double* x, double u, double h, double* x_new, double* D_x, double* D_u);
/* simulator.cc:11-36; state (4) updated in place */
void oracle_sim_step(const double* dyn, double* state, double dt, double u, const double* f_base,
                     const double* f_mass);

/* qp.cc:679-729.  G: N*N (lower triangle read), A: K*N.  Returns 0 success, 1 not PD, 2 rank deficient. */
int oracle_qp_nullspace(int N, int K, const double* G, const double* c, const double* A, const double* b,
                        double* x_out);
/* nonlinear.cc:531-538; returns 1 and writes *out when a value exists */
int oracle_quadratic_approx_minimum(double phi_0, double phi_prime_0, double alpha_0, double phi_alpha_0,
                                    double* out);
/* nonlinear.cc:563-578 */
void oracle_cubic_approx_coeffs(double phi_0, double phi_prime_0, double alpha_0, double phi_alpha_0,
                                double alpha_1, double phi_alpha_1, double* ab_out);
/* nonlinear.cc:594-605 */
int oracle_cubic_approx_minimum(double phi_prime_0, const double* ab, double* out);

/* Stateful single controller == pendulum::Optimization (optimization.hpp:73-108). Returns NULL on invalid params. */
void* oracle_opt_create(const double* opt);
void oracle_opt_destroy(void* handle);
void oracle_opt_reset(void* handle);
void oracle_opt_set_previous_solution(void* handle, const double* sol, int64_t n);
/* returns number of doubles written (0 when there is no previous solution) */
int64_t oracle_opt_get_previous_solution(void* handle, double* out, int64_t capacity);
/* One Step().  u_out[W], pred_out[W*4] (may be NULL), stats[ORACLE_STATS_STRIDE],
 * iter_log[max_iterations*ORACLE_ITERLOG_STRIDE] (may be NULL).  Returns 0, or -1 when the solver threw. */
int oracle_opt_step(void* handle, const double* state, const double* dyn, double set_point, double* u_out,
                    double* pred_out, double* stats, double* iter_log);

/* B independent Step() calls (one fresh controller per instance, one instance per OpenMP thread).
 * weights: NULL or B*ORACLE_WEIGHTS_STRIDE per-instance cost-weight overrides.
 * prev: NULL (all cold) or B*N warm-start solutions, used where has_prev[i] != 0.
 * Outputs may be NULL individually.  n_threads <= 0 means all cores.  Returns threads used. */
int oracle_batch_step(const double* opt, int64_t B, const double* states, const double* dyn,
                      const double* set_points, const double* weights, const double* prev,
                      const uint8_t* has_prev, double* u_out, double* pred_out, double* sol_out, double* stats,
                      double* iter_log, int n_threads);

/* Closed loop: n_steps x { Step(sim state); Simulator::Step(control_dt, u[0]) } per instance, warm-started
 * (optimization_test.cc:38-61).  states are updated in place; controls_out: B*n_steps (may be NULL);
 * counts_out[4] += {solves, sqp_iterations, line_search_steps, n_bad_termination(QP_INDEFINITE|MAX_LAMBDA)}.
 * term_hist_out: NULL or 7 counters by termination state. */
int oracle_closed_loop(const double* opt, int64_t B, int64_t n_steps, double* states, const double* dyn,
                       const double* set_points, double* controls_out, double* last_pred_terminal_out,
                       int64_t* counts_out, int64_t* term_hist_out, int n_threads);

/* ---- double pendulum on a cart: nx = 6 generalisation (oracle/double_cartpole.hpp; the reference has only the
 * physics spec symbolic/dynamics_double.py).  opt[16] = the first ten entries as above, then the six terminal
 * weights in state order [b_x, th_1, th_2, b_x', th_1', th_2'] (negative = hard equality).  dyn[6] =
 * {m_b, m_1, m_2, l_1, l_2, g}.  x[6] = [b_x, th_1, th_2, b_x', th_1', th_2']. */
#define ORACLE_DP_WEIGHTS_STRIDE 8 /* u, u_derivative, then the six terminal weights */
void oracle_dp_dynamics(const double* dyn, const double* x, double u, double* x_dot, double* J_x, double* J_u);
void oracle_dp_rk4(const double* dyn, const double* x, double u, double h, double* x_new);
void oracle_dp_rk4_jac(const double* dyn, const double* x, double u, double h, double* x_new, double* D_x,
                       double* D_u);
/* Same contract as oracle_batch_step with states B*6, dyn B*6, pred_out B*W*6, N = 6*NS + W. */
int oracle_dp_batch_step(const double* opt, int64_t B, const double* states, const double* dyn,
                         const double* set_points, const double* weights, const double* prev,
                         const uint8_t* has_prev, double* u_out, double* pred_out, double* sol_out, double* stats,
                         double* iter_log, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* ORACLE_CAPI_H_ */
