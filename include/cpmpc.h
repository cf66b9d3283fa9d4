/* cpmpc.h — C ABI of the B200-native batched cart-pole MPC solver (libcpmpc.so).
 *
 * This is the drop-in boundary for ONE hot path of gareth-cross/cart-pole-mpc: pendulum::Optimization::Step
 * (optimization/optimization.cc:39-97) run over a batch of independent instances on one GPU.  The reference has no
 * C ABI of its own; its boundary is the C++ class pendulum::Optimization (optimization/optimization.hpp:73-108) that
 * the nanobind wrapper (wrapper/wrapper.cc:87-90) and optimization_test.cc call.  Each entry point below names the
 * reference interface it replaces.  The C++ facade that keeps the reference's struct/class API on top of this ABI is
 * cart_pole_mpc_b200/host/optimization.hpp; the Python mirror of `pypendulum` is cart_pole_mpc_b200/pypendulum.py.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the ABI; every function returns a cpmpc_status
 *     (0 = success) and cpmpc_last_error() returns a thread-local message for the last failure.
 *   - all arrays are dense, row-major, FP64.  "B" is the batch size fixed at cpmpc_create().
 *   - W = window_length, S = state_spacing, NS = W/S + 1, N = 4*NS + W (variable layout
 *     [x(0) .. x(NS-1) | u(0) .. u(W-1)], optimization.cc:24-37).
 *   - a handle is bound to one device; calls on one handle must be serialised by the caller (the reference object is
 *     not thread safe either).  Use one handle per GPU for multi-GPU.
 *   - there is NO CPU fallback: every call fails with CPMPC_ERR_CUDA when no usable device exists.
 */
#ifndef CPMPC_H_
#define CPMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPMPC_VERSION 1

typedef enum cpmpc_status {
  CPMPC_OK = 0,
  CPMPC_ERR_INVALID_ARGUMENT = 1, /* F_ASSERT-class precondition failures (optimization.cc:13-22) */
  CPMPC_ERR_CUDA = 2,             /* CUDA runtime / launch failure, or no device */
  CPMPC_ERR_UNSUPPORTED = 3,      /* problem shape outside what the kernels were built for */
} cpmpc_status;

/* pendulum::OptimizationParams (optimization/optimization.hpp:12-48), field for field, plus the handful of
 * mini_opt solver parameters that Optimization::Step hard-codes (optimization.cc:73-79) so they stay visible. */
typedef struct cpmpc_params {
  double control_dt;                    /* 0.01 */
  int64_t window_length;                /* 40 */
  int64_t state_spacing;                /* 10 */
  int64_t max_iterations;               /* 8 */
  double relative_exit_tol;             /* 1e-5 */
  double absolute_first_derivative_tol; /* 1e-6 */
  double equality_penalty_initial;      /* 1.0 */
  double u_guess_sinusoid_amplitude;    /* 10.0 */
  double u_cost_weight;                 /* 0.1 */
  double u_derivative_cost_weight;      /* 0.1 */
  double b_x_final_cost_weight;         /* 150 */
  double th_final_cost_weight;          /* -1 (negative => equality constraint) */
  double b_x_dot_final_cost_weight;     /* -1 */
  double th_dot_final_cost_weight;      /* -1 */
} cpmpc_params;

/* Per-instance solver statistics: the fields of mini_opt::NLSSolverOutputs / NLSIteration that callers read
 * (dependencies/mini_opt/include/mini_opt/structs.hpp:277-347).  Enum values follow the reference declaration order. */
typedef struct cpmpc_stats {
  int32_t termination_state;      /* mini_opt::NLSTerminationState: 0 MAX_ITERATIONS, 1 SATISFIED_ABSOLUTE_TOL,
                                     2 SATISFIED_RELATIVE_TOL, 3 SATISFIED_FIRST_ORDER_TOL, 4 MAX_LAMBDA,
                                     5 QP_INDEFINITE; -1 = instance rejected (non-finite cost at the
                                     linearisation point: the reference throws, nonlinear.cc:356-357) */
  int32_t n_iterations;           /* solver_outputs.iterations.size() */
  int32_t n_line_search_steps;    /* NLSSolverOutputs::NumLineSearchSteps() */
  int32_t n_failed_line_searches; /* NLSSolverOutputs::NumFailedLineSearches() */
  int32_t last_step_result;       /* mini_opt::StepSizeSelectionResult of the last iteration */
  int32_t optimizer_state;        /* mini_opt::OptimizerState after the last iteration */
  double f;                       /* Errors::f at the final variables (0.5 * sum of squared costs) */
  double equality;                /* Errors::equality at the final variables (L1 norm) */
  double penalty;                 /* final equality penalty mu */
  double lambda;                  /* final Levenberg-Marquardt lambda */
} cpmpc_stats;

/* Layout of one record of the optional per-iteration log (doubles): optimizer_state, lambda (at entry),
 * f(0), c(0), d_f, d_equality, penalty, step_result, n_steps, then 6 x (alpha, f, equality) — the fields of
 * mini_opt::NLSIteration (structs.hpp:294-322).  The log holds max_iterations records per instance. */
#define CPMPC_ITERLOG_STRIDE 27
/* Per-instance cost-weight overrides: u, u_derivative, b_x_final, th_final, b_x_dot_final, th_dot_final. */
#define CPMPC_WEIGHTS_STRIDE 6
#define CPMPC_DYN_STRIDE 9   /* pendulum::SingleCartPoleParams, optimization/structs.hpp:8-26 */
#define CPMPC_STATE_STRIDE 4 /* pendulum::SingleCartPoleState,  optimization/structs.hpp:44-49 */

typedef struct cpmpc_solver* cpmpc_handle;

/* Thread-local description of the last error returned on this thread ("" when none). */
const char* cpmpc_last_error(void);
int cpmpc_version(void);

/* OptimizationParams{} defaults (optimization.hpp:12-48). */
void cpmpc_default_params(cpmpc_params* params);

/* Replaces `Optimization::Optimization(const OptimizationParams&)` (optimization.cc:13-22) for `batch` independent
 * controllers on CUDA device `device`.  Validates the same preconditions. */
int cpmpc_create(const cpmpc_params* params, int64_t batch, int device, cpmpc_handle* out);
/* Same with option flags.  The default mapping (flags == 0) gives results that do not depend on how instances are
 * grouped into batches: instance i alone and instance i inside any batch produce the same bits.
 *   CPMPC_FLAG_LOW_LATENCY  two lanes per shooting segment: the linearisation of a segment runs as two half-length
 *                           chunks side by side.  Shortens one solve's dependent chain (0.73 -> 0.60 ms for a batch
 *                           of one at the defaults) at the price of throughput on large batches, and rounds the
 *                           segment Jacobian products differently (parity with the reference is unchanged, bitwise
 *                           equality with the default mapping is not kept). */
#define CPMPC_FLAG_LOW_LATENCY 1u
/*   CPMPC_FLAG_MIN_LATENCY  one full warp per instance: every shooting segment is linearised as up to 32 / segments
 *                           chunks side by side and the QP loops run 32 lanes wide.  The mapping for single solves
 *                           and batches below a few hundred instances (batch of one at the defaults: 0.43 ms against
 *                           0.50 ms with LOW_LATENCY and 0.61 ms with the default mapping); large batches lose most
 *                           of their throughput.  Same rounding caveat as LOW_LATENCY.  Takes precedence over it. */
#define CPMPC_FLAG_MIN_LATENCY 2u
/*   CPMPC_FLAG_SPLIT_LAUNCH / CPMPC_FLAG_FUSED_LAUNCH  how a Step is launched; the results are bit-identical either
 *                           way.  Fused: one kernel per call (per closed loop), the two serial rollouts of a Step -
 *                           FillInitialGuess and ComputePredictedStates, optimization.cc:333-371 - on one lane of each
 *                           instance's lane group.  Split: those rollouts (and the closed loop's plant step) run in
 *                           their own kernels with one thread per instance before and after the solve kernel, three
 *                           stream-ordered launches per control tick: 9 % more solves/s at 65536 instances, 12 % at
 *                           W = 160, nothing to gain while the batch is a single wave.  Without either flag the
 *                           library picks by batch size (split beyond one wave of the persistent grid, from three
 *                           in the closed loop).  Ignored (fused) under LOW_LATENCY / MIN_LATENCY. */
#define CPMPC_FLAG_SPLIT_LAUNCH 4u
#define CPMPC_FLAG_FUSED_LAUNCH 8u
int cpmpc_create_ex(const cpmpc_params* params, int64_t batch, int device, uint32_t flags, cpmpc_handle* out);
int cpmpc_destroy(cpmpc_handle h);

/* Problem dimensions derived from the params of a handle. */
int64_t cpmpc_batch(cpmpc_handle h);
int64_t cpmpc_solution_dim(cpmpc_handle h); /* N */

/* Replaces `Optimization::Step` (optimization.cc:39-97) for all B instances, HOST buffers, synchronous.
 *   states      [B*4]   current states          dyn_params [B*9]   dynamics parameters
 *   set_points  [B]     b_x set points          weights    [B*6] or NULL (use params)
 *   u_out       [B*W]   OptimizationOutputs::u
 *   pred_out    [B*W*4] OptimizationOutputs::predicted_states, or NULL to skip computing them
 *   stats_out   [B]     or NULL
 * Warm-start state (Optimization::previous_solution_) is kept on the device inside the handle. */
int cpmpc_step(cpmpc_handle h, const double* states, const double* dyn_params, const double* set_points,
               const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out);

/* Pipelined variant of cpmpc_step for callers that issue step after step (a controller farm, the benchmark): the call
 * returns as soon as the work is enqueued, and up to TWO steps may be in flight, so that the device-to-host copy of
 * step k overlaps the host-to-device copy and the kernel of step k + 1 (the kernels themselves run one after the
 * other: step k + 1 warm-starts from step k, exactly as consecutive Optimization::Step calls do).
 *   - buffers are as for cpmpc_step and should be PINNED host memory (pageable memory makes the copies synchronous);
 *     inputs and outputs of a step must stay untouched until the cpmpc_wait that retires it;
 *   - cpmpc_wait blocks until the OLDEST step in flight has finished and its outputs are complete (no-op when
 *     nothing is in flight); a third cpmpc_step_async first waits for the oldest one itself;
 *   - flags: CPMPC_STEP_RESET = Optimization::Reset() of every instance, ordered before this step's solve;
 *   - every synchronous entry point of the handle (cpmpc_step, cpmpc_reset, cpmpc_get_previous_solution, ...)
 *     retires all steps in flight first. */
#define CPMPC_STEP_RESET 1u
int cpmpc_step_async(cpmpc_handle h, const double* states, const double* dyn_params, const double* set_points,
                     const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out, uint32_t flags);
int cpmpc_wait(cpmpc_handle h);

/* Same, with DEVICE pointers (on the handle's device), asynchronous on `stream` (a cudaStream_t, may be NULL for
 * the default stream).  The caller synchronises. */
int cpmpc_step_device(cpmpc_handle h, const double* states, const double* dyn_params, const double* set_points,
                      const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out,
                      void* stream);

/* Replaces `Optimization::Reset()` (optimization.hpp:83).  mask: HOST [B] bytes, non-zero = reset; NULL = all. */
int cpmpc_reset(cpmpc_handle h, const uint8_t* mask);

/* Asynchronous Reset() of every instance, ordered on `stream` (a cudaStream_t; NULL = default stream). */
int cpmpc_reset_device(cpmpc_handle h, void* stream);

/* Replaces `Optimization::SetPreviousSolution` (optimization.hpp:86-89).  solutions: HOST [B*N].
 * mask: HOST [B] (non-zero = set), NULL = all. */
int cpmpc_set_previous_solution(cpmpc_handle h, const double* solutions, const uint8_t* mask);

/* Reads back the warm-start buffer (what OptimizationOutputs::previous_solution will be on the NEXT Step,
 * optimization.cc:84-85).  solutions_out: HOST [B*N]; has_prev_out: HOST [B] or NULL. */
int cpmpc_get_previous_solution(cpmpc_handle h, double* solutions_out, uint8_t* has_prev_out);

/* Enables (1) / disables (0) the per-iteration log (NLSIteration records); off by default. */
int cpmpc_enable_iteration_log(cpmpc_handle h, int enable);
/* Copies the log of the last Step to HOST log_out [B * max_iterations * CPMPC_ITERLOG_STRIDE]. */
int cpmpc_get_iteration_log(cpmpc_handle h, double* log_out);

/* Replaces `Simulator::Step` (optimization/simulator.cc:11-23) for B plants, DEVICE pointers, async on `stream`.
 *   states [B*4] in/out, dyn_params [B*9], u [B], f_base/f_mass [B*2] or NULL (zero forces). */
int cpmpc_sim_step_device(int64_t batch, double* states, const double* dyn_params, const double* u, double dt,
                          const double* f_base, const double* f_mass, void* stream);
/* HOST-buffer variant of the above on CUDA device `device`, synchronous. */
int cpmpc_sim_step(int device, int64_t batch, double* states, const double* dyn_params, const double* u, double dt,
                   const double* f_base, const double* f_mass);

/* Closed loop of `n_steps` x { Step(states); Simulator::Step(control_dt, u[0]) } entirely on the device
 * (the loop of optimization_test.cc:38-61 / viz/src/application.ts:424-458), HOST buffers, synchronous.
 *   states [B*4] in/out; controls_out [B*n_steps] or NULL; stats_sum_out[4] or NULL receives
 *   {solves, sqp iterations, line-search steps, solves ending QP_INDEFINITE|MAX_LAMBDA|rejected}. */
int cpmpc_closed_loop(cpmpc_handle h, int64_t n_steps, double* states, const double* dyn_params,
                      const double* set_points, const double* weights, double* controls_out,
                      int64_t* stats_sum_out);
/* DEVICE-pointer variant, asynchronous on `stream`; stats_sum (device int64[4]) is accumulated into. */
int cpmpc_closed_loop_device(cpmpc_handle h, int64_t n_steps, double* states, const double* dyn_params,
                             const double* set_points, const double* weights, double* controls_out,
                             int64_t* stats_sum, void* stream);

/* Debug view INSIDE the first SQP iteration, for element-wise checks against the reference's own linearisation and
 * QP step (CreateDynamicalConstraint, optimization/optimization.cc:99-160; QPNullSpaceSolver::Solve,
 * dependencies/mini_opt/source/qp.cc:679-729).  Every instance builds the initial guess of Step() (warm start honoured,
 * not modified), linearises once, solves the QP with Levenberg-Marquardt damping `lambda` (nonlinear.cc:187-189) by
 * `qp_method` (0: what Step() would use, 1: segment-wise Riccati recursion, 2: condensed form) and stops.
 * blocks_out: HOST [B * cpmpc_debug_stride(h)] doubles per instance:
 *   Phi  [nseg][4][4]  d x_end / d x_start of every shooting segment (row-major; columns 0..3 of A_eq's block row)
 *   Gam  [4][W]        d x_end / d u_k of the segment that owns knot k
 *   e    [nseg][4]     shooting defects (b_eq of the block row)
 *   dxs  [NS][4], du [W]   the QP step in the reference's variable order
 *   qp_ok (1 = solved, 0 = not positive definite), f(0), c(0), lambda, 4 reserved. */
int64_t cpmpc_debug_stride(cpmpc_handle h);
int cpmpc_debug_first_qp(cpmpc_handle h, const double* states, const double* dyn_params, const double* set_points,
                         const double* weights, double lambda, int qp_method, double* blocks_out);

/* Per warp of the LAST launch on this handle, CPMPC_WARP_RECORD values: the SM cycles (clock64) it spent on its
 * instances, the first instance of the last batch it drew from the work queue, and the cycles per solver phase
 * (set-up + initial guess, linearisation, condensed QP, Riccati recursion, directional derivatives + penalty, line
 * search, update / exit tests, outputs).  Names the straggler warp, the instances it carried and where its time went
 * when a launch is one wave of warps and its duration is the slowest warp's (bench.py per_rank).
 * Synchronises the device.  cycles_out: HOST [CPMPC_WARP_RECORD * warps]; `capacity` counts warps.  Returns the number of warps; writes
 * nothing when cycles_out is NULL or capacity is too small; -1 on error. */
#define CPMPC_WARP_RECORD 10
int64_t cpmpc_get_warp_cycles(cpmpc_handle h, uint64_t* cycles_out, int64_t capacity);
/* The per-phase entries of that record are only filled by launches made after cpmpc_enable_phase_clocks(h, 1): eight
 * clock reads per SQP iteration cost 1.5 % of a step, so the kernels the library launches by default record the total
 * and the instance only (phases 0).  Enabling selects the general kernel instantiation (same results, bit for bit). */
int cpmpc_enable_phase_clocks(cpmpc_handle h, int enable);

/* Number of kernel launches issued by this library since load (for bench accounting). */
int64_t cpmpc_launch_count(void);

/* Measures the FP64 FMA throughput of the device with a register-resident DFMA chain microkernel and returns
 * it in *tflops_out (2 flops per FMA).  Used as the roofline denominator (MEASURED_PEAKS.json has no FP64 entry). */
int cpmpc_measure_fp64_peak(int device, double* tflops_out);

/* Debug view of the device's elementary functions (csrc/fastmath.cuh: the branch-free sin/cos, tanh, reciprocal and
 * square root the dynamics evaluate): for every x[i] writes {sin x, cos x, tanh x, 1/x, sqrt|x|, 1/sqrt|x|} to
 * out[6 i ...], computed ON THE DEVICE with the hardware seed instructions.  HOST buffers.  The reference has no
 * counterpart (it calls libm); tests compare with libm at the ulp level. */
int cpmpc_debug_fastmath(int device, int64_t n, const double* x, double* out);

/* ---------------------------------------------------------------------------------------------------------------------
 * Plant-generic entry points: the same controller over a plant with NX states per shooting node.  They exist for the
 * cart-mounted DOUBLE pendulum (BASELINE.json configs[4]).  The reference holds only its physics spec
 * (symbolic/dynamics_double.py:25-148; DoubleCartPoleParams at :13-22) and a TODO in the optimiser
 * (optimization/optimization.cc:197-199): there is no reference interface to replace, so these mirror the
 * single-pendulum entry points above with NX-wide arrays.  CPMPC_PLANT_SINGLE runs the single pendulum through the
 * same generic kernels (a cross-check of the generic code against the single-pendulum path; cpmpc_create is the tuned
 * path for it).
 *   state  x = [q, q']: single [b_x, th_1, b_x', th_1'] (NX = 4); double [b_x, th_1, th_2, b_x', th_1', th_2'] (NX = 6)
 *   dyn    single: SingleCartPoleParams (9 doubles); double: {m_b, m_1, m_2, l_1, l_2, g} (6 doubles)
 *   N = NX*NS + W; terminal targets: b_x -> set point, every angle -> +pi/2 (upright), every rate -> 0.
 * ------------------------------------------------------------------------------------------------------------------ */
#define CPMPC_PLANT_SINGLE 0
#define CPMPC_PLANT_DOUBLE 1
#define CPMPC_PLANT_MAX_NX 6
#define CPMPC_DP_STATE_STRIDE 6
#define CPMPC_DP_DYN_STRIDE 6   /* DoubleCartPoleParams, symbolic/dynamics_double.py:13-22 */
#define CPMPC_DP_WEIGHTS_STRIDE 8 /* u, u_derivative, then the six terminal weights in state order */

typedef struct cpmpc_plant_params {
  double control_dt;                    /* 0.01 */
  int64_t window_length;                /* 40 */
  int64_t state_spacing;                /* 10 */
  int64_t max_iterations;               /* 8 */
  double relative_exit_tol;             /* 1e-5 */
  double absolute_first_derivative_tol; /* 1e-6 */
  double equality_penalty_initial;      /* 1.0 */
  double u_guess_sinusoid_amplitude;    /* 10.0 */
  double u_cost_weight;                 /* 0.1 */
  double u_derivative_cost_weight;      /* 0.1 */
  double final_cost_weight[CPMPC_PLANT_MAX_NX]; /* state order; negative => hard equality; 150 for b_x, -1 else */
} cpmpc_plant_params;

typedef struct cpmpc_plant_solver* cpmpc_plant_handle;

int cpmpc_plant_state_dim(int plant);  /* NX, or -1 for an unknown plant */
int cpmpc_plant_dyn_dim(int plant);    /* doubles of dynamics parameters per instance */
void cpmpc_plant_default_params(int plant, cpmpc_plant_params* params);
int cpmpc_plant_create(int plant, const cpmpc_plant_params* params, int64_t batch, int device,
                       cpmpc_plant_handle* out);
int cpmpc_plant_destroy(cpmpc_plant_handle h);
int64_t cpmpc_plant_solution_dim(cpmpc_plant_handle h); /* N */
/* Step for all B instances, HOST buffers, synchronous: states [B*NX], dyn_params [B*dyn_dim], set_points [B],
 * weights [B*(2+NX)] or NULL, u_out [B*W], pred_out [B*W*NX] or NULL, stats_out [B] or NULL. */
int cpmpc_plant_step(cpmpc_plant_handle h, const double* states, const double* dyn_params, const double* set_points,
                     const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out);
/* DEVICE pointers, asynchronous on `stream`. */
int cpmpc_plant_step_device(cpmpc_plant_handle h, const double* states, const double* dyn_params,
                            const double* set_points, const double* weights, double* u_out, double* pred_out,
                            cpmpc_stats* stats_out, void* stream);
int cpmpc_plant_reset(cpmpc_plant_handle h);                      /* Reset() of every instance */
int cpmpc_plant_reset_device(cpmpc_plant_handle h, void* stream); /* same, ordered on `stream` */
int cpmpc_plant_set_previous_solution(cpmpc_plant_handle h, const double* solutions);          /* HOST [B*N] */
int cpmpc_plant_get_previous_solution(cpmpc_plant_handle h, double* solutions_out, uint8_t* has_prev_out);
int cpmpc_plant_enable_iteration_log(cpmpc_plant_handle h, int enable);
int cpmpc_plant_get_iteration_log(cpmpc_plant_handle h, double* log_out);

#ifdef __cplusplus
}
#endif
#endif /* CPMPC_H_ */
