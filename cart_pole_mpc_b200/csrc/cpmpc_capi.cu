// C ABI of libcpmpc.so (declared in include/cpmpc.h).  Host logic only: validation, device buffers, launch
// configuration, host<->device staging.  All arithmetic of the hot path is in solver_kernel.cuh.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cpmpc.h"
#include "capi_common.hpp"
#include "solver_kernel.cuh"

using namespace cpmpc;

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int Fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    const cudaError_t err_ = (expr);                                                                \
    if (err_ != cudaSuccess) {                                                                      \
      return Fail(CPMPC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err_));            \
    }                                                                                               \
  } while (0)

int Pow2Ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

namespace cpmpc {
namespace detail {
int Fail(int code, const std::string& msg) { return ::Fail(code, msg); }
void CountLaunches(int n) { g_launches.fetch_add(n); }
}  // namespace detail
}  // namespace cpmpc

struct cpmpc_solver {
  cpmpc_params params;
  long long B = 0;
  int device = 0;
  int lanes = 1;   // L: lanes per instance
  int grid = 0;    // persistent grid, in blocks
  int warps_per_block = 1;
  size_t smem_bytes = 0;
  KernelCfg cfg;
  cudaStream_t stream = nullptr;
  // persistent device state
  double* d_prev = nullptr;
  unsigned char* d_has_prev = nullptr;
  double* d_u_guess = nullptr;
  double* d_kk = nullptr;
  // split launch (large batches): shooting states of the initial guess [B, 4 NS] from cpmpc_guess_kernel, and the
  // controls for cpmpc_post_kernel when the caller asked for no u_out
  bool split = false, split_loop = false;  // Step / closed loop
  bool phase_clocks = false;  // cpmpc_enable_phase_clocks: launch the general kernel, which times its phases
  double* d_guess = nullptr;
  double* d_u_split = nullptr;
  unsigned long long* d_queue = nullptr;
  double* d_iter_log = nullptr;
  bool log_enabled = false;
  unsigned long long* d_warp_cycles = nullptr;  // [grid * warps_per_block] per-warp SM cycles of the last launch
  // staging for the host-buffer entry points
  double* d_states = nullptr;
  double* d_dyn = nullptr;
  double* d_sp = nullptr;
  double* d_weights = nullptr;
  double* d_u = nullptr;
  double* d_pred = nullptr;
  cpmpc_stats* d_stats = nullptr;
  long long* d_stats_sum = nullptr;
  double* d_controls = nullptr;
  long long controls_capacity = 0;
  // small batches: one packed pinned host buffer and its device twin, so that a Step() costs one copy each way
  unsigned char* h_pack = nullptr;
  unsigned char* d_pack = nullptr;
  // cpmpc_reset_device runs on the caller's stream; the host-buffer entry points run on `stream` and wait for it
  cudaEvent_t reset_event = nullptr;
  bool reset_pending = false;
  // cpmpc_step_async / cpmpc_wait: two slots, each with its own stream and staging buffers, so that the device-to-host
  // copy of step k overlaps the host-to-device copy and the kernel of step k + 1 (the kernels themselves are chained
  // by `kernel_done`: they share the warm start, the work queue and the Riccati scratch)
  struct AsyncSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t kernel_done = nullptr, all_done = nullptr;
    double *d_states = nullptr, *d_dyn = nullptr, *d_sp = nullptr, *d_weights = nullptr, *d_u = nullptr, *d_pred = nullptr;
    cpmpc_stats* d_stats = nullptr;
    bool busy = false;
  } slot[2];
  int next_slot = 0;      // slot the next cpmpc_step_async uses
  int in_flight = 0;      // 0..2
  cudaEvent_t last_kernel_done = nullptr;  // of the most recent asynchronous step (not owned: points into a slot)
};

namespace {

using KernelFn = void (*)(const KernelCfg, const KernelIO);

// mode 0: the general kernel.  1: the lean instantiation - no chunked linearisation, iteration log, debug view or
// in-kernel plant step (solver_kernel.cuh LEAN) - for every plain Step and the split closed loop.  2: lean with the
// plant step, for the fused closed loop.  fixed: additionally the default problem shape (W = 40, S = 10, four lanes)
// at compile time and the handle's weights for every instance (no per-instance weights in the call).  Launch() picks per launch.
template <bool LEAN, bool SIM>
KernelFn SelectKernelOf(int lanes) {
  switch (lanes) {
    case 1: return cpmpc_solve_kernel<1, LEAN, SIM>;
    case 2: return cpmpc_solve_kernel<2, LEAN, SIM>;
    case 4: return cpmpc_solve_kernel<4, LEAN, SIM>;
    case 8: return cpmpc_solve_kernel<8, LEAN, SIM>;
    case 16: return cpmpc_solve_kernel<16, LEAN, SIM>;
    default: return cpmpc_solve_kernel<32, LEAN, SIM>;
  }
}
KernelFn SelectKernel(int lanes, int mode = 0, bool fixed = false) {
  if (fixed && lanes == 4 && mode == 1) return cpmpc_solve_kernel<4, true, false, true>;
  if (fixed && lanes == 4 && mode == 2) return cpmpc_solve_kernel<4, true, true, true>;
  return mode == 2 ? SelectKernelOf<true, true>(lanes)
                   : mode == 1 ? SelectKernelOf<true, false>(lanes) : SelectKernelOf<false, true>(lanes);
}
bool DefaultShape(const KernelCfg& k) { return k.W == FX_W && k.S == FX_S && k.nseg == FX_NSEG && k.NS == FX_NS; }

// Shared-memory layout of one instance (doubles).  `k.cps` must be set.
void BuildLayout(KernelCfg& k) {
  int off = 0;
  auto take = [&off](int n) {
    const int o = off;
    off += n;
    return o;
  };
  const int W = k.W, S = k.S, NS = k.NS, nseg = k.nseg;
  k.nchunk = nseg * k.cps;
  const int n_xk = (k.cps > 1) ? 4 * k.nchunk : 0;
  k.off_xs = take(4 * NS);
  k.off_u = take(W);
  k.off_dxs = take(4 * NS);
  k.off_du = take(W);
  k.off_phi = take(16 * nseg);
  k.off_e = take(4 * nseg);
  k.off_xk = take(n_xk);
  k.off_phic = take((k.cps > 1) ? 16 * k.nchunk : 0);
  // the line-search candidates (shooting states, controls, chunk-start states) are only live after the QP has
  // been solved: they alias the Gamma block
  const int n_cand = 4 * NS + W + n_xk;
  k.off_gam = take(4 * W > n_cand ? 4 * W : n_cand);
  k.off_xc = k.off_gam;
  k.off_uc = k.off_xc + 4 * NS;
  k.off_xkc = k.off_uc + W;
  // QP scratch: the condensed form's Z [4][W] and tinv [W]; the Riccati recursion only stages the feedback rows
  // of ONE segment here for its forward sweep (its value function lives in registers)
  if (k.stage_knots < 1 || k.stage_knots > S) k.stage_knots = S;
  const int n_ric = k.stage_knots * GAIN_STRIDE;
  const int n_cond = k.condensed ? (k.z_global ? W : 5 * W) : 0;
  k.off_aug = take(n_ric > n_cond ? n_ric : n_cond);
  k.off_dyn = take(DYN_CONST_DOUBLES);
  // odd stride (in doubles) so that the same field of neighbouring instances falls into different banks
  k.stride = off | 1;
  k.kk_stride = W * GAIN_STRIDE;
}

int Launch(cpmpc_solver* h, KernelIO io, cudaStream_t stream) {
  io.B = h->B;
  io.u_guess = h->d_u_guess;
  io.prev = h->d_prev;
  io.has_prev = h->d_has_prev;
  io.kk_scratch = h->d_kk;
  io.queue = h->d_queue;
  io.iter_log = h->log_enabled ? h->d_iter_log : nullptr;
  io.warp_cycles = h->d_warp_cycles;
  const bool split = (io.sim ? h->split_loop : h->split) && io.dbg == nullptr;
  const bool lean = h->cfg.cps == 1 && io.iter_log == nullptr && io.dbg == nullptr && !h->phase_clocks;
  KernelFn fn = SelectKernel(h->lanes, lean ? ((io.sim != 0 && !split) ? 2 : 1) : 0, DefaultShape(h->cfg) && io.weights == nullptr);
  if (!split) {
    fn<<<h->grid, 32 * h->warps_per_block, h->smem_bytes, stream>>>(h->cfg, io);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return CPMPC_OK;
  }
  // Split launch: guess rollouts (one thread per instance) -> SQP -> predicted states / plant step (one thread per
  // instance), once per control tick, all stream-ordered.  Results are bit-identical to the fused kernel's.
  if (io.u_out == nullptr) io.u_out = h->d_u_split;
  const bool need_post = io.pred_out != nullptr || io.sim != 0;
  const long long n_ticks = io.n_steps;
  const unsigned wide_blocks = static_cast<unsigned>((h->B + 127) / 128);
  io.n_steps_total = n_ticks;
  io.n_steps = 1;
  io.split = 1;
  io.guess = h->d_guess;
  for (long long tick = 0; tick < n_ticks; ++tick) {
    io.step_index = tick;
    cpmpc_guess_kernel<<<wide_blocks, 128, 0, stream>>>(h->cfg, io, h->d_guess);
    fn<<<h->grid, 32 * h->warps_per_block, h->smem_bytes, stream>>>(h->cfg, io);
    if (need_post) cpmpc_post_kernel<<<wide_blocks, 128, 0, stream>>>(h->cfg, io, io.u_out);
    g_launches.fetch_add(need_post ? 3 : 2);
  }
  CUDA_TRY(cudaGetLastError());
  return CPMPC_OK;
}

// Staging buffers of the host-buffer entry points: allocated together on first use, all or nothing.
int EnsureStaging(cpmpc_solver* h) {
  if (h->d_stats_sum != nullptr) return CPMPC_OK;  // the last one allocated
  const long long B = h->B;
  const int W = h->cfg.W;
  auto alloc_all = [&]() -> int {
    CUDA_TRY(cudaMalloc(&h->d_states, sizeof(double) * B * 4));
    CUDA_TRY(cudaMalloc(&h->d_dyn, sizeof(double) * B * 9));
    CUDA_TRY(cudaMalloc(&h->d_sp, sizeof(double) * B));
    CUDA_TRY(cudaMalloc(&h->d_weights, sizeof(double) * B * CPMPC_WEIGHTS_STRIDE));
    CUDA_TRY(cudaMalloc(&h->d_u, sizeof(double) * B * W));
    CUDA_TRY(cudaMalloc(&h->d_pred, sizeof(double) * B * W * 4));
    CUDA_TRY(cudaMalloc(&h->d_stats, sizeof(cpmpc_stats) * B));
    CUDA_TRY(cudaMalloc(&h->d_stats_sum, sizeof(long long) * 4));
    return CPMPC_OK;
  };
  const int rc = alloc_all();
  if (rc != CPMPC_OK) {
    for (void** p : {reinterpret_cast<void**>(&h->d_states), reinterpret_cast<void**>(&h->d_dyn),
                     reinterpret_cast<void**>(&h->d_sp), reinterpret_cast<void**>(&h->d_weights),
                     reinterpret_cast<void**>(&h->d_u), reinterpret_cast<void**>(&h->d_pred),
                     reinterpret_cast<void**>(&h->d_stats), reinterpret_cast<void**>(&h->d_stats_sum)}) {
      cudaFree(*p);
      *p = nullptr;
    }
  }
  return rc;
}

int EnsureSlot(cpmpc_solver* h, cpmpc_solver::AsyncSlot& sl) {
  if (sl.d_stats != nullptr) return CPMPC_OK;  // the last one allocated
  const long long B = h->B;
  const int W = h->cfg.W;
  auto alloc_all = [&]() -> int {
    if (sl.stream == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    if (sl.kernel_done == nullptr) CUDA_TRY(cudaEventCreateWithFlags(&sl.kernel_done, cudaEventDisableTiming));
    if (sl.all_done == nullptr) CUDA_TRY(cudaEventCreateWithFlags(&sl.all_done, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&sl.d_states, sizeof(double) * B * 4));
    CUDA_TRY(cudaMalloc(&sl.d_dyn, sizeof(double) * B * 9));
    CUDA_TRY(cudaMalloc(&sl.d_sp, sizeof(double) * B));
    CUDA_TRY(cudaMalloc(&sl.d_weights, sizeof(double) * B * CPMPC_WEIGHTS_STRIDE));
    CUDA_TRY(cudaMalloc(&sl.d_u, sizeof(double) * B * W));
    CUDA_TRY(cudaMalloc(&sl.d_pred, sizeof(double) * B * W * 4));
    CUDA_TRY(cudaMalloc(&sl.d_stats, sizeof(cpmpc_stats) * B));
    return CPMPC_OK;
  };
  const int rc = alloc_all();
  if (rc != CPMPC_OK) {
    for (void** p : {reinterpret_cast<void**>(&sl.d_states), reinterpret_cast<void**>(&sl.d_dyn),
                     reinterpret_cast<void**>(&sl.d_sp), reinterpret_cast<void**>(&sl.d_weights),
                     reinterpret_cast<void**>(&sl.d_u), reinterpret_cast<void**>(&sl.d_pred),
                     reinterpret_cast<void**>(&sl.d_stats)}) {
      cudaFree(*p);
      *p = nullptr;
    }
  }
  return rc;
}

// Waits for the oldest asynchronous step in flight (no-op when there is none).
int WaitOldest(cpmpc_solver* h) {
  if (h->in_flight == 0) return CPMPC_OK;
  const int oldest = (h->next_slot + 2 - h->in_flight) % 2;
  cpmpc_solver::AsyncSlot& sl = h->slot[oldest];
  const cudaError_t e = cudaEventSynchronize(sl.all_done);
  sl.busy = false;
  h->in_flight -= 1;
  if (e != cudaSuccess) return Fail(CPMPC_ERR_CUDA, std::string("cpmpc_wait: ") + cudaGetErrorString(e));
  return CPMPC_OK;
}

int DrainAsync(cpmpc_solver* h) {
  while (h->in_flight > 0) {
    if (int rc = WaitOldest(h)) return rc;
  }
  return CPMPC_OK;
}

// Orders the handle's own stream behind a cpmpc_reset_device issued on another stream.
int WaitForReset(cpmpc_solver* h) {
  if (h->reset_pending) {
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->reset_event, 0));
    h->reset_pending = false;
  }
  return CPMPC_OK;
}

// Step() for small batches (single solves above all): the six small transfers of the general path cost more host time
// than the bytes they move, and pageable caller buffers make each of them a synchronous staged copy.  Inputs are
// packed into one pinned buffer, cross the bus as ONE copy, the kernel reads and writes the device twin, and the
// results come back as ONE copy.  Layout: [states B*4 | dyn B*9 | set_points B | weights B*6] then
// [u B*W | predicted B*W*4 | stats B].
constexpr long long kPackedBatchMax = 64;

int StepPacked(cpmpc_solver* h, const double* states, const double* dyn_params, const double* set_points,
               const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out) {
  const long long B = h->B;
  const int W = h->cfg.W;
  const size_t n_state = sizeof(double) * B * 4, n_dyn = sizeof(double) * B * CPMPC_DYN_STRIDE;
  const size_t n_sp = sizeof(double) * B, n_w = sizeof(double) * B * CPMPC_WEIGHTS_STRIDE;
  const size_t in_bytes = n_state + n_dyn + n_sp + n_w;
  const size_t n_u = sizeof(double) * B * W, n_pred = sizeof(double) * B * W * 4, n_stats = sizeof(cpmpc_stats) * B;
  const size_t out_bytes = n_u + n_pred + n_stats;
  if (h->h_pack == nullptr) {
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&h->h_pack), in_bytes + out_bytes));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->d_pack), in_bytes + out_bytes));
  }
  unsigned char* hp = h->h_pack;
  std::memcpy(hp, states, n_state);
  std::memcpy(hp + n_state, dyn_params, n_dyn);
  std::memcpy(hp + n_state + n_dyn, set_points, n_sp);
  if (weights != nullptr) std::memcpy(hp + n_state + n_dyn + n_sp, weights, n_w);
  cudaStream_t s = h->stream;
  unsigned char* dp = h->d_pack;
  CUDA_TRY(cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, s));
  unsigned char* dout = dp + in_bytes;
  if (int rc = cpmpc_step_device(h, reinterpret_cast<double*>(dp), reinterpret_cast<double*>(dp + n_state),
                                 reinterpret_cast<double*>(dp + n_state + n_dyn),
                                 weights ? reinterpret_cast<double*>(dp + n_state + n_dyn + n_sp) : nullptr,
                                 reinterpret_cast<double*>(dout),
                                 pred_out ? reinterpret_cast<double*>(dout + n_u) : nullptr,
                                 stats_out ? reinterpret_cast<cpmpc_stats*>(dout + n_u + n_pred) : nullptr, s)) {
    return rc;
  }
  // only what the caller asked for comes back: [u] or [u | predicted] or everything
  const size_t back = stats_out ? out_bytes : (pred_out ? n_u + n_pred : n_u);
  CUDA_TRY(cudaMemcpyAsync(hp + in_bytes, dout, back, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  std::memcpy(u_out, hp + in_bytes, n_u);
  if (pred_out != nullptr) std::memcpy(pred_out, hp + in_bytes + n_u, n_pred);
  if (stats_out != nullptr) std::memcpy(stats_out, hp + in_bytes + n_u + n_pred, n_stats);
  return CPMPC_OK;
}

}  // namespace

extern "C" {

const char* cpmpc_last_error(void) { return g_last_error.c_str(); }
int cpmpc_version(void) { return CPMPC_VERSION; }
int64_t cpmpc_launch_count(void) { return g_launches.load(); }

void cpmpc_default_params(cpmpc_params* p) {
  if (p == nullptr) return;
  p->control_dt = 0.01;
  p->window_length = 40;
  p->state_spacing = 10;
  p->max_iterations = 8;
  p->relative_exit_tol = 1.0e-5;
  p->absolute_first_derivative_tol = 1.0e-6;
  p->equality_penalty_initial = 1.0;
  p->u_guess_sinusoid_amplitude = 10.0;
  p->u_cost_weight = 0.1;
  p->u_derivative_cost_weight = 0.1;
  p->b_x_final_cost_weight = 150.0;
  p->th_final_cost_weight = -1.0;
  p->b_x_dot_final_cost_weight = -1.0;
  p->th_dot_final_cost_weight = -1.0;
}

int cpmpc_create(const cpmpc_params* params, int64_t batch, int device, cpmpc_handle* out) {
  return cpmpc_create_ex(params, batch, device, 0u, out);
}

int cpmpc_create_ex(const cpmpc_params* params, int64_t batch, int device, uint32_t flags, cpmpc_handle* out) {
  if (params == nullptr || out == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  // preconditions of Optimization::Optimization (optimization.cc:13-22) and CheckParams (nonlinear.cc:49-73)
  if (!(params->control_dt > 0)) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "control_dt must be > 0");
  if (params->window_length < 1) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "window_length must be >= 1");
  if (params->state_spacing < 1 || params->window_length % params->state_spacing != 0) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "state_spacing (" + std::to_string(params->state_spacing) +
                                                ") must divide into window_length (" +
                                                std::to_string(params->window_length) + ") cleanly");
  }
  if (params->max_iterations < 1) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "max_iterations must be >= 1");
  if (!(params->u_cost_weight >= 0.0)) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "u_cost_weight must be >= 0");
  if (!(params->u_derivative_cost_weight >= 0.0)) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "u_derivative_cost_weight must be >= 0");
  }
  if (!(params->relative_exit_tol >= 0 && params->relative_exit_tol <= 1)) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "relative_exit_tol must be in [0, 1]");
  }
  if (!(params->absolute_first_derivative_tol >= 0)) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "absolute_first_derivative_tol must be >= 0");
  }
  if (!(params->equality_penalty_initial >= 0)) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "equality_penalty_initial must be >= 0");
  }
  if (batch < 1) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "batch must be >= 1");
  if (params->window_length > 4096) return Fail(CPMPC_ERR_UNSUPPORTED, "window_length > 4096 is not supported");

  int n_dev = 0;
  CUDA_TRY(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) return Fail(CPMPC_ERR_CUDA, "no such CUDA device: " + std::to_string(device));
  CUDA_TRY(cudaSetDevice(device));

  cpmpc_solver* h = new cpmpc_solver();
  // any failure below releases the handle and whatever it already owns
  struct Guard {
    cpmpc_solver* h;
    ~Guard() {
      if (h != nullptr) cpmpc_destroy(h);
    }
  } guard{h};
  h->params = *params;
  h->B = batch;
  h->device = device;
  KernelCfg& k = h->cfg;
  std::memset(&k, 0, sizeof(k));
  k.W = static_cast<int>(params->window_length);
  k.S = static_cast<int>(params->state_spacing);
  k.nseg = k.W / k.S;
  k.NS = k.nseg + 1;
  k.N = 4 * k.NS + k.W;
  k.max_iter = static_cast<int>(params->max_iterations);
  k.max_ls = 5;  // optimization.cc:76
  k.dt = params->control_dt;
  k.rel_tol = params->relative_exit_tol;
  k.first_tol = params->absolute_first_derivative_tol;
  k.abs_tol = 1.0e-12;      // nonlinear.hpp:75
  k.pen_init = params->equality_penalty_initial;
  k.pen_scale = 1.01;       // nonlinear.hpp:98
  k.rho = 0.1;              // nonlinear.hpp:102
  k.armijo_c1 = 1.0e-4;     // nonlinear.cc:121
  k.lam_fail_init = 1.0e-2; // nonlinear.hpp:108
  k.lam_dec_success = 0.1;  // nonlinear.hpp:111
  k.lam_dec_restore = 0.8;  // nonlinear.hpp:114
  k.lam_max = 1.0;          // nonlinear.hpp:117
  k.lam_min = 0.0;          // nonlinear.hpp:120
  k.w_u = params->u_cost_weight;
  k.w_r = params->u_derivative_cost_weight;
  k.wt[0] = params->b_x_final_cost_weight;
  k.wt[1] = params->th_final_cost_weight;
  k.wt[2] = params->b_x_dot_final_cost_weight;
  k.wt[3] = params->th_dot_final_cost_weight;
  k.x_clamp = 5.0;    // optimization.cc:320
  k.u_clamp = 300.0;  // optimization.cc:327
  // Lanes per instance.  One lane per shooting segment is the minimum that keeps the merit rollouts parallel; twice
  // that (when a segment has at least two knots) lets the linearisation run two chunks per segment side by side and
  // halves its dependent chain.  Overridable for experiments.
  int base_lanes = Pow2Ceil(k.nseg);
  if (base_lanes > 32) base_lanes = 32;
  // At most 8 lanes per instance in the throughput mapping: the serial phases of a solve (knot-level Riccati recursion,
  // initial-guess and predicted-state rollouts) run on one lane per instance, so wider groups leave more of every warp
  // idle there (W = 160: 16 lanes 1.16 M solves/s, 8 lanes 1.26 M, 4 lanes 0.84 M on one B200,
  // profiles/r02_j_config4_lanes.txt); a lane then carries several shooting segments one after the other.
  int lanes = base_lanes < 8 ? base_lanes : 8;
  // CPMPC_FLAG_LOW_LATENCY: two lanes per segment.  It shortens one solve's dependent chain (batch-1 latency
  // 0.73 -> 0.60 ms at the defaults) but leaves half of every warp idle during the merit rollouts: 3.88 M solves/s
  // instead of 4.28 M at B = 4096, 3.6 M instead of 5.3 M at B = 65536 (profiles/r01_lane_sweep.jsonl).  It is
  // opt-in rather than picked from the batch size because the chunked linearisation rounds differently, and results
  // must not depend on how a caller happens to group instances into batches.
  if ((flags & CPMPC_FLAG_LOW_LATENCY) != 0u && 2 * base_lanes <= 32 && k.S >= 2 && (2 * base_lanes) / k.nseg >= 2) {
    lanes = 2 * base_lanes;
  }
  // CPMPC_FLAG_MIN_LATENCY: a whole warp per instance.  The chunk count is the smallest one that reaches the shortest
  // possible chunk (ceil(S / floor(32 / segments)) knots): 5 chunks of 2 knots at the defaults, not 8 ragged ones.
  int min_latency_cps = 0;
  if ((flags & CPMPC_FLAG_MIN_LATENCY) != 0u && k.S >= 2 && 32 / k.nseg >= 2) {
    const int max_cps = (32 / k.nseg < k.S) ? 32 / k.nseg : k.S;
    const int chain = (k.S + max_cps - 1) / max_cps;
    min_latency_cps = (k.S + chain - 1) / chain;
    lanes = 32;
  }
  if (const char* env = std::getenv("CPMPC_LANES")) {
    const int v = std::atoi(env);
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) lanes = v;
  }
  // QP method for the undamped iterations: condensed while the horizon is short enough for open-loop sensitivity
  // products to stay accurate, Riccati beyond (and always for Levenberg-Marquardt iterations)
  k.condensed = (k.W * k.dt <= 0.6) ? 1 : 0;
  if (const char* env = std::getenv("CPMPC_QP")) {
    if (std::strcmp(env, "riccati") == 0) k.condensed = 0;
    if (std::strcmp(env, "condensed") == 0) k.condensed = 1;
  }
  k.z_global = 0;
  k.stage_knots = k.S;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const bool chunked = lanes > base_lanes;
  size_t smem_per_warp = 0;
  for (;;) {
    // chunks per segment: only the low-latency mapping splits segments; lanes added below to fit shared memory
    // (long single-shooting segments need a large Riccati stage matrix per instance) just widen the QP loops
    k.cps = chunked ? (2 * base_lanes) / k.nseg : 1;
    if (k.cps < 1) k.cps = 1;
    if (k.cps > k.S) k.cps = k.S;
    if (min_latency_cps > 0 && min_latency_cps * k.nseg <= lanes) k.cps = min_latency_cps;
    if (const char* env = std::getenv("CPMPC_CHUNKS")) {
      const int v = std::atoi(env);
      if (v >= 1 && v <= k.S && v * k.nseg <= lanes) k.cps = v;
    }
    BuildLayout(k);
    smem_per_warp = sizeof(double) * static_cast<size_t>(k.stride) * (32 / lanes);
    if (smem_per_warp <= static_cast<size_t>(prop.sharedMemPerBlockOptin) || lanes == 32) break;
    lanes *= 2;
  }
  h->lanes = lanes;
  const int groups_per_warp = 32 / lanes;
  if (smem_per_warp > static_cast<size_t>(prop.sharedMemPerBlockOptin)) {
    return Fail(CPMPC_ERR_UNSUPPORTED, "problem needs " + std::to_string(smem_per_warp) +
                                           " bytes of shared memory per warp, device offers " +
                                           std::to_string(prop.sharedMemPerBlockOptin));
  }
  // One block per SM, up to 8 warps (two per SM sub-partition; 8 x 32 x 255 registers fill the register file).  A
  // batch that needs fewer warps than that is spread evenly over the SMs instead of filling some and idling others.
  int wpb_max = 8;
  if (const char* env = std::getenv("CPMPC_WARPS_PER_BLOCK")) {
    const int v = std::atoi(env);
    if (v >= 1 && v <= 8) wpb_max = v;
  }
  const int wpb_cap = wpb_max;
  while (wpb_max > 1 && smem_per_warp * wpb_max > static_cast<size_t>(prop.sharedMemPerBlockOptin)) --wpb_max;
  const long long warps_needed = (batch + groups_per_warp - 1) / groups_per_warp;
  // Compact layout for batches that do not fit one wave of the roomy one: the condensed QP's Z rows move from shared
  // memory to the group's L2-resident global scratch (the Riccati recursion's, never live in the same iteration) and
  // the Riccati forward sweep stages fewer knots at a time, until a warp more fits on the SM (defaults: 573 -> 437
  // doubles per instance, 6 -> 8 warps per SM, two per scheduler).  The arithmetic does not change: same bits.  At
  // equal warps per SM global Z rows cost 5 % (L2 round trips inside the substitutions), so a batch that fits one wave
  // anyway keeps them in shared memory (4096 instances: 0.704 against 0.745 ms); eight warps instead of six are worth
  // 11 % at 65536 instances (profiles/r02_w_layout_sweep.jsonl).
  const char* env_z = std::getenv("CPMPC_ZGLOBAL");
  const char* env_sk = std::getenv("CPMPC_STAGE_KNOTS");
  const bool want_compact = env_z != nullptr ? std::atoi(env_z) != 0
                                             : warps_needed > static_cast<long long>(wpb_max) * prop.multiProcessorCount;
  if (want_compact && k.condensed && wpb_max < wpb_cap) {
    KernelCfg best = k;
    int best_wpb = wpb_max;
    for (int sk = k.S; sk >= 1; --sk) {
      KernelCfg t = k;
      t.z_global = 1;
      t.stage_knots = env_sk != nullptr ? std::atoi(env_sk) : sk;
      BuildLayout(t);
      const size_t per_warp = sizeof(double) * static_cast<size_t>(t.stride) * groups_per_warp;
      int w = wpb_cap;
      while (w > 1 && per_warp * w > static_cast<size_t>(prop.sharedMemPerBlockOptin)) --w;
      if (w > best_wpb) {
        best = t;
        best_wpb = w;
      }
      if (best_wpb == wpb_cap || env_sk != nullptr) break;
    }
    if (best_wpb > wpb_max || env_z != nullptr) {
      k = best;
      wpb_max = best_wpb;
      smem_per_warp = sizeof(double) * static_cast<size_t>(k.stride) * groups_per_warp;
    }
  }
  int wpb = wpb_max;
  if (warps_needed <= static_cast<long long>(wpb_max) * prop.multiProcessorCount) {
    wpb = static_cast<int>((warps_needed + prop.multiProcessorCount - 1) / prop.multiProcessorCount);
    if (wpb < 1) wpb = 1;
  }
  h->warps_per_block = wpb;
  h->smem_bytes = smem_per_warp * wpb;
  KernelFn fn = SelectKernel(lanes);
  // The attribute belongs to the kernel instantiation (per device), not to the handle: handles that share an
  // instantiation but differ in warps per block must not lower each other's limit, so it is always set to the
  // device maximum.
  CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(prop.sharedMemPerBlockOptin)));
  for (int mode = 1; mode <= 2; ++mode) {
    for (int fixed = 0; fixed <= 1; ++fixed) {  // per-instance weights select the non-fixed instantiation at launch
      CUDA_TRY(cudaFuncSetAttribute(SelectKernel(lanes, mode, fixed != 0 && DefaultShape(k)),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(prop.sharedMemPerBlockOptin)));
    }
  }
  int blocks_per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fn, 32 * wpb, h->smem_bytes));
  if (blocks_per_sm < 1) {
    return Fail(CPMPC_ERR_UNSUPPORTED, "kernel does not fit on an SM");
  }
  const long long groups_per_block = static_cast<long long>(groups_per_warp) * wpb;
  const long long blocks_needed = (batch + groups_per_block - 1) / groups_per_block;
  const long long resident = static_cast<long long>(blocks_per_sm) * prop.multiProcessorCount;
  h->grid = static_cast<int>(blocks_needed < resident ? blocks_needed : resident);

  const long long n_groups = static_cast<long long>(h->grid) * groups_per_block;
  CUDA_TRY(cudaMalloc(&h->d_prev, sizeof(double) * batch * k.N));
  CUDA_TRY(cudaMalloc(&h->d_has_prev, batch));
  CUDA_TRY(cudaMemset(h->d_has_prev, 0, batch));
  CUDA_TRY(cudaMemset(h->d_prev, 0, sizeof(double) * batch * k.N));
  CUDA_TRY(cudaMalloc(&h->d_kk, sizeof(double) * n_groups * k.kk_stride));
  CUDA_TRY(cudaMalloc(&h->d_warp_cycles, sizeof(unsigned long long) * WARP_REC * h->grid * wpb));
  CUDA_TRY(cudaMemset(h->d_warp_cycles, 0, sizeof(unsigned long long) * WARP_REC * h->grid * wpb));
  // Batches beyond one wave of the persistent grid run the serial rollouts of a Step in their own one-thread-per-
  // instance kernels (Launch): within one wave the step is the latency of ONE solve and two extra launches only add to
  // it.  The closed loop gains less (its fused form has no tail per control tick: warps run on into the next tick) and
  // breaks even at about three waves (profiles/r02_y_closed_loop_split_vs_fused.jsonl).  Results do not depend on the choice.
  h->split = batch > n_groups;
  h->split_loop = batch >= 3 * n_groups;
  if ((flags & CPMPC_FLAG_SPLIT_LAUNCH) != 0u) h->split = h->split_loop = true;
  if ((flags & CPMPC_FLAG_FUSED_LAUNCH) != 0u) h->split = h->split_loop = false;
  if (const char* env = std::getenv("CPMPC_SPLIT")) h->split = h->split_loop = std::atoi(env) != 0;
  if (k.cps != 1) h->split = h->split_loop = false;  // the chunked linearisation wants the guess's chunk-start states
  if (h->split || h->split_loop) {
    CUDA_TRY(cudaMalloc(&h->d_guess, sizeof(double) * batch * 4 * k.NS));
    CUDA_TRY(cudaMalloc(&h->d_u_split, sizeof(double) * batch * k.W));
  }
  CUDA_TRY(cudaMalloc(&h->d_queue, sizeof(unsigned long long) * 2));
  CUDA_TRY(cudaMemset(h->d_queue, 0, sizeof(unsigned long long) * 2));
  CUDA_TRY(cudaMalloc(&h->d_u_guess, sizeof(double) * k.W));
  // cold-start sinusoid with the host's libm, exactly as optimization.cc:63-67 evaluates it
  std::vector<double> guess(k.W);
  for (int i = 0; i < k.W; ++i) {
    guess[i] = params->u_guess_sinusoid_amplitude *
               std::sin(static_cast<double>(i) / static_cast<double>(k.W) * 2 * M_PI);
  }
  CUDA_TRY(cudaMemcpy(h->d_u_guess, guess.data(), sizeof(double) * k.W, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->reset_event, cudaEventDisableTiming));
  guard.h = nullptr;
  *out = h;
  return CPMPC_OK;
}

int cpmpc_destroy(cpmpc_handle h) {
  if (h == nullptr) return CPMPC_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_prev);
  cudaFree(h->d_has_prev);
  cudaFree(h->d_u_guess);
  cudaFree(h->d_kk);
  cudaFree(h->d_guess);
  cudaFree(h->d_u_split);
  cudaFree(h->d_queue);
  cudaFree(h->d_warp_cycles);
  cudaFree(h->d_iter_log);
  cudaFree(h->d_states);
  cudaFree(h->d_dyn);
  cudaFree(h->d_sp);
  cudaFree(h->d_weights);
  cudaFree(h->d_u);
  cudaFree(h->d_pred);
  cudaFree(h->d_stats);
  cudaFree(h->d_stats_sum);
  cudaFree(h->d_controls);
  cudaFree(h->d_pack);
  if (h->h_pack) cudaFreeHost(h->h_pack);
  for (cpmpc_solver::AsyncSlot& sl : h->slot) {
    if (sl.stream) cudaStreamSynchronize(sl.stream);
    cudaFree(sl.d_states);
    cudaFree(sl.d_dyn);
    cudaFree(sl.d_sp);
    cudaFree(sl.d_weights);
    cudaFree(sl.d_u);
    cudaFree(sl.d_pred);
    cudaFree(sl.d_stats);
    if (sl.kernel_done) cudaEventDestroy(sl.kernel_done);
    if (sl.all_done) cudaEventDestroy(sl.all_done);
    if (sl.stream) cudaStreamDestroy(sl.stream);
  }
  if (h->reset_event) cudaEventDestroy(h->reset_event);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CPMPC_OK;
}

int64_t cpmpc_batch(cpmpc_handle h) { return h ? h->B : 0; }
int64_t cpmpc_solution_dim(cpmpc_handle h) { return h ? h->cfg.N : 0; }

int cpmpc_step_device(cpmpc_handle h, const double* states, const double* dyn_params, const double* set_points,
                      const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out,
                      void* stream) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  if (states == nullptr || dyn_params == nullptr || set_points == nullptr || u_out == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "states, dyn_params, set_points and u_out are required");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  KernelIO io;
  std::memset(&io, 0, sizeof(io));
  io.states = states;
  io.dyn = dyn_params;
  io.set_points = set_points;
  io.weights = weights;
  io.u_out = u_out;
  io.pred_out = pred_out;
  io.stats = stats_out;
  io.n_steps = 1;
  io.sim = 0;
  return Launch(h, io, static_cast<cudaStream_t>(stream));
}

int cpmpc_step(cpmpc_handle h, const double* states, const double* dyn_params, const double* set_points,
               const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  if (states == nullptr || dyn_params == nullptr || set_points == nullptr || u_out == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "states, dyn_params, set_points and u_out are required");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  if (int rc = DrainAsync(h)) return rc;
  if (int rc = WaitForReset(h)) return rc;
  if (h->B <= kPackedBatchMax) return StepPacked(h, states, dyn_params, set_points, weights, u_out, pred_out, stats_out);
  if (int rc = EnsureStaging(h)) return rc;
  const long long B = h->B;
  const int W = h->cfg.W;
  cudaStream_t s = h->stream;
  CUDA_TRY(cudaMemcpyAsync(h->d_states, states, sizeof(double) * B * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_dyn, dyn_params, sizeof(double) * B * 9, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_sp, set_points, sizeof(double) * B, cudaMemcpyHostToDevice, s));
  if (weights != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(h->d_weights, weights, sizeof(double) * B * CPMPC_WEIGHTS_STRIDE,
                             cudaMemcpyHostToDevice, s));
  }
  if (int rc = cpmpc_step_device(h, h->d_states, h->d_dyn, h->d_sp, weights ? h->d_weights : nullptr, h->d_u,
                                 pred_out ? h->d_pred : nullptr, stats_out ? h->d_stats : nullptr, s)) {
    return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(u_out, h->d_u, sizeof(double) * B * W, cudaMemcpyDeviceToHost, s));
  if (pred_out != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(pred_out, h->d_pred, sizeof(double) * B * W * 4, cudaMemcpyDeviceToHost, s));
  }
  if (stats_out != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(stats_out, h->d_stats, sizeof(cpmpc_stats) * B, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return CPMPC_OK;
}

int cpmpc_step_async(cpmpc_handle h, const double* states, const double* dyn_params, const double* set_points,
                     const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out, uint32_t flags) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  if (states == nullptr || dyn_params == nullptr || set_points == nullptr || u_out == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "states, dyn_params, set_points and u_out are required");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->in_flight == 2) {
    if (int rc = WaitOldest(h)) return rc;  // both slots taken: the oldest step has to finish first
  }
  cpmpc_solver::AsyncSlot& sl = h->slot[h->next_slot];
  if (int rc = EnsureSlot(h, sl)) return rc;
  const long long B = h->B;
  const int W = h->cfg.W;
  cudaStream_t s = sl.stream;
  // inputs: may overlap the previous step's kernel and its device-to-host copy
  CUDA_TRY(cudaMemcpyAsync(sl.d_states, states, sizeof(double) * B * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(sl.d_dyn, dyn_params, sizeof(double) * B * 9, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(sl.d_sp, set_points, sizeof(double) * B, cudaMemcpyHostToDevice, s));
  if (weights != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(sl.d_weights, weights, sizeof(double) * B * CPMPC_WEIGHTS_STRIDE, cudaMemcpyHostToDevice, s));
  }
  // the kernel: after everything that touches the warm start (previous asynchronous kernel, synchronous calls on the
  // handle's own stream, a reset issued on a caller stream)
  if (h->last_kernel_done != nullptr) CUDA_TRY(cudaStreamWaitEvent(s, h->last_kernel_done, 0));
  if (h->reset_pending) {
    CUDA_TRY(cudaStreamWaitEvent(s, h->reset_event, 0));
    h->reset_pending = false;
  }
  if ((flags & CPMPC_STEP_RESET) != 0u) CUDA_TRY(cudaMemsetAsync(h->d_has_prev, 0, h->B, s));
  if (int rc = cpmpc_step_device(h, sl.d_states, sl.d_dyn, sl.d_sp, weights ? sl.d_weights : nullptr, sl.d_u,
                                 pred_out ? sl.d_pred : nullptr, stats_out ? sl.d_stats : nullptr, s)) {
    return rc;
  }
  CUDA_TRY(cudaEventRecord(sl.kernel_done, s));
  h->last_kernel_done = sl.kernel_done;
  // results: overlap the next step's copy-in and kernel
  CUDA_TRY(cudaMemcpyAsync(u_out, sl.d_u, sizeof(double) * B * W, cudaMemcpyDeviceToHost, s));
  if (pred_out != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(pred_out, sl.d_pred, sizeof(double) * B * W * 4, cudaMemcpyDeviceToHost, s));
  }
  if (stats_out != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(stats_out, sl.d_stats, sizeof(cpmpc_stats) * B, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaEventRecord(sl.all_done, s));
  sl.busy = true;
  h->in_flight += 1;
  h->next_slot ^= 1;
  return CPMPC_OK;
}

int cpmpc_wait(cpmpc_handle h) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  return WaitOldest(h);
}

int cpmpc_reset(cpmpc_handle h, const uint8_t* mask) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (int rc = DrainAsync(h)) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (mask == nullptr) {
    CUDA_TRY(cudaMemset(h->d_has_prev, 0, h->B));
    return CPMPC_OK;
  }
  std::vector<unsigned char> flags(h->B);
  CUDA_TRY(cudaMemcpy(flags.data(), h->d_has_prev, h->B, cudaMemcpyDeviceToHost));
  for (long long i = 0; i < h->B; ++i) {
    if (mask[i]) flags[i] = 0;
  }
  CUDA_TRY(cudaMemcpy(h->d_has_prev, flags.data(), h->B, cudaMemcpyHostToDevice));
  return CPMPC_OK;
}

int cpmpc_reset_device(cpmpc_handle h, void* stream) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_has_prev, 0, h->B, static_cast<cudaStream_t>(stream)));
  // the host-buffer entry points run on the handle's own (non-blocking) stream: they wait for this event
  CUDA_TRY(cudaEventRecord(h->reset_event, static_cast<cudaStream_t>(stream)));
  h->reset_pending = true;
  return CPMPC_OK;
}

int cpmpc_set_previous_solution(cpmpc_handle h, const double* solutions, const uint8_t* mask) {
  if (h == nullptr || solutions == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  if (int rc = DrainAsync(h)) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  const int N = h->cfg.N;
  if (mask == nullptr) {
    CUDA_TRY(cudaMemcpy(h->d_prev, solutions, sizeof(double) * h->B * N, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(h->d_has_prev, 1, h->B));
    return CPMPC_OK;
  }
  std::vector<unsigned char> flags(h->B);
  CUDA_TRY(cudaMemcpy(flags.data(), h->d_has_prev, h->B, cudaMemcpyDeviceToHost));
  for (long long i = 0; i < h->B; ++i) {
    if (mask[i]) {
      CUDA_TRY(cudaMemcpy(h->d_prev + i * N, solutions + i * N, sizeof(double) * N, cudaMemcpyHostToDevice));
      flags[i] = 1;
    }
  }
  CUDA_TRY(cudaMemcpy(h->d_has_prev, flags.data(), h->B, cudaMemcpyHostToDevice));
  return CPMPC_OK;
}

int cpmpc_get_previous_solution(cpmpc_handle h, double* solutions_out, uint8_t* has_prev_out) {
  if (h == nullptr || solutions_out == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  if (int rc = DrainAsync(h)) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(solutions_out, h->d_prev, sizeof(double) * h->B * h->cfg.N, cudaMemcpyDeviceToHost));
  if (has_prev_out != nullptr) {
    CUDA_TRY(cudaMemcpy(has_prev_out, h->d_has_prev, h->B, cudaMemcpyDeviceToHost));
  }
  return CPMPC_OK;
}

int cpmpc_enable_iteration_log(cpmpc_handle h, int enable) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (enable && h->d_iter_log == nullptr) {
    const size_t bytes = sizeof(double) * h->B * h->cfg.max_iter * CPMPC_ITERLOG_STRIDE;
    CUDA_TRY(cudaMalloc(&h->d_iter_log, bytes));
    CUDA_TRY(cudaMemset(h->d_iter_log, 0, bytes));
  }
  h->log_enabled = enable != 0;
  return CPMPC_OK;
}

int cpmpc_get_iteration_log(cpmpc_handle h, double* log_out) {
  if (h == nullptr || log_out == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  if (h->d_iter_log == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "iteration log was never enabled");
  CUDA_TRY(cudaSetDevice(h->device));
  if (int rc = DrainAsync(h)) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(log_out, h->d_iter_log, sizeof(double) * h->B * h->cfg.max_iter * CPMPC_ITERLOG_STRIDE,
                      cudaMemcpyDeviceToHost));
  return CPMPC_OK;
}

int cpmpc_sim_step_device(int64_t batch, double* states, const double* dyn_params, const double* u, double dt,
                          const double* f_base, const double* f_mass, void* stream) {
  if (batch < 1 || states == nullptr || dyn_params == nullptr || u == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "batch, states, dyn_params and u are required");
  }
  if (!(dt >= 0.0)) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "dt must be >= 0");  // simulator.cc:13
  // runs on the CURRENT device of the calling thread: the pointers decide where the data lives
  const int threads = 128;
  const long long blocks = (batch + threads - 1) / threads;
  cpmpc_sim_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      batch, states, dyn_params, u, dt, f_base, f_mass);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return CPMPC_OK;
}

int cpmpc_sim_step(int device, int64_t batch, double* states, const double* dyn_params, const double* u, double dt,
                   const double* f_base, const double* f_mass) {
  if (batch < 1 || states == nullptr || dyn_params == nullptr || u == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "batch, states, dyn_params and u are required");
  }
  for (int64_t i = 0; i < batch; ++i) {
    if (!std::isfinite(u[i])) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "u must be finite");  // simulator.cc:14
  }
  CUDA_TRY(cudaSetDevice(device));
  double *d_s = nullptr, *d_d = nullptr, *d_u = nullptr, *d_fb = nullptr, *d_fm = nullptr;
  auto run = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_s, sizeof(double) * batch * 4));
    CUDA_TRY(cudaMalloc(&d_d, sizeof(double) * batch * 9));
    CUDA_TRY(cudaMalloc(&d_u, sizeof(double) * batch));
    CUDA_TRY(cudaMemcpy(d_s, states, sizeof(double) * batch * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_d, dyn_params, sizeof(double) * batch * 9, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_u, u, sizeof(double) * batch, cudaMemcpyHostToDevice));
    if (f_base != nullptr) {
      CUDA_TRY(cudaMalloc(&d_fb, sizeof(double) * batch * 2));
      CUDA_TRY(cudaMemcpy(d_fb, f_base, sizeof(double) * batch * 2, cudaMemcpyHostToDevice));
    }
    if (f_mass != nullptr) {
      CUDA_TRY(cudaMalloc(&d_fm, sizeof(double) * batch * 2));
      CUDA_TRY(cudaMemcpy(d_fm, f_mass, sizeof(double) * batch * 2, cudaMemcpyHostToDevice));
    }
    if (int rc = cpmpc_sim_step_device(batch, d_s, d_d, d_u, dt, d_fb, d_fm, nullptr)) return rc;
    CUDA_TRY(cudaMemcpy(states, d_s, sizeof(double) * batch * 4, cudaMemcpyDeviceToHost));
    return CPMPC_OK;
  };
  const int rc = run();  // temporaries are released on every path
  cudaFree(d_s);
  cudaFree(d_d);
  cudaFree(d_u);
  cudaFree(d_fb);
  cudaFree(d_fm);
  return rc;
}

int cpmpc_closed_loop_device(cpmpc_handle h, int64_t n_steps, double* states, const double* dyn_params,
                             const double* set_points, const double* weights, double* controls_out,
                             int64_t* stats_sum, void* stream) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  if (n_steps < 1 || states == nullptr || dyn_params == nullptr || set_points == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "n_steps >= 1, states, dyn_params and set_points are required");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  KernelIO io;
  std::memset(&io, 0, sizeof(io));
  io.states = states;
  io.states_io = states;
  io.dyn = dyn_params;
  io.set_points = set_points;
  io.weights = weights;
  io.n_steps = n_steps;
  io.sim = 1;
  io.controls_out = controls_out;
  io.stats_sum = reinterpret_cast<long long*>(stats_sum);
  return Launch(h, io, static_cast<cudaStream_t>(stream));
}

int cpmpc_closed_loop(cpmpc_handle h, int64_t n_steps, double* states, const double* dyn_params,
                      const double* set_points, const double* weights, double* controls_out,
                      int64_t* stats_sum_out) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  if (n_steps < 1 || states == nullptr || dyn_params == nullptr || set_points == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "n_steps >= 1, states, dyn_params and set_points are required");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  if (int rc = DrainAsync(h)) return rc;
  if (int rc = WaitForReset(h)) return rc;
  if (int rc = EnsureStaging(h)) return rc;
  const long long B = h->B;
  cudaStream_t s = h->stream;
  if (controls_out != nullptr && h->controls_capacity < B * n_steps) {
    cudaFree(h->d_controls);
    h->d_controls = nullptr;
    CUDA_TRY(cudaMalloc(&h->d_controls, sizeof(double) * B * n_steps));
    h->controls_capacity = B * n_steps;
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_states, states, sizeof(double) * B * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_dyn, dyn_params, sizeof(double) * B * 9, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_sp, set_points, sizeof(double) * B, cudaMemcpyHostToDevice, s));
  if (weights != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(h->d_weights, weights, sizeof(double) * B * CPMPC_WEIGHTS_STRIDE,
                             cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(cudaMemsetAsync(h->d_stats_sum, 0, sizeof(long long) * 4, s));
  if (int rc = cpmpc_closed_loop_device(h, n_steps, h->d_states, h->d_dyn, h->d_sp,
                                        weights ? h->d_weights : nullptr, controls_out ? h->d_controls : nullptr,
                                        reinterpret_cast<int64_t*>(h->d_stats_sum), s)) {
    return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(states, h->d_states, sizeof(double) * B * 4, cudaMemcpyDeviceToHost, s));
  if (controls_out != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(controls_out, h->d_controls, sizeof(double) * B * n_steps, cudaMemcpyDeviceToHost, s));
  }
  if (stats_sum_out != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(stats_sum_out, h->d_stats_sum, sizeof(long long) * 4, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return CPMPC_OK;
}

int cpmpc_enable_phase_clocks(cpmpc_handle h, int enable) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  h->phase_clocks = enable != 0;
  return CPMPC_OK;
}

int64_t cpmpc_get_warp_cycles(cpmpc_handle h, uint64_t* cycles_out, int64_t capacity) {
  if (h == nullptr) return 0;
  const int64_t n = static_cast<int64_t>(h->grid) * h->warps_per_block;
  if (cycles_out == nullptr || capacity < n) return n;
  if (cudaSetDevice(h->device) != cudaSuccess) return -1;
  if (cudaMemcpy(cycles_out, h->d_warp_cycles, sizeof(uint64_t) * WARP_REC * n, cudaMemcpyDeviceToHost) != cudaSuccess) {
    Fail(CPMPC_ERR_CUDA, "cudaMemcpy of the warp cycle counters failed");
    return -1;
  }
  return n;
}

int64_t cpmpc_debug_stride(cpmpc_handle h) {
  if (h == nullptr) return 0;
  const KernelCfg& k = h->cfg;
  return 16LL * k.nseg + 4LL * k.W + 4LL * k.nseg + 4LL * k.NS + k.W + 8;
}

int cpmpc_debug_first_qp(cpmpc_handle h, const double* states, const double* dyn_params, const double* set_points,
                         const double* weights, double lambda, int qp_method, double* blocks_out) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  if (states == nullptr || dyn_params == nullptr || set_points == nullptr || blocks_out == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "states, dyn_params, set_points and blocks_out are required");
  }
  if (!(lambda >= 0.0) || qp_method < 0 || qp_method > 2) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "lambda must be >= 0 and qp_method one of 0, 1, 2");
  }
  if (qp_method == 2 && !h->cfg.condensed) {
    return Fail(CPMPC_ERR_UNSUPPORTED, "this handle has no scratch for the condensed form (long horizon: W * dt > 0.6 s)");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  if (int rc = DrainAsync(h)) return rc;
  if (int rc = WaitForReset(h)) return rc;
  if (int rc = EnsureStaging(h)) return rc;
  const long long B = h->B;
  const long long stride = cpmpc_debug_stride(h);
  cudaStream_t s = h->stream;
  double* d_dbg = nullptr;
  CUDA_TRY(cudaMalloc(&d_dbg, sizeof(double) * B * stride));
  auto run = [&]() -> int {
    CUDA_TRY(cudaMemsetAsync(d_dbg, 0, sizeof(double) * B * stride, s));
    CUDA_TRY(cudaMemcpyAsync(h->d_states, states, sizeof(double) * B * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->d_dyn, dyn_params, sizeof(double) * B * 9, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->d_sp, set_points, sizeof(double) * B, cudaMemcpyHostToDevice, s));
    if (weights != nullptr) {
      CUDA_TRY(cudaMemcpyAsync(h->d_weights, weights, sizeof(double) * B * CPMPC_WEIGHTS_STRIDE,
                               cudaMemcpyHostToDevice, s));
    }
    KernelIO io;
    std::memset(&io, 0, sizeof(io));
    io.states = h->d_states;
    io.dyn = h->d_dyn;
    io.set_points = h->d_sp;
    io.weights = weights ? h->d_weights : nullptr;
    io.n_steps = 1;
    io.dbg = d_dbg;
    io.dbg_stride = stride;
    io.dbg_lambda = lambda;
    io.dbg_qp = qp_method;
    int rc = Launch(h, io, s);  // the kernel leaves the warm start alone when io.dbg is set
    if (rc == CPMPC_OK) {
      const cudaError_t e = cudaMemcpyAsync(blocks_out, d_dbg, sizeof(double) * B * stride, cudaMemcpyDeviceToHost, s);
      if (e != cudaSuccess) rc = Fail(CPMPC_ERR_CUDA, cudaGetErrorString(e));
    }
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess && rc == CPMPC_OK) rc = Fail(CPMPC_ERR_CUDA, cudaGetErrorString(e));
    return rc;
  };
  const int rc = run();
  cudaFree(d_dbg);
  return rc;
}

int cpmpc_debug_fastmath(int device, int64_t n, const double* x, double* out) {
  if (n < 0 || (n > 0 && (x == nullptr || out == nullptr))) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "bad arguments");
  if (n == 0) return CPMPC_OK;
  CUDA_TRY(cudaSetDevice(device));
  double *d_x = nullptr, *d_out = nullptr;
  struct Free {
    double*& a;
    double*& b;
    ~Free() {
      cudaFree(a);
      cudaFree(b);
    }
  } guard{d_x, d_out};
  CUDA_TRY(cudaMalloc(&d_x, sizeof(double) * n));
  CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * n * 6));
  CUDA_TRY(cudaMemcpy(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  cpmpc_fastmath_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(n, d_x, d_out);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(out, d_out, sizeof(double) * n * 6, cudaMemcpyDeviceToHost));
  return CPMPC_OK;
}

int cpmpc_measure_fp64_peak(int device, double* tflops_out) {
  if (tflops_out == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int threads = 256;
  const int blocks = prop.multiProcessorCount * 8;
  const int iters = 4096;
  double* d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best_ms = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    cpmpc_dfma_probe_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-7);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  g_launches.fetch_add(5);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  const double flops = 2.0 * 8.0 * 16.0 * static_cast<double>(iters) * threads * blocks;
  *tflops_out = flops / (best_ms * 1e-3) / 1e12;
  return CPMPC_OK;
}

}  // extern "C"
