// Plant-generic part of the C ABI (cpmpc_plant_* in include/cpmpc.h): host logic for plant_solver_kernel.cuh.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cpmpc.h"
#include "capi_common.hpp"
#include "plant_solver_kernel.cuh"

using namespace cpmpc;
using cpmpc::detail::Fail;

struct cpmpc_plant_solver {
  int plant = CPMPC_PLANT_DOUBLE;
  int nx = 6, np = 6;
  cpmpc_plant_params params;
  long long B = 0;
  int device = 0;
  int lanes = 4;
  int grid = 0;
  int warps_per_block = 1;
  size_t smem_bytes = 0;
  gp::GpCfg cfg;
  cudaStream_t stream = nullptr;
  double* d_prev = nullptr;
  unsigned char* d_has_prev = nullptr;
  double* d_u_guess = nullptr;
  double* d_kk = nullptr;
  bool split = false;         // serial rollouts in their own one-thread-per-instance kernels (large batches)
  double* d_guess = nullptr;  // [B, NX NS]
  double* d_u_split = nullptr;
  unsigned long long* d_queue = nullptr;
  double* d_iter_log = nullptr;
  bool log_enabled = false;
  double* d_states = nullptr;
  double* d_dyn = nullptr;
  double* d_sp = nullptr;
  double* d_weights = nullptr;
  double* d_u = nullptr;
  double* d_pred = nullptr;
  cpmpc_stats* d_stats = nullptr;
};

namespace {

using KernelFn = void (*)(const gp::GpCfg, const gp::GpIO);

template <class PL>
KernelFn SelectFor(int lanes) {
  switch (lanes) {
    case 1: return gp::plant_solve_kernel<1, PL>;
    case 2: return gp::plant_solve_kernel<2, PL>;
    case 4: return gp::plant_solve_kernel<4, PL>;
    case 8: return gp::plant_solve_kernel<8, PL>;
    case 16: return gp::plant_solve_kernel<16, PL>;
    default: return gp::plant_solve_kernel<32, PL>;
  }
}

KernelFn SelectKernel(int plant, int lanes) {
  return plant == CPMPC_PLANT_SINGLE ? SelectFor<gp::SinglePlant>(lanes) : SelectFor<gp::DoublePlant>(lanes);
}

int PlantNx(int plant) {
  if (plant == CPMPC_PLANT_SINGLE) return gp::SinglePlant::NX;
  if (plant == CPMPC_PLANT_DOUBLE) return gp::DoublePlant::NX;
  return -1;
}
int PlantNp(int plant) {
  if (plant == CPMPC_PLANT_SINGLE) return gp::SinglePlant::NP;
  if (plant == CPMPC_PLANT_DOUBLE) return gp::DoublePlant::NP;
  return -1;
}
int PlantNc(int plant) { return plant == CPMPC_PLANT_SINGLE ? gp::SinglePlant::NC : gp::DoublePlant::NC; }

int Pow2Ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Shared-memory layout of one instance (doubles).
void BuildLayout(gp::GpCfg& k, int nx, int nc) {
  int off = 0;
  auto take = [&off](int n) {
    const int o = off;
    off += n;
    return o;
  };
  const int W = k.W, S = k.S, NS = k.NS, nseg = k.nseg;
  const int xi = nx + 1, nrm = nx + 1, ncol = xi + nrm;
  k.off_xs = take(nx * NS);
  k.off_u = take(W);
  k.off_dxs = take(nx * NS);
  k.off_du = take(W);
  k.off_phi = take(nx * nx * nseg);
  k.off_e = take(nx * nseg);
  // line-search candidates alias the Gamma block (dead once the QP has been solved)
  const int n_cand = nx * NS + W;
  k.off_gam = take(nx * W > n_cand ? nx * W : n_cand);
  k.off_xc = k.off_gam;
  k.off_uc = k.off_xc + nx * NS;
  // Riccati stage matrix (packed lower triangle) or, aliased, the condensed-QP scratch Z [nx][W], tl [W], tinv [W]
  const int D = S + ncol;
  const int n_ric = D * (D + 1) / 2;
  const int n_cond = k.condensed ? (nx + 2) * W : 0;
  k.off_aug = take(n_ric > n_cond ? n_ric : n_cond);
  k.off_vbuf = take(xi * xi + xi * nrm);
  k.off_invd = take(S);
  k.off_dyn = take(nc);
  k.stride = off | 1;
  k.kk_stride = nseg * ncol * S;
}

int Launch(cpmpc_plant_solver* h, gp::GpIO io, cudaStream_t stream) {
  KernelFn fn = SelectKernel(h->plant, h->lanes);
  io.B = h->B;
  io.u_guess = h->d_u_guess;
  io.prev = h->d_prev;
  io.has_prev = h->d_has_prev;
  io.kk_scratch = h->d_kk;
  io.queue = h->d_queue;
  io.iter_log = h->log_enabled ? h->d_iter_log : nullptr;
  if (!h->split) {
    fn<<<h->grid, 32 * h->warps_per_block, h->smem_bytes, stream>>>(h->cfg, io);
    cpmpc::detail::CountLaunches(1);
    CPMPC_CUDA_TRY(cudaGetLastError());
    return CPMPC_OK;
  }
  // Split launch (large batches): guess rollouts -> SQP -> predicted states, the rollouts one thread per instance
  if (io.u_out == nullptr) io.u_out = h->d_u_split;
  io.guess = h->d_guess;
  io.split = 1;
  const unsigned wide_blocks = static_cast<unsigned>((h->B + 127) / 128);
  const bool single = h->plant == CPMPC_PLANT_SINGLE;
  if (single) gp::plant_guess_kernel<gp::SinglePlant><<<wide_blocks, 128, 0, stream>>>(h->cfg, io, h->d_guess);
  else gp::plant_guess_kernel<gp::DoublePlant><<<wide_blocks, 128, 0, stream>>>(h->cfg, io, h->d_guess);
  fn<<<h->grid, 32 * h->warps_per_block, h->smem_bytes, stream>>>(h->cfg, io);
  if (io.pred_out != nullptr) {
    if (single) gp::plant_post_kernel<gp::SinglePlant><<<wide_blocks, 128, 0, stream>>>(h->cfg, io, io.u_out);
    else gp::plant_post_kernel<gp::DoublePlant><<<wide_blocks, 128, 0, stream>>>(h->cfg, io, io.u_out);
  }
  cpmpc::detail::CountLaunches(io.pred_out != nullptr ? 3 : 2);
  CPMPC_CUDA_TRY(cudaGetLastError());
  return CPMPC_OK;
}

int EnsureStaging(cpmpc_plant_solver* h) {
  if (h->d_states != nullptr) return CPMPC_OK;
  const long long B = h->B;
  const int W = h->cfg.W;
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_states, sizeof(double) * B * h->nx));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_dyn, sizeof(double) * B * h->np));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_sp, sizeof(double) * B));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_weights, sizeof(double) * B * (2 + h->nx)));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_u, sizeof(double) * B * W));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_pred, sizeof(double) * B * W * h->nx));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_stats, sizeof(cpmpc_stats) * B));
  return CPMPC_OK;
}

}  // namespace

extern "C" {

int cpmpc_plant_state_dim(int plant) { return PlantNx(plant); }
int cpmpc_plant_dyn_dim(int plant) { return PlantNp(plant); }

void cpmpc_plant_default_params(int plant, cpmpc_plant_params* p) {
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
  for (int i = 0; i < CPMPC_PLANT_MAX_NX; ++i) p->final_cost_weight[i] = -1.0;
  p->final_cost_weight[0] = 150.0;
  (void)plant;
}

int cpmpc_plant_create(int plant, const cpmpc_plant_params* params, int64_t batch, int device,
                       cpmpc_plant_handle* out) {
  if (params == nullptr || out == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  const int nx = PlantNx(plant);
  if (nx < 0) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "unknown plant " + std::to_string(plant));
  // the preconditions of Optimization::Optimization (optimization.cc:13-22) and CheckParams (nonlinear.cc:49-73)
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
  CPMPC_CUDA_TRY(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) return Fail(CPMPC_ERR_CUDA, "no such CUDA device: " + std::to_string(device));
  CPMPC_CUDA_TRY(cudaSetDevice(device));

  cpmpc_plant_solver* h = new cpmpc_plant_solver();
  h->plant = plant;
  h->nx = nx;
  h->np = PlantNp(plant);
  h->params = *params;
  h->B = batch;
  h->device = device;
  gp::GpCfg& k = h->cfg;
  std::memset(&k, 0, sizeof(k));
  k.W = static_cast<int>(params->window_length);
  k.S = static_cast<int>(params->state_spacing);
  k.nseg = k.W / k.S;
  k.NS = k.nseg + 1;
  k.N = nx * k.NS + k.W;
  k.max_iter = static_cast<int>(params->max_iterations);
  k.max_ls = 5;  // optimization.cc:76
  k.dt = params->control_dt;
  k.rel_tol = params->relative_exit_tol;
  k.first_tol = params->absolute_first_derivative_tol;
  k.abs_tol = 1.0e-12;       // nonlinear.hpp:75
  k.pen_init = params->equality_penalty_initial;
  k.pen_scale = 1.01;        // nonlinear.hpp:98
  k.rho = 0.1;               // nonlinear.hpp:102
  k.armijo_c1 = 1.0e-4;      // nonlinear.cc:121
  k.lam_fail_init = 1.0e-2;  // nonlinear.hpp:108
  k.lam_dec_success = 0.1;   // nonlinear.hpp:111
  k.lam_dec_restore = 0.8;   // nonlinear.hpp:114
  k.lam_max = 1.0;           // nonlinear.hpp:117
  k.lam_min = 0.0;           // nonlinear.hpp:120
  k.w_u = params->u_cost_weight;
  k.w_r = params->u_derivative_cost_weight;
  for (int i = 0; i < nx; ++i) k.wt[i] = params->final_cost_weight[i];
  k.x_clamp = 5.0;    // optimization.cc:320
  k.u_clamp = 300.0;  // optimization.cc:327

  cudaDeviceProp prop;
  CPMPC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  int lanes = Pow2Ceil(k.nseg);
  if (lanes > 32) lanes = 32;
  if (const char* env = std::getenv("CPMPC_PLANT_LANES")) {
    const int v = std::atoi(env);
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) lanes = v;
  }
  // QP method for the undamped iterations: condensed while the horizon is short enough for open-loop sensitivity
  // products to stay accurate, Riccati beyond (and always for Levenberg-Marquardt iterations); same rule and the same
  // CPMPC_QP override as the single-pendulum handle (cpmpc_capi.cu)
  k.condensed = (k.W * k.dt <= 0.6) ? 1 : 0;
  if (const char* env = std::getenv("CPMPC_QP")) {
    if (std::strcmp(env, "riccati") == 0) k.condensed = 0;
    if (std::strcmp(env, "condensed") == 0) k.condensed = 1;
  }
  BuildLayout(k, nx, PlantNc(plant));
  size_t smem_per_warp = 0;
  for (;;) {
    smem_per_warp = sizeof(double) * static_cast<size_t>(k.stride) * (32 / lanes);
    if (smem_per_warp <= static_cast<size_t>(prop.sharedMemPerBlockOptin) || lanes == 32) break;
    lanes *= 2;
  }
  if (smem_per_warp > static_cast<size_t>(prop.sharedMemPerBlockOptin)) {
    delete h;
    return Fail(CPMPC_ERR_UNSUPPORTED, "problem needs " + std::to_string(smem_per_warp) +
                                           " bytes of shared memory per warp, device offers " +
                                           std::to_string(prop.sharedMemPerBlockOptin));
  }
  h->lanes = lanes;
  const int groups_per_warp = 32 / lanes;
  int wpb_max = 8;
  while (wpb_max > 1 && smem_per_warp * wpb_max > static_cast<size_t>(prop.sharedMemPerBlockOptin)) --wpb_max;
  const long long warps_needed = (batch + groups_per_warp - 1) / groups_per_warp;
  int wpb = wpb_max;
  if (warps_needed <= static_cast<long long>(wpb_max) * prop.multiProcessorCount) {
    wpb = static_cast<int>((warps_needed + prop.multiProcessorCount - 1) / prop.multiProcessorCount);
    if (wpb < 1) wpb = 1;
  }
  h->warps_per_block = wpb;
  h->smem_bytes = smem_per_warp * wpb;
  KernelFn fn = SelectKernel(plant, lanes);
  // per kernel instantiation and device, never lowered by a later handle (see cpmpc_capi.cu)
  CPMPC_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(prop.sharedMemPerBlockOptin)));
  int blocks_per_sm = 0;
  CPMPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fn, 32 * wpb, h->smem_bytes));
  if (blocks_per_sm < 1) {
    delete h;
    return Fail(CPMPC_ERR_UNSUPPORTED, "kernel does not fit on an SM");
  }
  const long long groups_per_block = static_cast<long long>(groups_per_warp) * wpb;
  const long long blocks_needed = (batch + groups_per_block - 1) / groups_per_block;
  const long long resident = static_cast<long long>(blocks_per_sm) * prop.multiProcessorCount;
  h->grid = static_cast<int>(blocks_needed < resident ? blocks_needed : resident);
  const long long n_groups = static_cast<long long>(h->grid) * groups_per_block;

  CPMPC_CUDA_TRY(cudaMalloc(&h->d_prev, sizeof(double) * batch * k.N));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_has_prev, batch));
  CPMPC_CUDA_TRY(cudaMemset(h->d_has_prev, 0, batch));
  CPMPC_CUDA_TRY(cudaMemset(h->d_prev, 0, sizeof(double) * batch * k.N));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_kk, sizeof(double) * n_groups * k.kk_stride));
  // split launch beyond one wave of the persistent grid, as in cpmpc_create_ex; results do not depend on it
  h->split = batch > n_groups;
  if (const char* env = std::getenv("CPMPC_SPLIT")) h->split = std::atoi(env) != 0;
  if (h->split) {
    CPMPC_CUDA_TRY(cudaMalloc(&h->d_guess, sizeof(double) * batch * h->nx * k.NS));
    CPMPC_CUDA_TRY(cudaMalloc(&h->d_u_split, sizeof(double) * batch * k.W));
  }
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_queue, sizeof(unsigned long long) * 2));
  CPMPC_CUDA_TRY(cudaMemset(h->d_queue, 0, sizeof(unsigned long long) * 2));
  CPMPC_CUDA_TRY(cudaMalloc(&h->d_u_guess, sizeof(double) * k.W));
  std::vector<double> guess(k.W);  // optimization.cc:63-67, host libm
  for (int i = 0; i < k.W; ++i) {
    guess[i] = params->u_guess_sinusoid_amplitude *
               std::sin(static_cast<double>(i) / static_cast<double>(k.W) * 2 * M_PI);
  }
  CPMPC_CUDA_TRY(cudaMemcpy(h->d_u_guess, guess.data(), sizeof(double) * k.W, cudaMemcpyHostToDevice));
  CPMPC_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  *out = h;
  return CPMPC_OK;
}

int cpmpc_plant_destroy(cpmpc_plant_handle h) {
  if (h == nullptr) return CPMPC_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_prev);
  cudaFree(h->d_has_prev);
  cudaFree(h->d_u_guess);
  cudaFree(h->d_kk);
  cudaFree(h->d_guess);
  cudaFree(h->d_u_split);
  cudaFree(h->d_queue);
  cudaFree(h->d_iter_log);
  cudaFree(h->d_states);
  cudaFree(h->d_dyn);
  cudaFree(h->d_sp);
  cudaFree(h->d_weights);
  cudaFree(h->d_u);
  cudaFree(h->d_pred);
  cudaFree(h->d_stats);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CPMPC_OK;
}

int64_t cpmpc_plant_solution_dim(cpmpc_plant_handle h) { return h ? h->cfg.N : 0; }

int cpmpc_plant_step_device(cpmpc_plant_handle h, const double* states, const double* dyn_params,
                            const double* set_points, const double* weights, double* u_out, double* pred_out,
                            cpmpc_stats* stats_out, void* stream) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  if (states == nullptr || dyn_params == nullptr || set_points == nullptr || u_out == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "states, dyn_params, set_points and u_out are required");
  }
  CPMPC_CUDA_TRY(cudaSetDevice(h->device));
  gp::GpIO io;
  std::memset(&io, 0, sizeof(io));
  io.states = states;
  io.dyn = dyn_params;
  io.set_points = set_points;
  io.weights = weights;
  io.u_out = u_out;
  io.pred_out = pred_out;
  io.stats = stats_out;
  return Launch(h, io, static_cast<cudaStream_t>(stream));
}

int cpmpc_plant_step(cpmpc_plant_handle h, const double* states, const double* dyn_params, const double* set_points,
                     const double* weights, double* u_out, double* pred_out, cpmpc_stats* stats_out) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  if (states == nullptr || dyn_params == nullptr || set_points == nullptr || u_out == nullptr) {
    return Fail(CPMPC_ERR_INVALID_ARGUMENT, "states, dyn_params, set_points and u_out are required");
  }
  CPMPC_CUDA_TRY(cudaSetDevice(h->device));
  if (int rc = EnsureStaging(h)) return rc;
  const long long B = h->B;
  const int W = h->cfg.W, nx = h->nx;
  cudaStream_t s = h->stream;
  CPMPC_CUDA_TRY(cudaMemcpyAsync(h->d_states, states, sizeof(double) * B * nx, cudaMemcpyHostToDevice, s));
  CPMPC_CUDA_TRY(cudaMemcpyAsync(h->d_dyn, dyn_params, sizeof(double) * B * h->np, cudaMemcpyHostToDevice, s));
  CPMPC_CUDA_TRY(cudaMemcpyAsync(h->d_sp, set_points, sizeof(double) * B, cudaMemcpyHostToDevice, s));
  if (weights != nullptr) {
    CPMPC_CUDA_TRY(
        cudaMemcpyAsync(h->d_weights, weights, sizeof(double) * B * (2 + nx), cudaMemcpyHostToDevice, s));
  }
  if (int rc = cpmpc_plant_step_device(h, h->d_states, h->d_dyn, h->d_sp, weights ? h->d_weights : nullptr, h->d_u,
                                       pred_out ? h->d_pred : nullptr, stats_out ? h->d_stats : nullptr, s)) {
    return rc;
  }
  CPMPC_CUDA_TRY(cudaMemcpyAsync(u_out, h->d_u, sizeof(double) * B * W, cudaMemcpyDeviceToHost, s));
  if (pred_out != nullptr) {
    CPMPC_CUDA_TRY(cudaMemcpyAsync(pred_out, h->d_pred, sizeof(double) * B * W * nx, cudaMemcpyDeviceToHost, s));
  }
  if (stats_out != nullptr) {
    CPMPC_CUDA_TRY(cudaMemcpyAsync(stats_out, h->d_stats, sizeof(cpmpc_stats) * B, cudaMemcpyDeviceToHost, s));
  }
  CPMPC_CUDA_TRY(cudaStreamSynchronize(s));
  return CPMPC_OK;
}

int cpmpc_plant_reset(cpmpc_plant_handle h) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  CPMPC_CUDA_TRY(cudaSetDevice(h->device));
  CPMPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
  CPMPC_CUDA_TRY(cudaMemset(h->d_has_prev, 0, h->B));
  return CPMPC_OK;
}

int cpmpc_plant_reset_device(cpmpc_plant_handle h, void* stream) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  CPMPC_CUDA_TRY(cudaSetDevice(h->device));
  CPMPC_CUDA_TRY(cudaMemsetAsync(h->d_has_prev, 0, h->B, static_cast<cudaStream_t>(stream)));
  return CPMPC_OK;
}

int cpmpc_plant_set_previous_solution(cpmpc_plant_handle h, const double* solutions) {
  if (h == nullptr || solutions == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  CPMPC_CUDA_TRY(cudaSetDevice(h->device));
  CPMPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
  CPMPC_CUDA_TRY(cudaMemcpy(h->d_prev, solutions, sizeof(double) * h->B * h->cfg.N, cudaMemcpyHostToDevice));
  CPMPC_CUDA_TRY(cudaMemset(h->d_has_prev, 1, h->B));
  return CPMPC_OK;
}

int cpmpc_plant_get_previous_solution(cpmpc_plant_handle h, double* solutions_out, uint8_t* has_prev_out) {
  if (h == nullptr || solutions_out == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  CPMPC_CUDA_TRY(cudaSetDevice(h->device));
  CPMPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
  CPMPC_CUDA_TRY(cudaMemcpy(solutions_out, h->d_prev, sizeof(double) * h->B * h->cfg.N, cudaMemcpyDeviceToHost));
  if (has_prev_out != nullptr) {
    CPMPC_CUDA_TRY(cudaMemcpy(has_prev_out, h->d_has_prev, h->B, cudaMemcpyDeviceToHost));
  }
  return CPMPC_OK;
}

int cpmpc_plant_enable_iteration_log(cpmpc_plant_handle h, int enable) {
  if (h == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null handle");
  CPMPC_CUDA_TRY(cudaSetDevice(h->device));
  if (enable && h->d_iter_log == nullptr) {
    const size_t bytes = sizeof(double) * h->B * h->cfg.max_iter * CPMPC_ITERLOG_STRIDE;
    CPMPC_CUDA_TRY(cudaMalloc(&h->d_iter_log, bytes));
    CPMPC_CUDA_TRY(cudaMemset(h->d_iter_log, 0, bytes));
  }
  h->log_enabled = enable != 0;
  return CPMPC_OK;
}

int cpmpc_plant_get_iteration_log(cpmpc_plant_handle h, double* log_out) {
  if (h == nullptr || log_out == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "null argument");
  if (h->d_iter_log == nullptr) return Fail(CPMPC_ERR_INVALID_ARGUMENT, "iteration log was never enabled");
  CPMPC_CUDA_TRY(cudaSetDevice(h->device));
  CPMPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
  CPMPC_CUDA_TRY(cudaMemcpy(log_out, h->d_iter_log, sizeof(double) * h->B * h->cfg.max_iter * CPMPC_ITERLOG_STRIDE,
                            cudaMemcpyDeviceToHost));
  return CPMPC_OK;
}

}  // extern "C"
