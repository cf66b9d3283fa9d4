// Rollout micro-probes for sm_100a: the two hot loops of cpmpc_solve_kernel in isolation (same dynamics.cuh / fastmath.cuh
// code), to answer what the solver profile cannot separate:
//   * how many cycles one dynamics evaluation costs a warp when nothing else is in the way (1 warp per scheduler),
//   * what a second / fourth resident warp per scheduler is worth for each loop (the FP64 pipe takes a warp-wide DFMA
//     every 2 cycles and returns it after 8: one warp alone cannot fill it),
//   * what two INDEPENDENT rollouts interleaved in one lane are worth (instruction-level instead of warp-level
//     parallelism: the same loop body, twice the independent dependent chains),
//   * whether the number of active lanes changes the cost of the loop.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cart_pole_mpc_b200/csrc -o rollout_probe rollout_probe.cu
//   ./rollout_probe            # one JSON line per configuration
#include <cuda_runtime.h>

#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "dynamics.cuh"

using namespace cpmpc;

__device__ __forceinline__ DynConst probe_const(int t) {
  const double p[9] = {1.0 + 1e-3 * (t & 7), 0.1, 0.25, 9.81, 0.05, 0.1, 0.02, 0.8, 100.0};
  return make_dyn_const(p);
}

// `steps` RK4 steps without Jacobians per thread (the merit rollout / initial guess / predicted states loop)
template <int ILP>
__global__ void __launch_bounds__(256, 1) merit_probe(double* out, long long* cycles, int steps, int active) {
  const int lane = threadIdx.x & 31;
  if (lane >= active) return;
  const DynConst P = probe_const(threadIdx.x);
  double x[ILP][4];
#pragma unroll
  for (int q = 0; q < ILP; ++q) {
    x[q][0] = 0.01 * q;
    x[q][1] = -1.5 + 1e-3 * threadIdx.x + 0.1 * q;
    x[q][2] = 0.1;
    x[q][3] = 0.2;
  }
  const long long t0 = clock64();
  for (int i = 0; i < steps; ++i) {
    const double u = 3.0 * ((i & 7) - 3.5);
    if (ILP == 1) {
      rk4_step<false>(P, x[0], u, 0.01);
    } else {
      // two rollouts through ONE rolled stage loop: the body holds both evaluations, so their chains interleave
      const double h = 0.01, hh = 0.005;
      double a[ILP][4], s[ILP][4];
#pragma unroll
      for (int q = 0; q < ILP; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          a[q][r] = x[q][r];
          s[q][r] = 0.0;
        }
#pragma unroll 1
      for (int st = 0; st < 4; ++st) {
        const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
        const double cst = (st == 2) ? h : hh;
#pragma unroll
        for (int q = 0; q < ILP; ++q) {
          double ax, at;
          dyn_eval<false>(P, a[q][0], a[q][1], a[q][2], a[q][3], u + q, 0.0, 0.0, 0.0, ax, at);
          s[q][0] += a[q][2] * wgt;
          s[q][1] += a[q][3] * wgt;
          s[q][2] += ax * wgt;
          s[q][3] += at * wgt;
          const double n0 = x[q][0] + a[q][2] * cst, n1 = x[q][1] + a[q][3] * cst, n2 = x[q][2] + ax * cst,
                       n3 = x[q][3] + at * cst;
          a[q][0] = n0;
          a[q][1] = n1;
          a[q][2] = n2;
          a[q][3] = n3;
        }
      }
      const double h6 = h / 6.0;
#pragma unroll
      for (int q = 0; q < ILP; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) x[q][r] += h6 * s[q][r];
    }
  }
  const long long t1 = clock64();
  double acc = 0.0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) acc += x[q][0] + x[q][1] + x[q][2] + x[q][3];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (lane == 0) cycles[(size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

// `steps` RK4 steps WITH the Jacobian chain and the forward-mode Phi product (the linearisation loop; Gamma, which lives
// in shared memory in the solver, is left out: this is the register part)
__global__ void __launch_bounds__(256, 1) lin_probe(double* out, long long* cycles, int steps, int active) {
  const int lane = threadIdx.x & 31;
  if (lane >= active) return;
  const DynConst P = probe_const(threadIdx.x);
  double x[4] = {0.0, -1.5 + 1e-3 * threadIdx.x, 0.1, 0.2};
  double Ph[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) Ph[q] = (q % 5 == 0) ? 1.0 : 0.0;
  double bsum = 0.0;
  const long long t0 = clock64();
  for (int i = 0; i < steps; ++i) {
    double A[16], Bv[4];
    rk4_step_jac(P, x, 3.0 * ((i & 7) - 3.5), 0.01, A, Bv);
    double T[16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        T[4 * r + j] = A[4 * r] * Ph[j] + A[4 * r + 1] * Ph[4 + j] + A[4 * r + 2] * Ph[8 + j] + A[4 * r + 3] * Ph[12 + j];
#pragma unroll
    for (int q = 0; q < 16; ++q) Ph[q] = T[q] * 0.999;
    bsum += Bv[0] + Bv[1] + Bv[2] + Bv[3];
  }
  const long long t1 = clock64();
  double acc = bsum + x[0] + x[1] + x[2] + x[3];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc += Ph[q];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (lane == 0) cycles[(size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

static double median_cycles(const std::vector<long long>& v) {
  std::vector<long long> s(v);
  std::sort(s.begin(), s.end());
  return (double)s[s.size() / 2];
}


int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms * 32);
  const int steps = 400;
  for (int kind = 0; kind < 3; ++kind) {          // 0 merit ILP1, 1 merit ILP2, 2 linearise
    for (int active : {32, 8}) {
      for (int warps : {4, 8, 16, 32}) {          // per block of one SM: 1, 2, 4, 8 warps per scheduler
        if (kind == 2 && warps > 8) continue;     // 255 registers: 8 warps per SM
        if (active != 32 && warps > 8) continue;
        std::vector<long long> h(sms * warps);
        for (int rep = 0; rep < 2; ++rep) {
          if (kind == 0) merit_probe<1><<<sms, 32 * warps>>>(out, cyc, steps, active);
          if (kind == 1) merit_probe<2><<<sms, 32 * warps>>>(out, cyc, steps, active);
          if (kind == 2) lin_probe<<<sms, 32 * warps>>>(out, cyc, steps, active);
          if (cudaDeviceSynchronize() != cudaSuccess) {
            std::printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
            return 1;
          }
        }
        cudaMemcpy(h.data(), cyc, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        const double c = median_cycles(h);
        const int rollouts = (kind == 1) ? 2 : 1;
        const double evals_per_warp = 4.0 * steps * rollouts;
        const double per_eval_warp = c / evals_per_warp;                          // latency view: one warp's progress
        const double per_eval_sched = c / (evals_per_warp * (warps / 4.0));       // throughput view: per scheduler
        std::printf("{\"probe\": \"%s\", \"active_lanes\": %d, \"warps_per_scheduler\": %.1f, "
                    "\"cycles_per_eval_per_warp\": %.1f, \"scheduler_cycles_per_warp_eval\": %.1f}\n",
                    kind == 0 ? "merit rollout" : kind == 1 ? "merit rollout, 2 interleaved per lane" : "linearise (RK4 + Jacobian chain + Phi)",
                    active, warps / 4.0, per_eval_warp, per_eval_sched);
      }
    }
  }
  cudaFree(out);
  cudaFree(cyc);
  return 0;
}
