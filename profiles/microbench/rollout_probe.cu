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

// The same linearisation loop software-pipelined by hand: iteration (step, stage) evaluates the dynamics for the NEXT
// stage (long dependent chains: sincos, tanh, sqrt, reciprocal) in the same basic block as the Jacobian chain-rule
// update of the CURRENT stage (many independent FMAs), so that ptxas can fill the stalls of the one with the other.
__global__ void __launch_bounds__(256, 1) lin_probe_swp(double* out, long long* cycles, int steps, int active) {
  const int lane = threadIdx.x & 31;
  if (lane >= active) return;
  const DynConst P = probe_const(threadIdx.x);
  double x[4] = {0.0, -1.5 + 1e-3 * threadIdx.x, 0.1, 0.2};
  double Ph[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) Ph[q] = (q % 5 == 0) ? 1.0 : 0.0;
  double bsum = 0.0;
  const double h = 0.01, hh = 0.005, h6 = h / 6.0;
  const long long t0 = clock64();
  // pipeline registers: derivatives of the stage about to be consumed
  double ax, at, fx[4], fth[4], gx, gth;
  double a2 = x[2], a3 = x[3];  // velocity arguments of that stage
  dyn_eval_jac(P, x[0], x[1], x[2], x[3], 3.0 * (-3.5), ax, at, fx, fth, gx, gth);
  double K[16], ku[4], A[16], Bv[4], s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    K[i] = 0.0;
    A[i] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ku[i] = 0.0;
    Bv[i] = 0.0;
  }
  double cprev = 0.0;
  const int total = 4 * steps;
#pragma unroll 1
  for (int it = 0; it < total; ++it) {
    const int st = it & 3;
    const int step = it >> 2;
    const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
    const double cst = (st == 2) ? h : hh;
    // (1) state sums with the current stage, and the argument of the NEXT evaluation
    s0 += a2 * wgt;
    s1 += a3 * wgt;
    s2 += ax * wgt;
    s3 += at * wgt;
    double n0, n1, n2, n3;
    if (st == 3) {
      n0 = x[0] + h6 * s0;
      n1 = x[1] + h6 * s1;
      n2 = x[2] + h6 * s2;
      n3 = x[3] + h6 * s3;
    } else {
      n0 = x[0] + a2 * cst;
      n1 = x[1] + a3 * cst;
      n2 = x[2] + ax * cst;
      n3 = x[3] + at * cst;
    }
    // (2) chain rule with the CURRENT stage's derivatives ...
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double k0 = K[j], k1 = K[4 + j], k2 = K[8 + j], k3 = K[12 + j];
      const double p0 = ((j == 0) ? 1.0 : 0.0) + k0 * cprev, p1 = ((j == 1) ? 1.0 : 0.0) + k1 * cprev;
      const double p2 = ((j == 2) ? 1.0 : 0.0) + k2 * cprev, p3 = ((j == 3) ? 1.0 : 0.0) + k3 * cprev;
      K[j] = p2;
      K[4 + j] = p3;
      K[8 + j] = fx[0] * p0 + fx[1] * p1 + fx[2] * p2 + fx[3] * p3;
      K[12 + j] = fth[0] * p0 + fth[1] * p1 + fth[2] * p2 + fth[3] * p3;
    }
    {
      const double k0 = ku[2] * cprev, k1 = ku[3] * cprev;
      const double k2 = (fx[0] * ku[0] + fx[1] * ku[1] + fx[2] * ku[2] + fx[3] * ku[3]) * cprev + gx;
      const double k3 = (fth[0] * ku[0] + fth[1] * ku[1] + fth[2] * ku[2] + fth[3] * ku[3]) * cprev + gth;
      ku[0] = k0;
      ku[1] = k1;
      ku[2] = k2;
      ku[3] = k3;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] += K[i] * wgt;
#pragma unroll
    for (int i = 0; i < 4; ++i) Bv[i] += ku[i] * wgt;
    // ... (3) in the same basic block as the NEXT evaluation
    const double u_next = 3.0 * ((((it + 1) >> 2) & 7) - 3.5);
    a2 = n2;
    a3 = n3;
    dyn_eval_jac(P, n0, n1, n2, n3, u_next, ax, at, fx, fth, gx, gth);
    cprev = cst;
    if (st == 3) {
      // step boundary: close the step (A = I + h/6 sum, Phi <- A Phi), restart the accumulators
      x[0] = n0;
      x[1] = n1;
      x[2] = n2;
      x[3] = n3;
      double T[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) A[i] = ((i % 5 == 0) ? 1.0 : 0.0) + h6 * A[i];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          T[4 * r + j] = A[4 * r] * Ph[j] + A[4 * r + 1] * Ph[4 + j] + A[4 * r + 2] * Ph[8 + j] + A[4 * r + 3] * Ph[12 + j];
#pragma unroll
      for (int q = 0; q < 16; ++q) Ph[q] = T[q] * 0.999;
      bsum += h6 * (Bv[0] + Bv[1] + Bv[2] + Bv[3]);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        K[i] = 0.0;
        A[i] = 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ku[i] = 0.0;
        Bv[i] = 0.0;
      }
      s0 = s1 = s2 = s3 = 0.0;
      cprev = 0.0;
      (void)step;
    }
  }
  const long long t1 = clock64();
  double acc = bsum + x[0] + x[1] + x[2] + x[3] + ax + at;
#pragma unroll
  for (int q = 0; q < 16; ++q) acc += Ph[q];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (lane == 0) cycles[(size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

// The solver's own linearisation loop (one lane per shooting segment, Phi and Gamma in shared memory with the solver's
// layout: 8 instances per warp, odd instance stride), outside the solver: what do the shared-memory read-modify-write
// of Gamma and the Phi round trip cost on top of lin_probe?  PHI_IN_SMEM selects where Phi is accumulated.
template <bool PHI_IN_SMEM>
__global__ void __launch_bounds__(256, 1) lin_probe_smem(double* out, long long* cycles, int reps, int active) {
  extern __shared__ double smem[];
  constexpr int W = 40, S = 10, L = 4, STRIDE = 613;
  const int lane = threadIdx.x & 31;
  if (lane >= active) return;
  const int g = (threadIdx.x >> 5) * 8 + lane / L;  // instance within the block
  const int s = lane % L;                           // this lane's segment
  double* sm = smem + (size_t)g * STRIDE;
  double* Gam = sm;            // [4][W]
  double* Phi = sm + 4 * W;    // [4 segments][16]
  double* uu = Phi + 64;       // [W]
  for (int q = s; q < W; q += L) uu[q] = 3.0 * ((q & 7) - 3.5);
  __syncwarp();
  const DynConst P = probe_const(threadIdx.x);
  double acc = 0.0;
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    double x[4] = {0.0, -1.5 + 1e-3 * threadIdx.x + 1e-4 * rep, 0.1, 0.2};
    double Phr[16];
    double* Ph = PHI_IN_SMEM ? (Phi + 16 * s) : Phr;
    for (int i = 0; i < S; ++i) {
      double A[16], Bv[4];
      const int col = s * S + i;
      rk4_step_jac(P, x, uu[col], 0.01, A, Bv);
      if (i == 0) {
#pragma unroll
        for (int q = 0; q < 16; ++q) Ph[q] = A[q];
      } else {
        double T[16];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            T[4 * r + j] = A[4 * r] * Ph[j] + A[4 * r + 1] * Ph[4 + j] + A[4 * r + 2] * Ph[8 + j] + A[4 * r + 3] * Ph[12 + j];
#pragma unroll
        for (int q = 0; q < 16; ++q) Ph[q] = T[q];
        for (int j = s * S; j < col; ++j) {
          const double g0 = Gam[j], g1 = Gam[W + j], g2 = Gam[2 * W + j], g3 = Gam[3 * W + j];
#pragma unroll
          for (int r = 0; r < 4; ++r)
            Gam[r * W + j] = A[4 * r] * g0 + A[4 * r + 1] * g1 + A[4 * r + 2] * g2 + A[4 * r + 3] * g3;
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) Gam[r * W + col] = Bv[r];
    }
    if (!PHI_IN_SMEM) {
#pragma unroll
      for (int q = 0; q < 16; ++q) Phi[16 * s + q] = Phr[q];
    }
    acc += x[0] + x[1] + x[2] + x[3];
  }
  const long long t1 = clock64();
  acc += Gam[lane] + Phi[lane];
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
  cudaFuncSetAttribute(lin_probe_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(lin_probe_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int kind = 0; kind < 6; ++kind) {          // 0 merit ILP1, 1 merit ILP2, 2 linearise, 3 linearise pipelined, 4/5 solver loop
    for (int active : {32, 8}) {
      for (int warps : {4, 8}) {                  // per block of one SM: 1, 2 warps per scheduler (launch bound 256)
        if (active != 32 && kind != 0) continue;
        std::vector<long long> h(sms * warps);
        for (int rep = 0; rep < 2; ++rep) {
          if (kind == 0) merit_probe<1><<<sms, 32 * warps>>>(out, cyc, steps, active);
          if (kind == 1) merit_probe<2><<<sms, 32 * warps>>>(out, cyc, steps, active);
          if (kind == 2) lin_probe<<<sms, 32 * warps>>>(out, cyc, steps, active);
          if (kind == 3) lin_probe_swp<<<sms, 32 * warps>>>(out, cyc, steps, active);
          const size_t smb = sizeof(double) * 613 * 8 * warps;
          if (kind == 4 && warps == 4) lin_probe_smem<true><<<sms, 32 * warps, smb>>>(out, cyc, steps / 10, active);
          if (kind == 5 && warps == 4) lin_probe_smem<false><<<sms, 32 * warps, smb>>>(out, cyc, steps / 10, active);
          if (kind >= 4 && warps != 4) continue;
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
                    kind == 0 ? "merit rollout" : kind == 1 ? "merit rollout, 2 interleaved per lane" :
                    kind == 2 ? "linearise (RK4 + Jacobian chain + Phi)" :
                    kind == 3 ? "linearise, next evaluation overlapped with the chain rule" :
                    kind == 4 ? "solver linearisation loop, Phi and Gamma in shared memory" :
                                "solver linearisation loop, Gamma in shared memory, Phi in registers",
                    active, warps / 4.0, per_eval_warp, per_eval_sched);
      }
    }
  }
  cudaFree(out);
  cudaFree(cyc);
  return 0;
}
