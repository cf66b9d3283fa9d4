// FP64 pipe micro-probes for sm_100a, used to pick the lanes-per-instance mapping of the solver kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu && ./fp64_probe
// Each probe runs `warps` warps on ONE SM sub-partition layout (blocks of 32*warps threads, one block per SM) and
// reports cycles per DFMA *per warp instruction*:
//   chains   independent dependent-DFMA chains per thread (ILP)
//   active   lanes of each warp that execute the loop (the others exit at once)
// Answers: dependent DFMA latency (chains=1, warps=1), issue cost of a warp-wide DFMA (chains=8), whether a
// half-populated warp costs half the pipe time (active=16 vs 32), and how warps per scheduler fill the pipe.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

template <int CHAINS>
__global__ void dfma_probe(double* out, long long* cycles, int iters, int active, double a, double b) {
  const int lane = threadIdx.x & 31;
  if (lane >= active) return;
  double v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x * 1e-3 + c;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) v[c] = fma(v[c], a, b);
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += v[c];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void shfl_probe(double* out, long long* cycles, int iters) {
  double v = threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1.0;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void lds_probe(double* out, long long* cycles, int iters) {
  __shared__ double sm[64];
  sm[threadIdx.x] = (double)((threadIdx.x + 1) & 31);
  sm[threadIdx.x + 32] = 0.0;
  __syncwarp();
  int idx = threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) idx = (int)sm[idx];
  }
  const long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void rcp_probe(double* out, long long* cycles, int iters) {
  double v = 1.5 + threadIdx.x * 1e-3;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      double y;
      asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
      v = y;
    }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int CHAINS>
static void run(int warps, int active, double* d_out, long long* d_cyc) {
  const int iters = 2000;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  for (int rep = 0; rep < 2; ++rep)
    dfma_probe<CHAINS><<<prop.multiProcessorCount, 32 * warps>>>(d_out, d_cyc, iters, active, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  const double per = (double)cyc / ((double)iters * 8 * CHAINS);
  // warps are spread over 4 sub-partitions: warps/4 (rounded up) share one FP64 pipe
  const int wps = (warps + 3) / 4;
  printf("{\"probe\":\"dfma\",\"chains\":%d,\"warps_per_block\":%d,\"warps_per_scheduler\":%d,\"active_lanes\":%d,"
         "\"cycles_per_dfma_per_warp\":%.3f,\"pipe_cycles_per_warp_dfma\":%.3f}\n",
         CHAINS, warps, wps, active, per, per / wps);
}

int main() {
  double* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, sizeof(double) * 148 * 1024 * 2);
  cudaMalloc(&d_cyc, sizeof(long long));
  const int warp_counts[] = {1, 4, 8, 16, 32};
  for (int w : warp_counts) {
    for (int active : {32, 16, 8}) {
      run<1>(w, active, d_out, d_cyc);
      run<2>(w, active, d_out, d_cyc);
      run<4>(w, active, d_out, d_cyc);
      run<8>(w, active, d_out, d_cyc);
    }
  }
  const int iters = 2000;
  long long cyc;
  shfl_probe<<<1, 32>>>(d_out, d_cyc, iters);
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  printf("{\"probe\":\"shfl_f64_plus_dadd\",\"cycles\":%.2f}\n", (double)cyc / (iters * 8));
  lds_probe<<<1, 32>>>(d_out, d_cyc, iters);
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  printf("{\"probe\":\"lds_f64_pointer_chase_plus_cvt\",\"cycles\":%.2f}\n", (double)cyc / (iters * 8));
  rcp_probe<<<1, 32>>>(d_out, d_cyc, iters);
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  printf("{\"probe\":\"mufu_rcp64h_dependent\",\"cycles\":%.2f}\n", (double)cyc / (iters * 8));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
