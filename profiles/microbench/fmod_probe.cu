// Is an out-of-line fmod(a, 2 pi) the same function as the inlined one?  (profiles/microbench: one-off probe.)
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstring>
static __device__ __noinline__ double fmod_two_pi(double a) { return fmod(a, 6.283185307179586); }
__global__ void k(const double* a, double* o1, double* o2, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr double two_pi = 6.283185307179586;
  double x = a[i], y = a[i];
  if (!(fabs(x) < two_pi)) x = fmod(x, two_pi);
  if (!(fabs(y) < two_pi)) y = fmod_two_pi(y);
  o1[i] = x;
  o2[i] = y;
}
int main() {
  const int n = 1 << 20;
  double* h = new double[n];
  uint64_t s = 88172645463325252ull;
  for (int i = 0; i < n; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    double u = (double)(s >> 11) / 9007199254740992.0;
    int e = (int)(s % 240) - 20;
    h[i] = ((s & 1) ? -1.0 : 1.0) * ldexp(1.0 + u, e);
  }
  h[0] = INFINITY; h[1] = -INFINITY; h[2] = NAN; h[3] = 6.283185307179586; h[4] = -6.283185307179586; h[5] = 1e300;
  double *d, *o1, *o2;
  cudaMalloc(&d, n * 8); cudaMalloc(&o1, n * 8); cudaMalloc(&o2, n * 8);
  cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(d, o1, o2, n);
  double* r1 = new double[n]; double* r2 = new double[n];
  cudaMemcpy(r1, o1, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(r2, o2, n * 8, cudaMemcpyDeviceToHost);
  int bad = 0, badh = 0;
  for (int i = 0; i < n; ++i) {
    if (memcmp(&r1[i], &r2[i], 8) != 0 && !(r1[i] != r1[i] && r2[i] != r2[i])) { if (bad < 5) printf("diff a=%.17g inline=%.17g call=%.17g\n", h[i], r1[i], r2[i]); ++bad; }
    double hh = (fabs(h[i]) < 6.283185307179586) ? h[i] : fmod(h[i], 6.283185307179586);
    if (memcmp(&r1[i], &hh, 8) != 0 && !(r1[i] != r1[i] && hh != hh)) { if (badh < 5) printf("host diff a=%.17g dev=%.17g host=%.17g\n", h[i], r1[i], hh); ++badh; }
  }
  printf("n=%d inline-vs-call mismatches=%d device-vs-host mismatches=%d\n", n, bad, badh);
  return 0;
}
