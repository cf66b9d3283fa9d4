// Host-side accuracy check of cart_pole_mpc_b200/csrc/fastmath.cuh (the seeds are float-precision on the host, the
// Newton/polynomial stages are the same code the device runs).  Prints max errors; exit code 1 if a bound is exceeded.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <random>

#include "../../cart_pole_mpc_b200/csrc/fastmath.cuh"

static double ulp_err(double got, double want) {
  if (got == want) return 0.0;
  const double u = std::nextafter(std::fabs(want), INFINITY) - std::fabs(want);
  return std::fabs(got - want) / u;
}

int main() {
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> ang(-40.0, 40.0), vel(-60.0, 60.0), logx(-30.0, 30.0);
  double max_sin = 0, max_cos = 0, max_tanh_abs = 0, max_rcp = 0, max_sqrt = 0, max_rsqrt = 0;
  for (int i = 0; i < 2000000; ++i) {
    const double a = ang(rng);
    double s, c;
    cpmpc::fm::sincos_bf(a, s, c);
    // compare in absolute terms relative to 1 ulp of the larger of the two (error near zeros of sin is absolute)
    max_sin = std::fmax(max_sin, std::fabs(s - std::sin(a)) / 1.1102230246251565e-16);
    max_cos = std::fmax(max_cos, std::fabs(c - std::cos(a)) / 1.1102230246251565e-16);
    const double v = vel(rng);
    max_tanh_abs = std::fmax(max_tanh_abs, std::fabs(cpmpc::fm::tanh_bf(v) - std::tanh(v)));
    const double tiny = v * 1e-6;
    max_tanh_abs = std::fmax(max_tanh_abs, std::fabs(cpmpc::fm::tanh_bf(tiny) - std::tanh(tiny)));
    const double x = std::exp(logx(rng));
    max_rcp = std::fmax(max_rcp, ulp_err(cpmpc::fm::rcp(x), 1.0 / x));
    max_rcp = std::fmax(max_rcp, ulp_err(cpmpc::fm::rcp(-x), -1.0 / x));
    double r, ir;
    cpmpc::fm::sqrt_rsqrt(x, r, ir);
    max_sqrt = std::fmax(max_sqrt, ulp_err(r, std::sqrt(x)));
    max_rsqrt = std::fmax(max_rsqrt, ulp_err(ir, 1.0 / std::sqrt(x)));
  }
  double s, c;
  cpmpc::fm::sincos_bf(INFINITY, s, c);
  const bool nan_ok = std::isnan(s) && std::isnan(c);
  cpmpc::fm::sincos_bf(0.0, s, c);
  const bool zero_ok = s == 0.0 && c == 1.0 && cpmpc::fm::tanh_bf(0.0) == 0.0 && cpmpc::fm::tanh_bf(1e300) == 1.0 &&
                       cpmpc::fm::tanh_bf(-1e300) == -1.0;
  std::printf("max |sin err| / 2^-53 = %.3f\nmax |cos err| / 2^-53 = %.3f\nmax tanh abs err = %.3e\n"
              "max rcp ulp = %.3f\nmax sqrt ulp = %.3f\nmax rsqrt ulp = %.3f\nnan_ok=%d zero_ok=%d\n",
              max_sin, max_cos, max_tanh_abs, max_rcp, max_sqrt, max_rsqrt, nan_ok, zero_ok);
  const bool ok = max_sin <= 2.0 && max_cos <= 2.0 && max_tanh_abs < 4e-16 && max_rcp <= 1.0 && max_sqrt <= 1.0 &&
                  max_rsqrt <= 2.0 && nan_ok && zero_ok;
  return ok ? 0 : 1;
}
