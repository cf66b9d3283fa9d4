// Branch-free FP64 elementary functions for the cart-pole dynamics on sm_100a.
//
// Why not libdevice: sin/cos/tanh/sqrt/1-over-x from the CUDA math library each carry a rarely-taken slow-path branch
// (huge arguments, denormals).  Those branches split a dynamics evaluation into many basic blocks, and ptxas only
// interleaves instructions inside a block — the four long dependent DFMA chains of one evaluation (sincos, tanh,
// reciprocal of the mass-matrix determinant, sqrt of the tip speed) ended up strictly one after the other: 50 % of all
// issue slots were `stall_wait` in the ncu profile of round 1 (profiles/r01_*).  The versions below are straight-line
// code, accurate to about 1 ulp on the argument ranges this path can reach, so the scheduler can overlap the chains.
//
// Every function also compiles for the host (seed instructions replaced by float-precision seeds) so that
// tests/test_fastmath.py can measure the ulp error without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define CPMPC_HD __device__ __forceinline__
#define CPMPC_FMA(a, b, c) fma((a), (b), (c))
#else
#ifdef __CUDACC__
#define CPMPC_HD __host__ __device__ inline
#else
#define CPMPC_HD inline
#endif
#define CPMPC_FMA(a, b, c) fma((a), (b), (c))
#endif

namespace cpmpc {
namespace fm {

// Polynomial / range-reduction constants.  On the device they live in constant memory: sm_100a has no 64-bit
// immediates, so a literal costs two UMOVs (and a uniform-datapath latency) in front of every DFMA that uses it,
// whereas a constant-bank table is fetched into uniform registers two doubles per LDCU.128.
#define CPMPC_FM_CONSTS(X) \
  X(pio2_2, 0x1.5110b4611a626p-22) \
  X(ln2_lo, -0x1.05c610ca86c39p-29) \
  X(s5, -2.50507602534068634195e-08) \
  X(c5, 2.08757232129817482790e-09) \
  X(s4, 2.75573137070700676789e-06) \
  X(c4, -2.75573143513906633035e-07) \
  X(s3, -1.98412698298579493134e-04) \
  X(c3, 2.48015872894767294178e-05) \
  X(s2, 8.33333333332248946124e-03) \
  X(c2, -1.38888888888741095749e-03) \
  X(s1, -1.66666666666666324348e-01) \
  X(c1, 4.16666666666666019037e-02) \
  X(x8, 0x1.a01991399f987p-16) \
  X(x9, 0x1.71ddf52ded199p-19) \
  X(x6, 0x1.6c16c18826db1p-10) \
  X(x7, 0x1.a01a01b15de9dp-13) \
  X(x4, 0x1.555555554f000p-5) \
  X(x5, 0x1.111111110f1e5p-7) \
  X(x2, 0x1.0000000000011p-1) \
  X(x3, 0x1.555555555555ap-3)
// (the table is ordered in 16-byte pairs of constants that are consumed at the same time — the sine and cosine chains
// and the even and odd halves of the exp polynomial advance in step — so that each LDCU.128 is live for one step only)
// Five FMAs of the chains below take TWO constants (the start of each Horner chain, fma(z, c_top, c_next), and the
// 1.5 * 2^52 rounding trick), and an FP64 instruction takes at most ONE operand from the constant bank or a uniform
// register: the other had to sit in an ordinary register, which ptxas either hoisted out of the caller's stage loop
// (ten registers held across the whole rollout) or, under register pressure, re-loaded with LDC.64 in every dynamics
// evaluation right in front of its use, on the critical path (the linearisation loop; and the merit loop flipped
// between the two with unrelated changes elsewhere in the solver).  A double whose low 32 bits are zero is encodable
// as an IMMEDIATE operand instead.  The magic number is such a value; the top coefficients of the sine and cosine
// polynomials are rounded to 21 significant bits here: each multiplies the highest power of a reduced argument, so the change is below 3e-18
// relative to the result (0.03 ulp; tests/test_fastmath.py holds the functions to the same bounds as before).
// The same encoding serves wherever 21 bits are enough or the remainder can go into the next term:
//   * 2/pi and log2(e) only pick the integer k of the range reduction; a relative error of 3e-7 moves the reduced
//     argument at most 1e-4 beyond its nominal interval, where the polynomials are as accurate as inside;
//   * pi/2 = kPio2Hi + pio2_2 and ln 2 = kLn2Hi + ln2_lo are split so that the leading part is an immediate
//     (k * leading part is exact in the FMA either way; both sums carry 74 bits);
//   * the two top coefficients of the exp polynomial multiply r^10 and r^11 with |r| <= ln2 / 2 (2e-18 relative).
constexpr double kTwoOverPi = 0x1.45f3p-1;          // 2 / pi to 21 bits
constexpr double kPio2Hi = 0x1.921fbp+0;            // pi / 2 to 21 bits
constexpr double kLog2e = 0x1.71547p+0;             // log2(e) to 21 bits
constexpr double kLn2Hi = 0x1.62e43p-1;             // ln 2 to 21 bits
constexpr double kX10 = 0x1.28b47p-22;              // degree-10 and -11 coefficients of the exp polynomial (below), 21 bits
constexpr double kX11 = 0x1.af63ap-26;
constexpr double kMagic = 0x1.8p+52;                // 1.5 * 2^52
constexpr double kS6 = 0x1.5d93ap-33;               // 1.58969099521155010221e-10 (fdlibm S6) to 21 bits
constexpr double kC6 = -0x1.8fae9p-37;              // -1.13596475577881948265e-11 (fdlibm C6) to 21 bits

struct alignas(16) FmConsts {
#define CPMPC_X(name, value) double name;
  CPMPC_FM_CONSTS(CPMPC_X)
#undef CPMPC_X
};
#define CPMPC_X(name, value) value,
#if defined(__CUDACC__)
__constant__ FmConsts kc_dev = {CPMPC_FM_CONSTS(CPMPC_X)};
#endif
static const FmConsts kc_host = {CPMPC_FM_CONSTS(CPMPC_X)};
#undef CPMPC_X
#if defined(__CUDA_ARCH__)
#define CPMPC_K(name) (::cpmpc::fm::kc_dev.name)
#else
#define CPMPC_K(name) (::cpmpc::fm::kc_host.name)
#endif

CPMPC_HD double rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
#else
  return (double)(1.0f / (float)x);
#endif
}

CPMPC_HD double rsqrt_seed(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
#else
  return (double)(1.0f / sqrtf((float)x));
#endif
}

// 1/x for normal, finite x (|x| in [1e-290, 1e290]); NaN/Inf propagate as NaN or 0.  The seed instruction carries about
// 20 bits; ONE cubic step y0 (1 + e + e^2), e = 1 - x y0, takes the error to e^3 < 2^-60, below the rounding of the last
// FMA: within 1 ulp (tests/native/fastmath_check.cc), three dependent FMAs instead of the five of cubic-plus-Newton.
CPMPC_HD double rcp(double x) {
  const double y0 = rcp_seed(x);
  const double e = CPMPC_FMA(-x, y0, 1.0);
  const double e2 = CPMPC_FMA(e, e, e);  // e + e^2
  return CPMPC_FMA(y0, e2, y0);
}

// sqrt(x) and 1/sqrt(x) for finite x > 0.  (x == 0 gives NaN: callers select it away.)  One coupled Newton step takes
// g ~ sqrt(x), h ~ 0.5 / sqrt(x) from the seed's 2^-20 to 2^-40; the root is then corrected with the residual
// x - g^2 (error ~2^-80 before rounding), the reciprocal root takes a second step on h.  A caller that uses only the
// root (the dynamics without Jacobians) pays seven FMAs, the last three are dead code there.
CPMPC_HD void sqrt_rsqrt(double x, double& root, double& inv_root) {
  const double y0 = rsqrt_seed(x);
  double g = x * y0;     // ~ sqrt(x)
  double h = 0.5 * y0;   // ~ 0.5 / sqrt(x)
  double r = CPMPC_FMA(-g, h, 0.5);
  g = CPMPC_FMA(g, r, g);
  h = CPMPC_FMA(h, r, h);
  const double d = CPMPC_FMA(-g, g, x);
  root = CPMPC_FMA(d, h, g);
  r = CPMPC_FMA(-g, h, 0.5);
  h = CPMPC_FMA(h, r, h);
  inv_root = h + h;
}

// x with its sign bit XORed with `sbit` (0 or 0x80000000).
CPMPC_HD double flip_sign(double x, uint32_t sbit) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(x) ^ (int)sbit, __double2loint(x));
#else
  uint64_t b;
  memcpy(&b, &x, sizeof(b));
  b ^= (uint64_t)sbit << 32;
  memcpy(&x, &b, sizeof(b));
  return x;
#endif
}

// sin and cos for |x| < 2^20 (the pole angle stays within a few turns between two mod_pi wraps).
CPMPC_HD void sincos_bf(double x, double& s_out, double& c_out) {
  // k = round(x * 2/pi) with the 1.5 * 2^52 trick; its low bits give the quadrant
  const double magic = kMagic;
  const double t = CPMPC_FMA(x, kTwoOverPi, magic);
  const double kd = t - magic;
  int64_t tb;
#if defined(__CUDA_ARCH__)
  tb = __double_as_longlong(t);
#else
  memcpy(&tb, &t, sizeof(tb));
#endif
  const int q = (int)(tb & 3);
  // Cody-Waite with pi/2 = kPio2Hi (21 bits: k * kPio2Hi is exact) + pio2_2 (53 bits): 74 bits of pi/2, i.e. an
  // absolute error of k * 4e-23 in the reduced argument (3e-17 at the |x| = 2^20 end of the supported range)
  double r = CPMPC_FMA(-kd, kPio2Hi, x);
  r = CPMPC_FMA(-kd, CPMPC_K(pio2_2), r);
  const double z = r * r;
  // fdlibm kernel polynomials on |r| <= pi/4
  double ps = CPMPC_FMA(z, kS6, CPMPC_K(s5));
  ps = CPMPC_FMA(z, ps, CPMPC_K(s4));
  ps = CPMPC_FMA(z, ps, CPMPC_K(s3));
  ps = CPMPC_FMA(z, ps, CPMPC_K(s2));
  ps = CPMPC_FMA(z, ps, CPMPC_K(s1));
  const double sn = CPMPC_FMA(z * r, ps, r);
  double pc = CPMPC_FMA(z, kC6, CPMPC_K(c5));
  pc = CPMPC_FMA(z, pc, CPMPC_K(c4));
  pc = CPMPC_FMA(z, pc, CPMPC_K(c3));
  pc = CPMPC_FMA(z, pc, CPMPC_K(c2));
  pc = CPMPC_FMA(z, pc, CPMPC_K(c1));
  const double zz = z * z;
  const double cs = CPMPC_FMA(zz, pc, CPMPC_FMA(-0.5, z, 1.0));
  const bool swap = (q & 1) != 0;
  const double sv = swap ? cs : sn;
  const double cv = swap ? sn : cs;
  // quadrant signs by flipping the sign bit (one integer instruction on the high word each, instead of an FP64
  // negation and a two-register select): bit 1 of q for the sine, bit 1 of q + 1 for the cosine
  s_out = flip_sign(sv, (uint32_t)(q & 2) << 30);
  c_out = flip_sign(cv, (uint32_t)((q + 1) & 2) << 30);
}

// tanh(x) with ABSOLUTE error ~2e-16 (enough here: it multiplies the friction force inside a sum of O(1) forces).
// A NaN argument gives +-1, not NaN (fmin drops it).  The only caller passes the cart velocity, which the same RK4 stage
// also adds to the cart position, so a NaN still reaches the state, the cost and the non-finite rejection
// (nonlinear.cc:356-357) without an extra select in the hottest function of the kernel.
CPMPC_HD double tanh_bf(double x) {
  const double ax = fmin(fabs(x), 20.0);
  const double y = -2.0 * ax;  // em = exp(y) in (4e-18, 1]
  const double magic = kMagic;
  const double t = CPMPC_FMA(y, kLog2e, magic);
  const double nd = t - magic;
  int64_t tb;
#if defined(__CUDA_ARCH__)
  tb = __double_as_longlong(t);
#else
  memcpy(&tb, &t, sizeof(tb));
#endif
  const int n = (int)(int32_t)(tb & 0xffffffff);  // in [-58, 0]
  double r = CPMPC_FMA(-nd, kLn2Hi, y);
  r = CPMPC_FMA(-nd, CPMPC_K(ln2_lo), r);
  // exp(r), |r| <= ln2/2: degree-11 polynomial through the Chebyshev nodes of the interval (relative error 1.4e-17
  // with these coefficients; tests/prototypes/exp_poly.py), even/odd split to halve the dependent depth
  const double r2 = r * r;
  double pe = CPMPC_FMA(r2, kX10, CPMPC_K(x8));
  pe = CPMPC_FMA(r2, pe, CPMPC_K(x6));
  pe = CPMPC_FMA(r2, pe, CPMPC_K(x4));
  pe = CPMPC_FMA(r2, pe, CPMPC_K(x2));
  pe = CPMPC_FMA(r2, pe, 1.0);
  double po = CPMPC_FMA(r2, kX11, CPMPC_K(x9));
  po = CPMPC_FMA(r2, po, CPMPC_K(x7));
  po = CPMPC_FMA(r2, po, CPMPC_K(x5));
  po = CPMPC_FMA(r2, po, CPMPC_K(x3));
  po = CPMPC_FMA(r2, po, 1.0);
  const double er = CPMPC_FMA(r, po, pe);
  // scale by 2^n (n in [-58, 0]: always a normal number)
  const int64_t sb = (int64_t)(n + 1023) << 52;
  double scale;
#if defined(__CUDA_ARCH__)
  scale = __longlong_as_double(sb);
#else
  memcpy(&scale, &sb, sizeof(scale));
#endif
  const double em = er * scale;
  // tanh = (1 - em) / (1 + em)
  const double th = (1.0 - em) * rcp(1.0 + em);
  return copysign(th, x);
}

}  // namespace fm
}  // namespace cpmpc
