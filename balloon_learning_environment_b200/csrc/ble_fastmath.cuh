// Branch-free arithmetic primitives for the production step kernel (ble_step_fused.cuh).
//
// nvcc expands an fp64 division / square root into ~30 instructions plus a CALL to an IEEE slow
// path, and an fp32 division into an FCHK + CALL pair; in the first-generation step kernel those
// expansions were ~1,900 SASS instructions per Euler sub-step with 51 CALLs and 70 BSSY/BSYNC
// pairs.  Everything the sub-step divides by or takes a root of is a positive, normal, well-scaled
// physical quantity (pressures, temperatures, volumes, densities), so the special cases can go:
//   fp64: MUFU.RCP64H / MUFU.RSQ64H seed (>= 20 bits) + two Newton steps + one residual correction
//         -> relative error <= 2^-52 (a few ulp in the worst case instead of correctly rounded);
//   fp32: MUFU.RCP / RSQ / LG2 / EX2 approximations (<= 2 ulp; lg2 abs error 2^-22).
// The host versions (tests/hostemu replays these headers with g++) use the libm equivalents, so
// the CPU replay checks the ALGEBRA of the kernel; the last-ulp behaviour of the MUFU seeds is
// covered by the GPU parity tests.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define BLE_FM __host__ __device__ __forceinline__
#else
#define BLE_FM inline
#endif

namespace ble {
namespace fm {

// ---- fp64 ---------------------------------------------------------------------------------------
BLE_FM double rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

// a / b with one residual correction (error <= 1 ulp for normal operands)
BLE_FM double div(double a, double b) {
#if defined(__CUDA_ARCH__)
  const double r = rcp(b);
  double q = a * r;
  const double rem = fma(-b, q, a);
  return fma(rem, r, q);
#else
  return a / b;
#endif
}

BLE_FM double rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double t = x * y;
  double e = fma(-t, y, 1.0);
  y = fma(0.5 * y, e, y);
  t = x * y;
  e = fma(-t, y, 1.0);
  y = fma(0.5 * y, e, y);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}

// sqrt(x) for x >= 0 (returns 0 for x == 0)
BLE_FM double sqrt_pos(double x) {
#if defined(__CUDA_ARCH__)
  const double xs = fmax(x, 1e-300);
  const double y = rsqrt(xs);
  double s = xs * y;
  const double rem = fma(-s, s, xs);
  s = fma(rem, 0.5 * y, s);
  return x > 0.0 ? s : 0.0;
#else
  return sqrt(x);
#endif
}

// ---- fp32 ---------------------------------------------------------------------------------------
BLE_FM float rcpf(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
BLE_FM float divf(float a, float b) { return a * rcpf(b); }
BLE_FM float sqrtf_pos(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
BLE_FM float lg2f(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return log2f(x);
#endif
}
BLE_FM float ex2f(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return exp2f(x);
#endif
}
// x^y for x > 0
BLE_FM float powf_pos(float x, float y) { return ex2f(y * lg2f(x)); }
BLE_FM float expf_fast(float x) { return ex2f(x * 1.4426950408889634f); }
// cube root of x > 0.  lg2.approx has an ABSOLUTE error of 2^-22, i.e. 5e-8 relative on the result
// for the volumes this is used on (lg2 ~ 10.8); one Newton step on y^3 = x brings it to fp32 rounding.
BLE_FM float cbrtf_pos(float x) {
#if defined(__CUDA_ARCH__)
  float y = ex2f(lg2f(x) * (1.0f / 3.0f));
  const float y2 = y * y;
  y = y - (y2 * y - x) * rcpf(3.0f * y2);
  return y;
#else
  return cbrtf(x);
#endif
}

// sin and cos of |x| <~ 100 in one go: Cody-Waite reduction by pi/2 in two steps, the classic single-precision
// minimax kernels on [-pi/4, pi/4], quadrant fix-up.  Plain arithmetic (identical on host and device); absolute
// error < 2e-7.  libm's sinf / cosf carry a Payne-Hanek slow path that nvcc inlines at every call site.
BLE_FM void sincosf_fast(float x, float* s, float* c) {
  const float k = rintf(x * 0.63661977236758134308f);
  float r = fmaf(k, -1.57079601287841796875f, x);
  r = fmaf(k, -3.1391647326017846353e-07f, r);
  r = fmaf(k, -5.3903025299577647655e-15f, r);
  const float r2 = r * r;
  const float sp = r + r * r2 * (-1.6666654611e-1f + r2 * (8.3321608736e-3f + r2 * (-1.9515295891e-4f)));
  const float cp = 1.0f + r2 * (-0.5f + r2 * (4.166664568298827e-2f + r2 * (-1.388731625493765e-3f + r2 * 2.443315711809948e-5f)));
  const int q = int(k) & 3;
  const float ss = (q & 1) ? cp : sp, cc = (q & 1) ? sp : cp;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}

// x mod m for m > 0 (result in [0, m)); exact enough for fp64 angles of a few thousand degrees
BLE_FM double mod_pos(double x, double m, double inv_m) { return fma(-m, floor(x * inv_m), x); }

BLE_FM float rsqrtf_pos(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}

}  // namespace fm
}  // namespace ble
