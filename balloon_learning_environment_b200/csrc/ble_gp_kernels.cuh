// Shared pieces of the WindGP kernels (env/wind_gp.py:98-241): the 8 x 8 blocked-lower layout of the kernel matrix /
// Cholesky factor, the fp64 tensor-core wrapper and two fp64 helpers.  The kernels themselves are in
// ble_gp_posterior.cuh (current) and ble_feature_kernels.cuh (first generation, kept as the A/B reference).
//
// Layout in HBM and shared memory ("blocked lower"): 8 x 8 blocks (b, j), j <= b, at (b (b + 1) / 2 + j) * 64 doubles;
// inside a block the elements are in mma.m8n8k4 A-fragment order (blk_inner), so a fragment load is 32 consecutive
// doubles.
// NOTE: included from inside `namespace ble` of ble_engine.cu after ble_feature_kernels.cuh.
#pragma once

constexpr int kGpBlk = 8;
constexpr int kGpNumBlk = kGpWindow / kGpBlk;                                   // 15
constexpr int kGpBlockedLower = kGpBlk * kGpBlk * (kGpNumBlk * (kGpNumBlk + 1) / 2);   // 7,680 doubles
constexpr int kGpFactorDoubles = kGpBlockedLower;                              // per balloon in d.gp_chol

__device__ __forceinline__ int blk_offset(int b, int j) { return (((b * (b + 1)) >> 1) + j) * (kGpBlk * kGpBlk); }
// inside a block: "A-fragment order" of mma.m8n8k4 -- element (r, c) at (c / 4) * 32 + r * 4 + (c % 4), i.e. for
// each half of the columns the 32 values sit in lane order (lane = 4 r + c % 4): a fragment load is 32 consecutive
// doubles, free of bank conflicts in both half-warps.
__device__ __forceinline__ int blk_inner(int r, int c) { return ((c >> 2) << 5) + (r << 2) + (c & 3); }
__device__ __forceinline__ int blocked_index(int i, int j) {
  return blk_offset(i >> 3, j >> 3) + blk_inner(i & 7, j & 7);
}

__device__ __forceinline__ double gp_rsqrt(double t) {       // t in [1e-30, 1e30]: float seed + two Newton steps
  double y = double(rsqrtf(float(t)));
  y = y * (1.5 - 0.5 * t * y * y);
  return y * (1.5 - 0.5 * t * y * y);
}


// sigma^2 exp(-sqrt(d2)) for d2 in [0, ~1e3] to ~2 ulp, without the range checks of the library calls:
// sqrt by one float-seeded Newton step, exp by 2^n * p(r), |r| <= ln2 / 2, degree-10 Taylor/Horner (|r|^11 / 11! < 3e-13).
__device__ __forceinline__ double gp_kernel_from_d2(double d2) {
  d2 = fmax(d2, 1e-30);                              // keeps the float seed finite; sqrt(1e-30) ~ 0
  double yr = double(rsqrtf(float(d2)));             // ~2^-22 relative
  yr = yr * (1.5 - 0.5 * d2 * yr * yr);              // ~2^-43: |error of dist| < 4e-12, far below what the features need
  const double dist = d2 * yr;
  const double t = -dist * 1.4426950408889634;       // log2(e)
  const double n = rint(t);
  const double r = fma(n, -1.9082149292705877e-10, fma(n, -0.6931471803691238, -dist));   // -dist - n ln2 (hi + lo)
  double p = 2.755731922398589e-07;                  // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);
  p = fma(p, r, 2.48015873015873e-05);
  p = fma(p, r, 1.984126984126984e-04);
  p = fma(p, r, 1.388888888888889e-03);
  p = fma(p, r, 8.333333333333333e-03);
  p = fma(p, r, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int ni = int(n);                             // >= -1100 here: d2 <= ~5e5 keeps the result normal
  return kGpSigma2 * __longlong_as_double(__double_as_longlong(p) + (static_cast<long long>(ni) << 52));
}


__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

