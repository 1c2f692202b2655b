// Wind field lookup: 4-D multilinear gather + OpenSimplex noise.
//
// Follows env/grid_based_wind_field.py:70-187 (+ generative/vae.py:26-93 grid geometry),
// env/wind_field.py:125-218 and env/simplex_wind_noise.py:50-211.  The noise function itself
// lives in the un-vendored third-party `opensimplex==0.3`; see oracle/opensimplex4.py for the
// restated definition this file implements (PARITY UNPINNED against the package).
#pragma once
#include "ble_physics.cuh"

#if !defined(__CUDACC__)
struct alignas(16) float4 { float x, y, z, w; };   // host replay only (tests/hostemu)
#endif

namespace ble {

// ---- field geometry (generative/vae.py:26-51) ----------------------------------------------------
constexpr int kNX = 21, kNY = 21, kNP = 10, kNT = 9;
constexpr int kFieldFloats = kNX * kNY * kNP * kNT * 2;        // 79,380 (native layout)
// "cell" layout: for every (x, y) column and every (pressure, time) cell the 2x2x2 corner
// block {p, p+1} x {t, t+1} x {u, v} is stored contiguously = 8 floats = one 32-byte sector.
// A lookup then touches exactly 4 aligned sectors (one per (x, y) corner).
constexpr int kPC = kNP - 1, kTC = kNT - 1;                    // 9 x 8 cells per column
constexpr int kCellFloats = 8;
constexpr int64_t kColumnFloats = int64_t(kPC) * kTC * kCellFloats;           // 576
constexpr int64_t kCellFieldFloats = int64_t(kNX) * kNY * kColumnFloats;      // 254,016 floats = 1,016,064 B

BLE_HD int64_t native_index(int ix, int iy, int ip, int it, int c) {
  return ((((int64_t(ix) * kNY + iy) * kNP + ip) * kNT + it) * 2 + c);
}
BLE_HD int64_t cell_index(int ix, int iy, int pc, int tc) {
  return ((int64_t(ix) * kNY + iy) * kPC + pc) * (kTC * kCellFloats) + int64_t(tc) * kCellFloats;
}

// Query point as the reference builds it (_prepare_get_forecast_inputs, :145-187): clip x, y
// to +-500 km and pressure to [5000, 14000] Pa, boomerang the time beyond 48 h, then ROUND
// EVERYTHING TO FP32 (the reference packs the point into a float32 array, :181).
struct FieldPoint { float x_km, y_km, p, t_h; };

BLE_HD double boomerang_hours(double hours) {
  if (hours < 48.0) return hours;                                          // :171-172
  const int64_t cycle = int64_t(hours / 48.0) % 2;                         // :136
  const double rem = fmod(hours, 48.0);                                    // :137
  return (cycle == 0) ? rem : 48.0 - rem;
}

BLE_HD FieldPoint make_field_point(double x_km, double y_km, double p, double hours) {
  FieldPoint q;
  q.x_km = float(fmin(fmax(x_km, -500.0), 500.0));
  q.y_km = float(fmin(fmax(y_km, -500.0), 500.0));
  q.p = float(fmin(fmax(p, 5000.0), 14000.0));
  q.t_h = float(boomerang_hours(hours));
  return q;
}

template <typename Real>
BLE_HD void axis_cell(float v, float g0, float step, int ncell, int* idx, Real* w) {
  // scipy interpn 'linear': i = clamp(searchsorted(grid, v) - 1, 0, n - 2), w = (v - g[i]) / step
  int i = int(floorf((v - g0) / step));
  i = i < 0 ? 0 : (i > ncell - 1 ? ncell - 1 : i);
  const Real gi = Real(g0) + Real(step) * Real(i);
  if (Real(v) < gi && i > 0) { --i; }                   // guard the float floor against 1-ulp slips
  const Real g = Real(g0) + Real(step) * Real(i);
  *idx = i;
  *w = (Real(v) - g) / Real(step);
}

struct float8 { float4 a, b; };

// 16-corner multilinear interpolation from the cell layout.  `ldcell` loads one 32-byte cell.
template <typename Real, typename LoadCell>
BLE_HD void interp_cells(const FieldPoint& q, LoadCell ldcell, Real* u, Real* v) {
  int ix, iy, pc, tc;
  Real wx, wy, wp, wt;
  axis_cell<Real>(q.x_km, -500.f, 50.f, kNX - 1, &ix, &wx);
  axis_cell<Real>(q.y_km, -500.f, 50.f, kNY - 1, &iy, &wy);
  axis_cell<Real>(q.p, 5000.f, 1000.f, kPC, &pc, &wp);
  axis_cell<Real>(q.t_h, 0.f, 6.f, kTC, &tc, &wt);
  const float8 c00 = ldcell(cell_index(ix, iy, pc, tc));
  const float8 c01 = ldcell(cell_index(ix, iy + 1, pc, tc));
  const float8 c10 = ldcell(cell_index(ix + 1, iy, pc, tc));
  const float8 c11 = ldcell(cell_index(ix + 1, iy + 1, pc, tc));
  const Real one = Real(1);
  // within a cell: a = {p0t0.u, p0t0.v, p0t1.u, p0t1.v}, b = {p1t0.u, p1t0.v, p1t1.u, p1t1.v}
  const Real w00 = (one - wp) * (one - wt), w01 = (one - wp) * wt, w10 = wp * (one - wt), w11 = wp * wt;
  auto cell_u = [&](const float8& c) {
    return w00 * Real(c.a.x) + w01 * Real(c.a.z) + w10 * Real(c.b.x) + w11 * Real(c.b.z);
  };
  auto cell_v = [&](const float8& c) {
    return w00 * Real(c.a.y) + w01 * Real(c.a.w) + w10 * Real(c.b.y) + w11 * Real(c.b.w);
  };
  const Real a00 = (one - wx) * (one - wy), a01 = (one - wx) * wy, a10 = wx * (one - wy), a11 = wx * wy;
  *u = a00 * cell_u(c00) + a01 * cell_u(c01) + a10 * cell_u(c10) + a11 * cell_u(c11);
  *v = a00 * cell_v(c00) + a01 * cell_v(c01) + a10 * cell_v(c10) + a11 * cell_v(c11);
}

// SimpleStaticWindField (env/wind_field.py:149-184): four 10 m/s sheets by pressure.
template <typename Real>
BLE_HD void static_wind(Real p, Real* u, Real* v) {
  if (p < Real(8000)) { *u = Real(10); *v = Real(0); }
  else if (p < Real(10000)) { *u = Real(0); *v = Real(10); }
  else if (p < Real(12000)) { *u = Real(-10); *v = Real(0); }
  else { *u = Real(0); *v = Real(-10); }
}

// ---- simplex noise -------------------------------------------------------------------------------------
constexpr double kStretch4 = -0.138196601125011;
constexpr double kSquish4 = 0.309016994374947;
constexpr double kNoiseMagnitude = 4.233932721683222;   // sqrt(1.02 / 0.0569), simplex_wind_noise.py:76

// weight, x, y, pressure, time spacings (simplex_wind_noise.py:50-64); index = component * 5 + harmonic
BLE_HD void harmonic_params(int h10, double* w, double* sx, double* sy, double* sp, double* st) {
  const double t[10][5] = {
      {0.1445, 702.269, 2116.987, 2587.802, 245.0},   {0.2766, 1483.570, 752.124, 646.208, 16.39},
      {0.2627, 276.810, 147.040, 587.702, 3.836},     {0.2137, 10214.525, 1512.216, 965.629, 41.780},
      {0.1025, 181.286, 420.942, 8500.0, 245.0},      {0.2716, 1974.228, 2028.814, 713.697, 26.435},
      {0.2684, 699.738, 541.845, 632.116, 9.530},     {0.2348, 217.750, 196.522, 686.825, 3.546},
      {0.1186, 47.500, 43.048, 66.553, 8.424},        {0.1066, 3663.291, 232.023, 7499.741, 225.0}};
  *w = t[h10][0]; *sx = t[h10][1]; *sy = t[h10][2]; *sp = t[h10][3]; *st = t[h10][4];
}

// Final blend of the 5 harmonics of one component (NoisyWindComponent.get_noise, :180-211):
// out = (sum w_h n_h / sum w) * sqrt(sum w / sum w^2)
BLE_HD void component_blend_constants(int comp, double* weights, double* scale) {
  double sw = 0, sw2 = 0;
  for (int h = 0; h < 5; ++h) {
    double w, a, b, c, d;
    harmonic_params(comp * 5 + h, &w, &a, &b, &c, &d);
    weights[h] = w; sw += w; sw2 += w * w;
  }
  *scale = sqrt(sw / sw2) / sw;
}

// 64 gradients = sign pattern (bit c of r negates component c) x position j of the "3".
template <typename Real>
BLE_HD Real gradient_dot(int hash, Real dx, Real dy, Real dz, Real dw) {
  const int g = hash >> 2;                // 0..63
  const int r = g >> 2, j = g & 3;
  Real ax = (j == 0) ? Real(3) * dx : dx;
  Real ay = (j == 1) ? Real(3) * dy : dy;
  Real az = (j == 2) ? Real(3) * dz : dz;
  Real aw = (j == 3) ? Real(3) * dw : dw;
  if (r & 1) ax = -ax;
  if (r & 2) ay = -ay;
  if (r & 4) az = -az;
  if (r & 8) aw = -aw;
  return (ax + ay) + (az + aw);
}

// noise4d(x, y, z, w) summed over every lattice vertex whose kernel (2 - |d|^2)^4 is non-zero.
// Candidate c in [0, 80): m = c / 5 is the cube corner (bit k of m = offset along axis k),
// e = c % 5: 0 = the corner itself, 1..4 = step one further out along axis e-1
// (offset 1 -> 2, offset 0 -> -1).  `perm` is this generator's 256-entry table.
template <typename Real, typename Perm>
BLE_HD Real simplex_noise4(const Perm& perm, double x, double y, double z, double w) {
  const double s = (x + y + z + w) * kStretch4;
  const double fx = floor(x + s), fy = floor(y + s), fz = floor(z + s), fw = floor(w + s);
  const double q = (fx + fy + fz + fw) * kSquish4;
  const Real dx0 = Real(x - (fx + q)), dy0 = Real(y - (fy + q));
  const Real dz0 = Real(z - (fz + q)), dw0 = Real(w - (fw + q));
  const int xb = int(int64_t(fx) & 255), yb = int(int64_t(fy) & 255);
  const int zb = int(int64_t(fz) & 255), wb = int(int64_t(fw) & 255);

  // pass 1: in-range mask over the 80 candidates
  uint64_t mask_lo = 0;   // candidates 0..63
  uint32_t mask_hi = 0;   // candidates 64..79
  const Real sq = Real(kSquish4);
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int i = m & 1, j = (m >> 1) & 1, k = (m >> 2) & 1, l = (m >> 3) & 1;
    const int pc = i + j + k + l;
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      int oi = i, oj = j, ok = k, ol = l, tot = pc;
      if (e == 1) { oi = i ? 2 : -1; tot += i ? 1 : -1; }
      if (e == 2) { oj = j ? 2 : -1; tot += j ? 1 : -1; }
      if (e == 3) { ok = k ? 2 : -1; tot += k ? 1 : -1; }
      if (e == 4) { ol = l ? 2 : -1; tot += l ? 1 : -1; }
      const Real t = Real(tot) * sq;
      const Real dx = dx0 - Real(oi) - t, dy = dy0 - Real(oj) - t;
      const Real dz = dz0 - Real(ok) - t, dw = dw0 - Real(ol) - t;
      const Real attn = Real(2) - dx * dx - dy * dy - dz * dz - dw * dw;
      const int c = m * 5 + e;
      if (attn > Real(0)) {
        if (c < 64) mask_lo |= (uint64_t(1) << c); else mask_hi |= (1u << (c - 64));
      }
    }
  }
  // pass 2: evaluate the (5..12) in-range vertices
  Real value = Real(0);
  while (mask_lo | mask_hi) {
    int c;
    if (mask_lo) {
      uint64_t low = mask_lo & (~mask_lo + 1);
      mask_lo ^= low;
#if defined(__CUDA_ARCH__)
      c = __ffsll((long long)low) - 1;
#else
      c = __builtin_ctzll(low);
#endif
    } else {
      uint32_t low = mask_hi & (~mask_hi + 1);
      mask_hi ^= low;
#if defined(__CUDA_ARCH__)
      c = 64 + __ffs(int(low)) - 1;
#else
      c = 64 + __builtin_ctz(low);
#endif
    }
    const int m = c / 5, e = c - m * 5;
    int o[4] = {m & 1, (m >> 1) & 1, (m >> 2) & 1, (m >> 3) & 1};
    int oi = o[0], oj = o[1], ok = o[2], ol = o[3];
    if (e == 1) oi = oi ? 2 : -1;
    if (e == 2) oj = oj ? 2 : -1;
    if (e == 3) ok = ok ? 2 : -1;
    if (e == 4) ol = ol ? 2 : -1;
    const Real t = Real(oi + oj + ok + ol) * sq;
    const Real dx = dx0 - Real(oi) - t, dy = dy0 - Real(oj) - t;
    const Real dz = dz0 - Real(ok) - t, dw = dw0 - Real(ol) - t;
    Real attn = Real(2) - dx * dx - dy * dy - dz * dz - dw * dw;
    int h = perm[(xb + oi) & 255];
    h = perm[(h + yb + oj) & 255];
    h = perm[(h + zb + ok) & 255];
    h = perm[(h + wb + ol) & 255];
    attn *= attn;
    value += attn * attn * gradient_dot<Real>(h, dx, dy, dz, dw);
  }
  return value / Real(30.0);
}

// OpenSimplex.__init__: 256-entry permutation from a 64-bit LCG (see oracle/opensimplex4.py).
BLE_HD void simplex_make_perm(int64_t seed, uint8_t* perm /*256, may be global*/, uint8_t* source /*256 scratch*/) {
  const uint64_t A = 6364136223846793005ull, Cc = 1442695040888963407ull;
  uint64_t s = uint64_t(seed);
  for (int i = 0; i < 256; ++i) source[i] = uint8_t(i);
  s = s * A + Cc; s = s * A + Cc; s = s * A + Cc;
  for (int i = 255; i >= 0; --i) {
    s = s * A + Cc;
    // Python: r = (int64(s) + 31) % (i + 1) with a non-negative result; the sum is taken in
    // unbounded integers, so do it in 128-bit-safe pieces.
    const int64_t v = int64_t(s);
    const int64_t n = i + 1;
    int64_t r = (v % n + 31 % n) % n;
    if (r < 0) r += n;
    perm[i] = source[r];
    source[r] = source[i];
  }
}

}  // namespace ble
