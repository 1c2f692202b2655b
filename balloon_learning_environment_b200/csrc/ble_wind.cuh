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
constexpr int kYC = kNY - 1, kPC = kNP - 1, kTC = kNT - 1;     // 20 x 9 x 8 (y, pressure, time) cells

BLE_HD int64_t native_index(int ix, int iy, int ip, int it, int c) {
  return ((((int64_t(ix) * kNY + iy) * kNP + ip) * kNT + it) * 2 + c);
}

// Device layout ("windows").  A lookup needs the 16 corners {x,x+1}x{y,y+1}x{p,p+1}x{t,t+1} x {u,v}
// = 32 floats = 128 bytes.  B200's L2 fills whole 128-byte lines from HBM (measured: 16.3 DRAM
// sectors per lookup with four scattered 32-byte cells), so the layout makes those 128 bytes
// CONTIGUOUS: for every (y, p, t) cell a row of per-x blocks, block(ix) = corners of column ix,
// ordered [dy][dp][dt][uv] (16 floats = 64 B).  The window of a lookup = block(ix) + block(ix+1).
//   BLE_LAYOUT_X64  : blocks packed every 64 B (21 per row): windows overlap, a window straddles two
//                     lines half of the time -> 1.5 lines / lookup, 1,935,360 B per field (6.1x).
//   BLE_LAYOUT_X128 : every window stored separately, 128 B aligned (20 per row): exactly 1 line /
//                     lookup, 3,686,400 B per field (11.6x).
// Within a window, 16-byte chunk j = dx*4 + dy*2 + dp holds (t0.u, t0.v, t1.u, t1.v).
struct FieldLayout {
  int32_t x_stride_floats;     // 16 (X64) or 32 (X128)
  int32_t row_floats;          // floats per (y,p,t) row: 21*16 = 336 or 20*32 = 640
  int64_t field_floats;        // kYC*kPC*kTC rows
};
BLE_HD FieldLayout make_layout(int kind) {
  FieldLayout l;
  if (kind == 1) { l.x_stride_floats = 32; l.row_floats = (kNX - 1) * 32; }
  else { l.x_stride_floats = 16; l.row_floats = kNX * 16; }
  l.field_floats = int64_t(kYC) * kPC * kTC * l.row_floats;
  return l;
}
BLE_HD int64_t window_index(const FieldLayout& l, int ix, int iy, int pc, int tc) {
  return ((int64_t(iy) * kPC + pc) * kTC + tc) * l.row_floats + int64_t(ix) * l.x_stride_floats;
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

// Cell of the lookup and the four interpolation weights.
template <typename Real>
struct FieldCell { int ix, iy, pc, tc; Real wx, wy, wp, wt; };

template <typename Real>
BLE_HD FieldCell<Real> locate(const FieldPoint& q) {
  FieldCell<Real> c;
  axis_cell<Real>(q.x_km, -500.f, 50.f, kNX - 1, &c.ix, &c.wx);
  axis_cell<Real>(q.y_km, -500.f, 50.f, kYC, &c.iy, &c.wy);
  axis_cell<Real>(q.p, 5000.f, 1000.f, kPC, &c.pc, &c.wp);
  axis_cell<Real>(q.t_h, 0.f, 6.f, kTC, &c.tc, &c.wt);
  return c;
}

// Contribution of 16-byte chunk j (= dx*4 + dy*2 + dp) of the window to (u, v).
template <typename Real>
BLE_HD void chunk_contribution(const FieldCell<Real>& c, int j, const float4& v4, Real* u, Real* v) {
  const Real one = Real(1);
  const Real w = ((j & 4) ? c.wx : one - c.wx) * ((j & 2) ? c.wy : one - c.wy) * ((j & 1) ? c.wp : one - c.wp);
  *u = w * ((one - c.wt) * Real(v4.x) + c.wt * Real(v4.z));
  *v = w * ((one - c.wt) * Real(v4.y) + c.wt * Real(v4.w));
}

// 16-corner multilinear interpolation (== scipy interpn 'linear', grid_based_wind_field.py:91);
// `ldchunk(j)` returns 16-byte chunk j of the lookup's 128-byte window.
template <typename Real, typename LoadChunk>
BLE_HD void interp_window(const FieldCell<Real>& c, LoadChunk ldchunk, Real* u, Real* v) {
  Real su = Real(0), sv = Real(0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < 8; ++j) {
    Real pu, pv;
    chunk_contribution<Real>(c, j, ldchunk(j), &pu, &pv);
    su += pu;
    sv += pv;
  }
  *u = su;
  *v = sv;
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
#define BLE_HARMONIC_TABLE                                                                        \
  {{0.1445, 702.269, 2116.987, 2587.802, 245.0},   {0.2766, 1483.570, 752.124, 646.208, 16.39},   \
   {0.2627, 276.810, 147.040, 587.702, 3.836},     {0.2137, 10214.525, 1512.216, 965.629, 41.780}, \
   {0.1025, 181.286, 420.942, 8500.0, 245.0},      {0.2716, 1974.228, 2028.814, 713.697, 26.435}, \
   {0.2684, 699.738, 541.845, 632.116, 9.530},     {0.2348, 217.750, 196.522, 686.825, 3.546},    \
   {0.1186, 47.500, 43.048, 66.553, 8.424},        {0.1066, 3663.291, 232.023, 7499.741, 225.0}}
static const double kHarmonicsHost[10][5] = BLE_HARMONIC_TABLE;
#if defined(__CUDACC__)
static __constant__ double kHarmonicsDev[10][5] = BLE_HARMONIC_TABLE;
#endif
// Reciprocal spacings in the units of the state: (1 / (1000 sx), 1 / (1000 sy), 1 / sp, 1 / (3600 st)) so that the
// lattice coordinate is x_m * inv + offset (production kernel; one multiply instead of two fp64 divisions).
#define BLE_HI(sx, sy, sp, st) {1.0 / (1000.0 * (sx)), 1.0 / (1000.0 * (sy)), 1.0 / (sp), 1.0 / (3600.0 * (st))}
#if defined(__CUDACC__)
static __constant__ double kHarmonicsInvDev[10][4] = {
    BLE_HI(702.269, 2116.987, 2587.802, 245.0),   BLE_HI(1483.570, 752.124, 646.208, 16.39),
    BLE_HI(276.810, 147.040, 587.702, 3.836),     BLE_HI(10214.525, 1512.216, 965.629, 41.780),
    BLE_HI(181.286, 420.942, 8500.0, 245.0),      BLE_HI(1974.228, 2028.814, 713.697, 26.435),
    BLE_HI(699.738, 541.845, 632.116, 9.530),     BLE_HI(217.750, 196.522, 686.825, 3.546),
    BLE_HI(47.500, 43.048, 66.553, 8.424),        BLE_HI(3663.291, 232.023, 7499.741, 225.0)};
// NoisyWindComponent.get_noise (:180-211): out = (sum w_h n_h) * sqrt(sum w / sum w^2) / sum w
static __constant__ float kBlendU[5] = {0.1445f, 0.2766f, 0.2627f, 0.2137f, 0.1025f};
static __constant__ float kBlendV[5] = {0.2716f, 0.2684f, 0.2348f, 0.1186f, 0.1066f};
constexpr float kBlendScaleU = 2.1196478805123253f, kBlendScaleV = 2.1018160718443704f;
#endif
BLE_HD void harmonic_params(int h10, double* w, double* sx, double* sy, double* sp, double* st) {
#if defined(__CUDA_ARCH__)
  const double* t = kHarmonicsDev[h10];
#else
  const double* t = kHarmonicsHost[h10];
#endif
  *w = t[0]; *sx = t[1]; *sy = t[2]; *sp = t[3]; *st = t[4];
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

// ---- A/B form: noise4d(x, y, z, w) summed over EVERY lattice vertex whose kernel (2 - |d|^2)^4 is non-zero --------
// (round-1 definition, oracle form='all'; the production kernels use the decision-tree selection further down and
// fall back to this sum only when built with -DBLE_NOISE_ALL_VERTICES).
// Candidate c in [0, 80): m = c / 5 is the cube corner (bit k of m = offset along axis k),
// e = c % 5: 0 = the corner itself, 1..4 = step one further out along axis e-1
// (offset 1 -> 2, offset 0 -> -1).  `perm` is this generator's 256-entry table.
template <typename Real, typename Perm>
BLE_HD Real simplex_noise4_all(const Perm& perm, double x, double y, double z, double w) {
  const double s = (x + y + z + w) * kStretch4;
  const double fx = floor(x + s), fy = floor(y + s), fz = floor(z + s), fw = floor(w + s);
  const double q = (fx + fy + fz + fw) * kSquish4;
  const Real dx0 = Real(x - (fx + q)), dy0 = Real(y - (fy + q));
  const Real dz0 = Real(z - (fz + q)), dw0 = Real(w - (fw + q));
  const int xb = int(int64_t(fx) & 255), yb = int(int64_t(fy) & 255);
  const int zb = int(int64_t(fz) & 255), wb = int(int64_t(fw) & 255);

  // pass 1: in-range mask over the 80 candidates.  With e_a = d0_a - o_a and T = tot * squish,
  //   |d|^2 = sum_a (e_a - T)^2 = sum e_a^2 - 2 T sum e_a + 4 T^2,
  // so a corner costs two 4-term sums and one FMA, and each "one step further" neighbour two more adds.
  uint64_t mask_lo = 0;   // candidates 0..63
  uint32_t mask_hi = 0;   // candidates 64..79
  const Real sq = Real(kSquish4);
  const Real d0[4] = {dx0, dy0, dz0, dw0};
  Real e1[4][2], q1[4][2], dq[4][2];                 // e, e^2 at offsets 0 / 1; e^2 change when stepping to -1 / 2
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    e1[a][0] = d0[a]; e1[a][1] = d0[a] - Real(1);
    q1[a][0] = e1[a][0] * e1[a][0]; q1[a][1] = e1[a][1] * e1[a][1];
    const Real em = d0[a] + Real(1), ep = d0[a] - Real(2);
    dq[a][0] = em * em - q1[a][0];                   // offset 0 -> -1
    dq[a][1] = ep * ep - q1[a][1];                   // offset 1 ->  2
  }
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int bit[4] = {m & 1, (m >> 1) & 1, (m >> 2) & 1, (m >> 3) & 1};
    const int pc = bit[0] + bit[1] + bit[2] + bit[3];
    const Real s1 = (e1[0][bit[0]] + e1[1][bit[1]]) + (e1[2][bit[2]] + e1[3][bit[3]]);
    const Real s2 = (q1[0][bit[0]] + q1[1][bit[1]]) + (q1[2][bit[2]] + q1[3][bit[3]]);
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      int tot = pc;
      Real t1 = s1, t2 = s2;
      if (e > 0) {
        const int a = e - 1;
        tot += bit[a] ? 1 : -1;
        t1 += bit[a] ? Real(-1) : Real(1);           // e_a changes by -(o' - o)
        t2 += dq[a][bit[a]];
      }
      const Real tt = Real(tot) * sq;
      const Real attn = (Real(2) - Real(4) * tt * tt) - t2 + Real(2) * tt * t1;
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
    attn = attn > Real(0) ? attn : Real(0);          // pass 1 decides the mask with a differently rounded sum
    int h = perm[(xb + oi) & 255];
    h = perm[(h + yb + oj) & 255];
    h = perm[(h + zb + ok) & 255];
    h = perm[(h + wb + ol) & 255];
    attn *= attn;
    value += attn * attn * gradient_dot<Real>(h, dx, dy, dz, dw);
  }
  return value / Real(30.0);
}

// ---- second-generation evaluation of the same all-vertices sum (A/B form) ------------------------------------
// simplex_noise4_all above spends ~620 instructions deciding which of the 80 candidates are in range and then loops
// over them one at a time (a warp iterates max-over-lanes ~12 times at ~105 instructions).  Two facts make it
// cheaper.  (1) In the skewed lattice coordinates u = x + STRETCH * sum(x) the kernel argument separates:
//     |d|^2 = sum_a (u_a - o_a)^2 + (sum(u) - sum(o))^2,
//   so the range test of a candidate is two table reads and two subtractions.  (2) The 16 cube corners are needed by
//   some lane of the warp almost always (7.6 of them are in range on average), so they are evaluated unconditionally
//   and branch-free -- compile-time offsets, the permutation look-ups shared as a binary tree (30 instead of 64) --
//   and only the "one step further out" neighbours (1.2 in range on average, never more than one direction per axis:
//   (u_a + 1)^2 < 2 needs u_a < 0.414, (u_a - 2)^2 < 2 needs u_a > 0.586) go through a mask + loop.
// Same definition, different summation order: agrees with simplex_noise4_all to fp32 rounding.
template <typename Real>
BLE_HD Real gradient_dot_fast(int hash, Real dx, Real dy, Real dz, Real dw) {
  // hash bits: [7..4] negate component 3..0, [3..2] position of the "3"
  const Real ax = (hash & 0x10) ? -dx : dx, ay = (hash & 0x20) ? -dy : dy;
  const Real az = (hash & 0x40) ? -dz : dz, aw = (hash & 0x80) ? -dw : dw;
  const int j = (hash >> 2) & 3;
  const Real lo = (j & 1) ? ay : ax, hi = (j & 1) ? aw : az;
  const Real pick = (j & 2) ? hi : lo;
  return ((ax + ay) + (az + aw)) + Real(2) * pick;
}

// `tables_done()` is called once, after the LAST read of the permutation table: the fused step kernels release their
// staging buffer there (and start the TMA copy of the next harmonic's tables), so that the copy runs under the 16 corner
// evaluations instead of being waited for (the wait was 5.6 % of k_step_warp's stall samples).  Hence the order: all the
// hashes of the 16 corners, the (few) extended neighbours completely, tables_done(), the corners' arithmetic.
struct NoTablesDone { BLE_HD void operator()() const {} };

template <typename Real, typename Perm, typename Done = NoTablesDone>
BLE_HD Real simplex_noise4_v2_all(const Perm& perm, double x, double y, double z, double w, Done tables_done = Done()) {
  const double s = (x + y + z + w) * kStretch4;
  const double xs = x + s, ys = y + s, zs = z + s, ws = w + s;
  const double fx = floor(xs), fy = floor(ys), fz = floor(zs), fw = floor(ws);
  const double q = (fx + fy + fz + fw) * kSquish4;
  const Real d0[4] = {Real(x - (fx + q)), Real(y - (fy + q)), Real(z - (fz + q)), Real(w - (fw + q))};
  const Real in[4] = {Real(xs - fx), Real(ys - fy), Real(zs - fz), Real(ws - fw)};   // inside the unit cell
  const Real S = (in[0] + in[1]) + (in[2] + in[3]);
  const int cb[4] = {int(int64_t(fx) & 255), int(int64_t(fy) & 255), int(int64_t(fz) & 255), int(int64_t(fw) & 255)};
  const Real sq = Real(kSquish4);

  // ---- hashes of the 16 corners ----
  int h1[2], h2[4], h3[8], hc[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 2; ++i) h1[i] = perm[(cb[0] + i) & 255];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 4; ++i) h2[i] = perm[(h1[i & 1] + cb[1] + (i >> 1)) & 255];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 8; ++i) h3[i] = perm[(h2[i & 3] + cb[2] + (i >> 2)) & 255];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int m = 0; m < 16; ++m) hc[m] = perm[(h3[m & 7] + cb[3] + (m >> 3)) & 255];

  // ---- neighbours one step further out along one axis: range test in skewed coordinates, then a loop ----
  Real A[4][2], Ae[4];
  int oe[4];                                           // the only feasible outward offset per axis: -1 or 2
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 4; ++a) {
    const Real u0 = in[a], u1 = in[a] - Real(1);
    A[a][0] = u0 * u0; A[a][1] = u1 * u1;
    const bool low = in[a] < Real(0.5);
    const Real ue = low ? in[a] + Real(1) : in[a] - Real(2);
    Ae[a] = ue * ue;
    oe[a] = low ? -1 : 2;
  }
  Real Q[7];                                           // 2 - (S - t)^2 for t = -1 .. 5
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int t = 0; t < 7; ++t) { const Real v = S - Real(t - 1); Q[t] = Real(2) - v * v; }
  uint32_t mask = 0;                                   // bit a * 8 + m8, m8 = offsets (0/1) of the other three axes
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 4; ++a) {
    const int b0 = a == 0 ? 1 : 0, b1 = a <= 1 ? 2 : 1, b2 = a <= 2 ? 3 : 2;
    const bool low = oe[a] < 0;
    Real QE[4];                                        // by the number of set bits among the other axes
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int n = 0; n < 4; ++n) QE[n] = (low ? Q[n] : Q[n + 3]) - Ae[a];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m8 = 0; m8 < 8; ++m8) {
      const int pc = (m8 & 1) + ((m8 >> 1) & 1) + (m8 >> 2);
      const Real attn = QE[pc] - A[b0][m8 & 1] - A[b1][(m8 >> 1) & 1] - A[b2][m8 >> 2];
      if (attn > Real(0)) mask |= 1u << (a * 8 + m8);
    }
  }
  const uint32_t oe_packed = uint32_t(oe[0] & 255) | (uint32_t(oe[1] & 255) << 8) | (uint32_t(oe[2] & 255) << 16) |
                             (uint32_t(oe[3] & 255) << 24);
  Real outer = Real(0);
  while (mask) {
#if defined(__CUDA_ARCH__)
    const int c = __ffs(int(mask)) - 1;
#else
    const int c = __builtin_ctz(mask);
#endif
    mask &= mask - 1;
    const int a = c >> 3, m8 = c & 7;
    const int m = (m8 & ((1 << a) - 1)) | ((m8 >> a) << (a + 1));        // offsets of the other axes, bit a cleared
    const int ext = int(int8_t((oe_packed >> (8 * a)) & 255));
    int o[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 4; ++b) o[b] = (b == a) ? ext : ((m >> b) & 1);
    const Real t = Real(o[0] + o[1] + o[2] + o[3]) * sq;
    const Real dx = d0[0] - Real(o[0]) - t, dy = d0[1] - Real(o[1]) - t;
    const Real dz = d0[2] - Real(o[2]) - t, dw = d0[3] - Real(o[3]) - t;
    Real attn = Real(2) - dx * dx - dy * dy - dz * dz - dw * dw;
    attn = attn > Real(0) ? attn : Real(0);            // the mask was decided with a differently rounded sum
    int h = perm[(cb[0] + o[0]) & 255];
    h = perm[(h + cb[1] + o[1]) & 255];
    h = perm[(h + cb[2] + o[2]) & 255];
    h = perm[(h + cb[3] + o[3]) & 255];
    attn *= attn;
    outer += attn * attn * gradient_dot_fast<Real>(h, dx, dy, dz, dw);
  }
  tables_done();

  // ---- the 16 corners, unconditionally ----
  Real e[4][2];                                        // real-space displacement per axis at offsets 0 / 1
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 4; ++a) { e[a][0] = d0[a]; e[a][1] = d0[a] - Real(1); }
  Real value = Real(0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int m = 0; m < 16; ++m) {
    const int pc = (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + (m >> 3);
    const Real t = Real(pc) * sq;
    const Real dx = e[0][m & 1] - t, dy = e[1][(m >> 1) & 1] - t, dz = e[2][(m >> 2) & 1] - t, dw = e[3][m >> 3] - t;
    Real attn = Real(2) - dx * dx - dy * dy - dz * dz - dw * dw;
    attn = attn > Real(0) ? attn : Real(0);
    attn *= attn;
    value += attn * attn * gradient_dot_fast<Real>(hc[m], dx, dy, dz, dw);
  }
  return (value + outer) / Real(30.0);
}

// ---- the package's vertex selection (oracle form='tree', the default) ------------------------------------------
// opensimplex.noise4d does not visit every in-range vertex: a decision tree on the position inside the unit cell
// picks the BASE vertices of one of four regions plus THREE extras (oracle/opensimplex4.py spells the rule out and
// cites what it restates).  Here the tree is evaluated branch-free:
//   * regions B / D (inSum > 2) are the point reflection v -> 1 - v of A / C, so the selection runs in the lower frame
//     u = 1 - ins and the result is reflected back;
//   * a vertex is an 8-bit code, 2 bits per axis holding offset + 1 (offsets are -1 .. 2), so that "lower a zero axis
//     to -1" clears one bit, "(0,0,0,2) on axis k" sets one bit, and the reflection is a bitwise NOT;
//   * the base vertices are cube corners: a 16-bit mask over the corner index m (bit a of m = offset along axis a).
BLE_HD uint32_t simplex_spread2(uint32_t m) {           // bit a of m -> bit 2a
  uint32_t t = (m | (m << 2)) & 0x33u;
  return (t | (t << 1)) & 0x55u;
}

template <typename Real>
struct SimplexTop2 {                                     // the two closest candidates so far
  Real a_sc, b_sc;
  uint32_t a_po, b_po;
  bool a_big, b_big;
  BLE_HD void offer(Real score, uint32_t point, bool big) {
    const bool into_b = (a_sc >= b_sc) && (score > b_sc);
    const bool into_a = (a_sc < b_sc) && (score > a_sc);
    b_sc = into_b ? score : b_sc; b_po = into_b ? point : b_po; b_big = into_b ? big : b_big;
    a_sc = into_a ? score : a_sc; a_po = into_a ? point : a_po; a_big = into_a ? big : a_big;
  }
};

// in[4] = position inside the unit cell.  Returns the corner mask of the region's base vertices and writes the three
// extras as codes (see above).
template <typename Real>
BLE_HD uint32_t simplex_tree_select(const Real in[4], uint32_t ext[3]) {
  const Real S = ((in[0] + in[1]) + in[2]) + in[3];
  const bool refl = S > Real(2);
  const bool penta = (S <= Real(1)) || (S >= Real(3));
  const Real u0 = refl ? Real(1) - in[0] : in[0], u1 = refl ? Real(1) - in[1] : in[1];
  const Real u2 = refl ? Real(1) - in[2] : in[2], u3 = refl ? Real(1) - in[3] : in[3];
  const Real T = ((u0 + u1) + u2) + u3;
  uint32_t cc, third_code;
  bool keep_first, three_lowered;
  if (penta) {
    // two closest of the unit vertices; is the origin closer than one of them?
    SimplexTop2<Real> t{u0, u1, 1u, 2u, true, true};
    t.offer(u2, 4u, true);
    t.offer(u3, 8u, true);
    const Real un = Real(1) - T;
    const bool origin_close = (un > t.a_sc) || (un > t.b_sc);
    cc = origin_close ? (t.b_sc > t.a_sc ? t.b_po : t.a_po) : (t.a_po | t.b_po);
    keep_first = !origin_close;        // extras = {cc, cc - e_z0, cc - e_z1}  or  {cc - e_z0, cc - e_z1, cc - e_z2}
    three_lowered = origin_close;
    third_code = 0;
  } else {
    // the closer of each complementary two-ones pair, then the unit vertices (scores as the package defines them)
    const bool p0 = (u0 + u1) > (u2 + u3), p1 = (u0 + u2) > (u1 + u3), p2 = (u0 + u3) > (u1 + u2);
    SimplexTop2<Real> t{p0 ? u0 + u1 : u2 + u3, p1 ? u0 + u2 : u1 + u3, p0 ? 0x3u : 0xCu, p1 ? 0x5u : 0xAu, true, true};
    t.offer(p2 ? u0 + u3 : u1 + u2, p2 ? 0x9u : 0x6u, true);
    const Real r = Real(2) - T;
    t.offer(r + u0, 1u, false);
    t.offer(r + u1, 2u, false);
    t.offer(r + u2, 4u, false);
    t.offer(r + u3, 8u, false);
    const bool both_big = t.a_big && t.b_big, both_small = !t.a_big && !t.b_big;
    const uint32_t big_po = t.a_big ? t.a_po : t.b_po, small_po = t.a_big ? t.b_po : t.a_po;
    cc = (both_big || both_small) ? (t.a_po | t.b_po) : big_po;
    keep_first = both_big;             // extras = {cc, cc - e_z0, 2 e_k}  or  {cc - e_z0, cc - e_z1, 2 e_k | origin}
    three_lowered = false;
    const uint32_t k = both_big ? (t.a_po & t.b_po) : small_po;          // one-hot axis of the (0,0,0,2) vertex
    third_code = both_small ? 0x55u : (0x55u | ((k * k) << 1));
  }
  const uint32_t code = 0x55u + simplex_spread2(cc);
  const uint32_t z = ~cc & 15u;
  const uint32_t h0 = z & (0u - z), zr = z ^ h0, h1 = zr & (0u - zr), h2 = zr ^ h1;   // one-hot zero axes, ascending
  const uint32_t l0 = code & ~(h0 * h0), l1 = code & ~(h1 * h1), l2 = code & ~(h2 * h2);
  uint32_t e0 = keep_first ? code : l0;
  uint32_t e1 = keep_first ? l0 : l1;
  uint32_t e2 = three_lowered ? l2 : (penta ? l1 : third_code);
  if (refl) { e0 = ~e0 & 0xFFu; e1 = ~e1 & 0xFFu; e2 = ~e2 & 0xFFu; }
  ext[0] = e0; ext[1] = e1; ext[2] = e2;
  return penta ? (refl ? 0xE880u : 0x0117u) : (refl ? 0x7EE8u : 0x177Eu);
}

// One selected vertex with run-time offsets (the extras): 4 dependent table reads + the kernel.
template <typename Real, typename Perm>
BLE_HD Real simplex_vertex(const Perm& perm, const int cb[4], const Real d0[4], uint32_t code) {
  const int o0 = int(code & 3u) - 1, o1 = int((code >> 2) & 3u) - 1, o2 = int((code >> 4) & 3u) - 1,
            o3 = int((code >> 6) & 3u) - 1;
  const Real t = Real(o0 + o1 + o2 + o3) * Real(kSquish4);
  const Real dx = d0[0] - Real(o0) - t, dy = d0[1] - Real(o1) - t, dz = d0[2] - Real(o2) - t, dw = d0[3] - Real(o3) - t;
  Real attn = Real(2) - dx * dx - dy * dy - dz * dz - dw * dw;
  attn = attn > Real(0) ? attn : Real(0);
  int h = perm[(cb[0] + o0) & 255];
  h = perm[(h + cb[1] + o1) & 255];
  h = perm[(h + cb[2] + o2) & 255];
  h = perm[(h + cb[3] + o3) & 255];
  attn *= attn;
  return attn * attn * gradient_dot_fast<Real>(h, dx, dy, dz, dw);
}

// Plain evaluation of the tree form (audit / host replay): base corners in index order, then the extras.
template <typename Real, typename Perm>
BLE_HD Real simplex_noise4(const Perm& perm, double x, double y, double z, double w) {
  const double s = (x + y + z + w) * kStretch4;
  const double xs = x + s, ys = y + s, zs = z + s, ws = w + s;
  const double fx = floor(xs), fy = floor(ys), fz = floor(zs), fw = floor(ws);
  const double q = (fx + fy + fz + fw) * kSquish4;
  const Real d0[4] = {Real(x - (fx + q)), Real(y - (fy + q)), Real(z - (fz + q)), Real(w - (fw + q))};
  const Real in[4] = {Real(xs - fx), Real(ys - fy), Real(zs - fz), Real(ws - fw)};
  const int cb[4] = {int(int64_t(fx) & 255), int(int64_t(fy) & 255), int(int64_t(fz) & 255), int(int64_t(fw) & 255)};
  uint32_t ext[3];
  const uint32_t cmask = simplex_tree_select<Real>(in, ext);
  Real value = Real(0);
  for (uint32_t m = 0; m < 16; ++m)
    if ((cmask >> m) & 1u) value += simplex_vertex<Real>(perm, cb, d0, 0x55u + simplex_spread2(m));
  for (int k = 0; k < 3; ++k) value += simplex_vertex<Real>(perm, cb, d0, ext[k]);
  return value / Real(30.0);
}

// Production evaluation of the tree form.  Everything runs in the LOWER frame (the unit cell reflected when inSum > 2,
// see simplex_tree_select): there the base vertices are always among 11 cube corners -- the 4 unit corners (every
// region), the 6 two-ones corners (regions C / D) and the origin (regions A / B) -- so 11 corners are evaluated
// branch-free with compile-time lower-frame offsets (the reflection only swaps which of two precomputed values an
// offset bit selects), the permutation look-ups shared as a pruned binary tree (24 reads), and the ones the region
// does not use masked out; the three extras follow with run-time offsets.  Order as in the A/B form so that the fused
// kernels can release the staged tables early: corner hashes, the extras completely, tables_done(), the corners'
// arithmetic.
template <typename Real, typename Perm, typename Done = NoTablesDone>
BLE_HD Real simplex_noise4_v2(const Perm& perm, double x, double y, double z, double w, Done tables_done = Done()) {
#if defined(BLE_NOISE_ALL_VERTICES)
  return simplex_noise4_v2_all<Real, Perm, Done>(perm, x, y, z, w, tables_done);
#else
  const double s = (x + y + z + w) * kStretch4;
  const double xs = x + s, ys = y + s, zs = z + s, ws = w + s;
  const double fx = floor(xs), fy = floor(ys), fz = floor(zs), fw = floor(ws);
  const double q = (fx + fy + fz + fw) * kSquish4;
  const Real d0[4] = {Real(x - (fx + q)), Real(y - (fy + q)), Real(z - (fz + q)), Real(w - (fw + q))};
  const Real in[4] = {Real(xs - fx), Real(ys - fy), Real(zs - fz), Real(ws - fw)};   // inside the unit cell
  const int cb[4] = {int(int64_t(fx) & 255), int(int64_t(fy) & 255), int(int64_t(fz) & 255), int(int64_t(fw) & 255)};
  const Real sq = Real(kSquish4);

  // ---- vertex selection ----
  uint32_t ext[3];
  const uint32_t cmask = simplex_tree_select<Real>(in, ext);
  const bool refl = (cmask & 0x8000u) != 0 || cmask == 0x7EE8u;   // regions B / D
  const bool penta = (cmask & 0x8001u) != 0;                      // regions A / B (the only masks with corner 0 or 15)

  // ---- hashes of the 11 lower-frame corners: lower-frame offset bit b of axis a is the real offset b ^ refl ----
  int ca[4][2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 4; ++a) { ca[a][0] = cb[a] + (refl ? 1 : 0); ca[a][1] = cb[a] + (refl ? 0 : 1); }
  int h1[2], h2[4], h3[7];
  h1[0] = perm[ca[0][0] & 255]; h1[1] = perm[ca[0][1] & 255];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 4; ++i) h2[i] = perm[(h1[i & 1] + ca[1][i >> 1]) & 255];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 7; ++i) h3[i] = perm[(h2[i & 3] + ca[2][i >> 2]) & 255];
  // corner list: origin, 4 units, 6 two-ones (lower-frame index m, bit a = offset along axis a)
  constexpr int kCorner[11] = {0, 1, 2, 4, 8, 3, 5, 6, 9, 10, 12};
  int hc[11];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int c = 0; c < 11; ++c) hc[c] = perm[(h3[kCorner[c] & 7] + ca[3][kCorner[c] >> 3]) & 255];

  // ---- the three extras ----
  Real outer = Real(0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 3; ++k) outer += simplex_vertex<Real>(perm, cb, d0, ext[k]);
  tables_done();

  // ---- the corners ----
  Real e[4][2];                                        // real-space displacement per axis at lower-frame offsets 0 / 1
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 4; ++a) { e[a][0] = refl ? d0[a] - Real(1) : d0[a]; e[a][1] = refl ? d0[a] : d0[a] - Real(1); }
  const Real tv[3] = {refl ? Real(4) * sq : Real(0), refl ? Real(3) * sq : sq, Real(2) * sq};   // by lower-frame popcount
  Real value = Real(0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int c = 0; c < 11; ++c) {
    const int m = kCorner[c];
    const int pc = (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + (m >> 3);
    const Real t = tv[pc];
    const Real dx = e[0][m & 1] - t, dy = e[1][(m >> 1) & 1] - t, dz = e[2][(m >> 2) & 1] - t, dw = e[3][m >> 3] - t;
    Real attn = Real(2) - dx * dx - dy * dy - dz * dz - dw * dw;
    const bool used = pc == 1 ? true : (pc == 0 ? penta : !penta);
    attn = (attn > Real(0) && used) ? attn : Real(0);
    attn *= attn;
    value += attn * attn * gradient_dot_fast<Real>(hc[c], dx, dy, dz, dw);
  }
  return (value + outer) / Real(30.0);
#endif
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
