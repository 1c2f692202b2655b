// Kernels of the Perciatelli observation surface (included by ble_engine.cu).
//
//   k_feat_observe   1 thread / balloon   WindGP.observe (env/wind_gp.py:98-119) into a 120-slot ring
//   k_feat_ambient   1 thread / balloon   features 0..15 (env/features.py:382-455)
//   k_feat_range     1 warp   / balloon   reachable pressure range (env/balloon/pressure_range_builder.py:203-275)
//   k_gp_factor      1 CTA    / balloon   K = kernel + noise, packed Cholesky (sklearn GPR.fit, wind_gp.py:179)
//   k_gp_column      1 CTA    / balloon   181 predictive means / variances + forecast column -> features 16..1098
//                                         (wind_gp.py:143-241, features.py:457-556)
// All fp64 (see ble_features.cuh).
#pragma once
// NOTE: included from inside `namespace ble` of ble_engine.cu (needs DevState, DD, RR, noise_at, forecast_at).

constexpr int kGpPacked = kGpWindow * (kGpWindow + 1) / 2;     // 7,260 doubles

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

// ---- observe -------------------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(128) k_feat_observe(DevState<Real> d) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const int32_t cnt = d.gp_count[e];
  const double t = double(d.t_elapsed[e]);
  double* ring = d.gp_obs + e * int64_t(kGpWindow * 6);
  if (cnt > 0 && ring[((cnt - 1) % kGpWindow) * 6 + 3] == t) return;      // this state is already in the history
  // measurement - forecast (wind_gp.py:112-116) == the simplex noise at this point
  Real nu = Real(0), nv = Real(0);
  if (d.enable_noise) noise_at<Real>(d, e, &nu, &nv);
  double* slot = ring + (cnt % kGpWindow) * 6;
  slot[0] = DD(d, D_X, e); slot[1] = DD(d, D_Y, e); slot[2] = DD(d, D_P, e); slot[3] = t;
  slot[4] = double(nu); slot[5] = double(nv);
  d.gp_count[e] = cnt + 1;
}

// ---- ambient features ----------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(128) k_feat_ambient(DevState<Real> d, float* __restrict__ obs) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const double x = DD(d, D_X, e), y = DD(d, D_Y, e), p = DD(d, D_P, e), sp = DD(d, D_SP, e);
  const int64_t ts = d.l[int64_t(L_DATE_TIME) * d.n + e];
  double lat, lng, el, flux;
  latlng_from_offset<double>(double(RR(d, R_LAT0, e)), double(RR(d, R_LNG0, e)), x, y, &lat, &lng);
  solar_calculator<double>(lat, lng, ts, &el, &flux);
  bool ok;
  const double st = sunrise_time(lat, lng, ts, &ok);
  const double soc = DD(d, D_CHARGE, e) / kBatteryCapacityWh;
  const double pr = (p + fmax(sp, 0.0)) / p;
  const double dist_km = sqrt(x * x + y * y) / 1000.0;
  const double heading = atan2(-x / 1000.0, -y / 1000.0);
  const uint32_t fl = d.flags[e];
  const int last_cmd = int((fl >> 2) & 3);
  const bool paused = ((fl >> 9) & 1u) || (((fl >> 4) & 7u) != 0u) || (((fl >> 7) & 3u) != 0u);
  const bool excess = (solar_power<double>(el, p) > kDayLoadW) && (soc > 0.99);
  const double power = fmin(fmax((power_table_lookup(pr, soc) - 100.0) / 200.0, 0.0), 1.0);
  float* o = obs + e * int64_t(kNumFeatures);
  o[0] = float(fmin(fmax((p - kLevelMin) / (kLevelMax - kLevelMin), 0.0), 1.0));
  o[1] = float(soc);
  o[2] = float(fmin(fmax((el + 90.0) / 180.0, 0.0), 1.0));
  o[3] = float(sin(st)); o[4] = float(cos(st));
  o[5] = float(sin(heading)); o[6] = float(cos(heading));
  o[7] = float(dist_km / (dist_km + 250.0));
  o[8] = last_cmd == kUp ? 1.f : 0.f;            // note the order UP, STAY, DOWN (features.py:426-435)
  o[9] = last_cmd == kStay ? 1.f : 0.f;
  o[10] = last_cmd == kDown ? 1.f : 0.f;
  o[11] = paused ? 1.f : 0.f; o[12] = paused ? 0.f : 1.f;
  o[13] = excess ? 1.f : 0.f;
  o[14] = float(power);
  o[15] = float(pr);
  if (!ok) atomicOr(&d.flags[e], 1u << 11);
}

// ---- reachable pressure range --------------------------------------------------------------------------
// get_pressure_range (env/balloon/pressure_range_builder.py:203-275) in two launches so that every lane works:
//   k_feat_range_levels  one thread per (balloon, scan level): the 20 stable-init solves of the scan (:231-233)
//   k_feat_range         one thread per balloon: the fully-vented pressure (:234-245), its own solve, and the
//                        two safe-pressure searches (:247-275)
template <typename Real>
__global__ void __launch_bounds__(128) k_feat_range_levels(DevState<Real> d, double* __restrict__ scratch /*[n][20][2]*/) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (t >= d.n * kRangeLevels) return;
  const int64_t e = t / kRangeLevels;
  const int lv = int(t - e * kRangeLevels);
  const double alpha = DD(d, D_ALPHA, e), mols_gas = double(RR(d, R_MOLS_GAS, e)), ir = double(RR(d, R_IR, e));
  const int64_t ts = d.l[int64_t(L_DATE_TIME) * d.n + e];
  double lat, lng;
  latlng_from_offset<double>(double(RR(d, R_LAT0, e)), double(RR(d, R_LNG0, e)), DD(d, D_X, e), DD(d, D_Y, e), &lat, &lng);
  double search_max, t_unused;
  atm_at_height_generic(alpha, kAltMin, &search_max, &t_unused);           // :230
  const double step = (search_max - 1000.0) / double(kRangeLevels - 1);    // np.linspace(1000, search_max, 20)
  const double level = (lv == kRangeLevels - 1) ? search_max : double(lv) * step + 1000.0;
  const StableParams sp = stable_params(alpha, level, mols_gas, lat, lng, ts, ir);
  scratch[t * 2] = level / sp.t_ambient;
  scratch[t * 2 + 1] = sp.superpressure;
}

template <typename Real>
__global__ void __launch_bounds__(128) k_feat_range(DevState<Real> d, const double* __restrict__ scratch) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const double alpha = DD(d, D_ALPHA, e), mols_gas = double(RR(d, R_MOLS_GAS, e)), ir = double(RR(d, R_IR, e));
  const int64_t ts = d.l[int64_t(L_DATE_TIME) * d.n + e];
  double lat, lng;
  latlng_from_offset<double>(double(RR(d, R_LAT0, e)), double(RR(d, R_LNG0, e)), DD(d, D_X, e), DD(d, D_Y, e), &lat, &lng);
  double search_max, t_unused;
  atm_at_height_generic(alpha, kAltMin, &search_max, &t_unused);
  const double step = (search_max - 1000.0) / double(kRangeLevels - 1);
  double levels[kRangeLevels], pt[kRangeLevels], sp[kRangeLevels + 1];
#pragma unroll
  for (int lv = 0; lv < kRangeLevels; ++lv) {
    levels[lv] = (lv == kRangeLevels - 1) ? search_max : double(lv) * step + 1000.0;
    pt[lv] = scratch[(e * kRangeLevels + lv) * 2];
    sp[lv] = scratch[(e * kRangeLevels + lv) * 2 + 1];
  }
  const double sig = min_float_pressure(levels, pt, mols_gas);
  sp[kRangeLevels] = stable_params(alpha, sig, mols_gas, lat, lng, ts, ir).superpressure;
  double pmin = 0.0, pmax = 0.0;
  const bool ok1 = search_safe_pressure(levels, sp, sig, sp[kRangeLevels], false, &pmin);
  const bool ok2 = search_safe_pressure(levels, sp, levels[kRangeLevels - 1], sp[kRangeLevels - 1], true, &pmax);
  d.feat_range[2 * e] = pmin;
  d.feat_range[2 * e + 1] = pmax;
  if (!(ok1 && ok2)) {                           // the reference raises ValueError here; flag it and mark all unreachable
    d.feat_range[2 * e] = 1.0; d.feat_range[2 * e + 1] = 0.0;
    atomicOr(&d.flags[e], 1u << 11);
  }
}

// The measurements inside the 6 h window (wind_gp.py:172-178), in ring order, as indices into the ring.
__device__ inline int gp_window_indices(const double* ring, int count, double now, int* idx) {
  const int stored = count < kGpWindow ? count : kGpWindow;
  int m = 0;
  for (int k = 0; k < stored; ++k) {
    const int slot = (count - stored + k) % kGpWindow;                      // chronological
    if (fabs(ring[slot * 6 + 3] - now) < kGpHorizonS) idx[m++] = slot;
  }
  return m;
}

// ---- K + Cholesky ----------------------------------------------------------------------------------------
constexpr int kFactorThreads = 128;
template <typename Real>
__global__ void __launch_bounds__(kFactorThreads) k_gp_factor(DevState<Real> d) {
  extern __shared__ __align__(16) double s_mem[];
  double* L = s_mem;                               // packed lower, kGpPacked
  double* a = L + kGpPacked;                       // [kGpWindow][4] scaled coordinates
  __shared__ int s_idx[kGpWindow];
  __shared__ int s_m;
  const int64_t e = blockIdx.x;
  const int tid = threadIdx.x;
  const double* ring = d.gp_obs + e * int64_t(kGpWindow * 6);
  if (tid == 0) s_m = gp_window_indices(ring, d.gp_count[e], double(d.t_elapsed[e]), s_idx);
  __syncthreads();
  const int m = s_m;
  if (tid == 0) d.gp_m[e] = m;
  if (m == 0) return;
  if (tid < m) {
    const double* o = ring + s_idx[tid] * 6;
    a[tid * 4 + 0] = o[0] / kGpScaleXY; a[tid * 4 + 1] = o[1] / kGpScaleXY;
    a[tid * 4 + 2] = o[2] / kGpScaleP;  a[tid * 4 + 3] = o[3] / kGpScaleT;
  }
  __syncthreads();
  const int total = tri(m);
  for (int k = tid; k < total; k += kFactorThreads) {
    int i = int((sqrt(8.0 * double(k) + 1.0) - 1.0) * 0.5);
    while (tri(i) > k) --i;
    while (tri(i + 1) <= k) ++i;
    const int j = k - tri(i);
    L[k] = gp_kernel(a + i * 4, a + j * 4) + (i == j ? kGpNoise : 0.0);    // kernel(X) + alpha I
  }
  __syncthreads();
  // left-looking Cholesky: thread i owns row i
  for (int j = 0; j < m; ++j) {
    double s = 0.0;
    if (tid >= j && tid < m) {
      const double* ri = L + tri(tid);
      const double* rj = L + tri(j);
      double s0 = 0.0, s1 = 0.0;
      int k = 0;
      for (; k + 1 < j; k += 2) { s0 += ri[k] * rj[k]; s1 += ri[k + 1] * rj[k + 1]; }
      if (k < j) s0 += ri[k] * rj[k];
      s = ri[j] - (s0 + s1);
    }
    __syncthreads();
    if (tid == j) L[tri(j) + j] = sqrt(s);
    __syncthreads();
    if (tid > j && tid < m) L[tri(tid) + j] = s / L[tri(j) + j];
    __syncthreads();                             // column j is complete before iteration j + 1 reads row j + 1
  }
  double* out = d.gp_chol + e * int64_t(kGpPacked);
  for (int k = tid; k < total; k += kFactorThreads) out[k] = L[k];
}

// ---- predictive column + feature assembly ------------------------------------------------------------------
// One CTA per balloon, one thread per query level (181) + one per target column (2).  Every thread
// runs a forward substitution v = L^-1 rhs (rhs = k* for a query, the measured errors for a target):
//   variance = sigma^2 - |v|^2,  mean = v . z  with z = L^-1 y      (sklearn GPR.predict)
// The substitution is blocked by 8 rows (left-looking): for row block b the thread keeps 8 fp64
// accumulators and streams v_j (own column, fp32 in shared memory) against the 8 x 1 slivers
// L[8b..8b+7][j], which are stored contiguously (block-column-major) so that a sliver is four
// broadcast LDS.128.  That is 14 instructions per 8 DFMA instead of 4 per DFMA.
constexpr int kColumnThreads = 192;              // 181 query levels + 2 right-hand sides (error u, v)
constexpr int kGpBlock = 8;
constexpr int kGpBlocks = kGpWindow / kGpBlock;  // 15
// Lb: for block b, columns j = 0 .. 8b+7, 8 rows each -> 64 * (1 + 2 + ... + 15) doubles
constexpr int kGpBlockedDoubles = kGpBlock * kGpBlock * (kGpBlocks * (kGpBlocks + 1) / 2);   // 7,680
constexpr size_t kColumnSmem = sizeof(double) * (kGpBlockedDoubles + kGpWindow * 4 + kGpWindow * 2 + kGpWindow) +
                               sizeof(float) * (size_t(kGpWindow) * kColumnThreads + kNumLevels * 3);

__device__ __forceinline__ int lb_offset(int b) { return kGpBlock * kGpBlock * ((b * (b + 1)) >> 1); }

template <typename Real>
__global__ void __launch_bounds__(kColumnThreads) k_gp_column(DevState<Real> d, float* __restrict__ obs) {
  extern __shared__ __align__(16) double s_mem[];
  double* Lb = s_mem;                              // blocked factor, kGpBlockedDoubles
  double* a = Lb + kGpBlockedDoubles;              // [m][4] scaled measurement coordinates
  double* z = a + kGpWindow * 4;                   // [m][2] = L^-1 y
  double* inv_diag = z + kGpWindow * 2;            // [m] 1 / L_ii
  float* V = reinterpret_cast<float*>(inv_diag + kGpWindow);   // [m][kColumnThreads]
  float* feat = V + size_t(kGpWindow) * kColumnThreads;          // [181][3]
  __shared__ int s_idx[kGpWindow];
  const int64_t e = blockIdx.x;
  const int tid = threadIdx.x;
  const int m = d.gp_m[e];
  const int nb = (m + kGpBlock - 1) / kGpBlock;    // row blocks actually used
  const double* ring = d.gp_obs + e * int64_t(kGpWindow * 6);
  const double x = DD(d, D_X, e), y = DD(d, D_Y, e), p_b = DD(d, D_P, e);
  const int32_t t_elapsed = d.t_elapsed[e];
  const double pmin = d.feat_range[2 * e], pmax = d.feat_range[2 * e + 1];
  if (m > 0) {
    if (tid == 0) gp_window_indices(ring, d.gp_count[e], double(t_elapsed), s_idx);
    // packed row-major (global) -> block-column-major (shared); rows >= m are identity padding
    const double* src = d.gp_chol + e * int64_t(kGpPacked);
    for (int k = tid; k < lb_offset(nb); k += kColumnThreads) {
      int b = 0;
      while (lb_offset(b + 1) <= k) ++b;
      const int rel = k - lb_offset(b);
      const int j = rel / kGpBlock, r = rel % kGpBlock;
      const int i = b * kGpBlock + r;
      double v = 0.0;
      if (i < m && j <= i) v = src[tri(i) + j];
      else if (i >= m && j == i) v = 1.0;
      Lb[k] = v;
    }
    __syncthreads();
    if (tid < nb * kGpBlock) {
      const int b = tid / kGpBlock, r = tid % kGpBlock;
      inv_diag[tid] = 1.0 / Lb[lb_offset(b) + tid * kGpBlock + r];
    }
    if (tid < m) {
      const double* o = ring + s_idx[tid] * 6;
      a[tid * 4 + 0] = o[0] / kGpScaleXY; a[tid * 4 + 1] = o[1] / kGpScaleXY;
      a[tid * 4 + 2] = o[2] / kGpScaleP;  a[tid * 4 + 3] = o[3] / kGpScaleT;
    }
    __syncthreads();
  }
  // thread q < 181: query level q; threads 181, 182: the two target columns (z = L^-1 y)
  double mean_u = 0.0, mean_v = 0.0, deviation = 0.0;
  const bool is_query = tid < kNumLevels;
  const bool is_rhs = tid == kNumLevels || tid == kNumLevels + 1;
  const double level_p = pressure_level(tid < kNumLevels ? tid : 0);
  const bool reachable = is_query && !(level_p < pmin || level_p > pmax);   // unreachable levels are not encoded
  if (m > 0 && (reachable || is_rhs)) {
    const double qa[4] = {x / kGpScaleXY, y / kGpScaleXY, level_p / kGpScaleP, double(t_elapsed) / kGpScaleT};
    double norm2 = 0.0;
    for (int b = 0; b < nb; ++b) {
      double acc[kGpBlock];
#pragma unroll
      for (int r = 0; r < kGpBlock; ++r) {
        const int i = b * kGpBlock + r;
        acc[r] = (i >= m) ? 0.0 : (is_query ? gp_kernel(qa, a + i * 4) : ring[s_idx[i] * 6 + 4 + (tid - kNumLevels)]);
      }
      const double* lb = Lb + lb_offset(b);
      const int jn = b * kGpBlock;
      for (int j = 0; j < jn; ++j) {
        const double vj = double(V[j * kColumnThreads + tid]);
        const double2* col = reinterpret_cast<const double2*>(lb + j * kGpBlock);
#pragma unroll
        for (int r2 = 0; r2 < kGpBlock / 2; ++r2) {
          const double2 l2 = col[r2];
          acc[2 * r2] -= l2.x * vj;
          acc[2 * r2 + 1] -= l2.y * vj;
        }
      }
      // 8 x 8 diagonal block
#pragma unroll
      for (int r = 0; r < kGpBlock; ++r) {
        const double v = acc[r] * inv_diag[jn + r];
#pragma unroll
        for (int r2 = r + 1; r2 < kGpBlock; ++r2) acc[r2] -= lb[(jn + r) * kGpBlock + r2] * v;
        const int i = jn + r;
        V[i * kColumnThreads + tid] = float(v);
        if (is_rhs && i < m) z[i * 2 + (tid - kNumLevels)] = v;
        norm2 += v * v;
      }
    }
    deviation = fmax(kGpSigma2 - norm2, 0.0) / kGpSigma2;                  // wind_gp.py:186-193
  }
  __syncthreads();
  if (reachable) {
    if (m > 0) {
      for (int i = 0; i < m; ++i) {                                        // mean = k*^T K^-1 y = v . z
        const double v = double(V[i * kColumnThreads + tid]);
        mean_u += v * z[i * 2]; mean_v += v * z[i * 2 + 1];
      }
    }
    double fu, fv;
    forecast_at<double, DevState<Real>>(d, e, x, y, level_p, t_elapsed, &fu, &fv);
    wind_level_features(mean_u + fu, mean_v + fv, deviation, x, y, &feat[tid * 3], &feat[tid * 3 + 1], &feat[tid * 3 + 2]);
  }
  __syncthreads();
  // centred, padded column (features.py:479-497, 536-556)
  const int lower = kNumLevels - nearest_pressure_level(p_b) - 1;
  float* o = obs + e * int64_t(kNumFeatures) + 16;
  for (int s = tid; s < 2 * kNumLevels - 1; s += kColumnThreads) {
    float f0 = 0.f, f1 = 1.f, f2 = 1.f;                                    // "unreachable" triple
    const int l = s - lower;
    if (l >= 0 && l < kNumLevels) {
      const double pl = pressure_level(l);
      if (!(pl < pmin || pl > pmax)) { f0 = feat[l * 3]; f1 = feat[l * 3 + 1]; f2 = feat[l * 3 + 2]; }
    }
    o[s * 3] = f0; o[s * 3 + 1] = f1; o[s * 3 + 2] = f2;
  }
}
