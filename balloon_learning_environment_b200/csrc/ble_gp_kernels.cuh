// WindGP posterior for the Perciatelli observation, second generation (env/wind_gp.py:98-241).
//
// The reference refits a GaussianProcessRegressor on the last 6 h of measurements at EVERY step
// (wind_gp.py:172-190): a 120 x 120 Cholesky (O(m^3)) followed by 181 predictive variances (a triangular
// solve per level).  Between two consecutive steps the measurement window only loses its oldest point
// and gains one new point, so here the factor is carried from step to step:
//
//   k_gp_update   CTA (128 threads, thread i = row i) per balloon.  Brings the lower Cholesky factor of
//                 K + alpha I from the window it was last computed for to the current one:
//                   drop oldest point  = rank-1 UPDATE of the trailing factor with its first column
//                                        (K22 = L22 L22^T + l21 l21^T; numerically benign, no downdate),
//                   append new point   = one forward substitution (the new row) + a square root,
//                 both O(m^2); falls back to the full left-looking factorisation when the windows do not
//                 chain (first call, history cleared, irregular use).  Also solves z = L^-1 y for the two
//                 error components.  Everything fp64, in shared memory.
//   k_gp_column2  CTA (15 warps) per balloon: V = L^-1 K*^T for the reachable pressure levels as a blocked
//                 right-looking triangular solve.  Warp b owns the 8-row block b of every column, a lane
//                 owns two columns, accumulators live in registers; the only shared traffic per step is the
//                 8 freshly solved entries of each column.  The owner of block j + 1 solves it inside step j
//                 (look-ahead), so the barrier never waits for a diagonal solve.  The factor arrives with ONE
//                 TMA bulk copy.
//
// Factor layout in HBM ("blocked lower"): 8 x 8 blocks (b, j), j <= b, at ((b (b + 1) / 2 + j) * 64 doubles,
// COLUMN-major inside a block (element (r, c) at c * 8 + r, so that the 8 rows of one column are two
// LDS.128 pairs); rows >= m are identity padding up to the next multiple of 8.
// NOTE: included from inside `namespace ble` of ble_engine.cu after ble_feature_kernels.cuh.
#pragma once

constexpr int kGpBlk = 8;
constexpr int kGpNumBlk = kGpWindow / kGpBlk;                                   // 15
constexpr int kGpBlockedLower = kGpBlk * kGpBlk * (kGpNumBlk * (kGpNumBlk + 1) / 2);   // 7,680 doubles
constexpr int kGpFactorDoubles = kGpBlockedLower;                              // per balloon in d.gp_chol

__device__ __forceinline__ int blk_offset(int b, int j) { return (((b * (b + 1)) >> 1) + j) * (kGpBlk * kGpBlk); }
__device__ __forceinline__ int blocked_index(int i, int j) {
  return blk_offset(i >> 3, j >> 3) + (j & 7) * kGpBlk + (i & 7);
}

// ---------------------------------------------------------------------------------------------------------
// k_gp_update
// ---------------------------------------------------------------------------------------------------------
constexpr int kUpdateThreads = 128;
constexpr size_t kUpdateSmem = sizeof(double) * (kGpBlockedLower + kGpWindow * 4 + kGpWindow * 2 + 8);
#define GP_L(i, j) Lp[blocked_index((i), (j))]

struct GpSweep {            // forward substitution with up to three right-hand sides held by thread i = row i
  double b0, b1, b2;
};

// Column-oriented forward substitution over rows [0, rows): on return thread j < rows holds
// (L^-1 rhs)_j in s.b*.  bc: shared double[2][6] broadcast slots.  One barrier per row.
__device__ __forceinline__ void gp_forward_sweep(const double* __restrict__ Lp, int rows, int tid, GpSweep* s,
                                                 double (*bc)[6]) {
  const double inv = tid < rows ? 1.0 / GP_L(tid, tid) : 0.0;          // off the dependent chain
  for (int j = 0; j < rows; ++j) {
    const int p = j & 1;
    if (tid == j) {
      s->b0 *= inv; s->b1 *= inv; s->b2 *= inv;
      bc[p][0] = s->b0; bc[p][1] = s->b1; bc[p][2] = s->b2;
    }
    __syncthreads();
    if (tid > j && tid < rows) {
      const double l = GP_L(tid, j);
      s->b0 -= l * bc[p][0]; s->b1 -= l * bc[p][1]; s->b2 -= l * bc[p][2];
    }
  }
}

__device__ __forceinline__ double gp_rsqrt(double t) {       // t in [1e-30, 1e30]: float seed + two Newton steps
  double y = double(rsqrtf(float(t)));
  y = y * (1.5 - 0.5 * t * y * y);
  return y * (1.5 - 0.5 * t * y * y);
}

__device__ __forceinline__ double block_sum(double v, double* red /* [4] shared */, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                   // red may still be read from a previous call
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// Row `row` of the shared factor becomes the identity row e_row (padding), all threads.
__device__ __forceinline__ void gp_identity_row(double* Lp, int row, int tid) {
  const int last = (row | 7);                        // last column of the diagonal block
  if (tid <= last) GP_L(row, tid) = tid == row ? 1.0 : 0.0;
}

template <typename Real>
__global__ void __launch_bounds__(kUpdateThreads) k_gp_update(DevState<Real> d) {
  extern __shared__ __align__(16) double s_mem[];
  double* Lp = s_mem;                              // blocked lower, same layout as in HBM
  double* a = Lp + kGpBlockedLower;                // [kGpWindow][4] scaled coordinates of the window's points
  double* y = a + kGpWindow * 4;                   // [kGpWindow][2] measured errors
  double* red = y + kGpWindow * 2;                 // [4] + spare
  __shared__ double bc[2][6];
  __shared__ int s_idx[kGpWindow];
  __shared__ int s_plan[6];                        // m_new, first_new, mode, drops, appends, m_old
  const int64_t e = blockIdx.x;
  const int tid = threadIdx.x;
  const double* ring = d.gp_obs + e * int64_t(kGpWindow * 6);
  double* factor = d.gp_chol + e * int64_t(kGpFactorDoubles);
  double* zout = d.gp_z + e * int64_t(kGpWindow * 2);

  // ---- plan: which measurements are inside the 6 h window now, and how does that chain to the factor ----
  const int count = d.gp_count[e];
  const int stored = count < kGpWindow ? count : kGpWindow;
  const double now = double(d.t_elapsed[e]);
  bool pass = false;                                 // thread k = k-th stored measurement, chronological
  if (tid < stored) pass = fabs(ring[((count - stored + tid) % kGpWindow) * 6 + 3] - now) < kGpHorizonS;
  const int m_par = __syncthreads_count(pass);
  // times only grow, so the window is normally the newest m measurements ("suffix")
  const bool suffix = __syncthreads_and(tid >= stored || pass == (tid >= stored - m_par));
  if (suffix) { if (tid < m_par) s_idx[tid] = (count - m_par + tid) % kGpWindow; }
  if (tid == 0) {
    const int m_new = suffix ? m_par : gp_window_indices(ring, count, now, s_idx);
    const int first_new = count - m_new;
    const int m_old = d.gp_m[e], first_old = d.gp_first[e];
    int mode = 2, drops = 0, appends = m_new;      // 0 unchanged, 1 incremental, 2 full factorisation
    if (suffix && m_old > 0 && first_old >= 0 && first_new >= first_old && first_new < first_old + m_old &&
        first_old + m_old <= count) {
      drops = first_new - first_old;
      appends = count - (first_old + m_old);
      mode = (drops == 0 && appends == 0) ? 0 : 1;
      if (drops + appends > 12) mode = 2;          // many steps since the last call: refactor instead
    }
    s_plan[0] = m_new; s_plan[1] = suffix ? first_new : -1; s_plan[2] = mode; s_plan[3] = drops; s_plan[4] = appends;
    s_plan[5] = m_old;
  }
  __syncthreads();
  const int m_new = s_plan[0], mode = s_plan[2], drops = s_plan[3], appends = s_plan[4], m_old = s_plan[5];
  if (mode == 0) return;                           // factor and z are already those of this window
  if (tid == 0) { d.gp_m[e] = m_new; d.gp_first[e] = s_plan[1]; }
  if (m_new == 0) return;
  const int nb_new = (m_new + 7) >> 3;
  if (mode == 1) {                                 // the old factor, as it lies in HBM (coalesced 16-byte copies)
    const int n2 = blk_offset((m_old + 7) >> 3, 0) >> 1;
    const double2* src = reinterpret_cast<const double2*>(factor);
    double2* dst = reinterpret_cast<double2*>(Lp);
    for (int k = tid; k < n2; k += kUpdateThreads) dst[k] = src[k];
  }
  if (tid < m_new) {
    const double* o = ring + s_idx[tid] * 6;
    a[tid * 4 + 0] = o[0] / kGpScaleXY; a[tid * 4 + 1] = o[1] / kGpScaleXY;
    a[tid * 4 + 2] = o[2] / kGpScaleP;  a[tid * 4 + 3] = o[3] / kGpScaleT;
    y[tid * 2] = o[4]; y[tid * 2 + 1] = o[5];
  }
  __syncthreads();
  int m_cur = 0;
  if (mode == 1 && drops == 1 && appends == 1 && m_old >= 2) {
    // ---- steady state, ONE sweep: iteration k finishes column k - 1 of the up-shifted factor (the rank-1
    // update that removes the oldest point) and at once uses it for step k - 1 of the forward substitutions
    // of the new point's kernel row and of the two error columns.  Thread i = old row i = new row i - 1.
    // In place: the slot written in iteration k was last read in iteration k - 1.
    const bool mine = tid >= 1 && tid < m_old;
    double x = mine ? GP_L(tid, 0) : 0.0;
    const double lkk = mine ? GP_L(tid, tid) : 1.0;
    const double inv_l = 1.0 / lkk;
    const int rows = m_old - 1;                    // size of the factor before the append
    GpSweep sw{0.0, 0.0, 0.0};
    if (mine) { sw.b0 = gp_kernel(a + rows * 4, a + (tid - 1) * 4); sw.b1 = y[(tid - 1) * 2]; sw.b2 = y[(tid - 1) * 2 + 1]; }
    __syncthreads();
    for (int k = 1; k < m_old; ++k) {
      const int p = k & 1;
      if (tid == k) {
        const double s = x * inv_l, t = fma(s, s, 1.0);
        const double inv_c = gp_rsqrt(t), c = t * inv_c;       // c = sqrt(1 + s^2) = r / l_kk
        const double inv_new = inv_l * inv_c;                  // 1 / new diagonal
        GP_L(k - 1, k - 1) = lkk * c;
        sw.b0 *= inv_new; sw.b1 *= inv_new; sw.b2 *= inv_new;
        bc[p][0] = inv_c; bc[p][1] = s; bc[p][2] = c; bc[p][3] = sw.b0; bc[p][4] = sw.b1; bc[p][5] = sw.b2;
      }
      __syncthreads();
      if (tid > k && tid < m_old) {
        const double inv_c = bc[p][0], s = bc[p][1], c = bc[p][2];
        const double lik = (GP_L(tid, k) + s * x) * inv_c;
        x = c * x - s * lik;
        GP_L(tid - 1, k - 1) = lik;
        sw.b0 -= lik * bc[p][3]; sw.b1 -= lik * bc[p][4]; sw.b2 -= lik * bc[p][5];
      }
    }
    const double r = mine ? sw.b0 : 0.0;
    const double rr = block_sum(r * r, red, tid);
    const double ru = block_sum(mine ? r * sw.b1 : 0.0, red, tid);
    const double rv = block_sum(mine ? r * sw.b2 : 0.0, red, tid);
    const double diag = sqrt(kGpSigma2 + kGpNoise - rr);
    if (mine) { GP_L(rows, tid - 1) = r; zout[(tid - 1) * 2] = sw.b1; zout[(tid - 1) * 2 + 1] = sw.b2; }
    if (tid == 0) {
      GP_L(rows, rows) = diag;
      zout[rows * 2] = (y[rows * 2] - ru) / diag; zout[rows * 2 + 1] = (y[rows * 2 + 1] - rv) / diag;
    }
    m_cur = m_new;
  } else if (mode == 1) {
    // ---- general incremental path: drops one by one, then appends one by one ------------------------
    m_cur = m_old;
    for (int dr = 0; dr < drops; ++dr) {
      // drop point 0: L22' L22'^T = L22 L22^T + x x^T with x = L[1:, 0]; thread i keeps x_i, results land one
      // row and one column up
      double x = (tid >= 1 && tid < m_cur) ? GP_L(tid, 0) : 0.0;
      __syncthreads();
      for (int k = 1; k < m_cur; ++k) {
        const int p = k & 1;
        if (tid == k) {
          const double lkk = GP_L(k, k);
          const double r = sqrt(lkk * lkk + x * x);
          bc[p][0] = r / lkk; bc[p][1] = x / lkk;
          GP_L(k - 1, k - 1) = r;
        }
        __syncthreads();
        if (tid > k && tid < m_cur) {
          const double c = bc[p][0], s = bc[p][1];
          const double lik = (GP_L(tid, k) + s * x) / c;
          x = c * x - s * lik;
          GP_L(tid - 1, k - 1) = lik;
        }
      }
      __syncthreads();
      --m_cur;
      gp_identity_row(Lp, m_cur, tid);             // the vacated last row is padding again
      __syncthreads();
    }
    const int n_sweeps = appends > 0 ? appends : 1;
    for (int ap = 0; ap < n_sweeps; ++ap) {
      const bool appending = appends > 0;
      const bool last = ap + 1 == n_sweeps;
      const int rows = m_cur;                      // rows already in the factor
      if (appending && (rows & 7) == 0) {          // the new row opens a new block-row: identity padding first
        for (int rr8 = 0; rr8 < 8; ++rr8) {
          for (int c = tid; c < rows + 8; c += kUpdateThreads) GP_L(rows + rr8, c) = c == rows + rr8 ? 1.0 : 0.0;
        }
        __syncthreads();
      }
      GpSweep sw{0.0, 0.0, 0.0};
      if (tid < rows) {
        if (appending) sw.b0 = gp_kernel(a + rows * 4, a + tid * 4);
        if (last) { sw.b1 = y[tid * 2]; sw.b2 = y[tid * 2 + 1]; }
      }
      gp_forward_sweep(Lp, rows, tid, &sw, bc);
      if (appending) {
        const double r = tid < rows ? sw.b0 : 0.0;
        const double rr = block_sum(r * r, red, tid);
        const double ru = block_sum(last ? r * sw.b1 : 0.0, red, tid);
        const double rv = block_sum(last ? r * sw.b2 : 0.0, red, tid);
        const double diag = sqrt(kGpSigma2 + kGpNoise - rr);
        if (tid < rows) GP_L(rows, tid) = r;
        if (tid == rows) {
          GP_L(rows, rows) = diag;
          if (last) { sw.b1 = (y[rows * 2] - ru) / diag; sw.b2 = (y[rows * 2 + 1] - rv) / diag; }
        }
        ++m_cur;
        __syncthreads();
      }
      if (last && tid < m_cur) { zout[tid * 2] = sw.b1; zout[tid * 2 + 1] = sw.b2; }
    }
  } else {
    // ---- full factorisation (left-looking, thread i owns row i) ------------------------------------
    for (int k = tid; k < blk_offset(nb_new, 0); k += kUpdateThreads) {      // K + alpha I, identity padding
      const int blk = k >> 6, c = (k >> 3) & 7, r = k & 7;
      int b = int((sqrtf(8.f * float(blk) + 1.f) - 1.f) * 0.5f);
      while (((b * (b + 1)) >> 1) > blk) --b;
      while ((((b + 1) * (b + 2)) >> 1) <= blk) ++b;
      const int j = blk - ((b * (b + 1)) >> 1);
      const int row = b * kGpBlk + r, col = j * kGpBlk + c;
      double v = 0.0;
      if (row < m_new) { if (col <= row) v = gp_kernel(a + row * 4, a + col * 4) + (row == col ? kGpNoise : 0.0); }
      else if (col == row) v = 1.0;
      Lp[k] = v;
    }
    __syncthreads();
    for (int j = 0; j < m_new; ++j) {
      double s = 0.0;
      if (tid >= j && tid < m_new) {
        double s0 = 0.0, s1 = 0.0;
        int k = 0;
        for (; k + 1 < j; k += 2) { s0 += GP_L(tid, k) * GP_L(j, k); s1 += GP_L(tid, k + 1) * GP_L(j, k + 1); }
        if (k < j) s0 += GP_L(tid, k) * GP_L(j, k);
        s = GP_L(tid, j) - (s0 + s1);
      }
      __syncthreads();
      if (tid == j) GP_L(j, j) = sqrt(s);
      __syncthreads();
      if (tid > j && tid < m_new) GP_L(tid, j) = s / GP_L(j, j);
      __syncthreads();
    }
    m_cur = m_new;
    GpSweep sw{0.0, tid < m_new ? y[tid * 2] : 0.0, tid < m_new ? y[tid * 2 + 1] : 0.0};
    gp_forward_sweep(Lp, m_new, tid, &sw, bc);
    if (tid < m_new) { zout[tid * 2] = sw.b1; zout[tid * 2 + 1] = sw.b2; }
  }
  __syncthreads();
  {                                                // shared -> HBM, same layout
    const int n2 = blk_offset(nb_new, 0) >> 1;
    const double2* src = reinterpret_cast<const double2*>(Lp);
    double2* dst = reinterpret_cast<double2*>(factor);
    for (int k = tid; k < n2; k += kUpdateThreads) dst[k] = src[k];
  }
}
#undef GP_L

// ---------------------------------------------------------------------------------------------------------
// k_gp_column2
// ---------------------------------------------------------------------------------------------------------
constexpr int kColWarps = kGpNumBlk;                  // 15: warp b owns the 8-row block b of every column
constexpr int kColThreads = 32 * kColWarps;
constexpr int kColPerPass = 64;                       // two columns per lane
struct ColumnSmem {
  double L[kGpBlockedLower];                          // 61,440 B, filled by one TMA bulk copy
  double xbuf[2][kGpBlk][kColPerPass];                // freshly solved 8 entries of every column, double-buffered
  double cxy[kGpWindow];                              // level-independent part of the squared distance
  double pz[kGpWindow];                               // scaled pressure of the measurements
  double z[kGpWindow][2];                             // L^-1 y
  double inv_diag[kGpWindow];                         // 1 / L_ii
  double red[kColWarps][kColPerPass][3];              // per-warp partial |v|^2, v.z_u, v.z_v
  float feat[kNumLevels * 3];
  int act[kNumLevels + 3];
  int n_act;
  unsigned long long bar;
};

// sigma^2 exp(-sqrt(d2)) for d2 in [0, ~1e3] to ~2 ulp, without the range checks of the library calls:
// sqrt by one float-seeded Newton step pair, exp by 2^n * p(r), |r| <= ln2 / 2, degree-11 Taylor/Horner.
__device__ __forceinline__ double gp_kernel_from_d2(double d2) {
  d2 = fmax(d2, 1e-30);                              // keeps the float seed finite; sqrt(1e-30) ~ 0
  double yr = double(rsqrtf(float(d2)));             // ~2^-22 relative
  yr = yr * (1.5 - 0.5 * d2 * yr * yr);              // ~2^-43
  yr = yr * (1.5 - 0.5 * d2 * yr * yr);              // full double
  const double dist = d2 * yr;
  const double t = -dist * 1.4426950408889634;       // log2(e)
  const double n = rint(t);
  const double r = fma(n, -1.9082149292705877e-10, fma(n, -0.6931471803691238, -dist));   // -dist - n ln2 (hi + lo)
  double p = 2.505210838544172e-08;                  // 1/11!
  p = fma(p, r, 2.755731922398589e-07);
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

// acc[r] -= sum_c L[r][c] v[c] for one 8 x 8 block (column-major in shared memory) and two columns.
__device__ __forceinline__ void gp_block_update(double (&a0)[kGpBlk], double (&a1)[kGpBlk], const double* __restrict__ Lb,
                                                const double* __restrict__ xb, int lane) {
#pragma unroll
  for (int c = 0; c < kGpBlk; ++c) {
    const double v0 = xb[c * kColPerPass + lane], v1 = xb[c * kColPerPass + lane + 32];
    const double2* col = reinterpret_cast<const double2*>(Lb + c * kGpBlk);
#pragma unroll
    for (int h = 0; h < kGpBlk / 2; ++h) {
      const double2 l = col[h];
      a0[2 * h] -= l.x * v0; a0[2 * h + 1] -= l.y * v0;
      a1[2 * h] -= l.x * v1; a1[2 * h + 1] -= l.y * v1;
    }
  }
}

// 8 x 8 lower-triangular solve of one column's block; the 8 results go to xout[r * kColPerPass].
__device__ __forceinline__ void gp_diag_solve(double (&ac)[kGpBlk], const double* __restrict__ Ld,
                                              const double* __restrict__ inv, const double* __restrict__ zj,
                                              double* __restrict__ xout, double* n2, double* mu, double* mv) {
#pragma unroll
  for (int r = 0; r < kGpBlk; ++r) {
    const double v = ac[r] * inv[r];
#pragma unroll
    for (int r2 = r + 1; r2 < kGpBlk; ++r2) ac[r2] -= Ld[r * kGpBlk + r2] * v;
    xout[r * kColPerPass] = v;
    *n2 += v * v;
    *mu += v * zj[r * 2];
    *mv += v * zj[r * 2 + 1];
  }
}

template <typename Real>
__global__ void __launch_bounds__(kColThreads, 1) k_gp_column2(DevState<Real> d, float* __restrict__ obs) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  ColumnSmem& S = *reinterpret_cast<ColumnSmem*>(s_raw);
  __shared__ int s_idx[kGpWindow];
  const int64_t e = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = d.gp_m[e];
  const int nb = (m + kGpBlk - 1) / kGpBlk;
  const double* ring = d.gp_obs + e * int64_t(kGpWindow * 6);
  const double x = DD(d, D_X, e), y = DD(d, D_Y, e), p_b = DD(d, D_P, e);
  const int32_t t_elapsed = d.t_elapsed[e];
  const double pmin = d.feat_range[2 * e], pmax = d.feat_range[2 * e + 1];
  const uint32_t bar = smem_u32(&S.bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    S.n_act = 0; S.act[kNumLevels] = kNumLevels; S.act[kNumLevels + 1] = -1;
  }
  __syncthreads();
  if (tid == 0 && m > 0) {
    const uint32_t bytes = uint32_t(blk_offset(nb, 0)) * 8u;
    const double* src = d.gp_chol + e * int64_t(kGpFactorDoubles);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(S.L)), "l"(src), "r"(bytes), "r"(bar) : "memory");
  }
  // while the factor is in flight: window slots and the (contiguous) range of reachable levels
  if (tid < kNumLevels) {
    const double pl = pressure_level(tid);
    if (!(pl < pmin || pl > pmax)) { atomicMin(&S.act[kNumLevels], tid); atomicMax(&S.act[kNumLevels + 1], tid); }
  }
  const int first_abs = d.gp_first[e];
  if (m > 0) {
    if (first_abs >= 0) { if (tid < m) s_idx[tid] = (first_abs + tid) % kGpWindow; }
    else if (tid == 0) gp_window_indices(ring, d.gp_count[e], double(t_elapsed), s_idx);   // irregular history
  }
  __syncthreads();
  {
    const int lo = S.act[kNumLevels], hi = S.act[kNumLevels + 1];
    const int n = hi >= lo ? hi - lo + 1 : 0;
    if (tid < n) S.act[tid] = lo + tid;
    if (tid == 0) S.n_act = n;
  }
  const double qx = x / kGpScaleXY, qy = y / kGpScaleXY, qt = double(t_elapsed) / kGpScaleT;
  if (tid < nb * kGpBlk) {
    double c = 0.0, pz = 0.0, zu = 0.0, zv = 0.0;
    if (tid < m) {
      const double* o = ring + s_idx[tid] * 6;
      const double dx = qx - o[0] / kGpScaleXY, dy = qy - o[1] / kGpScaleXY, dt = qt - o[3] / kGpScaleT;
      c = dx * dx + dy * dy + dt * dt;
      pz = o[2] / kGpScaleP;
      const double* zz = d.gp_z + e * int64_t(kGpWindow * 2) + tid * 2;
      zu = zz[0]; zv = zz[1];
    }
    S.cxy[tid] = c; S.pz[tid] = pz; S.z[tid][0] = zu; S.z[tid][1] = zv;
  }
  __syncthreads();
  const int n_act = S.n_act;
  const int b0 = warp;                               // the block this warp owns
  bool factor_ready = (m == 0);

  for (int first = 0; first < n_act && m > 0; first += kColPerPass) {
    const int c0 = first + lane, c1 = first + lane + 32;
    const bool on0 = c0 < n_act, on1 = c1 < n_act;
    const double pq0 = pressure_level(on0 ? S.act[c0] : 0) / kGpScaleP;
    const double pq1 = pressure_level(on1 ? S.act[c1] : 0) / kGpScaleP;
    double a0[kGpBlk], a1[kGpBlk];                   // rows of the owned block, two columns
#pragma unroll
    for (int r = 0; r < kGpBlk; ++r) {
      const int i0 = b0 * kGpBlk + r;
      a0[r] = a1[r] = 0.0;
      if (b0 < nb && i0 < m) {
        const double cx = S.cxy[i0], pi = S.pz[i0];
        if (on0) a0[r] = gp_kernel_from_d2(cx + (pq0 - pi) * (pq0 - pi));
        if (on1) a1[r] = gp_kernel_from_d2(cx + (pq1 - pi) * (pq1 - pi));
      }
    }
    if (!factor_ready) {                             // wait for the TMA transaction (phase 0), once
      asm volatile(
          "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
          "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar) : "memory");
      factor_ready = true;
      if (tid < nb * kGpBlk) S.inv_diag[tid] = 1.0 / S.L[blocked_index(tid, tid)];
      __syncthreads();
    }
    double n2[2] = {0.0, 0.0}, mu[2] = {0.0, 0.0}, mv[2] = {0.0, 0.0};
    // block 0 is solved up front; afterwards the owner of block j + 1 updates and solves it INSIDE step j
    // (look-ahead), so that nobody waits at the barrier for a diagonal solve
    if (warp == 0) {
      gp_diag_solve(a0, S.L + blk_offset(0, 0), S.inv_diag, &S.z[0][0], &S.xbuf[0][0][lane], &n2[0], &mu[0], &mv[0]);
      gp_diag_solve(a1, S.L + blk_offset(0, 0), S.inv_diag, &S.z[0][0], &S.xbuf[0][0][lane + 32], &n2[1], &mu[1], &mv[1]);
    }
    __syncthreads();
    for (int j = 0; j + 1 < nb; ++j) {
      if (b0 > j && b0 < nb) {
        gp_block_update(a0, a1, S.L + blk_offset(b0, j), &S.xbuf[j & 1][0][0], lane);
        if (b0 == j + 1) {
          const double* Ld = S.L + blk_offset(b0, b0);
          double* xo = &S.xbuf[b0 & 1][0][0];
          gp_diag_solve(a0, Ld, S.inv_diag + b0 * kGpBlk, &S.z[b0 * kGpBlk][0], xo + lane, &n2[0], &mu[0], &mv[0]);
          gp_diag_solve(a1, Ld, S.inv_diag + b0 * kGpBlk, &S.z[b0 * kGpBlk][0], xo + lane + 32, &n2[1], &mu[1], &mv[1]);
        }
      }
      __syncthreads();
    }
    // reduce the per-warp partials of every column, then the features of this pass's levels
#pragma unroll
    for (int col = 0; col < 2; ++col) {
      S.red[warp][lane + 32 * col][0] = n2[col]; S.red[warp][lane + 32 * col][1] = mu[col]; S.red[warp][lane + 32 * col][2] = mv[col];
    }
    __syncthreads();
    if (tid < kColPerPass && first + tid < n_act) {
      double norm2 = 0.0, mean_u = 0.0, mean_v = 0.0;
#pragma unroll
      for (int w = 0; w < kColWarps; ++w) { norm2 += S.red[w][tid][0]; mean_u += S.red[w][tid][1]; mean_v += S.red[w][tid][2]; }
      const int l = S.act[first + tid];
      const double deviation = fmax(kGpSigma2 - norm2, 0.0) / kGpSigma2;           // wind_gp.py:186-193
      double fu, fv;
      forecast_at<double, DevState<Real>>(d, e, x, y, pressure_level(l), t_elapsed, &fu, &fv);
      wind_level_features(mean_u + fu, mean_v + fv, deviation, x, y, &S.feat[l * 3], &S.feat[l * 3 + 1], &S.feat[l * 3 + 2]);
    }
    __syncthreads();
  }
  if (m == 0) {                                       // no measurement yet: zero mean and deviation (wind_gp.py:161-163)
    for (int k = tid; k < n_act; k += kColThreads) {
      const int l = S.act[k];
      double fu, fv;
      forecast_at<double, DevState<Real>>(d, e, x, y, pressure_level(l), t_elapsed, &fu, &fv);
      wind_level_features(fu, fv, 0.0, x, y, &S.feat[l * 3], &S.feat[l * 3 + 1], &S.feat[l * 3 + 2]);
    }
    __syncthreads();
  }
  // centred, padded column (features.py:479-497, 536-556)
  const int lower = kNumLevels - nearest_pressure_level(p_b) - 1;
  float* o = obs + e * int64_t(kNumFeatures) + 16;
  for (int s = tid; s < 2 * kNumLevels - 1; s += kColThreads) {
    float f0 = 0.f, f1 = 1.f, f2 = 1.f;                                            // "unreachable" triple
    const int l = s - lower;
    if (l >= 0 && l < kNumLevels) {
      const double pl = pressure_level(l);
      if (!(pl < pmin || pl > pmax)) { f0 = S.feat[l * 3]; f1 = S.feat[l * 3 + 1]; f2 = S.feat[l * 3 + 2]; }
    }
    o[s * 3] = f0; o[s * 3 + 1] = f1; o[s * 3 + 2] = f2;
  }
}
