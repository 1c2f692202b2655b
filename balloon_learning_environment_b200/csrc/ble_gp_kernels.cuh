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
//   k_gp_column4  CTA (8 warps) per balloon: V = L^-1 K*^T for the reachable pressure levels as a blocked
//                 right-looking triangular solve on the fp64 tensor cores, one warp per 8 columns (see the
//                 kernel's own header).  The factor arrives with ONE TMA bulk copy.
//
// Factor layout in HBM ("blocked lower"): 8 x 8 blocks (b, j), j <= b, at ((b (b + 1) / 2 + j) * 64 doubles,
// inside a block the elements are in mma A-fragment order (blk_inner); rows >= m are identity padding up to the
// next multiple of 8.
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
      const int blk = k >> 6, r = (k >> 2) & 7, c = ((k >> 5) & 1) * 4 + (k & 3);   // inverse of blk_inner
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


// ---------------------------------------------------------------------------------------------------------
// k_gp_column4: one warp per 8 columns, no dependency between warps
// ---------------------------------------------------------------------------------------------------------
// The columns of V = L^-1 K*^T are independent, and with DMMA an 8-column tile is exactly one n-tile: a warp
// keeps the whole 120 x 8 tile of its columns in registers (15 C fragments = 30 doubles per lane) and runs the
// complete blocked substitution alone --
//   step j:  V_j = inv(L_jj) C_j            2 DMMA (the C -> B re-layout goes through a 768-byte private tile)
//            C_b -= L_bj V_j,  b > j        2 DMMA per block, all independent
// -- so the sweep needs no CTA barrier, no published tiles and is perfectly balanced; the dependent chain per step
// is two DMMA pairs.  (Two earlier layouts -- row blocks owned by warps with V_j published through shared memory,
// first with DFMA, then with DMMA -- spent a third of their time at the per-step barrier behind the one warp that
// solved the diagonal block: 14.7 and 11.7 ms per 65,536 balloons against 9.4 ms here.)  DMMA runs at the DFMA
// rate on B200 (37 TFLOP/s, scripts/probes/dmma_probe.cu) but one instruction does the work of eight DFMA warp
// instructions and its operands are 256-byte conflict-free shared-memory fragments.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int kC4Warps = 8;
constexpr int kC4Threads = 32 * kC4Warps;
constexpr int kC4Stage = 12;                           // doubles per row of the private re-layout tile (conflict-free B loads)
struct Column4Smem {
  double L[kGpBlockedLower];                           // filled by one TMA bulk copy; the diagonal blocks are then
                                                       // replaced IN PLACE by their inverses (the sweep never needs L_jj)
  double stage[kC4Warps][kGpBlk * kC4Stage];
  double cxy[kGpWindow], pz[kGpWindow];
  double z[kGpWindow][2];
  float feat[kNumLevels * 3];
  int lo, hi;
  unsigned long long bar;
};

// C-fragment tile (row g, columns 2 tq, 2 tq + 1) -> the two B fragments (k = 4 h + tq, n = g) of the same 8 x 8 matrix
__device__ __forceinline__ void c4_c_to_b(double c0, double c1, double* __restrict__ st, int g, int tq, double* b0, double* b1) {
  *reinterpret_cast<double2*>(st + g * kC4Stage + 2 * tq) = make_double2(c0, c1);
  __syncwarp();
  *b0 = st[tq * kC4Stage + g];
  *b1 = st[(4 + tq) * kC4Stage + g];
  __syncwarp();
}

// Compile-time recursion over the pivot block J and the updated block B: every index of the register tile is a
// constant, so the 30 accumulators stay in registers (a `#pragma unroll` double loop was left partly rolled by
// nvcc, which moved the tile to local memory).
template <int J, int B>
struct C4Update {
  static __device__ __forceinline__ void run(double (&c)[kGpNumBlk][2], const double* __restrict__ L, int nb, int lane,
                                             double b0, double b1) {
    if (B < nb) {
      // the two k halves back to back on the same accumulator: measured faster on B200 than issuing all first
      // halves and then all second halves (10.5 vs 12.0 ms per 65,536 balloons)
      const double* Lb = L + (((B * (B + 1)) >> 1) + J) * (kGpBlk * kGpBlk);
      dmma884(c[B][0], c[B][1], -Lb[lane], b0);
      dmma884(c[B][0], c[B][1], -Lb[32 + lane], b1);
    }
    C4Update<J, B + 1>::run(c, L, nb, lane, b0, b1);
  }
};
template <int J>
struct C4Update<J, kGpNumBlk> {
  static __device__ __forceinline__ void run(double (&)[kGpNumBlk][2], const double*, int, int, double, double) {}
};
template <int J>
struct C4Sweep {
  static __device__ __forceinline__ void run(double (&c)[kGpNumBlk][2], const double* __restrict__ L,
                                             const double* __restrict__ z,
                                             double* __restrict__ st, int nb, int lane, int g, int tq,
                                             double (&n2)[2], double (&mu)[2], double (&mv)[2]) {
    if (J < nb) {
      double b0, b1, v0 = 0.0, v1 = 0.0;
      c4_c_to_b(c[J][0], c[J][1], st, g, tq, &b0, &b1);
      const double* Li = L + (((J * (J + 1)) >> 1) + J) * (kGpBlk * kGpBlk);   // inv(L_JJ), stored over L_JJ
      dmma884(v0, v1, Li[lane], b0);                                     // V_J = inv(L_JJ) C_J
      dmma884(v0, v1, Li[32 + lane], b1);
      const double zu = z[(J * kGpBlk + g) * 2], zv = z[(J * kGpBlk + g) * 2 + 1];
      n2[0] += v0 * v0; n2[1] += v1 * v1;
      mu[0] += v0 * zu; mu[1] += v1 * zu;
      mv[0] += v0 * zv; mv[1] += v1 * zv;
      if (J + 1 < nb) {
        c4_c_to_b(v0, v1, st, g, tq, &b0, &b1);
        C4Update<J, J + 1>::run(c, L, nb, lane, b0, b1);
      }
      C4Sweep<J + 1>::run(c, L, z, st, nb, lane, g, tq, n2, mu, mv);
    }
  }
};
template <>
struct C4Sweep<kGpNumBlk> {
  static __device__ __forceinline__ void run(double (&)[kGpNumBlk][2], const double*, const double*, double*,
                                             int, int, int, int, double (&)[2], double (&)[2], double (&)[2]) {}
};

template <typename Real>
__global__ void __launch_bounds__(kC4Threads, 2) k_gp_column4(DevState<Real> d, float* __restrict__ obs) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  Column4Smem& S = *reinterpret_cast<Column4Smem*>(s_raw);
  __shared__ int s_idx[kGpWindow];
  const int64_t e = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
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
    S.lo = kNumLevels; S.hi = -1;
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
    if (!(pl < pmin || pl > pmax)) { atomicMin(&S.lo, tid); atomicMax(&S.hi, tid); }
  }
  const int first_abs = d.gp_first[e];
  if (m > 0) {
    if (first_abs >= 0) { if (tid < m) s_idx[tid] = (first_abs + tid) % kGpWindow; }
    else if (tid == 0) gp_window_indices(ring, d.gp_count[e], double(t_elapsed), s_idx);   // irregular history
  }
  __syncthreads();
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
  const int lo = S.lo, hi = S.hi;
  const int n_act = hi >= lo ? hi - lo + 1 : 0;          // reachable levels lo .. hi
  if (m > 0) {
    asm volatile(                                     // wait for the TMA transaction (phase 0)
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar) : "memory");
    double xv[kGpBlk];
    const int jd = tid >> 3, cd = tid & 7;
    double* Ld = S.L + blk_offset(jd < nb ? jd : 0, jd < nb ? jd : 0);
    if (tid < nb * kGpBlk) {                          // column cd of inv(L_jj): forward substitution on e_cd
#pragma unroll
      for (int r = 0; r < kGpBlk; ++r) {
        double acc = r == cd ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < r; ++k) acc -= Ld[blk_inner(r, k)] * xv[k];
        xv[r] = acc / Ld[blk_inner(r, r)];
      }
    }
    __syncthreads();                                  // every column of every diagonal block has been read
    if (tid < nb * kGpBlk) {
#pragma unroll
      for (int r = 0; r < kGpBlk; ++r) Ld[blk_inner(r, cd)] = xv[r];
    }
  }
  __syncthreads();

  double* st = S.stage[warp];
  for (int tile = warp; tile * 8 < n_act && m > 0; tile += kC4Warps) {
    const int col0 = tile * 8 + 2 * tq;                  // this lane's two columns (C-fragment layout)
    const bool on0 = col0 < n_act, on1 = col0 + 1 < n_act;
    const double pq0 = pressure_level(lo + (on0 ? col0 : 0)) / kGpScaleP;
    const double pq1 = pressure_level(lo + (on1 ? col0 + 1 : 0)) / kGpScaleP;
    double c[kGpNumBlk][2];
#pragma unroll
    for (int b = 0; b < kGpNumBlk; ++b) {
      const int i = b * kGpBlk + g;
      c[b][0] = 0.0; c[b][1] = 0.0;
      if (b < nb && i < m) {
        const double cx = S.cxy[i], pi = S.pz[i];
        if (on0) c[b][0] = gp_kernel_from_d2(cx + (pq0 - pi) * (pq0 - pi));
        if (on1) c[b][1] = gp_kernel_from_d2(cx + (pq1 - pi) * (pq1 - pi));
      }
    }
    double n2[2] = {0.0, 0.0}, mu[2] = {0.0, 0.0}, mv[2] = {0.0, 0.0};
    C4Sweep<0>::run(c, S.L, &S.z[0][0], st, nb, lane, g, tq, n2, mu, mv);
    // sum over the 8 rows of the fragment (lanes with the same tq), then lane l < 8 finishes column tile * 8 + l
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        n2[i] += __shfl_xor_sync(0xffffffffu, n2[i], o);
        mu[i] += __shfl_xor_sync(0xffffffffu, mu[i], o);
        mv[i] += __shfl_xor_sync(0xffffffffu, mv[i], o);
      }
    }
    const int src = (lane >> 1) & 3;
    const double a0 = __shfl_sync(0xffffffffu, n2[0], src), a1 = __shfl_sync(0xffffffffu, n2[1], src);
    const double u0 = __shfl_sync(0xffffffffu, mu[0], src), u1 = __shfl_sync(0xffffffffu, mu[1], src);
    const double w0 = __shfl_sync(0xffffffffu, mv[0], src), w1 = __shfl_sync(0xffffffffu, mv[1], src);
    const int col = tile * 8 + lane;
    if (lane < 8 && col < n_act) {
      const double norm2 = (lane & 1) ? a1 : a0, mean_u = (lane & 1) ? u1 : u0, mean_v = (lane & 1) ? w1 : w0;
      const int l = lo + col;
      const double deviation = fmax(kGpSigma2 - norm2, 0.0) / kGpSigma2;             // wind_gp.py:186-193
      double fu, fv;
      forecast_at<double, DevState<Real>>(d, e, x, y, pressure_level(l), t_elapsed, &fu, &fv);
      wind_level_features(mean_u + fu, mean_v + fv, deviation, x, y, &S.feat[l * 3], &S.feat[l * 3 + 1], &S.feat[l * 3 + 2]);
    }
  }
  if (m == 0) {                                       // no measurement yet: zero mean and deviation (wind_gp.py:161-163)
    for (int k = tid; k < n_act; k += kC4Threads) {
      const int l = lo + k;
      double fu, fv;
      forecast_at<double, DevState<Real>>(d, e, x, y, pressure_level(l), t_elapsed, &fu, &fv);
      wind_level_features(fu, fv, 0.0, x, y, &S.feat[l * 3], &S.feat[l * 3 + 1], &S.feat[l * 3 + 2]);
    }
  }
  __syncthreads();
  // centred, padded column (features.py:479-497, 536-556)
  const int lower = kNumLevels - nearest_pressure_level(p_b) - 1;
  float* o = obs + e * int64_t(kNumFeatures) + 16;
  for (int s = tid; s < 2 * kNumLevels - 1; s += kC4Threads) {
    float f0 = 0.f, f1 = 1.f, f2 = 1.f;                                            // "unreachable" triple
    const int l = s - lower;
    if (l >= 0 && l < kNumLevels) {
      const double pl = pressure_level(l);
      if (!(pl < pmin || pl > pmax)) { f0 = S.feat[l * 3]; f1 = S.feat[l * 3 + 1]; f2 = S.feat[l * 3 + 2]; }
    }
    o[s * 3] = f0; o[s * 3 + 1] = f1; o[s * 3 + 2] = f2;
  }
}
