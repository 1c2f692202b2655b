// WindGP posterior for the Perciatelli observation, third generation (env/wind_gp.py:98-241).
//
// NOTE: included from inside `namespace ble` of ble_engine.cu after ble_gp_kernels.cuh (blk_offset, blk_inner,
// blocked_index, dmma884, smem_u32, forecast_at, wind_level_features ...).
//
// The reference refits sklearn's GaussianProcessRegressor on the measurements of the last 6 h at EVERY step
// (wind_gp.py:172-190: K = kernel(X) + alpha I, Cholesky, alpha = K^-1 y) and asks it for the posterior mean and
// standard deviation at the 181 pressure levels (:143-241).  This file does exactly that -- a full refit per call,
// nothing carried from step to step but the kernel matrix itself -- in one kernel per call:
//
//   k_feat_observe_k   one warp per balloon.  WindGP.observe (:98-119): appends the measurement to the 120-slot ring
//                      and writes ITS row / column of K + alpha I (120 kernel evaluations) into the balloon's kernel
//                      matrix in HBM.  K is kept in RING-SLOT order (the posterior does not depend on the order of the
//                      points), so nothing is ever shifted: 1 KB written per step instead of a 61 KB factor rewritten.
//   k_gp_posterior     one CTA (8 warps) per balloon.
//        1. K arrives with ONE TMA bulk copy (<= 61,440 B) while the query-dependent terms are prepared; slots that are
//           empty or older than 6 h become identity rows.
//        2. Blocked right-looking Cholesky, fp64, 8 x 8 blocks, the Schur updates on the fp64 tensor cores
//           (mma.sync.m8n8k4.f64, 2 DMMA per block pair); the diagonal block is factored and inverted by warp 0 while
//           the other warps are still in the previous trailing update (look-ahead), so a pivot costs two CTA barriers.
//           z = L^-1 y rides along.  The diagonal blocks are left INVERTED in place.
//        3. The factor is demoted in place to TF32 pairs {hi, lo} (8 bytes per entry, like the double it replaces).
//        4. V = L^-1 K*^T for the reachable levels as a blocked substitution on the tensor cores in split precision
//           (3 x mma.sync.m16n8k8.tf32 per tile: hi*hi + lo*hi + hi*lo, fp32 accumulate): one warp owns 8 levels, its
//           128 x 8 tile lives in registers (32 accumulators), the C -> B fragment re-layout is four shuffles.
//           Conditioning study (scripts/gp_precision_study.py): with the factor and z in fp64 and only this solve in
//           3 x TF32 the variance feature moves by <= 1e-6 and the mean features by <= 1e-5 (tolerance 1e-4).
//        5. deviation = (sigma^2 - |v|^2) / sigma^2, mean = v . z + forecast, the three features per level, the
//           centred / padded 361-level column (features.py:457-556).
// Measured on B200 (65,536 balloons, full window): see DESIGN.md section 4.
#pragma once

// ---- observe + kernel matrix row --------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(128) k_feat_observe_k(DevState<Real> d) {
  const int64_t e = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = int(threadIdx.x & 31u);
  if (e >= d.n) return;
  const int32_t cnt = d.gp_count[e];
  const double t = double(d.t_elapsed[e]);
  double* ring = d.gp_obs + e * int64_t(kGpWindow * 6);
  if (cnt > 0 && ring[((cnt - 1) % kGpWindow) * 6 + 3] == t) return;      // this state is already in the history
  const int s = cnt % kGpWindow;
  const double x = DD(d, D_X, e), y = DD(d, D_Y, e), p = DD(d, D_P, e);
  __syncwarp();                                                           // everybody has read the old slot s
  if (lane == 0) {
    // measurement - forecast (wind_gp.py:112-116) == the simplex noise at this point
    Real nu = Real(0), nv = Real(0);
    if (d.enable_noise) noise_at<Real>(d, e, &nu, &nv);
    double* slot = ring + s * 6;
    slot[0] = x; slot[1] = y; slot[2] = p; slot[3] = t; slot[4] = double(nu); slot[5] = double(nv);
    d.gp_count[e] = cnt + 1;
  }
  // row / column s of K + alpha I against every stored slot (the slot being replaced included: it is this point now)
  const int stored = cnt + 1 < kGpWindow ? cnt + 1 : kGpWindow;
  double* K = d.gp_chol + e * int64_t(kGpFactorDoubles);
  const double a[4] = {x / kGpScaleXY, y / kGpScaleXY, p / kGpScaleP, t / kGpScaleT};
  for (int j = lane; j < stored; j += 32) {
    double v;
    if (j == s) {
      v = kGpSigma2 + kGpNoise;
    } else {
      const double* o = ring + j * 6;
      const double b[4] = {o[0] / kGpScaleXY, o[1] / kGpScaleXY, o[2] / kGpScaleP, o[3] / kGpScaleT};
      v = gp_kernel(a, b);
    }
    K[j <= s ? blocked_index(s, j) : blocked_index(j, s)] = v;
  }
}

// ---- posterior -------------------------------------------------------------------------------------------------------
constexpr int kGpPWarps = 8;
constexpr int kGpPThreads = 32 * kGpPWarps;
constexpr int kGpTiles = (kGpWindow + 15) / 16;            // 8 row tiles of 16 for mma.m16n8k8

struct PosteriorSmem {
  double L[kGpBlockedLower];          // K -> L (fp64) -> {hi, lo} TF32 pairs; diagonal blocks hold inv(L_jj)
  float4 col[kPC * 8];                // the balloon's forecast COLUMN: the 9 lookup windows (one per pressure cell) of its
                                      // (x, y, t) cell, staged by ONE TMA tensor copy (box 128 B x 9 of the 5-D bank view);
                                      // 128-byte aligned (L is 61,440 B)
  double pz[kGpWindow + 16];         // measurement pressures / scale, padded like cxf
  float cxf[kGpWindow + 16];         // (x, y, t) part of the squared distance to the query, fp32; index = row + 8.  Rows that
                                     // are invalid, beyond the window, or the phantom row of an odd block count hold 3e38, so that
                                     // their K* entry evaluates to exactly 0 without a branch
  double yz[kGpWindow][2];            // errors y, overwritten by z = L^-1 y
  float feat[kNumLevels * 3];
  unsigned char valid[kGpWindow];
  FieldCell<double> cell;             // the balloon's (x, y, t) cell and weights (the pressure axis varies per level)
  int lo, hi, n_invalid;
  unsigned long long bar;
};

__device__ __forceinline__ uint32_t tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Inverse of blk_inner / blk_offset: flat index k of the blocked-lower layout -> (row, column).
__device__ __forceinline__ void blocked_coords(int k, int* row, int* col) {
  const int blk = k >> 6, r = (k >> 2) & 7, c = ((k >> 5) & 1) * 4 + (k & 3);
  int b = int((sqrtf(8.f * float(blk) + 1.f) - 1.f) * 0.5f);
  while (((b * (b + 1)) >> 1) > blk) --b;
  while ((((b + 1) * (b + 2)) >> 1) <= blk) ++b;
  const int j = blk - ((b * (b + 1)) >> 1);
  *row = b * kGpBlk + r; *col = j * kGpBlk + c;
}

// inv(L_jj) of the 8 x 8 diagonal block and z_j = inv(L_jj) y'_j, by one warp, for LATENCY: this is the serial part of
// the factorisation (15 of them per balloon, each waiting for the previous pivot; the first version -- one row per lane,
// a separate forward substitution for the inverse, ~1,100 dependent instructions -- held the other seven warps at the
// barrier for 27 % of the kernel, profiles/r02_k_gp_posterior_source_report.txt).
// The block arrives in the fp64 MMA's C-fragment layout (lane 4 g + tq holds (g, 2 tq), (g, 2 tq + 1)), i.e. straight
// from the DMMA that applied the last trailing update, and never goes through shared memory.  Right-looking Cholesky
// with the forward elimination of [I | y'] riding along: at pivot k every lane fetches a_kk, its row's a_gk and its two
// columns' a_ck by four independent shuffles (column k lives in component k & 1 of the lanes with tq == k >> 1), scales
// by 1 / l_kk (fp32 seed + one fp64 Newton step, 2^-44), updates its two entries of the trailing block, and applies the
// same row operation to its two entries of W (-> inv(L_jj)) and, for tq < 2, to y'.  Eight pivots x ~100 cycles.
// Entries above the diagonal are never read (they carry garbage).  On return blk holds inv(L_jj) (zeros above the
// diagonal) in A-fragment order and yz[0..8) holds z_j.
__device__ __forceinline__ void chol_diag_block(double a0, double a1, double* __restrict__ blk, double (*__restrict__ yz)[2], int lane) {
  constexpr unsigned kFull = 0xffffffffu;
  const int g = lane >> 2, tq = lane & 3, quad = lane & ~3;
  const int c0 = 2 * tq, c1 = c0 + 1;
  double w0 = g == c0 ? 1.0 : 0.0, w1 = g == c1 ? 1.0 : 0.0;
  double yv = tq < 2 ? yz[g][tq] : 0.0;
#pragma unroll
  for (int k = 0; k < kGpBlk; ++k) {
    const double v = (k & 1) ? a1 : a0;
    const int ks = k >> 1;
    const double dk = __shfl_sync(kFull, v, 4 * k + ks);               // a_kk
    const double agk = __shfl_sync(kFull, v, quad | ks);               // a_gk, own row
    const double ac0 = __shfl_sync(kFull, v, 4 * c0 + ks);             // a_ck for the two columns this lane updates
    const double ac1 = __shfl_sync(kFull, v, 4 * c1 + ks);
    const double wk0 = __shfl_sync(kFull, w0, 4 * k + tq), wk1 = __shfl_sync(kFull, w1, 4 * k + tq);   // row k of W
    const double yk = __shfl_sync(kFull, yv, 4 * k + tq);
    // In exact arithmetic every pivot of K + alpha I is >= alpha (a Schur complement of it).  The fp32 seed is clamped
    // there (one FMNMX; an fp64 clamp on dk itself cost 0.5 ms per 65,536 balloons on this dependent chain), so that a
    // rounding accident yields a finite factor instead of a NaN that would fill the whole observation.
    double inv = double(rsqrtf(fmaxf(float(dk), float(kGpNoise))));
    inv = inv * fma(-0.5 * dk * inv, inv, 1.5);                         // 1 / l_kk
    const double lgk = agk * inv;
    if (c0 > k) a0 = fma(-lgk, ac0 * inv, a0);
    if (c1 > k) a1 = fma(-lgk, ac1 * inv, a1);
    const double s0 = wk0 * inv, s1 = wk1 * inv, sy = yk * inv;         // row k of [W | y'] / l_kk
    if (g == k) { w0 = s0; w1 = s1; yv = sy; }
    else if (g > k) { w0 = fma(-lgk, s0, w0); w1 = fma(-lgk, s1, w1); yv = fma(-lgk, sy, yv); }
  }
  *reinterpret_cast<double2*>(blk + 4 * g + 2 * (tq & 1) + 32 * (tq >> 1)) = make_double2(w0, w1);
  if (tq < 2) yz[g][tq] = yv;
  __syncwarp();
}

// Column sweep: compile-time recursion over the pivot block J and the updated row tile T, so that every index into the
// register tile is a constant.
struct SweepCtx {
  const float2* L2;          // the factor as {hi, lo} pairs, blocked layout (off-diagonal blocks negated)
  const double (*yz)[2];
  int nb, lane, g, tq;
};

// Row tile T holds block rows (2 T - SHIFT, 2 T + 1 - SHIFT), SHIFT = nb & 1 (see the demotion step of the kernel).
template <int J, int T, int SHIFT>
struct SweepUpdate {
  static __device__ __forceinline__ void run(float (&c)[kGpTiles][4], const SweepCtx& s, uint32_t bh0, uint32_t bh1, uint32_t bl0,
                                             uint32_t bl1) {
    constexpr int b_top = 2 * T - SHIFT, b_bot = b_top + 1;
    if (b_bot < s.nb) {                                                   // (tiles are visited in order: nothing below either)
      if (b_top > J) {                                                    // full tile, fragment-major: two LDS.128
        const float4 hi = reinterpret_cast<const float4*>(s.L2 + blk_offset(b_top, J))[s.lane];
        const float4 lo = reinterpret_cast<const float4*>(s.L2 + blk_offset(b_bot, J))[s.lane];
        mma_tf32(c[T], __float_as_uint(hi.x), __float_as_uint(hi.y), __float_as_uint(hi.z), __float_as_uint(hi.w), bh0, bh1);
        mma_tf32(c[T], __float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w), bh0, bh1);
        mma_tf32(c[T], __float_as_uint(hi.x), __float_as_uint(hi.y), __float_as_uint(hi.z), __float_as_uint(hi.w), bl0, bl1);
      } else {                                                            // top half is the pivot block itself: bottom only
        const float2* blk = s.L2 + blk_offset(b_bot, J);
        const float2 a1 = blk[s.lane], a3 = blk[32 + s.lane];
        mma_tf32(c[T], 0u, __float_as_uint(a1.x), 0u, __float_as_uint(a3.x), bh0, bh1);
        mma_tf32(c[T], 0u, __float_as_uint(a1.y), 0u, __float_as_uint(a3.y), bh0, bh1);
        mma_tf32(c[T], 0u, __float_as_uint(a1.x), 0u, __float_as_uint(a3.x), bl0, bl1);
      }
    }
    SweepUpdate<J, T + 1, SHIFT>::run(c, s, bh0, bh1, bl0, bl1);
  }
};
template <int J, int SHIFT>
struct SweepUpdate<J, kGpTiles, SHIFT> {
  static __device__ __forceinline__ void run(float (&)[kGpTiles][4], const SweepCtx&, uint32_t, uint32_t, uint32_t, uint32_t) {}
};

// 8 x 8 tile held as C fragments (row g, columns 2 tq, 2 tq + 1) -> the B fragments (k = tq, tq + 4; n = g), split
__device__ __forceinline__ void c_to_b_split(float c0, float c1, int g, int tq, uint32_t* bh0, uint32_t* bh1, uint32_t* bl0,
                                             uint32_t* bl1) {
  const int src0 = 4 * tq + (g >> 1), src1 = 4 * (tq + 4) + (g >> 1);
  const float x0 = __shfl_sync(0xffffffffu, c0, src0), x1 = __shfl_sync(0xffffffffu, c1, src0);
  const float y0 = __shfl_sync(0xffffffffu, c0, src1), y1 = __shfl_sync(0xffffffffu, c1, src1);
  const float b0 = (g & 1) ? x1 : x0, b1 = (g & 1) ? y1 : y0;
  *bh0 = tf32_rna(b0); *bh1 = tf32_rna(b1);
  *bl0 = tf32_rna(b0 - __uint_as_float(*bh0)); *bl1 = tf32_rna(b1 - __uint_as_float(*bh1));
}

template <int J, int SHIFT>
struct Sweep {
  static __device__ __forceinline__ void run(float (&c)[kGpTiles][4], const SweepCtx& s, double (&n2)[2], double (&mu)[2],
                                             double (&mv)[2]) {
    if (J < s.nb) {
      constexpr int T = (J + SHIFT) >> 1, half = (J + SHIFT) & 1;
      uint32_t bh0, bh1, bl0, bl1;
      c_to_b_split(c[T][2 * half], c[T][2 * half + 1], s.g, s.tq, &bh0, &bh1, &bl0, &bl1);
      // V_J = inv(L_JJ) C_J: the inverse sits in rows 0..7 of the A tile, rows 8..15 are zero
      const float2* inv = s.L2 + blk_offset(J, J);
      const float2 a0 = inv[s.lane], a2 = inv[32 + s.lane];
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      mma_tf32(v, __float_as_uint(a0.x), 0u, __float_as_uint(a2.x), 0u, bh0, bh1);
      mma_tf32(v, __float_as_uint(a0.y), 0u, __float_as_uint(a2.y), 0u, bh0, bh1);
      mma_tf32(v, __float_as_uint(a0.x), 0u, __float_as_uint(a2.x), 0u, bl0, bl1);
      const double zu = s.yz[J * kGpBlk + s.g][0], zv = s.yz[J * kGpBlk + s.g][1];
      const double v0 = double(v[0]), v1 = double(v[1]);
      n2[0] = fma(v0, v0, n2[0]); n2[1] = fma(v1, v1, n2[1]);
      mu[0] = fma(v0, zu, mu[0]); mu[1] = fma(v1, zu, mu[1]);
      mv[0] = fma(v0, zv, mv[0]); mv[1] = fma(v1, zv, mv[1]);
      if (J + 1 < s.nb) {
        c_to_b_split(v[0], v[1], s.g, s.tq, &bh0, &bh1, &bl0, &bl1);
        SweepUpdate<J, (J + SHIFT + 1) / 2, SHIFT>::run(c, s, bh0, bh1, bl0, bl1);   // the tiles holding block rows > J
      }
      Sweep<J + 1, SHIFT>::run(c, s, n2, mu, mv);
    }
  }
};
template <int SHIFT>
struct Sweep<kGpNumBlk, SHIFT> {
  static __device__ __forceinline__ void run(float (&)[kGpTiles][4], const SweepCtx&, double (&)[2], double (&)[2], double (&)[2]) {}
};

template <typename Real>
__global__ void __launch_bounds__(kGpPThreads, 3)
k_gp_posterior(DevState<Real> d, float* __restrict__ obs, const __grid_constant__ CUtensorMap bank_map, int tile_column) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  PosteriorSmem& S = *reinterpret_cast<PosteriorSmem*>(s_raw);
  const int64_t e = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int count = d.gp_count[e];
  const int stored = count < kGpWindow ? count : kGpWindow;
  const int nb = (stored + kGpBlk - 1) / kGpBlk;
  const int rows = nb * kGpBlk;
  const double* ring = d.gp_obs + e * int64_t(kGpWindow * 6);
  const double x = DD(d, D_X, e), y = DD(d, D_Y, e), p_b = DD(d, D_P, e);
  const int32_t t_elapsed = d.t_elapsed[e];
  const double pmin = d.feat_range[2 * e], pmax = d.feat_range[2 * e + 1];
  const uint32_t bar = smem_u32(&S.bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    S.lo = kNumLevels; S.hi = -1; S.n_invalid = 0;
  }
  __syncthreads();
  // The forecast column (features.py:457-497 asks the forecast at 181 pressure levels of ONE (x, y, t)): every level
  // interpolates inside the same (x, y, t) cell, so the column's data is the 9 windows of that cell, 1,152 B.  They come
  // as one box of the bank's 5-D tensor view [field][y-cell][p-cell][t-cell][row] (cp.async.bulk.tensor, UTMALDG in
  // SASS) instead of one 128-byte gather per level (181 x 128 B requested for the same 1,152 B).
  const bool use_tile = tile_column != 0 && d.wind_model != BLE_WIND_SIMPLE_STATIC;
  if (tid == 0 && (nb > 0 || use_tile)) {                   // 1. the kernel matrix (one TMA bulk copy) and the column tile
    const FieldCell<double> cell0 = locate<double>(make_field_point(x / 1000.0, y / 1000.0, p_b, double(t_elapsed) / 3600.0));
    S.cell = cell0;
    const uint32_t k_bytes = nb > 0 ? uint32_t(blk_offset(nb, 0)) * 8u : 0u;
    const uint32_t bytes = k_bytes + (use_tile ? uint32_t(sizeof(S.col)) : 0u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    if (nb > 0) {
      const double* src = d.gp_chol + e * int64_t(kGpFactorDoubles);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(S.L)), "l"(src), "r"(k_bytes), "r"(bar) : "memory");
    }
    if (use_tile) {
      const int c0 = cell0.ix * d.layout.x_stride_floats, c1 = cell0.tc, c2 = 0, c3 = cell0.iy, c4 = d.env_field[e];
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                   ::"r"(smem_u32(S.col)), "l"(reinterpret_cast<uint64_t>(&bank_map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
                     "r"(bar) : "memory");
    }
  }
  // while it is in flight: reachable levels, validity of the slots, the query's distance terms, the targets
  if (tid < kNumLevels) {
    const double pl = pressure_level(tid);
    if (!(pl < pmin || pl > pmax)) { atomicMin(&S.lo, tid); atomicMax(&S.hi, tid); }
  }
  const double qx = x / kGpScaleXY, qy = y / kGpScaleXY, qt = double(t_elapsed) / kGpScaleT;
  if (tid < kGpWindow + 16 && (tid < 8 || tid >= rows + 8)) { S.cxf[tid] = 3e38f; S.pz[tid] = 0.0; }   // padding rows
  if (tid < rows) {
    double c = 0.0, pz = 0.0, eu = 0.0, ev = 0.0;
    bool ok = false;
    if (tid < stored) {
      const double* o = ring + tid * 6;
      ok = fabs(o[3] - double(t_elapsed)) < kGpHorizonS;    // wind_gp.py:172-178
      if (ok) {
        const double dx = qx - o[0] / kGpScaleXY, dy = qy - o[1] / kGpScaleXY, dt = qt - o[3] / kGpScaleT;
        c = dx * dx + dy * dy + dt * dt;
        pz = o[2] / kGpScaleP;
        eu = o[4]; ev = o[5];
      }
    }
    S.cxf[tid + 8] = ok ? float(c) : 3e38f; S.pz[tid + 8] = pz; S.yz[tid][0] = eu; S.yz[tid][1] = ev; S.valid[tid] = ok ? 1 : 0;
    if (!ok) atomicAdd(&S.n_invalid, 1);
  }
  __syncthreads();
  const int lo = S.lo, hi = S.hi;
  const int n_act = hi >= lo ? hi - lo + 1 : 0;             // reachable levels lo .. hi
  const int m_valid = rows - S.n_invalid;

  if (nb > 0 || use_tile) {
    asm volatile(                                           // wait for the TMA transactions (phase 0)
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar) : "memory");
  }
  // forecast at level l: from the staged tile (same interp_window as every other lookup: bit-identical) or by a gather
  auto column_forecast = [&](int l, double* fu, double* fv) {
    if (use_tile) {
      FieldCell<double> c = S.cell;                          // x, y, t axes: the balloon's; pressure axis: this level's
      axis_cell<double>(make_field_point(0.0, 0.0, pressure_level(l), 0.0).p, 5000.f, 1000.f, kPC, &c.pc, &c.wp);
      const float4* w = S.col + c.pc * 8;
      interp_window<double>(c, [w](int j) { return w[j]; }, fu, fv);
    } else {
      forecast_at<double, DevState<Real>>(d, e, x, y, pressure_level(l), t_elapsed, fu, fv);
    }
  };
  if (m_valid > 0) {
    if (S.n_invalid > 0) {                                  // empty / expired slots: identity rows and columns
      for (int k = tid; k < blk_offset(nb, 0); k += kGpPThreads) {
        int row, col;
        blocked_coords(k, &row, &col);
        if (col > row) S.L[k] = 0.0;                        // upper part of a diagonal block (never written in HBM)
        else if (!S.valid[row] || !S.valid[col]) S.L[k] = row == col ? 1.0 : 0.0;
      }
    } else {
      for (int j = warp; j < nb; j += kGpPWarps) {          // zero the upper parts of the diagonal blocks
        double* blk = S.L + blk_offset(j, j);
        for (int q = lane; q < 64; q += 32) { const int r = q >> 3, c = q & 7; if (c > r) blk[blk_inner(r, c)] = 0.0; }
      }
    }
    __syncthreads();

    // ---- 2. blocked Cholesky (fp64), z = L^-1 y alongside ----
    const int cpos = 4 * g + 2 * (tq & 1) + 32 * (tq >> 1);                 // this lane's (g, 2 tq), (g, 2 tq + 1) of a C tile
    if (warp == 0) {
      const double2 a = *reinterpret_cast<const double2*>(S.L + blk_offset(0, 0) + cpos);
      __syncwarp();
      chol_diag_block(a.x, a.y, S.L + blk_offset(0, 0), &S.yz[0], lane);
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
      const double* X = S.L + blk_offset(j, j);             // inv(L_jj)
      // panel: L_bj = K_bj inv(L_jj)^T for b > j; y'_b -= L_bj z_j
      for (int b = j + 1 + warp; b < nb; b += kGpPWarps) {
        double* Kb = S.L + blk_offset(b, j);
        double c0 = 0.0, c1 = 0.0;
        dmma884(c0, c1, Kb[lane], X[lane]);                 // B fragment of X^T == A-fragment order of X
        dmma884(c0, c1, Kb[32 + lane], X[32 + lane]);
        __syncwarp();
        *reinterpret_cast<double2*>(Kb + 4 * g + 2 * (tq & 1) + 32 * (tq >> 1)) = make_double2(c0, c1);   // (g, 2 tq), (g, 2 tq + 1)
        __syncwarp();
        if (lane < 16) {
          const int row = lane & 7, col = lane >> 3;
          double acc = 0.0;
#pragma unroll
          for (int c = 0; c < kGpBlk; ++c) acc = fma(Kb[blk_inner(row, c)], S.yz[j * kGpBlk + c][col], acc);
          S.yz[b * kGpBlk + row][col] -= acc;
        }
      }
      __syncthreads();
      // trailing update K_bb' -= L_bj L_b'j^T for j < b' <= b.  Warp 0 takes (j + 1, j + 1) and factors it at once
      // (look-ahead); warps 1..7 own block COLUMNS b' cyclically: the B operand L_b'j stays in registers while b runs down
      // the column, and the block addresses advance by increments.
      if (warp == 0) {
        if (j + 1 < nb) {
          double* C = S.L + blk_offset(j + 1, j + 1);
          const double* A = S.L + blk_offset(j + 1, j);
          double2 cc = *reinterpret_cast<double2*>(C + cpos);
          dmma884(cc.x, cc.y, -A[lane], A[lane]);
          dmma884(cc.x, cc.y, -A[32 + lane], A[32 + lane]);
          __syncwarp();
          chol_diag_block(cc.x, cc.y, C, &S.yz[(j + 1) * kGpBlk], lane);     // the updated block stays in registers
        }
      } else {
        // Warps 1..7 own block COLUMNS b2 of the trailing matrix: the B operand L_b2j stays in registers while b runs down
        // the column and the block addresses advance by increments (18 instructions per block; dealing single blocks
        // round-robin balanced the warps perfectly but cost 41 instructions per block and no time).  The columns are
        // dealt serpentine -- j + 1 .. j + 7 to warps 1 .. 7, j + 8 .. j + 14 to warps 7 .. 1 -- so every warp gets
        // one long and one short column.
#pragma unroll 1
        for (int round = 0; round < 2; ++round) {
          const int b2 = round == 0 ? j + warp : j + 15 - warp;
          if (b2 >= nb) continue;
          const double* B = S.L + blk_offset(b2, j);
          const double bf0 = B[lane], bf1 = B[32 + lane];
          int b = b2 == j + 1 ? b2 + 1 : b2;                                // (j + 1, j + 1) belongs to warp 0
          int offA = blk_offset(b, j), offC = blk_offset(b, b2);
          for (; b < nb; ++b) {
            double2 cc = *reinterpret_cast<double2*>(S.L + offC + cpos);
            dmma884(cc.x, cc.y, -S.L[offA + lane], bf0);
            dmma884(cc.x, cc.y, -S.L[offA + 32 + lane], bf1);
            *reinterpret_cast<double2*>(S.L + offC + cpos) = cc;
            offA += (b + 1) * (kGpBlk * kGpBlk); offC += (b + 1) * (kGpBlk * kGpBlk);
          }
        }
      }
      __syncthreads();
    }

    // ---- 3. demote the factor in place to TF32 pairs (off-diagonal blocks negated) ----
    // The sweep reads 16 x 8 tiles of L as mma.m16n8k8 A fragments: block rows are paired (2 T - shift, 2 T + 1 - shift)
    // with shift = nb & 1, so that the LAST block row is always the bottom of a tile (for an odd nb the top of tile 0 is
    // a phantom row that no update ever touches).  A tile whose two blocks both lie below the pivot column is stored
    // FRAGMENT-MAJOR over the 1,024 bytes of its two blocks: lane l's four hi parts {(g, tq), (g + 8, tq), (g, tq + 4),
    // (g + 8, tq + 4)} as one float4 in the top block, its four lo parts as one float4 in the bottom block -- one
    // LDS.128 each, landing in the register quad the MMA wants.  Diagonal blocks (the inverses) and the block right
    // under a diagonal block whose tile partner is that diagonal block stay element-wise {hi, lo} float2.
    {
      // Work units of pivot column j, in order: u = 0 the diagonal block; then, if block row j + 1 is the BOTTOM of a tile
      // (its tile partner is the diagonal block), that block alone; then the full tiles (b, b + 1), b = b0, b0 + 2, ...
      // Unit u of column j belongs to warp (u + 3 j) mod 8, so a warp walks each column with stride 8 (the first version
      // scanned all 120 (j, b) pairs in every warp: 17 % of the kernel's instructions).
      const int shift = nb & 1;
      for (int j = 0; j < nb; ++j) {
        const bool under = j + 1 < nb && ((j + 1 + shift) & 1) != 0;          // (j + 1, j) is a single block
        const int b0 = j + 1 + (under ? 1 : 0);                              // first full tile's top block row
        const int n_units = 1 + (under ? 1 : 0) + (nb > b0 ? (nb - b0) >> 1 : 0);
        for (int u = (warp - 3 * j) & (kGpPWarps - 1); u < n_units; u += kGpPWarps) {
          const bool single = u == 0 || (under && u == 1);
          if (!single) {
            const int b = b0 + 2 * (u - 1 - (under ? 1 : 0));
            double* t0 = S.L + blk_offset(b, j);
            double* t1 = S.L + blk_offset(b + 1, j);
            const float e0 = float(-t0[lane]), e2 = float(-t0[32 + lane]), e1 = float(-t1[lane]), e3 = float(-t1[32 + lane]);
            __syncwarp();
            float4 hi, lo4;
            hi.x = __uint_as_float(tf32_rna(e0)); hi.y = __uint_as_float(tf32_rna(e1));
            hi.z = __uint_as_float(tf32_rna(e2)); hi.w = __uint_as_float(tf32_rna(e3));
            lo4.x = __uint_as_float(tf32_rna(e0 - hi.x)); lo4.y = __uint_as_float(tf32_rna(e1 - hi.y));
            lo4.z = __uint_as_float(tf32_rna(e2 - hi.z)); lo4.w = __uint_as_float(tf32_rna(e3 - hi.w));
            reinterpret_cast<float4*>(t0)[lane] = hi;
            reinterpret_cast<float4*>(t1)[lane] = lo4;
          } else {
            const int b = j + u;
            double* t0 = S.L + blk_offset(b, j);
            const float sgn = b == j ? 1.f : -1.f;
            const float e0 = sgn * float(t0[lane]), e2 = sgn * float(t0[32 + lane]);
            __syncwarp();
            const float h0 = __uint_as_float(tf32_rna(e0)), h2 = __uint_as_float(tf32_rna(e2));
            reinterpret_cast<float2*>(t0)[lane] = make_float2(h0, __uint_as_float(tf32_rna(e0 - h0)));
            reinterpret_cast<float2*>(t0)[32 + lane] = make_float2(h2, __uint_as_float(tf32_rna(e2 - h2)));
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- 4. column sweep ----
  SweepCtx sc{reinterpret_cast<const float2*>(S.L), S.yz, nb, lane, g, tq};
  for (int tile = warp; tile * 8 < n_act && m_valid > 0; tile += kGpPWarps) {
    const int col0 = tile * 8 + 2 * tq;                     // this lane's two columns (C-fragment layout)
    const bool on0 = col0 < n_act, on1 = col0 + 1 < n_act;
    const double pq0 = pressure_level(lo + (on0 ? col0 : 0)) / kGpScaleP;
    const double pq1 = pressure_level(lo + (on1 ? col0 + 1 : 0)) / kGpScaleP;
    float c[kGpTiles][4];
    const int shift = nb & 1;
#pragma unroll
    for (int t = 0; t < kGpTiles; ++t) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 8 * (2 * t + h - shift) + g + 8;      // padded index of block row 2 t + h - shift (7 + g: the phantom row)
        // K*[i][l] = sigma^2 exp(-sqrt(c_i + (p_l - p_i)^2)): the pressure difference in fp64 (it cancels), the rest in fp32
        // with the approximate root and exponential (the solve that consumes it is 3 x TF32: ~1e-7 relative either way)
        const double pi = S.pz[i];
        const float cx = S.cxf[i];
        const float e0 = float(pq0 - pi), e1 = float(pq1 - pi);
        float k0 = float(kGpSigma2) * fm::ex2f(-1.4426950408889634f * fm::sqrtf_pos(fmaf(e0, e0, cx)));
        float k1 = float(kGpSigma2) * fm::ex2f(-1.4426950408889634f * fm::sqrtf_pos(fmaf(e1, e1, cx)));
        if (!on0) k0 = 0.f;
        if (!on1) k1 = 0.f;
        c[t][2 * h] = k0; c[t][2 * h + 1] = k1;
      }
    }
    double n2[2] = {0.0, 0.0}, mu[2] = {0.0, 0.0}, mv[2] = {0.0, 0.0};
    if (shift) Sweep<0, 1>::run(c, sc, n2, mu, mv); else Sweep<0, 0>::run(c, sc, n2, mu, mv);
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
      column_forecast(l, &fu, &fv);
      wind_level_features(mean_u + fu, mean_v + fv, deviation, x, y, &S.feat[l * 3], &S.feat[l * 3 + 1], &S.feat[l * 3 + 2]);
    }
  }
  if (m_valid == 0) {                                       // no measurement yet: zero mean and deviation (wind_gp.py:161-163)
    for (int k = tid; k < n_act; k += kGpPThreads) {
      const int l = lo + k;
      double fu, fv;
      column_forecast(l, &fu, &fv);
      wind_level_features(fu, fv, 0.0, x, y, &S.feat[l * 3], &S.feat[l * 3 + 1], &S.feat[l * 3 + 2]);
    }
  }
  __syncthreads();
  // ---- 5. centred, padded column (features.py:479-497, 536-556) ----
  const int lower = kNumLevels - nearest_pressure_level(p_b) - 1;
  float* o = obs + e * int64_t(kNumFeatures) + 16;
  for (int s = tid; s < 2 * kNumLevels - 1; s += kGpPThreads) {
    float f0 = 0.f, f1 = 1.f, f2 = 1.f;                                            // "unreachable" triple
    const int l = s - lower;
    if (l >= 0 && l < kNumLevels) {
      const double pl = pressure_level(l);
      if (!(pl < pmin || pl > pmax)) { f0 = S.feat[l * 3]; f1 = S.feat[l * 3 + 1]; f2 = S.feat[l * 3 + 2]; }
    }
    o[s * 3] = f0; o[s * 3 + 1] = f1; o[s * 3 + 2] = f2;
  }
}
