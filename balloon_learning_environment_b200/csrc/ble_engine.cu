// CUDA engine + C ABI (include/ble_b200.h) for the batched BLE transition function.
// Target: sm_100a (B200).  One handle per GPU.  The production step kernels live in ble_step_fused.cu (one launch per
// BalloonEnv.step); this file holds the engine (memory, launches, error handling), the reset / IO / wind-query / feature
// kernels, the generation path, the legacy step kernels (fp64 audit build, A/B) and the C entry points.
#include <cublasLt.h>
#include <cuda.h>           // CUtensorMap types only: the encoder is fetched with cudaGetDriverEntryPoint (no -lcuda)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>

#include "../../include/ble_b200.h"
#include "ble_devstate.cuh"
#include "ble_fastmath.cuh"
#include "ble_step_fused.h"
#include "ble_agents.cuh"
#include "ble_features.cuh"

namespace ble {

// ---------------------------------------------------------------------------------------------
// State upload / download (get/set_balloon_state, env/balloon_arena.py:213-220)
// ---------------------------------------------------------------------------------------------
template <typename Real>
__global__ void k_state_upload(DevState<Real> d, const double* __restrict__ f, const int64_t* __restrict__ iv) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const int64_t n = d.n;
  auto F = [&](int row) { return f[int64_t(row) * n + e]; };
  auto I = [&](int row) { return iv[int64_t(row) * n + e]; };
  DD(d, D_X, e) = F(BLE_F_X); DD(d, D_Y, e) = F(BLE_F_Y); DD(d, D_P, e) = F(BLE_F_PRESSURE);
  DD(d, D_TAMB, e) = F(BLE_F_AMBIENT_TEMPERATURE); DD(d, D_TINT, e) = F(BLE_F_INTERNAL_TEMPERATURE);
  DD(d, D_VOL, e) = F(BLE_F_ENVELOPE_VOLUME); DD(d, D_SP, e) = F(BLE_F_SUPERPRESSURE);
  DD(d, D_MOLS_AIR, e) = F(BLE_F_MOLS_AIR); DD(d, D_CHARGE, e) = F(BLE_F_BATTERY_CHARGE);
  RR(d, R_ACS_W, e) = Real(F(BLE_F_ACS_POWER)); RR(d, R_ACS_FLOW, e) = Real(F(BLE_F_ACS_MASS_FLOW));
  RR(d, R_SOLAR_W, e) = Real(F(BLE_F_SOLAR_CHARGING)); RR(d, R_LOAD_W, e) = Real(F(BLE_F_POWER_LOAD));
  RR(d, R_LAT0, e) = Real(F(BLE_F_CENTER_LAT)); RR(d, R_LNG0, e) = Real(F(BLE_F_CENTER_LNG));
  RR(d, R_IR, e) = Real(F(BLE_F_UPWELLING_INFRARED)); RR(d, R_MOLS_GAS, e) = Real(F(BLE_F_MOLS_LIFT_GAS));
  Atmosphere atm; atm.init(F(BLE_F_ATMOSPHERE_ALPHA));
  store_atmosphere(d, e, atm);
  d.l[int64_t(L_DATE_TIME) * n + e] = I(BLE_I_DATE_TIME);
  d.l[int64_t(L_SUNRISE_H) * n + e] = I(BLE_I_SUNRISE_H);
  d.l[int64_t(L_SUNSET) * n + e] = I(BLE_I_SUNSET);
  d.t_elapsed[e] = int32_t(I(BLE_I_TIME_ELAPSED));
  d.flags[e] = pack_flags(int(I(BLE_I_STATUS)), int(I(BLE_I_LAST_COMMAND)), int(I(BLE_I_ENVELOPE_STATE)),
                          int(I(BLE_I_ALTITUDE_STATE)), int(I(BLE_I_POWER_PAUSED)),
                          int(I(BLE_I_POWER_SAFETY_ENABLED)), 0);
}

template <typename Real>
__global__ void k_state_download(DevState<Real> d, double* __restrict__ f, int64_t* __restrict__ iv) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const int64_t n = d.n;
  auto F = [&](int row) -> double& { return f[int64_t(row) * n + e]; };
  auto I = [&](int row) -> int64_t& { return iv[int64_t(row) * n + e]; };
  F(BLE_F_X) = DD(d, D_X, e); F(BLE_F_Y) = DD(d, D_Y, e); F(BLE_F_PRESSURE) = DD(d, D_P, e);
  F(BLE_F_AMBIENT_TEMPERATURE) = DD(d, D_TAMB, e); F(BLE_F_INTERNAL_TEMPERATURE) = DD(d, D_TINT, e);
  F(BLE_F_ENVELOPE_VOLUME) = DD(d, D_VOL, e); F(BLE_F_SUPERPRESSURE) = DD(d, D_SP, e);
  F(BLE_F_MOLS_AIR) = DD(d, D_MOLS_AIR, e); F(BLE_F_MOLS_LIFT_GAS) = double(RR(d, R_MOLS_GAS, e));
  F(BLE_F_BATTERY_CHARGE) = DD(d, D_CHARGE, e); F(BLE_F_ACS_POWER) = double(RR(d, R_ACS_W, e));
  F(BLE_F_ACS_MASS_FLOW) = double(RR(d, R_ACS_FLOW, e)); F(BLE_F_SOLAR_CHARGING) = double(RR(d, R_SOLAR_W, e));
  F(BLE_F_POWER_LOAD) = double(RR(d, R_LOAD_W, e)); F(BLE_F_CENTER_LAT) = double(RR(d, R_LAT0, e));
  F(BLE_F_CENTER_LNG) = double(RR(d, R_LNG0, e)); F(BLE_F_UPWELLING_INFRARED) = double(RR(d, R_IR, e));
  F(BLE_F_ATMOSPHERE_ALPHA) = DD(d, D_ALPHA, e);
  const uint32_t fl = d.flags[e];
  I(BLE_I_DATE_TIME) = d.l[int64_t(L_DATE_TIME) * n + e];
  I(BLE_I_TIME_ELAPSED) = d.t_elapsed[e];
  I(BLE_I_LAST_COMMAND) = (fl >> 2) & 3; I(BLE_I_STATUS) = fl & 3;
  I(BLE_I_ENVELOPE_STATE) = (fl >> 4) & 7; I(BLE_I_ALTITUDE_STATE) = (fl >> 7) & 3;
  I(BLE_I_POWER_PAUSED) = (fl >> 9) & 1;
  I(BLE_I_SUNRISE_H) = d.l[int64_t(L_SUNRISE_H) * n + e];
  I(BLE_I_SUNSET) = d.l[int64_t(L_SUNSET) * n + e];
  I(BLE_I_POWER_SAFETY_ENABLED) = (fl >> 10) & 1;
}

// ---------------------------------------------------------------------------------------------
// Wind field re-layout: native [F,21,21,10,9,2] -> 128-byte lookup windows (ble_wind.cuh)
// ---------------------------------------------------------------------------------------------
// One thread per 64-byte block = one grid column (ix) of one (y,p,t) cell, ordered [dy][dp][dt][uv].
// X64: 21 blocks per row; X128: 20 windows per row, each holding block(ix) then block(ix+1).
__global__ void k_fields_to_windows(const float* __restrict__ native, float* __restrict__ cells,
                                    FieldLayout layout, int64_t first_field, int64_t n_fields,
                                    const int32_t* __restrict__ dst_index /* nullptr: first_field + f */) {
  const int blocks_per_row = layout.row_floats / 16;
  const int64_t per_field = int64_t(kYC) * kPC * kTC * blocks_per_row;
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (t >= per_field * n_fields) return;
  const int64_t f = t / per_field;
  int64_t r = t - f * per_field;
  const int b = int(r % blocks_per_row); r /= blocks_per_row;
  const int tc = int(r % kTC); r /= kTC;
  const int pc = int(r % kPC);
  const int iy = int(r / kPC);
  const int ix = (layout.x_stride_floats == 32) ? (b >> 1) + (b & 1) : b;   // X128: window b/2, half b&1
  const float* src = native + f * kFieldFloats;
  const int64_t dst_field = dst_index != nullptr ? int64_t(dst_index[f]) : first_field + f;
  float4* dst = reinterpret_cast<float4*>(cells + dst_field * layout.field_floats +
                                          ((int64_t(iy) * kPC + pc) * kTC + tc) * layout.row_floats + int64_t(b) * 16);
#pragma unroll
  for (int c = 0; c < 4; ++c) {                       // chunk c = dy*2 + dp
    const int y = iy + (c >> 1), p = pc + (c & 1);
    dst[c] = make_float4(src[native_index(ix, y, p, tc, 0)], src[native_index(ix, y, p, tc, 1)],
                         src[native_index(ix, y, p, tc + 1, 0)], src[native_index(ix, y, p, tc + 1, 1)]);
  }
}

// GridBasedWindField.get_forecast for M arbitrary points (C ABI ble_wind_gather).
// One thread locates one lookup (clip, boomerang, fp32 point, cell + weights); the 128-byte
// windows are then loaded COOPERATIVELY: in round r, the 8 lanes of group g = lane/8 read the
// 8 x 16 B chunks of the window owned by lane 4r + g, i.e. one warp-wide LDG.128 covers four
// whole 128-byte lines.  Each lane weights its chunk and the group reduces with 3 xor-shuffles.
template <typename Real>
__global__ void __launch_bounds__(256)
k_wind_gather(const float* __restrict__ cells, FieldLayout layout, const float4* __restrict__ xyzt,
              const int32_t* __restrict__ field_idx, float2* __restrict__ uv, int64_t m) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const bool valid = i < m;
  FieldCell<Real> c{};
  int64_t base = -1;                                   // in float4 units; -1 = no lookup
  if (valid) {
    const float4 q4 = __ldg(xyzt + i);
    const FieldPoint q = make_field_point(double(q4.x), double(q4.y), double(q4.z), double(q4.w));
    c = locate<Real>(q);
    base = (int64_t(__ldg(field_idx + i)) * layout.field_floats + window_index(layout, c.ix, c.iy, c.pc, c.tc)) >> 2;
  }
  const float4* cells4 = reinterpret_cast<const float4*>(cells);
  const int j = int(lane & 7u);                        // my chunk within the group's window
  const unsigned g = lane >> 3;
  Real my_u = Real(0), my_v = Real(0);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int src = 4 * r + int(g);
    const int64_t b = __shfl_sync(0xffffffffu, base, src);
    FieldCell<Real> o;
    o.wx = __shfl_sync(0xffffffffu, c.wx, src); o.wy = __shfl_sync(0xffffffffu, c.wy, src);
    o.wp = __shfl_sync(0xffffffffu, c.wp, src); o.wt = __shfl_sync(0xffffffffu, c.wt, src);
    const float4 v4 = (b >= 0) ? __ldg(cells4 + b + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    Real pu, pv;
    chunk_contribution<Real>(o, j, v4, &pu, &pv);
    pu += __shfl_xor_sync(0xffffffffu, pu, 1); pv += __shfl_xor_sync(0xffffffffu, pv, 1);
    pu += __shfl_xor_sync(0xffffffffu, pu, 2); pv += __shfl_xor_sync(0xffffffffu, pv, 2);
    pu += __shfl_xor_sync(0xffffffffu, pu, 4); pv += __shfl_xor_sync(0xffffffffu, pv, 4);
    // the sum for lookup `src` now sits in group g; its owner (lane src) picks it up
    const Real ru = __shfl_sync(0xffffffffu, pu, int(lane & 3u) * 8);
    const Real rv = __shfl_sync(0xffffffffu, pv, int(lane & 3u) * 8);
    if (int(lane >> 2) == r) { my_u = ru; my_v = rv; }
  }
  if (valid) uv[i] = make_float2(float(my_u), float(my_v));
}

// ---------------------------------------------------------------------------------------------
// Noise: permutation tables (reset time) and the per-step noise kernel
// ---------------------------------------------------------------------------------------------
// One thread per table, 64 consecutive balloons of ONE harmonic per CTA: the 64 tables are built in shared memory (the
// Fisher-Yates walk indexes its arrays dynamically; as per-thread local arrays they lived in local memory and the kernel
// took 3.3 ms for 65,536 balloons) and leave as one contiguous, coalesced 16 KB block.  Lane-interleaved layout
// (byte i of table t at i * 64 + t) for the scratch array; the finished table is written at its rotated position
// (bank-conflict-free rotation by 4 (e mod 32) bytes, undone at lookup time) of a table-major staging area.
constexpr int kPermCta = 64;
__global__ void __launch_bounds__(kPermCta)
k_make_perms(int64_t n, const int64_t* __restrict__ seeds /*[n,2,5]*/, const float* __restrict__ offsets_in /*[n,2,5,4]*/,
             const uint8_t* __restrict__ mask, uint8_t* __restrict__ perm /*[10][n][256]*/, float* __restrict__ offsets /*[10][4][n]*/) {
  __shared__ uint8_t s_source[256 * kPermCta];              // [i][t]
  __shared__ __align__(16) uint8_t s_table[kPermCta * 256]; // [t][rotated j]
  const int h10 = blockIdx.y, t = threadIdx.x;
  const int64_t e0 = int64_t(blockIdx.x) * kPermCta, e = e0 + t;
  const bool live = e < n && (mask == nullptr || mask[e] != 0);
  if (live) {
    const uint64_t A = 6364136223846793005ull, Cc = 1442695040888963407ull;
    uint64_t s = uint64_t(seeds[e * 10 + h10]);
    for (int i = 0; i < 256; ++i) s_source[i * kPermCta + t] = uint8_t(i);
    s = s * A + Cc; s = s * A + Cc; s = s * A + Cc;
    const int rot = int(e & 31) * 4;
    uint8_t* table = s_table + t * 256;
    for (int i = 255; i >= 0; --i) {
      s = s * A + Cc;
      // Python: r = (int64(s) + 31) % (i + 1), floor modulo (simplex_make_perm in ble_wind.cuh is the plain statement);
      // here in 32-bit pieces: s = hi 2^32 + lo (unsigned) - 2^64 [s < 0]
      const uint32_t m = uint32_t(i + 1);
      const uint32_t hi = uint32_t(s >> 32), lo = uint32_t(s);
      const uint32_t p32 = uint32_t((uint64_t(1) << 32) % m);          // 2^32 mod m
      uint32_t r = ((hi % m) * p32 + lo % m + 31u) % m;                // < 256 * 256 + 256 + 31: no overflow
      if (int64_t(s) < 0) { const uint32_t p64 = (p32 * p32) % m; r = (r + m - p64) % m; }
      table[(i + rot) & 255] = s_source[r * kPermCta + t];
      s_source[r * kPermCta + t] = s_source[i * kPermCta + t];
    }
    for (int c = 0; c < 4; ++c) offsets[(int64_t(h10) * 4 + c) * n + e] = offsets_in[(e * 10 + h10) * 4 + c];
  }
  __syncthreads();
  // copy-out: the CTA's tables are contiguous in HBM; a masked-out table keeps what it had
  uint4* dst = reinterpret_cast<uint4*>(perm + (int64_t(h10) * n + e0) * 256);
  const uint4* src = reinterpret_cast<const uint4*>(s_table);
  for (int k = t; k < kPermCta * 16; k += kPermCta) {
    const int64_t ek = e0 + (k >> 4);
    if (ek < n && (mask == nullptr || mask[ek] != 0)) dst[k] = src[k];
  }
}

// ble_config.auto_reset: the balloons whose step returned done start a new episode; its seed continues a splitmix64 chain
// from the seed of the episode that ended.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void k_auto_reset_prepare(int64_t n, const uint8_t* __restrict__ done, uint64_t* __restrict__ episode_seed,
                                     uint8_t* __restrict__ mask) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n) return;
  const bool m = done[e] != 0;
  mask[e] = m ? 1 : 0;
  if (m) episode_seed[e] = splitmix64(episode_seed[e]);
}
__global__ void k_keep_episode_seeds(int64_t n, const uint64_t* __restrict__ seeds, const uint8_t* __restrict__ mask,
                                     uint64_t* __restrict__ episode_seed) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e < n && (mask == nullptr || mask[e] != 0)) episode_seed[e] = seeds[e];
}

constexpr int kNoiseBlock = 128;

// One thread per (balloon, harmonic).  blockIdx.y = harmonic (0..9), blockIdx.x = block of 128
// consecutive balloons, whose 128 x 256 B permutation tables are contiguous in HBM and are
// staged into shared memory with ONE TMA bulk copy (cp.async.bulk) tracked by an mbarrier.
template <typename Real>
__global__ void __launch_bounds__(kNoiseBlock)
k_noise(DevState<Real> d) {
  extern __shared__ __align__(128) uint8_t s_perm[];     // kNoiseBlock * 256 bytes
  __shared__ __align__(8) uint64_t s_bar;
  const int h10 = blockIdx.y;
  const int64_t e0 = int64_t(blockIdx.x) * kNoiseBlock;
  const int count = int(min(int64_t(kNoiseBlock), d.n - e0));
  const uint32_t bar = smem_u32(&s_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = uint32_t(count) * 256u;
    const uint8_t* src = d.perm + (int64_t(h10) * d.n + e0) * 256;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(s_perm)), "l"(src), "r"(bytes), "r"(bar) : "memory");
  }
  // coordinates while the copy is in flight (NoisyWindHarmonic.get_noise, simplex_wind_noise.py:116-146)
  const int64_t e = e0 + threadIdx.x;
  const bool live = threadIdx.x < count;
  double X = 0, Y = 0, Z = 0, W = 0;
  if (live) {
    // same expression as k_step_fused's in-kernel evaluation: both paths give bit-identical noise
    const double* hp = kHarmonicsInvDev[h10];
    X = fma(double(DD(d, D_X, e)), hp[0], double(d.offsets[(int64_t(h10) * 4 + 0) * d.n + e]));
    Y = fma(double(DD(d, D_Y, e)), hp[1], double(d.offsets[(int64_t(h10) * 4 + 1) * d.n + e]));
    Z = fma(double(DD(d, D_P, e)), hp[2], double(d.offsets[(int64_t(h10) * 4 + 2) * d.n + e]));
    W = fma(double(d.t_elapsed[e]), hp[3], double(d.offsets[(int64_t(h10) * 4 + 3) * d.n + e]));
  }
  // wait for the TMA transaction (phase 0)
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar) : "memory");
  if (live) {
    RotatedPerm perm{s_perm + threadIdx.x * 256, int(e & 31) * 4};
    const Real v = simplex_noise4_v2<Real>(perm, X, Y, Z, W);
    d.noise_partial[int64_t(h10) * d.n + e] = Real(kNoiseMagnitude) * v;
  }
}

// Simplex noise at balloon e's current state, from the harmonic partials written by k_noise:
// NoisyWindComponent.get_noise (:180-211) = weighted mean of 5 harmonics, variance-rescaled.
template <typename Real>
__device__ __forceinline__ void noise_at(const DevState<Real>& d, int64_t e, Real* nu_out, Real* nv_out) {
  const double wu[5] = {0.1445, 0.2766, 0.2627, 0.2137, 0.1025};
  const double wv[5] = {0.2716, 0.2684, 0.2348, 0.1186, 0.1066};
  Real nu = Real(0), nv = Real(0);
  double swu = 0, swu2 = 0, swv = 0, swv2 = 0;
#pragma unroll
  for (int h = 0; h < 5; ++h) {
    nu += d.noise_partial[int64_t(h) * d.n + e] * Real(wu[h]);
    nv += d.noise_partial[int64_t(5 + h) * d.n + e] * Real(wv[h]);
    swu += wu[h]; swu2 += wu[h] * wu[h]; swv += wv[h]; swv2 += wv[h] * wv[h];
  }
  *nu_out = nu / Real(swu) * Real(sqrt(swu / swu2));
  *nv_out = nv / Real(swv) * Real(sqrt(swv / swv2));
}

// forecast + noise at the balloon's current state (WindField.get_ground_truth, env/wind_field.py:125-145)
template <typename Real>
__device__ __forceinline__ void wind_at_balloon(const DevState<Real>& d, int64_t e, double x, double y, double p,
                                                int32_t t_elapsed, Real* u, Real* v) {
  forecast_at<Real, DevState<Real>>(d, e, x, y, p, t_elapsed, u, v);
  if (d.enable_noise) {
    Real nu, nv;
    noise_at<Real>(d, e, &nu, &nv);
    *u += nu;
    *v += nv;
  }
}

template <typename Real>
__global__ void __launch_bounds__(128)
k_wind_at_balloon(DevState<Real> d, float2* __restrict__ uv) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  Real u, v;
  wind_at_balloon<Real>(d, e, DD(d, D_X, e), DD(d, D_Y, e), DD(d, D_P, e), d.t_elapsed[e], &u, &v);
  uv[e] = make_float2(float(u), float(v));
}

// WindField.get_forecast / get_ground_truth (env/wind_field.py:69-145) at ARBITRARY points of a balloon's own wind field
// and noise generators -- what SimulatorState.wind_field hands to a reference consumer.  One thread per query; the
// permutation tables are read in place (this is the N = 1 adaptor's path, not the step's).
template <typename Real>
__global__ void __launch_bounds__(128)
k_wind_query(DevState<Real> d, const double* __restrict__ xyzt /*[m][4]: x m, y m, Pa, elapsed s*/,
             const int32_t* __restrict__ env_idx, int with_noise, float2* __restrict__ uv, int64_t m) {
  const int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (q >= m) return;
  const int64_t e = env_idx[q];
  const double x = xyzt[4 * q], y = xyzt[4 * q + 1], p = xyzt[4 * q + 2], t_s = xyzt[4 * q + 3];
  Real u, v;
  if (d.wind_model == BLE_WIND_SIMPLE_STATIC) {
    static_wind<Real>(Real(p), &u, &v);
  } else {
    const FieldPoint pt = make_field_point(x / 1000.0, y / 1000.0, p, t_s / 3600.0);
    const FieldCell<Real> c = locate<Real>(pt);
    WindowLoader ld{reinterpret_cast<const float4*>(d.cells + int64_t(d.env_field[e]) * d.layout.field_floats +
                                                    window_index(d.layout, c.ix, c.iy, c.pc, c.tc))};
    interp_window<Real>(c, ld, &u, &v);
  }
  if (with_noise && d.enable_noise) {                      // SimplexWindNoise.get_wind_noise (:209-218)
    Real comp[2] = {Real(0), Real(0)};
    for (int h10 = 0; h10 < 10; ++h10) {
      const double* hp = kHarmonicsInvDev[h10];
      const double X = fma(x, hp[0], double(d.offsets[(int64_t(h10) * 4 + 0) * d.n + e]));
      const double Y = fma(y, hp[1], double(d.offsets[(int64_t(h10) * 4 + 1) * d.n + e]));
      const double Z = fma(p, hp[2], double(d.offsets[(int64_t(h10) * 4 + 2) * d.n + e]));
      const double W = fma(t_s, hp[3], double(d.offsets[(int64_t(h10) * 4 + 3) * d.n + e]));
      RotatedPerm perm{d.perm + (int64_t(h10) * d.n + e) * 256, int(e & 31) * 4};
      const Real nh = Real(kNoiseMagnitude) * simplex_noise4_v2<Real>(perm, X, Y, Z, W);
      comp[h10 / 5] += nh * Real(h10 < 5 ? kBlendU[h10] : kBlendV[h10 - 5]);
    }
    u += comp[0] * Real(kBlendScaleU);
    v += comp[1] * Real(kBlendScaleV);
  }
  uv[q] = make_float2(float(u), float(v));
}

// Atmosphere.at_pressure / at_height (env/balloon/standard_atmosphere.py:89-154) of a balloon's atmosphere.
// out[q] = (height m, temperature K, pressure Pa, density kg/m^3); NaN where the reference asserts (:95-96, :126-127).
template <typename Real>
__global__ void __launch_bounds__(128)
k_atmosphere_query(DevState<Real> d, int which, const double* __restrict__ q, const int32_t* __restrict__ env_idx,
                   double* __restrict__ out, int64_t m) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  const double alpha = DD(d, D_ALPHA, env_idx[i]);
  double height, pressure, temperature;
  bool ok;
  if (which == 0) { pressure = q[i]; ok = atm_at_pressure_generic(alpha, pressure, &height, &temperature); }
  else { height = q[i]; ok = atm_at_height_generic(alpha, height, &pressure, &temperature); }
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  out[4 * i] = ok ? height : nan;
  out[4 * i + 1] = ok ? temperature : nan;
  out[4 * i + 2] = ok ? pressure : nan;
  out[4 * i + 3] = ok ? pressure * kMAir / (kR * temperature) : nan;      // :150-153
}

// ---------------------------------------------------------------------------------------------
// The fused physics step: wind at balloon -> safety layers -> 18 Euler sub-steps -> reward/done
// ---------------------------------------------------------------------------------------------
#ifndef BLE_STEP_MIN_BLOCKS
#define BLE_STEP_MIN_BLOCKS 4     // 128 registers: measured 0.225 ms/step vs 0.310 ms at 184 registers (N = 65,536)
#endif
template <typename Real>
__global__ void __launch_bounds__(128, BLE_STEP_MIN_BLOCKS)
k_step(DevState<Real> d, const int32_t* __restrict__ actions, float* __restrict__ reward,
       uint8_t* __restrict__ done, float2* __restrict__ wind_uv) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const uint32_t fl = d.flags[e];
  if ((fl & 3u) != uint32_t(kOk)) {          // finished balloon: no-op (documented divergence)
    reward[e] = 0.f;
    done[e] = 1;
    if (wind_uv != nullptr) wind_uv[e] = make_float2(0.f, 0.f);
    return;
  }
  BalloonState<Real> s;
  s.x = DD(d, D_X, e); s.y = DD(d, D_Y, e); s.pressure = DD(d, D_P, e);
  s.t_ambient = DD(d, D_TAMB, e); s.t_internal = DD(d, D_TINT, e); s.volume = DD(d, D_VOL, e);
  s.superpressure = DD(d, D_SP, e); s.mols_air = DD(d, D_MOLS_AIR, e); s.charge = DD(d, D_CHARGE, e);
  s.acs_power = RR(d, R_ACS_W, e); s.acs_flow = RR(d, R_ACS_FLOW, e);
  s.solar_w = RR(d, R_SOLAR_W, e); s.load_w = RR(d, R_LOAD_W, e);
  s.lat0 = RR(d, R_LAT0, e); s.lng0 = RR(d, R_LNG0, e); s.ir = RR(d, R_IR, e); s.mols_gas = RR(d, R_MOLS_GAS, e);
  s.date_time = d.l[int64_t(L_DATE_TIME) * d.n + e];
  s.time_elapsed = d.t_elapsed[e];
  s.status = kOk;
  Atmosphere atm = load_atmosphere(d, e);
  SafetyState ss;
  ss.sunrise_h = d.l[int64_t(L_SUNRISE_H) * d.n + e];
  ss.sunset = d.l[int64_t(L_SUNSET) * d.n + e];
  ss.last_command = int((fl >> 2) & 3); ss.envelope_state = int((fl >> 4) & 7);
  ss.altitude_state = int((fl >> 7) & 3); ss.power_paused = int((fl >> 9) & 1);
  ss.power_safety_enabled = int((fl >> 10) & 1);

  Real u, v;
  wind_at_balloon<Real>(d, e, s.x, s.y, s.pressure, s.time_elapsed, &u, &v);   // PRE-step lookup
  int action = actions[e];
  action = action < 0 ? 0 : (action > 2 ? 2 : action);
  int eff;
  const Real r = agent_step<Real>(s, atm, ss, action, double(u), double(v), &eff);

  DD(d, D_X, e) = s.x; DD(d, D_Y, e) = s.y; DD(d, D_P, e) = s.pressure;
  DD(d, D_TAMB, e) = s.t_ambient; DD(d, D_TINT, e) = s.t_internal; DD(d, D_VOL, e) = s.volume;
  DD(d, D_SP, e) = s.superpressure; DD(d, D_MOLS_AIR, e) = s.mols_air; DD(d, D_CHARGE, e) = s.charge;
  RR(d, R_ACS_W, e) = s.acs_power; RR(d, R_ACS_FLOW, e) = s.acs_flow;
  RR(d, R_SOLAR_W, e) = s.solar_w; RR(d, R_LOAD_W, e) = s.load_w;
  d.l[int64_t(L_DATE_TIME) * d.n + e] = s.date_time;
  d.l[int64_t(L_SUNRISE_H) * d.n + e] = ss.sunrise_h;
  d.l[int64_t(L_SUNSET) * d.n + e] = ss.sunset;
  d.t_elapsed[e] = s.time_elapsed;
  d.flags[e] = pack_flags(s.status, ss.last_command, ss.envelope_state, ss.altitude_state, ss.power_paused,
                          ss.power_safety_enabled, (atm.ok ? 0 : 1) | int((fl >> 11) & 1u));   // the error bit is sticky
  reward[e] = float(r);
  done[e] = (s.status != kOk) ? 1 : 0;
  if (wind_uv != nullptr) wind_uv[e] = make_float2(float(u), float(v));
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised physics step (production fp32 build).
//
// One block = 4 warps = 32 balloons; lane l of EVERY warp works on balloon 32*block + l and warp w
// plays role w of the sub-step (ble_physics.cuh): P pressure/atmosphere, T thermal body, E envelope +
// ACS, S sun + power.  All four keep the balloon's state in registers; per sub-step each computes
// only its own right-hand side, publishes it in shared memory (double-buffered, ONE barrier per
// sub-step) and then everybody applies the same update.  The dependent instruction chain per
// sub-step drops from ~1,400 to ~300 instructions, and the kernel has 4x the warps to hide latency.
// ---------------------------------------------------------------------------------------------
struct WsExchange {
  double p[2][32], tamb[2][32], vol[2][32], sp[2][32], mols[2][32], charge[2][32];
  float dt_body[2][32], dt_solar[2][32];
  int st_env[2][32], st_pwr[2][32];
  int eff[32];
  float cz[3][32], flux[2][32];
};

__global__ void __launch_bounds__(128, 4)
k_step_ws(DevState<float> d, const int32_t* __restrict__ actions, float* __restrict__ reward,
          uint8_t* __restrict__ done, float2* __restrict__ wind_uv) {
  using Real = float;
  __shared__ WsExchange ex;
  const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = int64_t(blockIdx.x) * 32 + lane;
  const bool valid = e < d.n;
  const int64_t ec = valid ? e : d.n - 1;                   // clamp: every thread stays for the barriers
  const uint32_t fl = d.flags[ec];
  bool live = valid && (fl & 3u) == uint32_t(kOk);
  if (valid && !live && role == 0) {                        // finished balloon: no-op (documented divergence)
    reward[e] = 0.f;
    done[e] = 1;
    if (wind_uv != nullptr) wind_uv[e] = make_float2(0.f, 0.f);
  }
  // state in registers (every role)
  BalloonState<Real> s;
  s.x = DD(d, D_X, ec); s.y = DD(d, D_Y, ec); s.pressure = DD(d, D_P, ec);
  s.t_ambient = DD(d, D_TAMB, ec); s.t_internal = DD(d, D_TINT, ec); s.volume = DD(d, D_VOL, ec);
  s.superpressure = DD(d, D_SP, ec); s.mols_air = DD(d, D_MOLS_AIR, ec); s.charge = DD(d, D_CHARGE, ec);
  s.acs_power = RR(d, R_ACS_W, ec); s.acs_flow = RR(d, R_ACS_FLOW, ec);
  s.solar_w = RR(d, R_SOLAR_W, ec); s.load_w = RR(d, R_LOAD_W, ec);
  s.lat0 = RR(d, R_LAT0, ec); s.lng0 = RR(d, R_LNG0, ec); s.ir = RR(d, R_IR, ec); s.mols_gas = RR(d, R_MOLS_GAS, ec);
  s.date_time = d.l[int64_t(L_DATE_TIME) * d.n + ec];
  s.time_elapsed = d.t_elapsed[ec];
  s.status = kOk;
  Real uf, vf;
  wind_at_balloon<Real>(d, ec, s.x, s.y, s.pressure, s.time_elapsed, &uf, &vf);   // PRE-step lookup (all roles)
  const double u = double(uf), v = double(vf);
  int action = actions[ec];
  action = action < 0 ? 0 : (action > 2 ? 2 : action);

  // ---- prologue: role 0 runs the safety layers, roles 1..3 one exact sun evaluation each ----
  Atmosphere atm;
  SafetyState ss;
  if (role == 0) {
    atm = load_atmosphere(d, ec);
    atm.incremental = true;
    ss.sunrise_h = d.l[int64_t(L_SUNRISE_H) * d.n + ec];
    ss.sunset = d.l[int64_t(L_SUNSET) * d.n + ec];
    ss.last_command = action; ss.envelope_state = int((fl >> 4) & 7);
    ss.altitude_state = int((fl >> 7) & 3); ss.power_paused = int((fl >> 9) & 1);
    ss.power_safety_enabled = int((fl >> 10) & 1);
    int eff = action;
    if (ss.power_safety_enabled) {
      eff = power_safety<double>(eff, s.date_time, s.charge, &ss.sunrise_h, &ss.sunset, &ss.power_paused);
    }
    eff = envelope_safety<double>(eff, s.superpressure, &ss.envelope_state);
    double altitude, t_unused;
    atm.at_pressure(s.pressure, &altitude, &t_unused);
    eff = altitude_safety<double>(eff, altitude, &ss.altitude_state);
    ex.eff[lane] = eff;
  } else {
    const int point = role - 1;                             // 0, 1, 2 -> t0, t0 + 90 s, t0 + 180 s
    const double dt_s = 90.0 * point;
    Real cz, flux;
    SunTrack<Real>::exact(s, s.x + u * dt_s, s.y + v * dt_s, s.date_time + int64_t(90 * point), &cz, &flux);
    ex.cz[point][lane] = cz;
    if (point != 1) ex.flux[point >> 1][lane] = flux;
  }
  __syncthreads();
  const int eff = ex.eff[lane];
  SunTrack<Real> sun;
  sun.c0 = ex.cz[0][lane]; sun.c1 = ex.cz[1][lane]; sun.c2 = ex.cz[2][lane];
  sun.f0 = ex.flux[0][lane]; sun.f2 = ex.flux[1][lane];
  const Real earth_per_area = earth_heat_per_area<Real>(s.ir);

  // ---- 18 sub-steps ----
  int n_done = 0, status = kOk;
  Real acs_power = s.acs_power, acs_flow = s.acs_flow, solar_w = s.solar_w, load_w = s.load_w;
#pragma unroll 1
  for (int k = 0; k < kSubSteps; ++k) {
    const int b = k & 1;
    if (live) {
      if (role == 0) {
        double np, nt;
        role_pressure<Real>(atm, s.pressure, s.t_ambient, s.volume, s.mols_air, double(s.mols_gas), &np, &nt);
        ex.p[b][lane] = np; ex.tamb[b][lane] = nt;
      } else if (role == 1) {
        ex.dt_body[b][lane] = role_thermal_body<Real>(s.volume, s.t_internal, s.t_ambient, s.pressure, earth_per_area);
      } else if (role == 2) {
        double nv, nsp, nm;
        int st;
        role_envelope_acs<Real>(double(s.mols_gas), s.mols_air, s.t_internal, s.pressure, s.superpressure, eff,
                                &nv, &nsp, &nm, &acs_power, &acs_flow, &st);
        ex.vol[b][lane] = nv; ex.sp[b][lane] = nsp; ex.mols[b][lane] = nm; ex.st_env[b][lane] = st;
      } else {
        SunAngles<Real> ang;
        Real flux, dts;
        double nc;
        int oop;
        sun.at(s, k, &ang, &flux);
        role_sun_power<Real>(ang, flux, s.volume, s.pressure, s.superpressure, s.charge, eff, &dts, &solar_w, &load_w,
                             &nc, &oop, &acs_power);
        ex.dt_solar[b][lane] = dts; ex.charge[b][lane] = nc; ex.st_pwr[b][lane] = oop;
      }
    }
    __syncthreads();
    if (live) {
      s.pressure = ex.p[b][lane]; s.t_ambient = ex.tamb[b][lane];
      s.t_internal = s.t_internal + double(ex.dt_body[b][lane] + ex.dt_solar[b][lane]) * double(kStrideS);
      s.volume = ex.vol[b][lane]; s.superpressure = ex.sp[b][lane]; s.mols_air = ex.mols[b][lane];
      s.charge = ex.charge[b][lane];
      ++n_done;
      const int st_env = ex.st_env[b][lane];
      status = ex.st_pwr[b][lane] ? int(kOutOfPower) : st_env;            // later assignment wins (:541-542)
      if (status != kOk) live = false;                                     // break (:327-328)
    }
  }

  // ---- epilogue: each role stores what it owns ----
  const bool stepped = valid && (fl & 3u) == uint32_t(kOk);
  if (!stepped) return;
  if (role == 0) {
    DD(d, D_X, e) = s.x + u * double(kStrideS) * double(n_done);
    DD(d, D_Y, e) = s.y + v * double(kStrideS) * double(n_done);
    DD(d, D_P, e) = s.pressure; DD(d, D_TAMB, e) = s.t_ambient;
    d.l[int64_t(L_DATE_TIME) * d.n + e] = s.date_time + int64_t(kStrideS) * n_done;
    d.l[int64_t(L_SUNRISE_H) * d.n + e] = ss.sunrise_h;
    d.l[int64_t(L_SUNSET) * d.n + e] = ss.sunset;
    d.t_elapsed[e] = s.time_elapsed + kStrideS * n_done;
    d.flags[e] = pack_flags(status, ss.last_command, ss.envelope_state, ss.altitude_state, ss.power_paused,
                            ss.power_safety_enabled, (atm.ok ? 0 : 1) | int((fl >> 11) & 1u));
    if (wind_uv != nullptr) wind_uv[e] = make_float2(uf, vf);
  } else if (role == 1) {
    DD(d, D_TINT, e) = s.t_internal;
  } else if (role == 2) {
    DD(d, D_VOL, e) = s.volume; DD(d, D_SP, e) = s.superpressure; DD(d, D_MOLS_AIR, e) = s.mols_air;
    RR(d, R_ACS_W, e) = acs_power; RR(d, R_ACS_FLOW, e) = acs_flow;
  } else {
    DD(d, D_CHARGE, e) = s.charge;
    RR(d, R_SOLAR_W, e) = solar_w; RR(d, R_LOAD_W, e) = load_w;
    // reward on the post-step state (env/balloon_env.py:44-102)
    s.x += u * double(kStrideS) * double(n_done);
    s.y += v * double(kStrideS) * double(n_done);
    s.acs_power = acs_power;
    SunAngles<Real> ang;
    ang.el = Real(0);
    Real flux;
    if (action == kDown) sun.at(s, n_done, &ang, &flux);
    reward[e] = perciatelli_reward<Real>(s, action, ang.el);
    done[e] = (status != kOk) ? 1 : 0;
  }
}

// Derived BalloonState properties (env/balloon/balloon.py:217-250) for the N = 1 adaptor / features.
template <typename Real>
__global__ void __launch_bounds__(128) k_derived(DevState<Real> d, double* __restrict__ out) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const int64_t n = d.n;
  const double x = DD(d, D_X, e), y = DD(d, D_Y, e), p = DD(d, D_P, e), sp = DD(d, D_SP, e);
  Real lat, lng, el, flux;
  latlng_from_offset<Real>(RR(d, R_LAT0, e), RR(d, R_LNG0, e), Real(x), Real(y), &lat, &lng);
  solar_calculator<Real>(lat, lng, d.l[int64_t(L_DATE_TIME) * n + e], &el, &flux);
  const double soc = DD(d, D_CHARGE, e) / kBatteryCapacityWh;
  const bool excess = (solar_power<Real>(el, Real(p)) > Real(kDayLoadW)) && (soc > 0.99);
  const uint32_t fl = d.flags[e];
  const bool paused = ((fl >> 9) & 1u) || (((fl >> 4) & 7u) != 0u) || (((fl >> 7) & 3u) != 0u);
  Atmosphere atm = load_atmosphere(d, e);
  double h, t;
  atm.at_pressure(p, &h, &t);
  out[int64_t(BLE_D_LAT) * n + e] = double(lat);
  out[int64_t(BLE_D_LNG) * n + e] = double(lng);
  out[int64_t(BLE_D_SOLAR_ELEVATION) * n + e] = double(el);
  out[int64_t(BLE_D_SOLAR_FLUX) * n + e] = double(flux);
  out[int64_t(BLE_D_EXCESS_ENERGY) * n + e] = excess ? 1.0 : 0.0;
  out[int64_t(BLE_D_NAVIGATION_IS_PAUSED) * n + e] = paused ? 1.0 : 0.0;
  out[int64_t(BLE_D_PRESSURE_RATIO) * n + e] = (p + fmax(sp, 0.0)) / p;
  out[int64_t(BLE_D_BATTERY_SOC) * n + e] = soc;
  out[int64_t(BLE_D_ALTITUDE) * n + e] = h;
}

#include "ble_feature_kernels.cuh"
#include "ble_gp_kernels.cuh"
#include "ble_gp_posterior.cuh"
#include "ble_decoder.cuh"

// ---------------------------------------------------------------------------------------------
// Evaluation surface (eval/eval_lib.py:123-211, agents/station_seeker_agent.py:72-113)
// ---------------------------------------------------------------------------------------------
// One warp per balloon: lanes score levels lane, lane + 32, ... of the 361-level column, then the
// warp reduces to the FIRST level holding the largest score (the reference's strict '>' scan).
__global__ void __launch_bounds__(128)
k_agent_station_seeker(const float* __restrict__ obs, int64_t n, int32_t* __restrict__ actions,
                       int32_t* __restrict__ best_level) {
  const int64_t e = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = int(threadIdx.x & 31u);
  if (e >= n) return;
  const float* o = obs + e * int64_t(kNumFeatures);
  const SeekerDistanceTerms t = seeker_distance_terms(o[7]);
  double score = 0.0;
  int best = kColumnLevels;
  for (int l = lane; l < kColumnLevels; l += 32) {
    const float w0 = o[16 + 3 * l], w1 = o[17 + 3 * l], w2 = o[18 + 3 * l];
    if (!level_is_valid(w0, w1, w2)) continue;
    const double sc = seeker_altitude_score(t, w0, w1, w2, l);
    if (sc > score) { score = sc; best = l; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double os = __shfl_down_sync(0xffffffffu, score, off);
    const int ob = __shfl_down_sync(0xffffffffu, best, off);
    if (os > score || (os == score && ob < best)) { score = os; best = ob; }
  }
  if (lane == 0) {
    const bool none = best >= kColumnLevels;       // the reference asserts here (:109-110); callers check best < 0
    actions[e] = none ? 1 : seeker_action_for_level(best);
    if (best_level != nullptr) best_level[e] = none ? -1 : best;
  }
}

// RandomWalkAgent (agents/random_walk_agent.py:35-94): the target pressure performs a Gaussian random
// walk whose step grows with the time elapsed in the episode (:82-90); Philox replaces jax.random.
struct EvalBuffers { double* reward; int32_t* within; int32_t* steps; uint8_t* active; };

template <typename Real>
__global__ void __launch_bounds__(128) k_eval_begin(DevState<Real> d, EvalBuffers ev) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  ev.reward[e] = 0.0; ev.within[e] = 0; ev.steps[e] = 0;
  ev.active[e] = (d.flags[e] & 3u) == uint32_t(kOk) ? 1 : 0;
}

// One pass of eval_agent's loop body (:166-185) for every balloon still flying: reward sum, steps within
// the station-keeping radius (:119-121), step count, flight-path sample (:64-80), stop at a terminal state.
template <typename Real>
__global__ void __launch_bounds__(128)
k_eval_accumulate(DevState<Real> d, EvalBuffers ev, const float* __restrict__ reward, double radius_m,
                  float* __restrict__ path /* [6][n] or nullptr */) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const bool live = ev.active[e] != 0;
  const double x = DD(d, D_X, e), y = DD(d, D_Y, e);
  if (path != nullptr) {
    const int64_t n = d.n;
    const float nanv = __int_as_float(0x7fc00000);
    path[0 * n + e] = live ? float(x * 1e-3) : nanv;
    path[1 * n + e] = live ? float(y * 1e-3) : nanv;
    path[2 * n + e] = live ? float(DD(d, D_P, e)) : nanv;
    path[3 * n + e] = live ? float(DD(d, D_SP, e)) : nanv;
    path[4 * n + e] = live ? float(d.t_elapsed[e]) : nanv;
    path[5 * n + e] = live ? float(DD(d, D_CHARGE, e) / kBatteryCapacityWh) : nanv;
  }
  if (!live) return;
  ev.reward[e] += double(reward[e]);
  ev.within[e] += (sqrt(x * x + y * y) <= radius_m) ? 1 : 0;
  ev.steps[e] += 1;
  if ((d.flags[e] & 3u) != uint32_t(kOk)) ev.active[e] = 0;
}

template <typename Real>
__global__ void __launch_bounds__(128) k_eval_results(DevState<Real> d, EvalBuffers ev, double* __restrict__ out) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const int64_t n = d.n;
  const uint32_t status = d.flags[e] & 3u;
  const int steps = ev.steps[e];
  out[int64_t(BLE_E_CUMULATIVE_REWARD) * n + e] = ev.reward[e];
  out[int64_t(BLE_E_TIME_WITHIN_RADIUS) * n + e] = steps > 0 ? double(ev.within[e]) / double(steps) : 0.0;
  out[int64_t(BLE_E_OUT_OF_POWER) * n + e] = status == uint32_t(kOutOfPower) ? 1.0 : 0.0;
  out[int64_t(BLE_E_ENVELOPE_BURST) * n + e] = status == uint32_t(kBurst) ? 1.0 : 0.0;
  out[int64_t(BLE_E_ZEROPRESSURE) * n + e] = status == uint32_t(kZeroPressure) ? 1.0 : 0.0;
  out[int64_t(BLE_E_FINAL_TIMESTEP) * n + e] = double(steps);
  out[int64_t(BLE_E_ACTIVE) * n + e] = ev.active[e] ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Reset (env/balloon_arena.py:161-182,228-268; utils/sampling.py:37-152)
// ---------------------------------------------------------------------------------------------
#include "ble_rng.cuh"

// Deterministic derived state: power-safety sunrise/sunset (+30 min hysteresis) and, optionally,
// the stable-init solve.  Always fp64 (see ble_physics.cuh).
template <typename Real>
__device__ void init_derived_one(DevState<Real>& d, int64_t e, bool run_stable_init) {
  const double alpha = DD(d, D_ALPHA, e);
  const double x = DD(d, D_X, e), y = DD(d, D_Y, e);
  const double lat0 = double(RR(d, R_LAT0, e)), lng0 = double(RR(d, R_LNG0, e));
  const int64_t ts = d.l[int64_t(L_DATE_TIME) * d.n + e];
  double lat, lng;
  latlng_from_offset<double>(lat0, lng0, x, y, &lat, &lng);
  int64_t sunrise, sunset;
  const bool ok = next_sunrise_sunset(lat, lng, ts, &sunrise, &sunset);   // PowerSafetyLayer.__init__
  d.l[int64_t(L_SUNRISE_H) * d.n + e] = sunrise + 30 * 60;                // power_safety.py:45-50
  d.l[int64_t(L_SUNSET) * d.n + e] = sunset;
  uint32_t fl = d.flags[e];
  fl &= ~((7u << 4) | (3u << 7) | (1u << 9));     // envelope/altitude NOMINAL, not paused (balloon.py:210-215)
  if (!ok) fl |= (1u << 11);
  if (run_stable_init) {                           // stable_init.cold_start_to_stable_params :132-157
    const StableParams sp = stable_params(alpha, DD(d, D_P, e), double(RR(d, R_MOLS_GAS, e)),
                                          lat, lng, ts, double(RR(d, R_IR, e)));
    DD(d, D_TAMB, e) = sp.t_ambient; DD(d, D_TINT, e) = sp.t_internal;
    DD(d, D_MOLS_AIR, e) = sp.mols_air; DD(d, D_VOL, e) = sp.volume;
    DD(d, D_SP, e) = sp.superpressure;
    if (!sp.ok) fl |= (1u << 11);
  }
  d.flags[e] = fl;
}

template <typename Real>
__global__ void __launch_bounds__(128) k_init_derived(DevState<Real> d, int run_stable_init) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  init_derived_one<Real>(d, e, run_stable_init != 0);
}

template <typename Real>
__global__ void __launch_bounds__(128)
k_reset(DevState<Real> d, const uint64_t* __restrict__ seeds, const uint8_t* __restrict__ mask,
        int64_t* __restrict__ noise_seeds /*[n,2,5]*/, float* __restrict__ noise_offsets /*[n,2,5,4]*/) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  if (mask != nullptr && mask[e] == 0) return;
  Philox rng;
  rng.init(seeds[e], 0);            // the episode is a function of its seed alone (eval suites shard by seed)
  // Atmosphere.reset: alpha ~ U(0,1)  (standard_atmosphere.py:82)
  const double alpha = rng.uniform();
  // sampling.sample_time: uniform second in [2011-01-01, 2014-12-31)  (utils/sampling.py:65-83)
  const int64_t t0 = 1293840000;               // 2011-01-01T00:00:00Z
  const int64_t range = 126144000;             // 1460 days
  const int64_t ts = t0 + int64_t(rng.uniform() * double(range));
  // _initialize_balloon (env/balloon_arena.py:228-268)
  const double ga = rng.gamma(1.2), gb = rng.gamma(2.0);
  const double radius_m = 200.0e3 * (ga / (ga + gb));                      // 200 km * Beta(1.2, 2.0)
  const double theta = rng.uniform() * 2.0 * kPi;
  const double x = cos(theta) * radius_m, y = sin(theta) * radius_m;
  const double lat_deg = -10.0 + 20.0 * rng.uniform();                     // sample_location :37-61
  const double lng_deg = -175.0 + 350.0 * rng.uniform();
  double pmax, tmp;
  atm_at_height_generic(alpha, kAltMin, &pmax, &tmp);                      // sample_pressure :86-111
  const double pressure = 6500.0 + (pmax - 6500.0) * rng.uniform();
  double ir;
  do {                                                                     // sample_upwelling_infrared :114-152
    const double z = 2.0 + 315.0 * rng.normal();
    ir = 315.0 * (1.0 / (1.0 + exp(-z)));
  } while (!(ir >= 225.0));

  Atmosphere atm; atm.init(alpha);
  store_atmosphere(d, e, atm);
  // BalloonState defaults (env/balloon/balloon.py:175-208)
  DD(d, D_X, e) = x; DD(d, D_Y, e) = y; DD(d, D_P, e) = pressure;
  DD(d, D_TAMB, e) = 206.0; DD(d, D_TINT, e) = 206.0; DD(d, D_VOL, e) = 1804.0;
  DD(d, D_SP, e) = 0.0; DD(d, D_MOLS_AIR, e) = 0.0; DD(d, D_CHARGE, e) = 2905.6;
  RR(d, R_ACS_W, e) = Real(0); RR(d, R_ACS_FLOW, e) = Real(0); RR(d, R_SOLAR_W, e) = Real(0);
  RR(d, R_LOAD_W, e) = Real(0);
  RR(d, R_LAT0, e) = Real(lat_deg * (kPi / 180.0)); RR(d, R_LNG0, e) = Real(lng_deg * (kPi / 180.0));
  RR(d, R_IR, e) = Real(ir); RR(d, R_MOLS_GAS, e) = Real(6830.0);
  d.l[int64_t(L_DATE_TIME) * d.n + e] = ts;
  d.t_elapsed[e] = 0;
  d.flags[e] = pack_flags(kOk, kStay, kEnvNominal, kAltNominal, 0, 1, 0);
  if (d.gp_count != nullptr) { d.gp_count[e] = 0; d.gp_m[e] = 0; d.gp_first[e] = 0; }   // new FeatureConstructor (env/balloon_arena.py:179-182)
  init_derived_one<Real>(d, e, true);
  // SimplexWindNoise.reset_wind_noise (simplex_wind_noise.py:98-114)
  for (int h = 0; h < 10; ++h) {
    noise_seeds[e * 10 + h] = int64_t(rng.uniform() * 1634753849.0);
    for (int c = 0; c < 4; ++c) noise_offsets[(e * 10 + h) * 4 + c] = rng.uniform_f32() * 2.0f - 1.0f;
  }
}

// GenerativeWindFieldSampler.sample_field (env/generative_wind_field.py:52-62): z ~ N(0, I_64) per seed.
__global__ void __launch_bounds__(256)
k_sample_latents(const uint64_t* __restrict__ seeds, int64_t count, float* __restrict__ latents /*[count,64]*/) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= count * kDecLatents) return;
  Philox rng;
  rng.init(seeds[i / kDecLatents], 1 + uint64_t(i % kDecLatents));        // stream 0 belongs to k_reset
  latents[i] = float(rng.normal());
}

// RandomWalkAgent (agents/random_walk_agent.py:35-94): begin_episode draws the target pressure
// (sampling.sample_pressure without atmosphere: U(6500, 11400), utils/sampling.py:84-111); every later
// step adds time_elapsed_s * 0.1666 * N(0, 1) (:82-90).  Philox keyed by (seed, step) replaces jax.random.
__global__ void __launch_bounds__(128)
k_agent_random_walk(const float* __restrict__ obs, int64_t n, double* __restrict__ target,
                    const uint64_t* __restrict__ seeds, int32_t step_index, int32_t* __restrict__ actions) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n) return;
  Philox rng;
  rng.init(seeds[e], (uint64_t(1) << 32) + uint64_t(step_index));
  double t;
  if (step_index == 0) t = 6500.0 + (11400.0 - 6500.0) * rng.uniform();
  else t = target[e] + double(step_index) * kAgentStepS * 0.1666 * rng.normal();
  target[e] = t;
  actions[e] = random_walk_action(obs[e * int64_t(kNumFeatures)], t);
}

// info of BalloonEnv.step (env/balloon_env.py:280-290) for the first-generation kernels (k_step_fused writes it itself)
template <typename Real>
__global__ void __launch_bounds__(128)
k_step_info(DevState<Real> d, uint8_t* __restrict__ status, int32_t* __restrict__ time_elapsed, uint8_t* __restrict__ sim_error) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= d.n) return;
  const uint32_t fl = d.flags[e];
  if (status != nullptr) status[e] = uint8_t(fl & 3u);
  if (time_elapsed != nullptr) time_elapsed[e] = d.t_elapsed[e];
  if (sim_error != nullptr) sim_error[e] = uint8_t((fl >> 11) & 1u);
}

// ---------------------------------------------------------------------------------------------
// Host-side engine
// ---------------------------------------------------------------------------------------------
// Entry points run on the handle's device and hand the calling thread's current device back on return.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
struct EngineBase {
  virtual ~EngineBase() {}
  virtual int upload_fields(const float*, int64_t, const int32_t*, cudaStream_t) = 0;
  virtual int alloc_fields(int64_t, cudaStream_t) = 0;
  virtual int write_fields(const float*, int64_t, int64_t, cudaStream_t) = 0;
  virtual int set_field_map(const int32_t*, cudaStream_t) = 0;
  virtual int set_noise(const int64_t*, const float*, const uint8_t*, cudaStream_t) = 0;
  virtual int state_upload(const ble_state_soa*, cudaStream_t) = 0;
  virtual int state_download(ble_state_soa*, cudaStream_t) = 0;
  virtual int reset(const uint64_t*, const uint8_t*, cudaStream_t) = 0;
  virtual int init_derived(int, cudaStream_t) = 0;
  virtual int step(const int32_t*, float*, uint8_t*, float*, cudaStream_t) = 0;
  virtual int step_ex(const int32_t*, const ble_step_out*, int, cudaStream_t) = 0;
  virtual int step_host(const int32_t*, float*, uint8_t*, cudaStream_t) = 0;
  virtual int wind_at(float*, cudaStream_t) = 0;
  virtual int wind_gather(const float*, const int32_t*, float*, int64_t, cudaStream_t) = 0;
  virtual int derived(double*, cudaStream_t) = 0;
  virtual int wind_query(const double*, const int32_t*, int, float*, int64_t, cudaStream_t) = 0;
  virtual int atmosphere_query(int, const double*, const int32_t*, double*, int64_t, cudaStream_t) = 0;
  virtual int set_decoder(const float* const*, const float* const*, cudaStream_t) = 0;
  virtual int decode(const float*, int64_t, float*, cudaStream_t) = 0;
  virtual int generate_fields(const uint64_t*, int64_t, int64_t, cudaStream_t) = 0;
  virtual int generate_fields_at(const uint64_t*, int64_t, int64_t, const int32_t*, cudaStream_t) = 0;
  virtual int sample_latents(const uint64_t*, int64_t, float*, cudaStream_t) = 0;
  virtual int agent_station_seeker(const float*, int32_t*, int32_t*, cudaStream_t) = 0;
  virtual int agent_random_walk(const float*, const uint64_t*, int32_t, int32_t*, cudaStream_t) = 0;
  virtual int eval_begin(cudaStream_t) = 0;
  virtual int eval_accumulate(const float*, float*, cudaStream_t) = 0;
  virtual int eval_results(double*, cudaStream_t) = 0;
  virtual int features_observe(cudaStream_t) = 0;
  virtual int features(float*, cudaStream_t) = 0;
  virtual int features_clear(const uint8_t*, cudaStream_t) = 0;
  virtual int features_track(int) = 0;
  int64_t n = 0;
  int64_t launches = 0;
  std::string err;
};

#define BLE_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
      return e_ == cudaErrorMemoryAllocation ? BLE_ERR_OUT_OF_MEMORY : BLE_ERR_CUDA;       \
    }                                                                                      \
  } while (0)

#define BLE_DEVICE_GUARD()                                                                 \
  DeviceGuard device_guard_(device);                                                       \
  if (!device_guard_.ok) { err = "cudaSetDevice failed"; return BLE_ERR_CUDA; }

template <typename Real>
struct Engine : EngineBase {
  DevState<Real> d{};
  ble_config cfg{};
  int device = 0;
  float* cells = nullptr; int64_t n_fields = 0;
  CUtensorMap bank_map{};          // 5-D TMA view of the field bank (k_gp_posterior's column tile); valid iff tile_column
  int tile_column = 0;
  int32_t* env_field = nullptr;
  uint8_t* perm = nullptr; float* offsets = nullptr;
  int64_t* noise_seeds = nullptr; float* noise_offsets_in = nullptr;
  bool have_state = false, have_fields = false, have_noise = false;
  bool noise_valid = false;      // noise_partial matches the current state
  cudaStream_t noise_stream = nullptr;   // stream the valid noise_partial was (or is being) produced on
  double* gp_obs = nullptr; int32_t* gp_count = nullptr; double* gp_chol = nullptr; int32_t* gp_m = nullptr;
  int32_t* gp_first = nullptr; double* gp_z = nullptr; double* range_scratch = nullptr;
  bool host_zero_copy = true;            // BLE_HOST_ZERO_COPY=0: staged copies in ble_step_host instead of mapped pinned memory
  bool host_prefetch_noise = true;       // BLE_HOST_PREFETCH_NOISE=0: ble_step_host does not queue the next step's noise kernel
  bool track_measurements = true;        // ble_features_track: append a WindGP measurement after every reset / step
  uint64_t* episode_seed = nullptr; uint8_t* auto_mask = nullptr;   // ble_config.auto_reset: seed chain and the done mask of the last step
  bool gp_refit_every_step = false;      // BLE_GP_REFIT=1: the first-generation kernels (full refit per call), kept for A/B checks
  double* feat_range = nullptr;
  // VAE decoder (reset path)
  float* dec_w[4] = {nullptr, nullptr, nullptr, nullptr};
  float* dec_b[4] = {nullptr, nullptr, nullptr, nullptr};
  float* dec_act[2] = {nullptr, nullptr};
  void* dec_workspace = nullptr;
  cublasLtHandle_t lt = nullptr;
  bool have_decoder = false;
  static constexpr int64_t kDecChunk = 4096;
  static constexpr int64_t kGenChunk = 2048;                 // fields decoded per pass of generate_fields
  float* gen_latents = nullptr; float* gen_flow[2] = {nullptr, nullptr};
  cudaStream_t gen_stream = nullptr, feat_stream = nullptr;
  cudaEvent_t ev_feat_fork = nullptr, ev_feat_join = nullptr;
  cudaEvent_t ev_flow[2] = {nullptr, nullptr}, ev_written[2] = {nullptr, nullptr}, ev_fork = nullptr;
  // evaluation surface
  EvalBuffers ev{nullptr, nullptr, nullptr, nullptr};
  double* walk_target = nullptr;
  static constexpr size_t kDecWorkspace = size_t(32) << 20;
  // ble_step_host staging
  int32_t* h_actions = nullptr; float* h_reward = nullptr; uint8_t* h_done = nullptr;
  int32_t* d_actions = nullptr; float* d_reward = nullptr; uint8_t* d_done = nullptr;

  int create(int dev, int64_t n_envs, const ble_config& c) {
    device = dev; n = n_envs; cfg = c;
    BLE_DEVICE_GUARD();
    if (const char* z = std::getenv("BLE_HOST_ZERO_COPY")) host_zero_copy = std::atoi(z) != 0;
    if (const char* z = std::getenv("BLE_HOST_PREFETCH_NOISE")) host_prefetch_noise = std::atoi(z) != 0;
    if (const char* g = std::getenv("BLE_L2_FETCH_GRANULARITY")) {     // experiment knob: 32 / 64 / 128
      BLE_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(std::atoi(g))));
    }
    d.n = n;
    BLE_CUDA(cudaMalloc(&d.dd, sizeof(double) * D_COUNT * n));
    BLE_CUDA(cudaMalloc(&d.r, sizeof(Real) * R_COUNT * n));
    BLE_CUDA(cudaMalloc(&d.l, sizeof(int64_t) * L_COUNT * n));
    BLE_CUDA(cudaMalloc(&d.t_elapsed, sizeof(int32_t) * n));
    BLE_CUDA(cudaMalloc(&d.flags, sizeof(uint32_t) * n));
    BLE_CUDA(cudaMalloc(&env_field, sizeof(int32_t) * n));
    BLE_CUDA(cudaMemset(env_field, 0, sizeof(int32_t) * n));
    BLE_CUDA(cudaMalloc(&d.noise_partial, sizeof(Real) * 10 * n));
    BLE_CUDA(cudaMemset(d.noise_partial, 0, sizeof(Real) * 10 * n));
    BLE_CUDA(cudaMalloc(&d_actions, sizeof(int32_t) * n));
    BLE_CUDA(cudaMalloc(&d_reward, sizeof(float) * n + sizeof(uint8_t) * n));      // reward [n] then done [n]: one D2H copy
    d_done = reinterpret_cast<uint8_t*>(d_reward + n);
    BLE_CUDA(cudaMallocHost(&h_actions, sizeof(int32_t) * n));
    BLE_CUDA(cudaMallocHost(&h_reward, sizeof(float) * n + sizeof(uint8_t) * n));
    h_done = reinterpret_cast<uint8_t*>(h_reward + n);
    d.env_field = env_field;
    d.wind_model = cfg.wind_model;
    d.layout = make_layout(cfg.field_layout);
    d.enable_noise = 0;
    BLE_CUDA(cudaFuncSetAttribute(k_noise<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNoiseBlock * 256));
    BLE_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    if (std::is_same<Real, float>::value) BLE_CUDA(fused_setup(fused_blocks_per_sm));
    if (cfg.enable_features) {
      BLE_CUDA(cudaMalloc(&gp_obs, sizeof(double) * size_t(kGpWindow) * 6 * n));
      BLE_CUDA(cudaMalloc(&gp_count, sizeof(int32_t) * n));
      BLE_CUDA(cudaMemset(gp_count, 0, sizeof(int32_t) * n));
      BLE_CUDA(cudaMalloc(&gp_chol, sizeof(double) * size_t(kGpFactorDoubles) * n));
      BLE_CUDA(cudaMalloc(&gp_m, sizeof(int32_t) * n));
      BLE_CUDA(cudaMemset(gp_m, 0, sizeof(int32_t) * n));
      BLE_CUDA(cudaMalloc(&gp_first, sizeof(int32_t) * n));
      BLE_CUDA(cudaMemset(gp_first, 0, sizeof(int32_t) * n));
      BLE_CUDA(cudaMalloc(&gp_z, sizeof(double) * size_t(kGpWindow) * 2 * n));
      BLE_CUDA(cudaMalloc(&range_scratch, sizeof(double) * size_t(kRangeLevels) * 2 * n));
      d.gp_first = gp_first; d.gp_z = gp_z;
      if (const char* g = std::getenv("BLE_GP_REFIT")) gp_refit_every_step = std::atoi(g) != 0;
      BLE_CUDA(cudaFuncSetAttribute(k_gp_posterior<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(PosteriorSmem))));
      BLE_CUDA(cudaFuncSetAttribute(k_gp_posterior<Real>, cudaFuncAttributePreferredSharedMemoryCarveout, int(cudaSharedmemCarveoutMaxShared)));


      BLE_CUDA(cudaMalloc(&feat_range, sizeof(double) * 2 * n));
      d.gp_obs = gp_obs; d.gp_count = gp_count; d.gp_chol = gp_chol; d.gp_m = gp_m; d.feat_range = feat_range;
      BLE_CUDA(cudaFuncSetAttribute(k_gp_factor<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    int(sizeof(double) * (kGpPacked + kGpWindow * 4))));
      BLE_CUDA(cudaFuncSetAttribute(k_gp_column<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kColumnSmem)));
    }
    return BLE_OK;
  }

  ~Engine() override {
    DeviceGuard device_guard_(device);
    cudaFree(d.dd); cudaFree(d.r); cudaFree(d.l); cudaFree(d.t_elapsed); cudaFree(d.flags);
    cudaFree(env_field); cudaFree(d.noise_partial); cudaFree(cells); cudaFree(perm); cudaFree(offsets);
    cudaFree(noise_seeds); cudaFree(noise_offsets_in);
    for (int i = 0; i < 4; ++i) { cudaFree(dec_w[i]); cudaFree(dec_b[i]); }
    cudaFree(dec_act[0]); cudaFree(dec_act[1]); cudaFree(dec_workspace);
    if (lt != nullptr) cublasLtDestroy(lt);
    cudaFree(gp_obs); cudaFree(gp_count); cudaFree(gp_chol); cudaFree(gp_m); cudaFree(feat_range);
    cudaFree(gp_first); cudaFree(gp_z); cudaFree(range_scratch);
    cudaFree(gen_latents); cudaFree(gen_flow[0]); cudaFree(gen_flow[1]); cudaFree(episode_seed); cudaFree(auto_mask);
    for (int b = 0; b < 2; ++b) { if (ev_flow[b]) cudaEventDestroy(ev_flow[b]); if (ev_written[b]) cudaEventDestroy(ev_written[b]); }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (gen_stream) cudaStreamDestroy(gen_stream);
    if (feat_stream) cudaStreamDestroy(feat_stream);
    if (ev_feat_fork) cudaEventDestroy(ev_feat_fork);
    if (ev_feat_join) cudaEventDestroy(ev_feat_join);
    cudaFree(ev.reward); cudaFree(ev.within); cudaFree(ev.steps); cudaFree(ev.active); cudaFree(walk_target);
    cudaFree(d_actions); cudaFree(d_reward);
    cudaFreeHost(h_actions); cudaFreeHost(h_reward);
  }

  static unsigned grid_for(int64_t items, int block) { return unsigned((items + block - 1) / block); }

  int alloc_fields(int64_t nf, cudaStream_t s) override {
    if (nf <= 0) { err = "alloc_fields: n_fields must be > 0"; return BLE_ERR_INVALID_ARGUMENT; }
    BLE_DEVICE_GUARD();
    if (nf != n_fields) {
      BLE_CUDA(cudaStreamSynchronize(s));
      cudaFree(cells); cells = nullptr; n_fields = 0; have_fields = false;
      BLE_CUDA(cudaMalloc(&cells, sizeof(float) * size_t(d.layout.field_floats) * size_t(nf)));
      n_fields = nf;
    }
    d.cells = cells;
    return encode_bank_map();
  }

  // TMA tensor map of the bank as [field][y-cell][p-cell][t-cell][row floats]: a box {128 B, 1, 9, 1, 1} is the forecast
  // column of one (x, y, t) cell -- the 9 lookup windows every pressure level of a balloon's column interpolates in.
  int encode_bank_map() {
    tile_column = 0;
    const char* env = std::getenv("BLE_COLUMN_TILE");            // "0": per-level gathers instead (A/B)
    if (env != nullptr && std::strcmp(env, "0") == 0) return BLE_OK;
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult found;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &found) != cudaSuccess ||
        found != cudaDriverEntryPointSuccess || fn == nullptr) {
      err = "alloc_fields: the driver does not export cuTensorMapEncodeTiled";
      return BLE_ERR_CUDA;
    }
    const cuuint64_t row = cuuint64_t(d.layout.row_floats);
    const cuuint64_t dims[5] = {row, cuuint64_t(kTC), cuuint64_t(kPC), cuuint64_t(kYC), cuuint64_t(n_fields)};
    const cuuint64_t strides[4] = {row * 4, row * 4 * kTC, row * 4 * kTC * kPC, cuuint64_t(d.layout.field_floats) * 4};   // bytes
    const cuuint32_t box[5] = {32, 1, cuuint32_t(kPC), 1, 1};                                                           // 128 B x 9
    const cuuint32_t elem[5] = {1, 1, 1, 1, 1};
    const CUresult rc = reinterpret_cast<EncodeTiled>(fn)(&bank_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, cells, dims, strides, box, elem,
                                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { err = "alloc_fields: cuTensorMapEncodeTiled failed (code " + std::to_string(int(rc)) + ")"; return BLE_ERR_CUDA; }
    tile_column = 1;
    return BLE_OK;
  }

  int write_fields(const float* fields, int64_t first, int64_t count, cudaStream_t s) override {
    return write_fields_at(fields, first, count, nullptr, s);
  }

  // dst_index (device int32 [count], every entry in [0, n_fields)) scatters the fields; nullptr = contiguous
  int write_fields_at(const float* fields, int64_t first, int64_t count, const int32_t* dst_index, cudaStream_t s) {
    if (fields == nullptr || first < 0 || count <= 0 || (dst_index == nullptr && first + count > n_fields) || cells == nullptr) {
      err = "write_fields: range outside the allocated fields (call ble_alloc_fields first)";
      return BLE_ERR_INVALID_ARGUMENT;
    }
    BLE_DEVICE_GUARD();
    const int64_t threads = count * int64_t(kYC) * kPC * kTC * (d.layout.row_floats / 16);
    k_fields_to_windows<<<grid_for(threads, 256), 256, 0, s>>>(fields, cells, d.layout, first, count, dst_index);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    have_fields = true;
    return BLE_OK;
  }

  int set_field_map(const int32_t* map, cudaStream_t s) override {
    if (map == nullptr) { err = "set_field_map: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    BLE_DEVICE_GUARD();
    BLE_CUDA(cudaMemcpyAsync(env_field, map, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, s));
    return BLE_OK;
  }

  int upload_fields(const float* fields, int64_t nf, const int32_t* map, cudaStream_t s) override {
    if (fields == nullptr || nf <= 0) { err = "upload_fields: fields must be non-null, n_fields > 0"; return BLE_ERR_INVALID_ARGUMENT; }
    int rc = alloc_fields(nf, s);
    if (rc != BLE_OK) return rc;
    rc = write_fields(fields, 0, nf, s);
    if (rc != BLE_OK) return rc;
    return map != nullptr ? set_field_map(map, s) : BLE_OK;
  }

  int ensure_noise_buffers() {
    if (perm == nullptr) {
      BLE_CUDA(cudaMalloc(&perm, size_t(10) * n * 256));
      BLE_CUDA(cudaMalloc(&offsets, sizeof(float) * 40 * n));
      d.perm = perm; d.offsets = offsets;
    }
    return BLE_OK;
  }

  int set_noise(const int64_t* seeds, const float* offs, const uint8_t* mask, cudaStream_t s) override {
    if (seeds == nullptr || offs == nullptr) { err = "set_noise: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    BLE_DEVICE_GUARD();
    int rc = ensure_noise_buffers();
    if (rc != BLE_OK) return rc;
    k_make_perms<<<dim3(grid_for(n, kPermCta), 10), kPermCta, 0, s>>>(n, seeds, offs, mask, perm, offsets);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    have_noise = true;
    d.enable_noise = cfg.enable_noise ? 1 : 0;
    noise_valid = false;
    return BLE_OK;
  }

  int state_upload(const ble_state_soa* st, cudaStream_t s) override {
    if (st == nullptr || st->f64 == nullptr || st->i64 == nullptr) { err = "state_upload: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    BLE_DEVICE_GUARD();
    k_state_upload<Real><<<grid_for(n, 128), 128, 0, s>>>(d, st->f64, st->i64);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    have_state = true;
    noise_valid = false;
    return BLE_OK;
  }

  int state_download(ble_state_soa* st, cudaStream_t s) override {
    if (st == nullptr || st->f64 == nullptr || st->i64 == nullptr) { err = "state_download: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    if (!have_state) { err = "state_download: no state (call ble_reset or ble_state_upload first)"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    k_state_download<Real><<<grid_for(n, 128), 128, 0, s>>>(d, st->f64, st->i64);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int reset(const uint64_t* seeds, const uint8_t* mask, cudaStream_t s) override {
    if (seeds == nullptr) { err = "reset: seeds must be non-null"; return BLE_ERR_INVALID_ARGUMENT; }
    if (mask != nullptr && !have_state) { err = "reset: a masked reset needs an initial full reset"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    int rc = ensure_noise_buffers();
    if (rc != BLE_OK) return rc;
    if (noise_seeds == nullptr) {
      BLE_CUDA(cudaMalloc(&noise_seeds, sizeof(int64_t) * 10 * n));
      BLE_CUDA(cudaMalloc(&noise_offsets_in, sizeof(float) * 40 * n));
    }
    if (cfg.auto_reset) {
      if (episode_seed == nullptr) {
        BLE_CUDA(cudaMalloc(&episode_seed, sizeof(uint64_t) * n));
        BLE_CUDA(cudaMalloc(&auto_mask, n));
      }
      if (seeds != episode_seed) {
        k_keep_episode_seeds<<<grid_for(n, 256), 256, 0, s>>>(n, seeds, mask, episode_seed);
        ++launches;
      }
    }
    k_reset<Real><<<grid_for(n, 128), 128, 0, s>>>(d, seeds, mask, noise_seeds, noise_offsets_in);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    k_make_perms<<<dim3(grid_for(n, kPermCta), 10), kPermCta, 0, s>>>(n, noise_seeds, noise_offsets_in, mask, perm, offsets);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    have_state = true; have_noise = true;
    d.enable_noise = cfg.enable_noise ? 1 : 0;
    noise_valid = false;
    // arena.reset ends with feature_constructor.observe(get_measurements()) (env/balloon_arena.py:179-182)
    if (cfg.enable_features && track_measurements && (cfg.wind_model != BLE_WIND_GRID || have_fields)) return features_observe(s);
    return BLE_OK;
  }

  int init_derived(int run_stable_init, cudaStream_t s) override {
    if (!have_state) { err = "init_derived: no state uploaded"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    k_init_derived<Real><<<grid_for(n, 128), 128, 0, s>>>(d, run_stable_init);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int check_ready(const char* who) {
    if (!have_state) { err = std::string(who) + ": no balloon state (call ble_reset or ble_state_upload)"; return BLE_ERR_NOT_READY; }
    if (cfg.wind_model == BLE_WIND_GRID && !have_fields) {
      err = std::string(who) + ": no wind fields (call ble_upload_fields before stepping)";   // grid_based_wind_field.py:86-87
      return BLE_ERR_NOT_READY;
    }
    if (cfg.enable_noise && !have_noise) { err = std::string(who) + ": noise enabled but not seeded (ble_set_noise / ble_reset)"; return BLE_ERR_NOT_READY; }
    return BLE_OK;
  }

  int launch_noise(cudaStream_t s) {
    if (!d.enable_noise) return BLE_OK;
    if (noise_valid) {
      // produced ahead of time (features_observe, step_host's prefetch): a consumer on another stream waits for it
      if (noise_stream != s && cudaStreamSynchronize(noise_stream) != cudaSuccess) {
        cudaGetLastError();                                   // the producing stream was destroyed meanwhile
        BLE_CUDA(cudaDeviceSynchronize());
      }
      return BLE_OK;
    }
    dim3 grid(grid_for(n, kNoiseBlock), 10);
    k_noise<Real><<<grid, kNoiseBlock, kNoiseBlock * 256, s>>>(d);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    noise_valid = true;
    noise_stream = s;
    return BLE_OK;
  }

  int step(const int32_t* actions, float* reward, uint8_t* done, float* wind_uv, cudaStream_t s) override {
    ble_step_out out{};
    out.reward = reward; out.done = done; out.wind_uv = wind_uv;
    return step_ex(actions, &out, 1, s);
  }

  // BalloonEnv.step with the info outputs (env/balloon_env.py:280-290) written by the step kernel itself;
  // n_steps > 1 = ble_rollout: actions / reward / done are [n_steps][N] and the steps run inside ONE launch.
  int step_ex(const int32_t* actions, const ble_step_out* o, int n_steps, cudaStream_t s) override {
    if (actions == nullptr || o == nullptr || o->reward == nullptr || o->done == nullptr || n_steps < 1) {
      err = "step: null argument"; return BLE_ERR_INVALID_ARGUMENT;
    }
    int rc = check_ready("step");
    if (rc != BLE_OK) return rc;
    const bool observe = cfg.enable_features && track_measurements;
    if (n_steps > 1 && observe) {
      err = "rollout: the WindGP measurement history cannot be tracked inside a multi-step launch (ble_features_track(0) first)";
      return BLE_ERR_UNSUPPORTED;
    }
    if (cfg.auto_reset && n_steps > 1) { err = "rollout: not available with ble_config.auto_reset"; return BLE_ERR_UNSUPPORTED; }
    if (cfg.auto_reset && episode_seed == nullptr) { err = "step: auto_reset needs a ble_reset first (it continues that seed chain)"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    FusedOut fo{o->reward, o->done, reinterpret_cast<float2*>(o->wind_uv), o->status, o->time_elapsed, o->sim_error};
    if (use_fused()) {
      int mode = 0;
      if (d.enable_noise) {
        mode = 1;
        if (noise_valid && n_steps == 1) {      // evaluated ahead of time (features_observe, step_host's prefetch)
          rc = launch_noise(s);                 // no launch: only orders this stream behind the producer
          if (rc != BLE_OK) return rc;
          mode = 2;
        }
      }
      launches += launch_fused(actions, fo, mode, n_steps, s);
      BLE_CUDA(cudaGetLastError());
      noise_valid = false;
    } else {
      for (int k = 0; k < n_steps; ++k) {
        rc = launch_noise(s);
        if (rc != BLE_OK) return rc;
        launch_step(actions + int64_t(k) * n, fo.reward + int64_t(k) * n, fo.done + int64_t(k) * n,
                    k == n_steps - 1 ? fo.wind_uv : nullptr, s);
        ++launches;
        BLE_CUDA(cudaGetLastError());
        noise_valid = false;
      }
      if (fo.status != nullptr || fo.time_elapsed != nullptr || fo.sim_error != nullptr) {
        k_step_info<Real><<<grid_for(n, 128), 128, 0, s>>>(d, fo.status, fo.time_elapsed, fo.sim_error);
        ++launches;
        BLE_CUDA(cudaGetLastError());
      }
    }
    if (cfg.auto_reset) {
      // the balloons that just returned done = 1 start a new episode; the masked reset ends with the measurement append
      // for EVERY balloon (post-step state of the ones that fly on, first state of the new episodes), so nothing more here
      k_auto_reset_prepare<<<grid_for(n, 256), 256, 0, s>>>(n, fo.done, episode_seed, auto_mask);
      ++launches;
      BLE_CUDA(cudaGetLastError());
      return reset(episode_seed, auto_mask, s);
    }
    // arena.step ends with feature_constructor.observe(get_measurements()) (env/balloon_arena.py:201):
    // the noise evaluated for it at the post-step state is also next step's pre-step wind.
    if (observe) return features_observe(s);
    return BLE_OK;
  }

  // Which step kernel: the fused one-launch kernel (production fp32 build) unless BLE_STEP_KERNEL = "thread" | "ws"
  // asks for a first-generation kernel (read per call: the parity tests run all of them on the same recorded states).
  bool use_fused() const {
    if constexpr (!std::is_same<Real, float>::value) return false;
    const char* env = std::getenv("BLE_STEP_KERNEL");
    return env == nullptr || (std::strcmp(env, "thread") != 0 && std::strcmp(env, "ws") != 0);
  }

  // ---- VAE decoder: Dense 64 -> 1000 -> 1000 -> 1000 -> 4410 (flax kernels are [in, out] row-major) ----
  int set_decoder(const float* const* kernels, const float* const* biases, cudaStream_t s) override {
    if (kernels == nullptr || biases == nullptr) { err = "set_decoder: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    BLE_DEVICE_GUARD();
    const int dims[5] = {kDecLatents, kDecHidden, kDecHidden, kDecHidden, kDecOut};
    const int ld[5] = {kDecLatents, kDecHidden, kDecHidden, kDecHidden, kDecOutPad};     // row pitch of layer i's [in, out] weights
    if (lt == nullptr && cublasLtCreate(&lt) != CUBLAS_STATUS_SUCCESS) { err = "set_decoder: cublasLtCreate failed"; return BLE_ERR_CUDA; }
    for (int i = 0; i < 4; ++i) {
      if (kernels[i] == nullptr || biases[i] == nullptr) { err = "set_decoder: null layer"; return BLE_ERR_INVALID_ARGUMENT; }
      if (dec_w[i] == nullptr) {
        BLE_CUDA(cudaMalloc(&dec_w[i], sizeof(float) * size_t(dims[i]) * ld[i + 1]));
        BLE_CUDA(cudaMalloc(&dec_b[i], sizeof(float) * ld[i + 1]));
        BLE_CUDA(cudaMemsetAsync(dec_w[i], 0, sizeof(float) * size_t(dims[i]) * ld[i + 1], s));
        BLE_CUDA(cudaMemsetAsync(dec_b[i], 0, sizeof(float) * ld[i + 1], s));
      }
      BLE_CUDA(cudaMemcpy2DAsync(dec_w[i], sizeof(float) * ld[i + 1], kernels[i], sizeof(float) * dims[i + 1],
                                 sizeof(float) * dims[i + 1], dims[i], cudaMemcpyDeviceToDevice, s));
      BLE_CUDA(cudaMemcpyAsync(dec_b[i], biases[i], sizeof(float) * dims[i + 1], cudaMemcpyDeviceToDevice, s));
    }
    if (dec_act[0] == nullptr) {
      BLE_CUDA(cudaMalloc(&dec_act[0], sizeof(float) * kDecChunk * kDecOutPad));
      BLE_CUDA(cudaMalloc(&dec_act[1], sizeof(float) * kDecChunk * kDecOutPad));
      BLE_CUDA(cudaMalloc(&dec_workspace, kDecWorkspace));
    }
    have_decoder = true;
    return BLE_OK;
  }

  // out[F, n_out] = act(in[F, k] @ W[k, n_out] + b), row-major == column-major (n_out x F) = W^T-free GEMM
  int dense(const float* in, float* out, int layer, int64_t f, int k, int n_out, bool relu, cudaStream_t s) {
    cublasLtMatmulDesc_t op = nullptr;
    cublasLtMatrixLayout_t la = nullptr, lb = nullptr, lc = nullptr;
    cublasLtMatmulPreference_t pref = nullptr;
    auto cleanup = [&]() {
      if (pref) cublasLtMatmulPreferenceDestroy(pref);
      if (la) cublasLtMatrixLayoutDestroy(la);
      if (lb) cublasLtMatrixLayoutDestroy(lb);
      if (lc) cublasLtMatrixLayoutDestroy(lc);
      if (op) cublasLtMatmulDescDestroy(op);
    };
    bool ok = cublasLtMatmulDescCreate(&op, cfg.decoder_fp32 ? CUBLAS_COMPUTE_32F : CUBLAS_COMPUTE_32F_FAST_TF32,
                                       CUDA_R_32F) == CUBLAS_STATUS_SUCCESS;
    const cublasLtEpilogue_t epi = relu ? CUBLASLT_EPILOGUE_RELU_BIAS : CUBLASLT_EPILOGUE_BIAS;
    const float* bias = dec_b[layer];
    ok = ok && cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE, &epi, sizeof(epi)) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatrixLayoutCreate(&la, CUDA_R_32F, n_out, k, n_out) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatrixLayoutCreate(&lb, CUDA_R_32F, k, f, k) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatrixLayoutCreate(&lc, CUDA_R_32F, n_out, f, n_out) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulPreferenceCreate(&pref) == CUBLAS_STATUS_SUCCESS;
    size_t ws = kDecWorkspace;
    ok = ok && cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws, sizeof(ws)) == CUBLAS_STATUS_SUCCESS;
    cublasLtMatmulHeuristicResult_t heur;
    int found = 0;
    ok = ok && cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, lc, lc, pref, 1, &heur, &found) == CUBLAS_STATUS_SUCCESS && found > 0;
    const float one = 1.f, zero = 0.f;
    ok = ok && cublasLtMatmul(lt, op, &one, dec_w[layer], la, in, lb, &zero, out, lc, out, lc, &heur.algo,
                              dec_workspace, kDecWorkspace, s) == CUBLAS_STATUS_SUCCESS;
    cleanup();
    if (!ok) { err = "decode: cuBLASLt GEMM failed"; return BLE_ERR_CUDA; }
    ++launches;
    return BLE_OK;
  }

  // the four GEMMs for c <= kDecChunk latents; the Dense_3 output [c, 4416] lands in `flow` (default dec_act[1])
  int decoder_gemms(const float* latents, int64_t c, cudaStream_t s, float* flow = nullptr) {
    int rc = dense(latents, dec_act[0], 0, c, kDecLatents, kDecHidden, true, s);
    if (rc == BLE_OK) rc = dense(dec_act[0], dec_act[1], 1, c, kDecHidden, kDecHidden, true, s);
    if (rc == BLE_OK) rc = dense(dec_act[1], dec_act[0], 2, c, kDecHidden, kDecHidden, true, s);
    if (rc == BLE_OK) rc = dense(dec_act[0], flow != nullptr ? flow : dec_act[1], 3, c, kDecHidden, kDecOutPad, false, s);
    return rc;
  }

  int decode(const float* latents, int64_t f, float* fields, cudaStream_t s) override {
    if (latents == nullptr || fields == nullptr || f <= 0) { err = "decode: bad argument"; return BLE_ERR_INVALID_ARGUMENT; }
    if (!have_decoder) { err = "decode: no decoder weights (call ble_set_decoder first)"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    const ResizeTaps taps = make_resize_taps();
    for (int64_t first = 0; first < f; first += kDecChunk) {
      const int64_t c = std::min<int64_t>(kDecChunk, f - first);
      const int rc = decoder_gemms(latents + first * kDecLatents, c, s);
      if (rc != BLE_OK) return rc;
      const int64_t threads = c * int64_t(kNX) * kNY * kDecFlows;
      k_decode_epilogue<<<grid_for(threads, 256), 256, 0, s>>>(dec_act[1], fields + first * kFieldFloats, taps, c);
      ++launches;
      BLE_CUDA(cudaGetLastError());
    }
    return BLE_OK;
  }

  // sample_field for `count` seeds straight into the field bank: latents -> decoder -> lookup windows
  int generate_fields(const uint64_t* seeds, int64_t first, int64_t count, cudaStream_t s) override {
    return generate_fields_at(seeds, first, count, nullptr, s);
  }

  int generate_fields_at(const uint64_t* seeds, int64_t first, int64_t count, const int32_t* dst_index, cudaStream_t s) override {
    if (seeds == nullptr || first < 0 || count <= 0 || (dst_index == nullptr && first + count > n_fields) || count > n_fields) {
      err = "generate_fields: range outside the allocated fields (call ble_alloc_fields first)";
      return BLE_ERR_INVALID_ARGUMENT;
    }
    if (!have_decoder) { err = "generate_fields: no decoder weights (call ble_set_decoder first)"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    if (gen_latents == nullptr) {
      BLE_CUDA(cudaMalloc(&gen_latents, sizeof(float) * kGenChunk * kDecLatents));
      for (int b = 0; b < 2; ++b) {
        BLE_CUDA(cudaMalloc(&gen_flow[b], sizeof(float) * kGenChunk * kDecOutPad));
        BLE_CUDA(cudaEventCreateWithFlags(&ev_flow[b], cudaEventDisableTiming));
        BLE_CUDA(cudaEventCreateWithFlags(&ev_written[b], cudaEventDisableTiming));
      }
      BLE_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      BLE_CUDA(cudaStreamCreateWithFlags(&gen_stream, cudaStreamNonBlocking));
      BLE_CUDA(cudaFuncSetAttribute(k_flow_to_windows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(float) * kFlowSmemFloats)));
      BLE_CUDA(cudaFuncSetAttribute(k_flow_to_windows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(float) * kFlowSmemFloats)));
    }
    const ResizeTaps taps = make_resize_taps();
    // Two streams: the decoder GEMMs of chunk k + 1 (tensor cores, caller's stream) run under the window writer of
    // chunk k (HBM-write bound, gen_stream).  Dense_3 writes into one of two flow buffers; the writer's stream forks
    // from the caller's stream here and joins it again below, so the caller sees ordinary stream semantics.
    BLE_CUDA(cudaEventRecord(ev_fork, s));
    BLE_CUDA(cudaStreamWaitEvent(gen_stream, ev_fork, 0));
    int64_t chunk_no = 0;
    for (int64_t done_f = 0; done_f < count; done_f += kGenChunk, ++chunk_no) {
      const int64_t c = std::min<int64_t>(kGenChunk, count - done_f);
      const int b = int(chunk_no & 1);
      if (chunk_no >= 2) BLE_CUDA(cudaStreamWaitEvent(s, ev_written[b], 0));          // flow buffer b is free again
      // The GEMMs always run on a full chunk (zero latents beyond c) so that cuBLASLt picks the same
      // algorithm whatever the batch: a seed's field is then bit-identical however a suite is sharded.
      if (c < kGenChunk) BLE_CUDA(cudaMemsetAsync(gen_latents, 0, sizeof(float) * kGenChunk * kDecLatents, s));
      k_sample_latents<<<grid_for(c * kDecLatents, 256), 256, 0, s>>>(seeds + done_f, c, gen_latents);
      ++launches;
      BLE_CUDA(cudaGetLastError());
      const int rc = decoder_gemms(gen_latents, kGenChunk, s, gen_flow[b]);
      if (rc != BLE_OK) return rc;
      BLE_CUDA(cudaEventRecord(ev_flow[b], s));
      BLE_CUDA(cudaStreamWaitEvent(gen_stream, ev_flow[b], 0));
      // resize + curl + window layout in one pass: the windows are the only HBM traffic of the field writer
      auto writer = d.layout.x_stride_floats == 32 ? k_flow_to_windows<true> : k_flow_to_windows<false>;
      writer<<<unsigned(c * kYC), kFlowThreads, sizeof(float) * kFlowSmemFloats, gen_stream>>>(
          gen_flow[b], cells, d.layout, taps, first + done_f, c, dst_index != nullptr ? dst_index + done_f : nullptr);
      ++launches;
      BLE_CUDA(cudaGetLastError());
      BLE_CUDA(cudaEventRecord(ev_written[b], gen_stream));
      have_fields = true;
    }
    BLE_CUDA(cudaStreamWaitEvent(s, ev_written[0], 0));
    if (chunk_no >= 2) BLE_CUDA(cudaStreamWaitEvent(s, ev_written[1], 0));
    return BLE_OK;
  }

  int sample_latents(const uint64_t* seeds, int64_t count, float* latents, cudaStream_t s) override {
    if (seeds == nullptr || latents == nullptr || count <= 0) { err = "sample_latents: bad argument"; return BLE_ERR_INVALID_ARGUMENT; }
    BLE_DEVICE_GUARD();
    k_sample_latents<<<grid_for(count * kDecLatents, 256), 256, 0, s>>>(seeds, count, latents);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int agent_station_seeker(const float* obs, int32_t* actions, int32_t* best, cudaStream_t s) override {
    if (obs == nullptr || actions == nullptr) { err = "agent_station_seeker: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    BLE_DEVICE_GUARD();
    k_agent_station_seeker<<<grid_for(n * 32, 128), 128, 0, s>>>(obs, n, actions, best);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int agent_random_walk(const float* obs, const uint64_t* seeds, int32_t step_index, int32_t* actions, cudaStream_t s) override {
    if (obs == nullptr || seeds == nullptr || actions == nullptr || step_index < 0) {
      err = "agent_random_walk: bad argument"; return BLE_ERR_INVALID_ARGUMENT;
    }
    BLE_DEVICE_GUARD();
    if (walk_target == nullptr) {
      if (step_index != 0) { err = "agent_random_walk: step_index 0 (begin_episode) must come first"; return BLE_ERR_NOT_READY; }
      BLE_CUDA(cudaMalloc(&walk_target, sizeof(double) * n));
    }
    k_agent_random_walk<<<grid_for(n, 128), 128, 0, s>>>(obs, n, walk_target, seeds, step_index, actions);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int eval_begin(cudaStream_t s) override {
    if (!have_state) { err = "eval_begin: no balloon state"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    if (ev.reward == nullptr) {
      BLE_CUDA(cudaMalloc(&ev.reward, sizeof(double) * n));
      BLE_CUDA(cudaMalloc(&ev.within, sizeof(int32_t) * n));
      BLE_CUDA(cudaMalloc(&ev.steps, sizeof(int32_t) * n));
      BLE_CUDA(cudaMalloc(&ev.active, sizeof(uint8_t) * n));
    }
    k_eval_begin<Real><<<grid_for(n, 128), 128, 0, s>>>(d, ev);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int eval_accumulate(const float* reward, float* path, cudaStream_t s) override {
    if (reward == nullptr) { err = "eval_accumulate: null reward"; return BLE_ERR_INVALID_ARGUMENT; }
    if (ev.reward == nullptr) { err = "eval_accumulate: call ble_eval_begin first"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    k_eval_accumulate<Real><<<grid_for(n, 128), 128, 0, s>>>(d, ev, reward, 50000.0, path);   // env.radius, balloon_env.py:136
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int eval_results(double* out, cudaStream_t s) override {
    if (out == nullptr) { err = "eval_results: null output"; return BLE_ERR_INVALID_ARGUMENT; }
    if (ev.reward == nullptr) { err = "eval_results: call ble_eval_begin first"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    k_eval_results<Real><<<grid_for(n, 128), 128, 0, s>>>(d, ev, out);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  void launch_step(const int32_t* actions, float* reward, uint8_t* done, float2* wind_uv, cudaStream_t s) {
    if constexpr (std::is_same<Real, float>::value) {
      // k_step_ws needs 4 threads per balloon: at 128 registers an SM holds 4 blocks = 128 balloons, so
      // one wave covers 148 * 128 = 18,944 balloons.  Below that it is ~2x faster than one thread per
      // balloon (shorter dependent chain); above it the extra waves cost more than they save.
      // BLE_STEP_KERNEL = "ws" | "thread" overrides the choice (read per call: the parity tests run both
      // kernels on the same recorded states)
      const char* env = std::getenv("BLE_STEP_KERNEL");
      const std::string forced(env != nullptr ? env : "");
      const bool use_ws = forced == "ws" || (forced != "thread" && n <= int64_t(148) * 128);
      if (use_ws) {
        k_step_ws<<<grid_for(n, 32), 128, 0, s>>>(d, actions, reward, done, wind_uv);
        return;
      }
    }
    k_step<Real><<<grid_for(n, 128), 128, 0, s>>>(d, actions, reward, done, wind_uv);
  }

  // Shape of the fused step kernel (ble_step_fused.h): the widest latency shape whose CTAs all fit in ONE wave (so that
  // every phase-1 task has its own warp), else the throughput shape.  BLE_STEP_WARPS = 0 | 4 | 8 | 14 overrides.
  int fused_blocks_per_sm[4] = {0, 0, 0, 0};
  int sm_count = 148;
  int fused_shape() const {
    if (const char* w = std::getenv("BLE_STEP_WARPS")) {
      if (*w != 0) {
        const int v = std::atoi(w);
        for (int shape : kFusedShapes) if (v == shape) return v;
      }
    }
    const int64_t blocks = (n + 31) / 32;
    for (int i = 3; i >= 1; --i) {
      if (blocks <= int64_t(fused_blocks_per_sm[i]) * sm_count) return kFusedShapes[i];
    }
    return 0;
  }
  int launch_fused(const int32_t* actions, const FusedOut& fo, int mode, int n_steps, cudaStream_t s) {
    if constexpr (std::is_same<Real, float>::value) return fused_launch(fused_shape(), d, actions, fo, mode, n_steps, s);
    return 0;
  }

  int features_observe(cudaStream_t s) override {
    if (!cfg.enable_features) { err = "features_observe: handle was created with enable_features = 0"; return BLE_ERR_UNSUPPORTED; }
    int rc = check_ready("features_observe");
    if (rc != BLE_OK) return rc;
    BLE_DEVICE_GUARD();
    rc = launch_noise(s);
    if (rc != BLE_OK) return rc;
    k_feat_observe_k<Real><<<grid_for(n * 32, 128), 128, 0, s>>>(d);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int features_clear(const uint8_t* mask, cudaStream_t s) override {
    if (!cfg.enable_features) { err = "features_clear: handle was created with enable_features = 0"; return BLE_ERR_UNSUPPORTED; }
    BLE_DEVICE_GUARD();
    if (mask == nullptr) {
      BLE_CUDA(cudaMemsetAsync(gp_count, 0, sizeof(int32_t) * n, s));
      BLE_CUDA(cudaMemsetAsync(gp_m, 0, sizeof(int32_t) * n, s));
    } else {
      err = "features_clear: masked clear is done by ble_reset"; return BLE_ERR_UNSUPPORTED;
    }
    return BLE_OK;
  }

  int features_track(int on) override {
    if (!cfg.enable_features) { err = "features_track: handle was created with enable_features = 0"; return BLE_ERR_UNSUPPORTED; }
    track_measurements = on != 0;
    return BLE_OK;
  }

  int features(float* obs, cudaStream_t s) override {
    if (obs == nullptr) { err = "features: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    if (!cfg.enable_features) { err = "features: handle was created with enable_features = 0"; return BLE_ERR_UNSUPPORTED; }
    int rc = check_ready("features");
    if (rc != BLE_OK) return rc;
    BLE_DEVICE_GUARD();
    // The 16 ambient features (one latency-bound sunrise search per balloon, 0.4 ms at 65,536) touch nothing the wind
    // column needs: they run on a side stream, forked from and joined to the caller's stream, under k_gp_posterior.
    if (feat_stream == nullptr) {
      BLE_CUDA(cudaStreamCreateWithFlags(&feat_stream, cudaStreamNonBlocking));
      BLE_CUDA(cudaEventCreateWithFlags(&ev_feat_fork, cudaEventDisableTiming));
      BLE_CUDA(cudaEventCreateWithFlags(&ev_feat_join, cudaEventDisableTiming));
    }
    BLE_CUDA(cudaEventRecord(ev_feat_fork, s));
    BLE_CUDA(cudaStreamWaitEvent(feat_stream, ev_feat_fork, 0));
    k_feat_ambient<Real><<<grid_for(n, 128), 128, 0, feat_stream>>>(d, obs);
    BLE_CUDA(cudaEventRecord(ev_feat_join, feat_stream));
    k_feat_range_levels<Real><<<grid_for(n * kRangeLevels, 128), 128, 0, s>>>(d, range_scratch);
    k_feat_range<Real><<<grid_for(n, 128), 128, 0, s>>>(d, range_scratch);
    if (gp_refit_every_step) {
      k_gp_factor<Real><<<unsigned(n), kFactorThreads, sizeof(double) * (kGpPacked + kGpWindow * 4), s>>>(d);
      k_gp_column<Real><<<unsigned(n), kColumnThreads, kColumnSmem, s>>>(d, obs);
    } else {
      k_gp_posterior<Real><<<unsigned(n), kGpPThreads, sizeof(PosteriorSmem), s>>>(d, obs, bank_map, tile_column);
    }
    BLE_CUDA(cudaStreamWaitEvent(s, ev_feat_join, 0));
    launches += gp_refit_every_step ? 5 : 4;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  // Device-side alias of a page-locked, mapped host allocation; nullptr for pageable memory.
  static const void* mapped_host_pointer(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
  }

  int step_host(const int32_t* actions_host, float* reward_host, uint8_t* done_host, cudaStream_t s) override {
    if (actions_host == nullptr || reward_host == nullptr || done_host == nullptr) { err = "step_host: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    int rc = check_ready("step_host");
    if (rc != BLE_OK) return rc;
    BLE_DEVICE_GUARD();
    rc = launch_noise(s);                  // needs no actions: runs while the host stages them (no-op when prefetched)
    if (rc != BLE_OK) return rc;
    if (host_zero_copy) {
      // caller's buffers already page-locked and mapped (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory):
      // the step kernel reads / writes THEM over PCIe and the two staging copies drop out
      const void* da = mapped_host_pointer(actions_host);
      void* dr = const_cast<void*>(mapped_host_pointer(reward_host));
      void* dd = const_cast<void*>(mapped_host_pointer(done_host));
      if (da != nullptr && dr != nullptr && dd != nullptr) {
        rc = step(static_cast<const int32_t*>(da), static_cast<float*>(dr), static_cast<uint8_t*>(dd), nullptr, s);
        if (rc != BLE_OK) return rc;
        BLE_CUDA(cudaStreamSynchronize(s));
        return host_prefetch_noise ? launch_noise(s) : BLE_OK;
      }
    }
    std::memcpy(h_actions, actions_host, sizeof(int32_t) * n);
    if (host_zero_copy) {
      // pinned host memory is device-accessible under UVA: the step kernel reads the actions and writes reward /
      // done straight over PCIe (4 + 5 bytes per balloon), which saves three copy launches and their latencies
      rc = step(h_actions, h_reward, h_done, nullptr, s);
      if (rc != BLE_OK) return rc;
    } else {
      BLE_CUDA(cudaMemcpyAsync(d_actions, h_actions, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
      rc = step(d_actions, d_reward, d_done, nullptr, s);
      if (rc != BLE_OK) return rc;
      BLE_CUDA(cudaMemcpyAsync(h_reward, d_reward, (sizeof(float) + sizeof(uint8_t)) * n, cudaMemcpyDeviceToHost, s));
    }
    BLE_CUDA(cudaStreamSynchronize(s));
    // The wind at the post-step state is the next step's pre-step wind (env/balloon_arena.py:184-202), so its noise
    // kernel is queued now and runs while the host copies the results out and prepares the next actions; any call
    // that changes the state (reset, state_upload, set_noise) invalidates it.
    if (host_prefetch_noise) {
      rc = launch_noise(s);
      if (rc != BLE_OK) return rc;
    }
    std::memcpy(reward_host, h_reward, sizeof(float) * n);
    std::memcpy(done_host, h_done, sizeof(uint8_t) * n);
    return BLE_OK;
  }

  int wind_at(float* uv, cudaStream_t s) override {
    if (uv == nullptr) { err = "wind_at_balloon: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    int rc = check_ready("wind_at_balloon");
    if (rc != BLE_OK) return rc;
    BLE_DEVICE_GUARD();
    rc = launch_noise(s);
    if (rc != BLE_OK) return rc;
    k_wind_at_balloon<Real><<<grid_for(n, 128), 128, 0, s>>>(d, reinterpret_cast<float2*>(uv));
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int derived(double* out, cudaStream_t s) override {
    if (out == nullptr) { err = "derived: null argument"; return BLE_ERR_INVALID_ARGUMENT; }
    if (!have_state) { err = "derived: no balloon state"; return BLE_ERR_NOT_READY; }
    BLE_DEVICE_GUARD();
    k_derived<Real><<<grid_for(n, 128), 128, 0, s>>>(d, out);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int wind_query(const double* xyzt, const int32_t* env_idx, int with_noise, float* uv, int64_t m, cudaStream_t s) override {
    if (m < 0 || (m > 0 && (xyzt == nullptr || env_idx == nullptr || uv == nullptr))) { err = "wind_query: bad argument"; return BLE_ERR_INVALID_ARGUMENT; }
    if (cfg.wind_model == BLE_WIND_GRID && !have_fields) { err = "wind_query: no wind fields (call ble_upload_fields first)"; return BLE_ERR_NOT_READY; }
    if (with_noise && cfg.enable_noise && !have_noise) { err = "wind_query: noise enabled but not seeded"; return BLE_ERR_NOT_READY; }
    if (m == 0) return BLE_OK;
    BLE_DEVICE_GUARD();
    k_wind_query<Real><<<grid_for(m, 128), 128, 0, s>>>(d, xyzt, env_idx, with_noise, reinterpret_cast<float2*>(uv), m);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int atmosphere_query(int which, const double* q, const int32_t* env_idx, double* out, int64_t m, cudaStream_t s) override {
    if (m < 0 || (which != 0 && which != 1) || (m > 0 && (q == nullptr || env_idx == nullptr || out == nullptr))) {
      err = "atmosphere_query: bad argument"; return BLE_ERR_INVALID_ARGUMENT;
    }
    if (!have_state) { err = "atmosphere_query: no balloon state"; return BLE_ERR_NOT_READY; }
    if (m == 0) return BLE_OK;
    BLE_DEVICE_GUARD();
    k_atmosphere_query<Real><<<grid_for(m, 128), 128, 0, s>>>(d, which, q, env_idx, out, m);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }

  int wind_gather(const float* xyzt, const int32_t* fidx, float* uv, int64_t m, cudaStream_t s) override {
    if (m < 0 || (m > 0 && (xyzt == nullptr || fidx == nullptr || uv == nullptr))) { err = "wind_gather: bad argument"; return BLE_ERR_INVALID_ARGUMENT; }
    if (!have_fields) { err = "wind_gather: no wind fields (call ble_upload_fields first)"; return BLE_ERR_NOT_READY; }
    if (m == 0) return BLE_OK;
    BLE_DEVICE_GUARD();
    k_wind_gather<Real><<<grid_for(m, 256), 256, 0, s>>>(cells, d.layout, reinterpret_cast<const float4*>(xyzt), fidx,
                                                       reinterpret_cast<float2*>(uv), m);
    ++launches;
    BLE_CUDA(cudaGetLastError());
    return BLE_OK;
  }
};

}  // namespace ble

// -----------------------------------------------------------------------------------------------
// C ABI
// -----------------------------------------------------------------------------------------------
struct ble_handle {
  ble::EngineBase* eng;
};

static thread_local std::string g_create_error;

extern "C" {

int ble_create(int device, int64_t n_envs, const ble_config* config, ble_handle** out) {
  if (out == nullptr || config == nullptr || n_envs <= 0) {
    g_create_error = "ble_create: out/config must be non-null and n_envs > 0";
    return BLE_ERR_INVALID_ARGUMENT;
  }
  if (config->precision != BLE_PRECISION_FP32 && config->precision != BLE_PRECISION_FP64) {
    g_create_error = "ble_create: unknown precision"; return BLE_ERR_INVALID_ARGUMENT;
  }
  if (config->field_layout != BLE_LAYOUT_X64 && config->field_layout != BLE_LAYOUT_X128) {
    g_create_error = "ble_create: unknown field_layout"; return BLE_ERR_INVALID_ARGUMENT;
  }
  if (config->wind_model != BLE_WIND_GRID && config->wind_model != BLE_WIND_SIMPLE_STATIC) {
    g_create_error = "ble_create: unknown wind_model"; return BLE_ERR_INVALID_ARGUMENT;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    g_create_error = "ble_create: CUDA device not available (this library has no CPU fallback)";
    return BLE_ERR_CUDA;
  }
  ble::EngineBase* eng = nullptr;
  int rc;
  if (config->precision == BLE_PRECISION_FP64) {
    auto* e = new (std::nothrow) ble::Engine<double>();
    if (e == nullptr) { g_create_error = "ble_create: host allocation failed"; return BLE_ERR_OUT_OF_MEMORY; }
    rc = e->create(device, n_envs, *config); eng = e;
  } else {
    auto* e = new (std::nothrow) ble::Engine<float>();
    if (e == nullptr) { g_create_error = "ble_create: host allocation failed"; return BLE_ERR_OUT_OF_MEMORY; }
    rc = e->create(device, n_envs, *config); eng = e;
  }
  if (rc != BLE_OK) { g_create_error = eng->err; delete eng; return rc; }
  auto* h = new (std::nothrow) ble_handle{eng};
  if (h == nullptr) { delete eng; g_create_error = "ble_create: host allocation failed"; return BLE_ERR_OUT_OF_MEMORY; }
  *out = h;
  return BLE_OK;
}

int ble_destroy(ble_handle* h) {
  if (h == nullptr) return BLE_ERR_INVALID_ARGUMENT;
  delete h->eng;
  delete h;
  return BLE_OK;
}

const char* ble_last_error(const ble_handle* h) { return h == nullptr ? g_create_error.c_str() : h->eng->err.c_str(); }
int64_t ble_num_envs(const ble_handle* h) { return h == nullptr ? 0 : h->eng->n; }
int64_t ble_launch_count(const ble_handle* h) { return h == nullptr ? 0 : h->eng->launches; }

#define BLE_H(h) if ((h) == nullptr) return BLE_ERR_INVALID_ARGUMENT

int ble_upload_fields(ble_handle* h, const float* fields, int64_t n_fields, const int32_t* env_to_field, void* stream) {
  BLE_H(h); return h->eng->upload_fields(fields, n_fields, env_to_field, cudaStream_t(stream));
}
int ble_alloc_fields(ble_handle* h, int64_t n_fields, void* stream) {
  BLE_H(h); return h->eng->alloc_fields(n_fields, cudaStream_t(stream));
}
int ble_write_fields(ble_handle* h, const float* fields, int64_t first_field, int64_t count, void* stream) {
  BLE_H(h); return h->eng->write_fields(fields, first_field, count, cudaStream_t(stream));
}
int ble_set_field_map(ble_handle* h, const int32_t* env_to_field, void* stream) {
  BLE_H(h); return h->eng->set_field_map(env_to_field, cudaStream_t(stream));
}
int ble_set_noise(ble_handle* h, const int64_t* seeds, const float* offsets, void* stream) {
  BLE_H(h); return h->eng->set_noise(seeds, offsets, nullptr, cudaStream_t(stream));
}
int ble_state_upload(ble_handle* h, const ble_state_soa* state, void* stream) {
  BLE_H(h); return h->eng->state_upload(state, cudaStream_t(stream));
}
int ble_state_download(ble_handle* h, ble_state_soa* state, void* stream) {
  BLE_H(h); return h->eng->state_download(state, cudaStream_t(stream));
}
int ble_reset(ble_handle* h, const uint64_t* seeds, const uint8_t* mask, void* stream) {
  BLE_H(h); return h->eng->reset(seeds, mask, cudaStream_t(stream));
}
int ble_init_derived(ble_handle* h, int32_t run_stable_init, void* stream) {
  BLE_H(h); return h->eng->init_derived(run_stable_init, cudaStream_t(stream));
}
int ble_step(ble_handle* h, const int32_t* actions, float* reward, uint8_t* done, float* wind_uv, void* stream) {
  BLE_H(h); return h->eng->step(actions, reward, done, wind_uv, cudaStream_t(stream));
}
int ble_step_ex(ble_handle* h, const int32_t* actions, const ble_step_out* out, void* stream) {
  BLE_H(h); return h->eng->step_ex(actions, out, 1, cudaStream_t(stream));
}
int ble_rollout(ble_handle* h, const int32_t* actions, int32_t n_steps, const ble_step_out* out, void* stream) {
  BLE_H(h); return h->eng->step_ex(actions, out, n_steps, cudaStream_t(stream));
}
int ble_step_host(ble_handle* h, const int32_t* actions_host, float* reward_host, uint8_t* done_host, void* stream) {
  BLE_H(h); return h->eng->step_host(actions_host, reward_host, done_host, cudaStream_t(stream));
}
int ble_wind_at_balloon(ble_handle* h, float* wind_uv, void* stream) {
  BLE_H(h); return h->eng->wind_at(wind_uv, cudaStream_t(stream));
}
int ble_set_decoder(ble_handle* h, const float* const* kernels, const float* const* biases, void* stream) {
  BLE_H(h); return h->eng->set_decoder(kernels, biases, cudaStream_t(stream));
}
int ble_decode_fields(ble_handle* h, const float* latents, int64_t n_fields, float* fields, void* stream) {
  BLE_H(h); return h->eng->decode(latents, n_fields, fields, cudaStream_t(stream));
}
int ble_generate_fields(ble_handle* h, const uint64_t* seeds, int64_t first_field, int64_t count, void* stream) {
  BLE_H(h); return h->eng->generate_fields(seeds, first_field, count, cudaStream_t(stream));
}
int ble_generate_fields_at(ble_handle* h, const uint64_t* seeds, const int32_t* field_index, int64_t count, void* stream) {
  BLE_H(h);
  if (field_index == nullptr) { h->eng->err = "generate_fields_at: null field_index"; return BLE_ERR_INVALID_ARGUMENT; }
  return h->eng->generate_fields_at(seeds, 0, count, field_index, cudaStream_t(stream));
}
int ble_sample_latents(ble_handle* h, const uint64_t* seeds, int64_t count, float* latents, void* stream) {
  BLE_H(h); return h->eng->sample_latents(seeds, count, latents, cudaStream_t(stream));
}
int ble_agent_station_seeker(ble_handle* h, const float* obs, int32_t* actions, int32_t* best_level, void* stream) {
  BLE_H(h); return h->eng->agent_station_seeker(obs, actions, best_level, cudaStream_t(stream));
}
int ble_agent_random_walk(ble_handle* h, const float* obs, const uint64_t* seeds, int32_t step_index, int32_t* actions,
                          void* stream) {
  BLE_H(h); return h->eng->agent_random_walk(obs, seeds, step_index, actions, cudaStream_t(stream));
}
int ble_eval_begin(ble_handle* h, void* stream) {
  BLE_H(h); return h->eng->eval_begin(cudaStream_t(stream));
}
int ble_eval_accumulate(ble_handle* h, const float* reward, float* flight_path, void* stream) {
  BLE_H(h); return h->eng->eval_accumulate(reward, flight_path, cudaStream_t(stream));
}
int ble_eval_results(ble_handle* h, double* out, void* stream) {
  BLE_H(h); return h->eng->eval_results(out, cudaStream_t(stream));
}
int ble_features_observe(ble_handle* h, void* stream) {
  BLE_H(h); return h->eng->features_observe(cudaStream_t(stream));
}
int ble_features_perciatelli(ble_handle* h, float* obs, void* stream) {
  BLE_H(h); return h->eng->features(obs, cudaStream_t(stream));
}
int ble_features_track(ble_handle* h, int32_t on) {
  BLE_H(h); return h->eng->features_track(on);
}
int ble_features_clear(ble_handle* h, void* stream) {
  BLE_H(h); return h->eng->features_clear(nullptr, cudaStream_t(stream));
}
int ble_derived(ble_handle* h, double* out, void* stream) {
  BLE_H(h); return h->eng->derived(out, cudaStream_t(stream));
}
int ble_wind_query(ble_handle* h, const double* xyzt, const int32_t* env_idx, int32_t with_noise, float* uv, int64_t m,
                   void* stream) {
  BLE_H(h); return h->eng->wind_query(xyzt, env_idx, with_noise, uv, m, cudaStream_t(stream));
}
int ble_atmosphere_query(ble_handle* h, int32_t which, const double* q, const int32_t* env_idx, double* out, int64_t m,
                         void* stream) {
  BLE_H(h); return h->eng->atmosphere_query(which, q, env_idx, out, m, cudaStream_t(stream));
}
int ble_wind_gather(ble_handle* h, const float* xyzt, const int32_t* field_idx, float* uv, int64_t m, void* stream) {
  BLE_H(h); return h->eng->wind_gather(xyzt, field_idx, uv, m, cudaStream_t(stream));
}

}  // extern "C"
