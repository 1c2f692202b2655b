// Device-side state of one handle (struct of arrays) and the small helpers every kernel file shares.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/ble_b200.h"
#include "ble_physics.cuh"
#include "ble_wind.cuh"

namespace ble {

// ---------------------------------------------------------------------------------------------
// Device-side state layout (struct of arrays, one row per field, N columns)
// ---------------------------------------------------------------------------------------------
enum DRow : int {     // fp64 rows: the stiff integrator variables (see ble_physics.cuh) + atmosphere
  D_X = 0, D_Y, D_P, D_TAMB, D_TINT, D_VOL, D_SP, D_MOLS_AIR, D_CHARGE,   // 9 dynamic rows (read + written every step)
  D_ALPHA, D_L0, D_L1, D_L2, D_T1, D_T2, D_P1, D_P2, D_P3,              // per-episode atmosphere (layers 0..2)
  D_COUNT
};
enum RRow : int {     // `Real` rows (fp32 in production)
  R_ACS_W = 0, R_ACS_FLOW, R_SOLAR_W, R_LOAD_W,      // diagnostics written every step
  R_LAT0, R_LNG0, R_IR, R_MOLS_GAS,                  // per-episode constants
  R_COUNT
};
enum LRow : int { L_DATE_TIME = 0, L_SUNRISE_H, L_SUNSET, L_COUNT };

// flags word: status[0:2) last_command[2:4) envelope[4:7) altitude[7:9) paused[9] psl[10] atm_err[11]
__host__ __device__ inline uint32_t pack_flags(int status, int last_cmd, int env, int alt, int paused,
                                               int psl, int atm_err) {
  return uint32_t(status) | (uint32_t(last_cmd) << 2) | (uint32_t(env) << 4) | (uint32_t(alt) << 7) |
         (uint32_t(paused) << 9) | (uint32_t(psl) << 10) | (uint32_t(atm_err) << 11);
}

template <typename Real>
struct DevState {
  int64_t n;
  double* dd;         // [D_COUNT][n]
  Real* r;            // [R_COUNT][n]
  int64_t* l;         // [L_COUNT][n]
  int32_t* t_elapsed; // [n]
  uint32_t* flags;    // [n]
  // wind
  const float* cells;        // [F][layout.field_floats], 128-byte windows (ble_wind.cuh)
  FieldLayout layout;
  const int32_t* env_field;  // [n]
  const uint8_t* perm;       // [10][n][256], each table rotated by 4*(env%32) bytes
  const float* offsets;      // [10][4][n]
  Real* noise_partial;       // [10][n]
  int wind_model, enable_noise;
  // observation surface (WindGP history ring, env/wind_gp.py:98-119): last kGpWindow measurements
  double* gp_obs;            // [n][kGpWindow][6] = x, y, pressure, t, error_u, error_v
  int32_t* gp_count;         // [n] measurements seen so far (ring slot = count % kGpWindow)
  double* gp_chol;           // [n][7,680] lower Cholesky factor of the window, 8 x 8 blocked (ble_gp_kernels.cuh)
  int32_t* gp_m;             // [n] number of measurements the factor was computed for (0 = none)
  int32_t* gp_first;         // [n] index (in measurements seen) of the factor's first point; -1 = not a suffix
  double* gp_z;              // [n][kGpWindow][2] L^-1 (error_u, error_v)
  double* feat_range;        // [n][2] reachable pressure range
};

template <typename Real>
__device__ __forceinline__ Real& RR(const DevState<Real>& d, int row, int64_t e) { return d.r[int64_t(row) * d.n + e]; }
template <typename Real>
__device__ __forceinline__ double& DD(const DevState<Real>& d, int row, int64_t e) { return d.dd[int64_t(row) * d.n + e]; }

template <typename Real>
__device__ __forceinline__ void store_atmosphere(const DevState<Real>& d, int64_t e, const Atmosphere& atm) {
  DD(d, D_ALPHA, e) = atm.alpha;
  DD(d, D_L0, e) = atm.l0; DD(d, D_L1, e) = atm.l1; DD(d, D_L2, e) = atm.l2;
  DD(d, D_T1, e) = atm.t1; DD(d, D_T2, e) = atm.t2;
  DD(d, D_P1, e) = atm.p1; DD(d, D_P2, e) = atm.p2; DD(d, D_P3, e) = atm.p3;
}
template <typename Real>
__device__ __forceinline__ Atmosphere load_atmosphere(const DevState<Real>& d, int64_t e) {
  return Atmosphere::from_rows(DD(d, D_ALPHA, e), DD(d, D_L0, e), DD(d, D_L1, e), DD(d, D_L2, e), DD(d, D_T1, e),
                               DD(d, D_T2, e), DD(d, D_P1, e), DD(d, D_P2, e), DD(d, D_P3, e));
}

struct WindowLoader {        // per-thread path: the 8 chunks of one window
  const float4* base;
  __device__ __forceinline__ float4 operator()(int j) const { return __ldg(base + j); }
};


struct RotatedPerm {      // view of one rotated table in shared memory
  const uint8_t* t;
  int rot;
  __device__ __forceinline__ int operator[](int i) const { return t[(i + rot) & 255]; }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }


// Forecast wind at an arbitrary point of balloon e's field (WindField.get_forecast).
template <typename Real, typename State>
__device__ __forceinline__ void forecast_at(const State& d, int64_t e, double x, double y, double p,
                                            int32_t t_elapsed, Real* u, Real* v) {
  if (d.wind_model == BLE_WIND_SIMPLE_STATIC) {
    static_wind<Real>(Real(p), u, v);
  } else {
    const FieldPoint q = make_field_point(x / 1000.0, y / 1000.0, p, double(t_elapsed) / 3600.0);
    const FieldCell<Real> c = locate<Real>(q);
    WindowLoader ld{reinterpret_cast<const float4*>(d.cells + int64_t(d.env_field[e]) * d.layout.field_floats +
                                                    window_index(d.layout, c.ix, c.iy, c.pc, c.tc))};
    interp_window<Real>(c, ld, u, v);
  }
}

}  // namespace ble
