// TEST-ONLY host replay of balloon_learning_environment_b200/csrc/ble_physics.cuh + ble_wind.cuh.
//
// There is no GPU in the build container, so the numerics of the templated device functions
// (fp32 production path and fp64 audit path) are checked here by compiling the SAME headers
// with g++ and driving them from pytest through ctypes.  This library is never loaded by the
// product package; the product path has no CPU fallback.
#include <cstdint>
#include <cstring>
#include "../../balloon_learning_environment_b200/csrc/ble_physics.cuh"
#include "../../balloon_learning_environment_b200/csrc/ble_step_roles.cuh"
#include "../../balloon_learning_environment_b200/csrc/ble_wind.cuh"
#include "../../balloon_learning_environment_b200/csrc/ble_features.cuh"
#include "../../balloon_learning_environment_b200/csrc/ble_agents.cuh"
#include "../../include/ble_b200.h"

using namespace ble;

// The kernels never call Atmosphere::init on the step path: k_state_upload / k_reset store the nine table rows
// and load_atmosphere() rebuilds the struct with Atmosphere::from_rows.  The replay takes the same road.
static Atmosphere atmosphere_as_the_kernels_load_it(double alpha) {
  Atmosphere t; t.init(alpha);
  return Atmosphere::from_rows(t.alpha, t.l0, t.l1, t.l2, t.t1, t.t2, t.p1, t.p2, t.p3);
}

// precision 2 = the production kernel's arithmetic: fp32 right-hand sides on the branch-free role functions
// of ble_step_roles.cuh (what k_step_fused runs, one role per warp), replayed here in sequence.
template <typename Real, bool kRoles = false>
static void step_impl(int64_t n, double* f, int64_t* iv, const int32_t* actions, const double* wind,
                      double* reward, int32_t* eff_out) {
  auto F = [&](int r, int64_t e) -> double& { return f[int64_t(r) * n + e]; };
  auto I = [&](int r, int64_t e) -> int64_t& { return iv[int64_t(r) * n + e]; };
  for (int64_t e = 0; e < n; ++e) {
    reward[e] = 0.0;
    eff_out[e] = kStay;
    if (I(BLE_I_STATUS, e) != kOk) continue;
    BalloonState<Real> s;
    s.x = F(BLE_F_X, e); s.y = F(BLE_F_Y, e); s.pressure = F(BLE_F_PRESSURE, e);
    s.t_ambient = F(BLE_F_AMBIENT_TEMPERATURE, e); s.t_internal = F(BLE_F_INTERNAL_TEMPERATURE, e);
    s.volume = F(BLE_F_ENVELOPE_VOLUME, e); s.superpressure = F(BLE_F_SUPERPRESSURE, e);
    s.mols_air = F(BLE_F_MOLS_AIR, e); s.charge = F(BLE_F_BATTERY_CHARGE, e);
    s.acs_power = Real(F(BLE_F_ACS_POWER, e)); s.acs_flow = Real(F(BLE_F_ACS_MASS_FLOW, e));
    s.solar_w = Real(F(BLE_F_SOLAR_CHARGING, e)); s.load_w = Real(F(BLE_F_POWER_LOAD, e));
    s.lat0 = Real(F(BLE_F_CENTER_LAT, e)); s.lng0 = Real(F(BLE_F_CENTER_LNG, e));
    s.ir = Real(F(BLE_F_UPWELLING_INFRARED, e)); s.mols_gas = Real(F(BLE_F_MOLS_LIFT_GAS, e));
    s.date_time = I(BLE_I_DATE_TIME, e); s.time_elapsed = int32_t(I(BLE_I_TIME_ELAPSED, e));
    s.status = int(I(BLE_I_STATUS, e));
    Atmosphere atm = atmosphere_as_the_kernels_load_it(F(BLE_F_ATMOSPHERE_ALPHA, e));
    SafetyState ss;
    ss.sunrise_h = I(BLE_I_SUNRISE_H, e); ss.sunset = I(BLE_I_SUNSET, e);
    ss.envelope_state = int(I(BLE_I_ENVELOPE_STATE, e)); ss.altitude_state = int(I(BLE_I_ALTITUDE_STATE, e));
    ss.power_paused = int(I(BLE_I_POWER_PAUSED, e)); ss.power_safety_enabled = int(I(BLE_I_POWER_SAFETY_ENABLED, e));
    ss.last_command = int(I(BLE_I_LAST_COMMAND, e));
    int eff;
    Real r;
    if constexpr (kRoles) r = roles::agent_step_roles(s, atm, ss, actions[e], wind[2 * e], wind[2 * e + 1], &eff);
    else r = agent_step<Real>(s, atm, ss, actions[e], wind[2 * e], wind[2 * e + 1], &eff);
    reward[e] = double(r); eff_out[e] = eff;
    F(BLE_F_X, e) = s.x; F(BLE_F_Y, e) = s.y; F(BLE_F_PRESSURE, e) = s.pressure;
    F(BLE_F_AMBIENT_TEMPERATURE, e) = s.t_ambient; F(BLE_F_INTERNAL_TEMPERATURE, e) = s.t_internal;
    F(BLE_F_ENVELOPE_VOLUME, e) = s.volume; F(BLE_F_SUPERPRESSURE, e) = s.superpressure;
    F(BLE_F_MOLS_AIR, e) = s.mols_air; F(BLE_F_BATTERY_CHARGE, e) = s.charge;
    F(BLE_F_ACS_POWER, e) = s.acs_power; F(BLE_F_ACS_MASS_FLOW, e) = s.acs_flow;
    F(BLE_F_SOLAR_CHARGING, e) = s.solar_w; F(BLE_F_POWER_LOAD, e) = s.load_w;
    I(BLE_I_DATE_TIME, e) = s.date_time; I(BLE_I_TIME_ELAPSED, e) = s.time_elapsed;
    I(BLE_I_STATUS, e) = s.status; I(BLE_I_LAST_COMMAND, e) = ss.last_command;
    I(BLE_I_ENVELOPE_STATE, e) = ss.envelope_state; I(BLE_I_ALTITUDE_STATE, e) = ss.altitude_state;
    I(BLE_I_POWER_PAUSED, e) = ss.power_paused; I(BLE_I_SUNRISE_H, e) = ss.sunrise_h;
    I(BLE_I_SUNSET, e) = ss.sunset;
  }
}

template <typename Real>
static void substeps_impl(int64_t n, double* f, int64_t* iv, const int32_t* eff, const double* wind, int nsub) {
  auto F = [&](int r, int64_t e) -> double& { return f[int64_t(r) * n + e]; };
  auto I = [&](int r, int64_t e) -> int64_t& { return iv[int64_t(r) * n + e]; };
  for (int64_t e = 0; e < n; ++e) {
    BalloonState<Real> s;
    s.x = F(BLE_F_X, e); s.y = F(BLE_F_Y, e); s.pressure = F(BLE_F_PRESSURE, e);
    s.t_ambient = F(BLE_F_AMBIENT_TEMPERATURE, e); s.t_internal = F(BLE_F_INTERNAL_TEMPERATURE, e);
    s.volume = F(BLE_F_ENVELOPE_VOLUME, e); s.superpressure = F(BLE_F_SUPERPRESSURE, e);
    s.mols_air = F(BLE_F_MOLS_AIR, e); s.charge = F(BLE_F_BATTERY_CHARGE, e);
    s.acs_power = 0; s.acs_flow = 0; s.solar_w = 0; s.load_w = 0;
    s.lat0 = Real(F(BLE_F_CENTER_LAT, e)); s.lng0 = Real(F(BLE_F_CENTER_LNG, e));
    s.ir = Real(F(BLE_F_UPWELLING_INFRARED, e)); s.mols_gas = Real(F(BLE_F_MOLS_LIFT_GAS, e));
    s.date_time = I(BLE_I_DATE_TIME, e); s.time_elapsed = int32_t(I(BLE_I_TIME_ELAPSED, e));
    s.status = 0;
    Atmosphere atm = atmosphere_as_the_kernels_load_it(F(BLE_F_ATMOSPHERE_ALPHA, e));
    SunTrack<Real> sun; sun.init(s, wind[2 * e], wind[2 * e + 1]);
    atm.incremental = !is_double<Real>::value;
    const Real epa = earth_heat_per_area<Real>(s.ir);
    for (int k = 0; k < nsub; ++k) {
      SunAngles<Real> ang; Real flux; sun.at(s, k, &ang, &flux);
      euler_substep<Real>(s, atm, wind[2 * e], wind[2 * e + 1], eff[e], ang, flux, epa);
    }
    F(BLE_F_X, e) = s.x; F(BLE_F_Y, e) = s.y; F(BLE_F_PRESSURE, e) = s.pressure;
    F(BLE_F_AMBIENT_TEMPERATURE, e) = s.t_ambient; F(BLE_F_INTERNAL_TEMPERATURE, e) = s.t_internal;
    F(BLE_F_ENVELOPE_VOLUME, e) = s.volume; F(BLE_F_SUPERPRESSURE, e) = s.superpressure;
    F(BLE_F_MOLS_AIR, e) = s.mols_air; F(BLE_F_BATTERY_CHARGE, e) = s.charge;
    I(BLE_I_DATE_TIME, e) = s.date_time; I(BLE_I_TIME_ELAPSED, e) = s.time_elapsed;
  }
}

extern "C" {

void emu_substeps(int precision, int64_t n, double* f, int64_t* iv, const int32_t* eff, const double* wind, int nsub) {
  if (precision == BLE_PRECISION_FP64) substeps_impl<double>(n, f, iv, eff, wind, nsub);
  else substeps_impl<float>(n, f, iv, eff, wind, nsub);
}

void emu_step(int precision, int64_t n, double* f, int64_t* iv, const int32_t* actions,
              const double* wind, double* reward, int32_t* eff) {
  if (precision == BLE_PRECISION_FP64) step_impl<double>(n, f, iv, actions, wind, reward, eff);
  else if (precision == 2) step_impl<float, true>(n, f, iv, actions, wind, reward, eff);
  else step_impl<float>(n, f, iv, actions, wind, reward, eff);
}

void emu_solar(int precision, int64_t n, const double* lat, const double* lng, const int64_t* ts,
               double* el, double* flux) {
  for (int64_t i = 0; i < n; ++i) {
    if (precision == BLE_PRECISION_FP64) {
      solar_calculator<double>(lat[i], lng[i], ts[i], &el[i], &flux[i]);
    } else {
      float e, fl; solar_calculator<float>(float(lat[i]), float(lng[i]), ts[i], &e, &fl);
      el[i] = e; flux[i] = fl;
    }
  }
}

void emu_noise(int precision, int64_t n, const int64_t* seeds, const double* xyzw, double* out) {
  uint8_t perm[256], scratch[256];
  for (int64_t i = 0; i < n; ++i) {
    simplex_make_perm(seeds[i], perm, scratch);
    const double* p = xyzw + 4 * i;
    out[i] = precision == BLE_PRECISION_FP64 ? simplex_noise4<double>((const uint8_t*)perm, p[0], p[1], p[2], p[3])
                                             : double(simplex_noise4<float>((const uint8_t*)perm, p[0], p[1], p[2], p[3]));
  }
}

// the A/B all-vertices forms (oracle form='all'): which = 0 round-1 evaluation, 1 second generation
void emu_noise_all(int precision, int which, int64_t n, const int64_t* seeds, const double* xyzw, double* out) {
  uint8_t perm[256], scratch[256];
  for (int64_t i = 0; i < n; ++i) {
    simplex_make_perm(seeds[i], perm, scratch);
    const double* p = xyzw + 4 * i;
    const uint8_t* pm = perm;
    if (precision == BLE_PRECISION_FP64)
      out[i] = which ? simplex_noise4_v2_all<double>(pm, p[0], p[1], p[2], p[3]) : simplex_noise4_all<double>(pm, p[0], p[1], p[2], p[3]);
    else
      out[i] = which ? double(simplex_noise4_v2_all<float>(pm, p[0], p[1], p[2], p[3]))
                     : double(simplex_noise4_all<float>(pm, p[0], p[1], p[2], p[3]));
  }
}

// production evaluation (tree form, corners unrolled); precision as emu_noise
void emu_noise_v2(int precision, int64_t n, const int64_t* seeds, const double* xyzw, double* out) {
  uint8_t perm[256], scratch[256];
  for (int64_t i = 0; i < n; ++i) {
    simplex_make_perm(seeds[i], perm, scratch);
    const double* p = xyzw + 4 * i;
    out[i] = precision == BLE_PRECISION_FP64 ? simplex_noise4_v2<double>((const uint8_t*)perm, p[0], p[1], p[2], p[3])
                                             : double(simplex_noise4_v2<float>((const uint8_t*)perm, p[0], p[1], p[2], p[3]));
  }
}

void emu_perm(int64_t seed, uint8_t* perm) { uint8_t scratch[256]; simplex_make_perm(seed, perm, scratch); }

// fields: native layout [F,21,21,10,9,2]; xyzt: (x km, y km, p, hours) float32
void emu_interp(int precision, int64_t m, const float* fields, const int32_t* fidx, const float* xyzt,
                double* uv) {
  for (int64_t i = 0; i < m; ++i) {
    const float* base = fields + int64_t(fidx[i]) * kFieldFloats;
    const FieldPoint q = make_field_point(xyzt[4 * i], xyzt[4 * i + 1], xyzt[4 * i + 2], xyzt[4 * i + 3]);
    auto run = [&](auto tag) {
      using R = decltype(tag);
      const FieldCell<R> c = locate<R>(q);
      auto ld = [&](int j) {      // chunk j = dx*4 + dy*2 + dp, gathered from the native layout
        const int ix = c.ix + ((j >> 2) & 1), iy = c.iy + ((j >> 1) & 1), ip = c.pc + (j & 1);
        return float4{base[native_index(ix, iy, ip, c.tc, 0)], base[native_index(ix, iy, ip, c.tc, 1)],
                      base[native_index(ix, iy, ip, c.tc + 1, 0)], base[native_index(ix, iy, ip, c.tc + 1, 1)]};
      };
      R u, v; interp_window<R>(c, ld, &u, &v); uv[2 * i] = u; uv[2 * i + 1] = v;
    };
    if (precision == BLE_PRECISION_FP64) run(double(0)); else run(float(0));
  }
}

void emu_sunrise_sunset(int64_t n, const double* lat, const double* lng, const int64_t* ts,
                        int64_t* sunrise, int64_t* sunset) {
  for (int64_t i = 0; i < n; ++i) next_sunrise_sunset(lat[i], lng[i], ts[i], &sunrise[i], &sunset[i]);
}

void emu_stable(int64_t n, const double* alpha, const double* p, const double* lat, const double* lng,
                const int64_t* ts, const double* ir, double* out /*[n,5]*/) {
  for (int64_t i = 0; i < n; ++i) {
    const StableParams s = stable_params(alpha[i], p[i], 6830.0, lat[i], lng[i], ts[i], ir[i]);
    out[5 * i] = s.t_ambient; out[5 * i + 1] = s.t_internal; out[5 * i + 2] = s.mols_air;
    out[5 * i + 3] = s.volume; out[5 * i + 4] = s.superpressure;
  }
}

// Reachable pressure range exactly as the feature kernels compute it (20 levels + min-float pressure).
void emu_pressure_range(int64_t n, const double* alpha, const double* mols_gas, const double* lat, const double* lng,
                        const int64_t* ts, const double* ir, double* out /*[n,2]*/, int32_t* ok) {
  for (int64_t e = 0; e < n; ++e) {
    double search_max, t_unused;
    atm_at_height_generic(alpha[e], kAltMin, &search_max, &t_unused);
    double levels[kRangeLevels], p_over_t[kRangeLevels], sp[kRangeLevels];
    for (int j = 0; j < kRangeLevels; ++j) {
      levels[j] = 1000.0 + (search_max - 1000.0) * double(j) / double(kRangeLevels - 1);
      if (j == kRangeLevels - 1) levels[j] = search_max;
      const StableParams s = stable_params(alpha[e], levels[j], mols_gas[e], lat[e], lng[e], ts[e], ir[e]);
      p_over_t[j] = levels[j] / s.t_ambient;
      sp[j] = s.superpressure;
    }
    const double pmin_sig = min_float_pressure(levels, p_over_t, mols_gas[e]);
    const double sp_min_sig = stable_params(alpha[e], pmin_sig, mols_gas[e], lat[e], lng[e], ts[e], ir[e]).superpressure;
    bool ok1 = search_safe_pressure(levels, sp, pmin_sig, sp_min_sig, false, &out[2 * e]);
    bool ok2 = search_safe_pressure(levels, sp, levels[kRangeLevels - 1], sp[kRangeLevels - 1], true, &out[2 * e + 1]);
    ok[e] = ok1 && ok2;
  }
}

void emu_sunrise_time(int64_t n, const double* lat, const double* lng, const int64_t* ts, double* out) {
  for (int64_t e = 0; e < n; ++e) { bool ok; out[e] = sunrise_time(lat[e], lng[e], ts[e], &ok); }
}

double emu_power_table(double pr, double soc) { return power_table_lookup(pr, soc); }
int emu_nearest_level(double p) { return nearest_pressure_level(p); }

// StationSeeker / RandomWalk action rules (ble_agents.cuh) on [n, 1099] observations.
void emu_station_seeker(int64_t n, const float* obs, int32_t* actions, int32_t* best, double* scores /* [n,361] */) {
  for (int64_t e = 0; e < n; ++e) {
    const int b = seeker_best_level(obs + e * kNumFeatures, scores + e * kColumnLevels);
    best[e] = b;
    actions[e] = b < 0 ? 1 : seeker_action_for_level(b);
  }
}
void emu_random_walk(int64_t n, const float* obs, const double* target, int32_t* actions) {
  for (int64_t e = 0; e < n; ++e) actions[e] = random_walk_action(obs[e * kNumFeatures], target[e]);
}

}  // extern "C"
