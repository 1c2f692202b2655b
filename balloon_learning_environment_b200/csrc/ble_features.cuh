// Perciatelli observation surface: scalar building blocks (host + device).
//
// Follows env/features.py:56-103,269-581, env/wind_gp.py:33-241,
// env/balloon/pressure_range_builder.py:43-275 and env/balloon/power_table.py:21-38.
// Everything here is fp64: the observation is produced once per agent step and several of its
// pieces are discrete searches (sunrise/sunset, reachable pressure range, level rounding).
#pragma once
#include "ble_physics.cuh"

namespace ble {

constexpr int kNumLevels = 181;                       // features.py:283
constexpr int kNumFeatures = 3 * (2 * kNumLevels - 1) + 16;   // 1099, features.py:291
constexpr double kLevelMin = 5000.0, kLevelMax = 14000.0;     // utils/constants.py:37-38
constexpr double kLevelStep = (kLevelMax - kLevelMin) / (kNumLevels - 1);   // 50 Pa
constexpr double kToleranceM = 1e-5;                  // features.py:53
constexpr int kGpWindow = 120;                        // 6 h of 3-minute observations (wind_gp.py:66, 172-178)
constexpr double kGpHorizonS = 6.0 * 3600.0;
constexpr double kGpSigma2 = 3.6 * 3.6;               // wind_gp.py:37
constexpr double kGpNoise = 0.05;                     // wind_gp.py:38
constexpr double kGpScaleXY = 357000.0, kGpScaleP = 326.0, kGpScaleT = 34560.0;   // wind_gp.py:33-35

BLE_HD double pressure_level(int l) { return kLevelMin + kLevelStep * double(l); }

// env/balloon/power_table.py:21-38 (bisect.bisect == upper bound).
BLE_HD double power_table_lookup(double pressure_ratio, double soc) {
  const double pr_edges[7] = {1.08, 1.11, 1.14, 1.17, 1.2, 1.23, 1.26};
  int pr_id = 0;
  while (pr_id < 7 && pressure_ratio >= pr_edges[pr_id]) ++pr_id;
  double e0, e1, e2 = 2.0, v0 = 0, v1, v2, v3 = 0;
  switch (pr_id) {
    case 0: e0 = 0.3; e1 = 0.4; e2 = 0.5; v1 = 150; v2 = 175; v3 = 200; break;
    case 1: e0 = 0.3; e1 = 0.4; e2 = 0.7; v1 = 200; v2 = 200; v3 = 225; break;
    case 2: e0 = 0.3; e1 = 0.4; e2 = 0.6; v1 = 225; v2 = 225; v3 = 250; break;
    case 3: e0 = 0.3; e1 = 0.4; e2 = 0.5; v1 = 200; v2 = 225; v3 = 250; break;
    case 4: e0 = 0.3; e1 = 0.4; e2 = 0.5; v1 = 225; v2 = 250; v3 = 275; break;
    case 5: e0 = 0.4; e1 = 0.5; v1 = 275; v2 = 300; break;
    case 6: e0 = 0.5; e1 = 0.6; v1 = 300; v2 = 325; break;
    default: e0 = 0.5; e1 = 0.6; v1 = 325; v2 = 350; break;
  }
  if (soc < e0) return v0;
  if (soc < e1) return v1;
  if (soc < e2) return v2;
  return pr_id <= 4 ? v3 : v2;
}

// compute_sunrise_time (env/features.py:72-103): [sunrise, sunset] -> [0, pi], [sunset, sunrise] -> [pi, 2 pi].
BLE_HD double sunrise_time(double lat, double lng, int64_t ts, bool* ok) {
  int64_t sunrise, sunset;
  *ok = next_sunrise_sunset(lat, lng, ts, &sunrise, &sunset);
  const int64_t day = 86400;
  if (sunset < sunrise) {                              // day time: sunset is next
    const int64_t prev_sunrise = sunrise - day;
    return kPi * double(ts - prev_sunrise) / double(sunset - prev_sunrise);
  }
  const int64_t prev_sunset = sunset - day;
  return kPi + kPi * double(ts - prev_sunset) / double(sunrise - prev_sunset);
}

// ---- reachable pressure range (env/balloon/pressure_range_builder.py) -------------------------------
constexpr int kRangeLevels = 20;                      // :231
constexpr double kRangeMinSp = 250.0;                 // envelope_safety.BUFFER (:220-224)
constexpr double kRangeMaxSp = kMaxSuperpressure - 250.0;

// _compute_safe_pressure / _compute_x_crossing (:43-102).  Returns false where the reference raises.
BLE_HD bool safe_pressure(double p1, double sp1, double p2, double sp2, double* out) {
  if (!(p1 < p2) || sp1 == sp2) return false;
  double y_star;
  if ((sp1 < kRangeMinSp && sp2 >= kRangeMinSp) || (sp1 >= kRangeMinSp && sp2 < kRangeMinSp)) y_star = kRangeMinSp;
  else if ((sp1 > kRangeMaxSp && sp2 <= kRangeMaxSp) || (sp1 <= kRangeMaxSp && sp2 > kRangeMaxSp)) y_star = kRangeMaxSp;
  else return false;
  if (y_star < fmin(sp1, sp2) || y_star > fmax(sp1, sp2)) return false;
  const double alpha = fabs((y_star - sp1) / (sp2 - sp1));
  *out = alpha * (p2 - p1) + p1;
  return true;
}

// _search_for_safe_pressure (:105-182) with the superpressures of the 20 levels and of the
// "significant" pressure already evaluated.  direction_min: scan downwards from the top level
// ('min' in the reference = looking for the MAX safe pressure), else upwards.
BLE_HD bool search_safe_pressure(const double* levels, const double* sp_levels, double significant,
                                 double sp_significant, bool direction_min, double* out) {
  if (sp_significant >= kRangeMinSp && sp_significant <= kRangeMaxSp) { *out = significant; return true; }
  double last_p = significant, last_sp = sp_significant;
  for (int n = 0; n < kRangeLevels; ++n) {
    const int j = direction_min ? kRangeLevels - 1 - n : n;
    const double p = levels[j];
    if (direction_min ? (p > significant) : (p < significant)) continue;
    const double sp = sp_levels[j];
    if (sp > kRangeMaxSp || sp < kRangeMinSp) { last_p = p; last_sp = sp; continue; }
    return direction_min ? safe_pressure(p, sp, last_p, last_sp, out) : safe_pressure(last_p, last_sp, p, sp, out);
  }
  return false;
}

// The pressure whose p/T equals that of the fully-vented balloon (:234-245): linear interpolation
// of pressure over p/T with extrapolation from the end segments (scipy interp1d).
BLE_HD double min_float_pressure(const double* levels, const double* p_over_t, double mols_gas) {
  const double empty_mass = kPayloadMass + kEnvelopeMass + mols_gas * kMHe;
  const double target = empty_mass * kR / (kMAir * kVolBase);
  int i = 0;
  while (i < kRangeLevels - 2 && !(target <= p_over_t[i + 1])) ++i;       // searchsorted(left) - 1, clipped
  const double slope = (levels[i + 1] - levels[i]) / (p_over_t[i + 1] - p_over_t[i]);
  return slope * (target - p_over_t[i]) + levels[i];
}

// ---- wind column encoding (features.py:457-581) --------------------------------------------------------
// _nearest_pressure_level (:354-380): Python round() == round-half-to-even == rint().
BLE_HD int nearest_pressure_level(double pressure) {
  const double p = fmin(fmax(pressure, kLevelMin), kLevelMax);
  return int(rint((p - kLevelMin) / kLevelStep));
}

// (uncertainty, angle error / pi, magnitude squash) of one level (:500-549).
BLE_HD void wind_level_features(double mean_u, double mean_v, double deviation, double x, double y,
                                float* f_unc, float* f_angle, float* f_mag) {
  const double dist = sqrt(x * x + y * y);
  const double sx = -x / (dist + kToleranceM), sy = -y / (dist + kToleranceM);
  const double mag = sqrt(mean_u * mean_u + mean_v * mean_v);
  const double ux = mean_u / (mag + kToleranceM), uy = mean_v / (mag + kToleranceM);
  double angle;
  if (dist < kToleranceM) {
    angle = 0.0;
  } else {
    const double c = fmin(fmax(ux * sx + uy * sy, -1.0), 1.0);
    angle = (mag < kToleranceM) ? kPi : acos(c);
  }
  *f_unc = float(deviation);
  *f_angle = float(angle / kPi);
  *f_mag = float(mag / (mag + 30.0));
}

// Matern-1/2 kernel of env/wind_gp.py:66-72 on pre-scaled coordinates (x/357km, y/357km, p/326, t/34560).
BLE_HD double gp_kernel(const double* a, const double* b) {
  const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2], d3 = a[3] - b[3];
  return kGpSigma2 * exp(-sqrt(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3));
}

}  // namespace ble
