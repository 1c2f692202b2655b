// Evaluation surface: baseline controllers acting on the Perciatelli observation (host + device).
//
// Follows agents/station_seeker_agent.py:37-178 (StationSeeker score and action rule),
// agents/random_walk_agent.py:62-73 (hysteresis band around a target pressure),
// env/features.py:148-266 (validity of a level, de-normalisation of bearing / magnitude) and
// utils/transforms.py:45-94.  The reference evaluates the score in fp64 on the float32 features
// (numpy 1.19 scalar promotion), so everything here is double.
#pragma once
#include "ble_features.cuh"

namespace ble {

constexpr int kColumnLevels = 2 * kNumLevels - 1;      // 361 relative levels (features.py:291)
constexpr int kColumnCenter = kColumnLevels / 2;       // wind_column_center (features.py:253-256)

// station_seeker_agent.py:44-57
constexpr double kSeekHalfRadius = 35.0, kSeekMagnitudeWeight = 0.07;
constexpr double kSeekCloseBearingWeight = 0.6, kSeekFarBearingWeight = 0.45;
constexpr double kSeekCloseBearing = 250.0, kSeekFarBearing = 500.0;
constexpr double kSeekDefaultScore = 0.5, kSeekK2 = 0.05, kSeekK3 = 0.001, kSeekEpsilon = 0.01;

// PerciatelliWindFeature.is_valid_wind (features.py:155-160)
BLE_HD bool level_is_valid(float uncertainty, float bearing, float magnitude) {
  return magnitude != 1.0f || bearing != 1.0f || uncertainty != 0.0f;
}

// Terms of wind_score that depend only on the distance feature (station_seeker_agent.py:152-172).
struct SeekerDistanceTerms { double bearing_weight, alpha_delta; };
BLE_HD SeekerDistanceTerms seeker_distance_terms(float distance_feature) {
  const double d = double(distance_feature);
  const double distance = d * 250.0 / (1.0 - d);                         // undo_squash (transforms.py:88-94)
  const double coeff = fmin(fmax((distance - kSeekCloseBearing) / (kSeekFarBearing - kSeekCloseBearing), 0.0), 1.0);
  SeekerDistanceTerms t;
  t.bearing_weight = kSeekCloseBearingWeight + coeff * (kSeekFarBearingWeight - kSeekCloseBearingWeight);
  t.alpha_delta = exp(-distance / kSeekHalfRadius);
  return t;
}

// altitude_score (:115-150) of one valid level.
BLE_HD double seeker_altitude_score(const SeekerDistanceTerms& t, float uncertainty, float bearing, float magnitude,
                                    int level) {
  const double unc = double(uncertainty);
  const double bearing_rad = double(bearing) * kPi;                      // features.py:264-265
  const double m = double(magnitude);
  const double speed = m * 30.0 / (1.0 - m);                             // features.py:266
  const double wind_score = (1.0 - t.alpha_delta) * exp(-t.bearing_weight * bearing_rad)
                            + t.alpha_delta * exp(-kSeekMagnitudeWeight * speed);
  const int dist = level > kColumnCenter ? level - kColumnCenter : kColumnCenter - level;
  const double hysteresis = kSeekK2 * exp(-kSeekK3 * double(dist));
  return (1.0 - unc + kSeekEpsilon) * wind_score + unc * kSeekDefaultScore + hysteresis;
}

// pick_action (:72-88): DOWN = 0, STAY = 1, UP = 2; lower index = lower pressure = higher altitude.
BLE_HD int seeker_action_for_level(int best_level) {
  return best_level < kColumnCenter ? 2 : (best_level > kColumnCenter ? 0 : 1);
}

// Scalar scan exactly as find_best_pressure_level (:90-113); returns -1 when no level is valid.
BLE_HD int seeker_best_level(const float* obs, double* scores /* [361] or nullptr */) {
  const SeekerDistanceTerms t = seeker_distance_terms(obs[7]);
  int best = -1;
  double best_score = 0.0;
  for (int l = 0; l < kColumnLevels; ++l) {
    const float* w = obs + 16 + 3 * l;
    double sc = 0.0;
    if (level_is_valid(w[0], w[1], w[2])) sc = seeker_altitude_score(t, w[0], w[1], w[2], l);
    if (scores != nullptr) scores[l] = sc;
    if (sc > best_score) { best_score = sc; best = l; }
  }
  return best;
}

// RandomWalkAgent._select_action (random_walk_agent.py:62-73).
BLE_HD int random_walk_action(float pressure_feature, double target_pressure) {
  const double p = double(pressure_feature) * (kLevelMax - kLevelMin) + kLevelMin;   // features.py:192-195
  if (p - 100.0 > target_pressure) return 2;
  if (p + 100.0 < target_pressure) return 0;
  return 1;
}

}  // namespace ble
