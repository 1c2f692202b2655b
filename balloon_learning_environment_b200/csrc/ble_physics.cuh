// Balloon physics for one balloon, templated on the arithmetic type.
//
// Real = float  : the production path (fp32 math; fp64 only where fp32 cannot
//                 hold the reference's quantity: Julian time, the +-1 Pa height
//                 secant, the buoyancy difference).
// Real = double : audit path, follows the reference's fp64 arithmetic order.
//
// Everything here is a per-balloon scalar function (__host__ __device__) so the
// kernels in ble_kernels.cu stay thin and tests/hostemu can replay the exact
// same source on the CPU when no GPU is present.  Citations are to
// /root/reference/balloon_learning_environment/<file>:<line>.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define BLE_HD __host__ __device__ __forceinline__
#define BLE_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define BLE_HD inline
#define BLE_HD_NOINLINE inline
#endif

namespace ble {

// ---- constants (utils/constants.py:23-38, env/balloon/balloon.py:156-172) ----
constexpr double kGravity = 9.80665;
constexpr double kR = 8.3144621;
constexpr double kMAir = 0.028964922481160;
constexpr double kMHe = 0.004002602;
constexpr double kRAir = kR / kMAir;
constexpr double kVolBase = 1804.0;
constexpr double kDvDp = 0.0199;
constexpr double kEnvelopeMass = 68.5;
constexpr double kMaxSuperpressure = 2380.0;
constexpr double kCod = 0.25;
constexpr double kPayloadMass = 92.5;
constexpr double kNightLoadW = 183.7;
constexpr double kDayLoadW = 120.4;
constexpr double kValveDiameter = 0.04;
constexpr double kBatteryCapacityWh = 3058.56;
constexpr double kValveCd = 0.62;
constexpr double kMinSolarElDeg = -4.242;            // env/balloon/solar.py:38
constexpr double kEarthRadiusM = 6371000.0;          // utils/spherical_geometry.py:29
constexpr double kPi = 3.14159265358979323846;
constexpr int kStrideS = 10;                         // env/balloon/balloon.py:269
constexpr int kAgentStepS = 180;                     // utils/constants.py:35
constexpr int kSubSteps = kAgentStepS / kStrideS;

enum Command : int { kDown = 0, kStay = 1, kUp = 2 };                    // env/balloon/control.py:21-25
enum Status : int { kOk = 0, kOutOfPower = 1, kBurst = 2, kZeroPressure = 3 };  // balloon.py:66-70
enum EnvState : int { kEnvNominal = 0, kEnvLowCritical = 1, kEnvLow = 2, kEnvHigh = 3, kEnvHighCritical = 4 };
enum AltState : int { kAltNominal = 0, kAltLow = 1, kAltVeryLow = 2 };

// ---- type-dispatched math ----------------------------------------------------
BLE_HD float r_sin(float x) { return sinf(x); }
BLE_HD double r_sin(double x) { return sin(x); }
BLE_HD float r_cos(float x) { return cosf(x); }
BLE_HD double r_cos(double x) { return cos(x); }
BLE_HD float r_tan(float x) { return tanf(x); }
BLE_HD double r_tan(double x) { return tan(x); }
BLE_HD float r_asin(float x) { return asinf(x); }
BLE_HD double r_asin(double x) { return asin(x); }
BLE_HD float r_acos(float x) { return acosf(x); }
BLE_HD double r_acos(double x) { return acos(x); }
BLE_HD float r_atan2(float y, float x) { return atan2f(y, x); }
BLE_HD double r_atan2(double y, double x) { return atan2(y, x); }
BLE_HD float r_sqrt(float x) { return sqrtf(x); }
BLE_HD double r_sqrt(double x) { return sqrt(x); }
BLE_HD float r_cbrt(float x) { return cbrtf(x); }
BLE_HD double r_cbrt(double x) { return cbrt(x); }
BLE_HD float r_exp(float x) { return expf(x); }
BLE_HD double r_exp(double x) { return exp(x); }
BLE_HD float r_log(float x) { return logf(x); }
BLE_HD double r_log(double x) { return log(x); }
BLE_HD float r_pow(float x, float y) { return powf(x, y); }
BLE_HD double r_pow(double x, double y) { return pow(x, y); }
// x^y for the smooth fp32 right-hand sides: ex2(y * lg2(x)) on the SFU (relative error ~3e-7 * y),
// full-precision pow in fp64 / on the host.
BLE_HD float r_pow_fast(float x, float y) {
#if defined(__CUDA_ARCH__)
  return __powf(x, y);
#else
  return powf(x, y);
#endif
}
BLE_HD double r_pow_fast(double x, double y) { return pow(x, y); }
BLE_HD float r_abs(float x) { return fabsf(x); }
BLE_HD double r_abs(double x) { return fabs(x); }
BLE_HD float r_min(float a, float b) { return fminf(a, b); }
BLE_HD double r_min(double a, double b) { return fmin(a, b); }
BLE_HD float r_max(float a, float b) { return fmaxf(a, b); }
BLE_HD double r_max(double a, double b) { return fmax(a, b); }
template <typename Real> BLE_HD Real r_rad(Real deg) { return deg * Real(kPi / 180.0); }
template <typename Real> BLE_HD Real r_deg(Real rad) { return rad * Real(180.0 / kPi); }

template <typename Real> struct is_double { static constexpr bool value = false; };
template <> struct is_double<double> { static constexpr bool value = true; };

// ---- standard atmosphere (env/balloon/standard_atmosphere.py:60-202) ----------
// Per-balloon atmosphere = lapse blend alpha.  Layers 0..2 (up to 32 km, i.e. down
// to ~750 Pa) are kept in registers; anything higher takes the generic fp64 path.
constexpr double kAtmH[8] = {-610.0, 17000.0, 21000.0, 32000.0, 47000.0, 51000.0, 71000.0, 85000.0};
constexpr double kLapseLow[7] = {-0.007, 0.006, 0.001, 0.0028, 0.0, -0.0028, -0.002};
constexpr double kLapseHigh[7] = {-0.0058, 0.005, 0.001, 0.0028, 0.0, -0.0028, -0.002};
constexpr double kAtmT0 = 300.0;
constexpr double kAtmP0 = 108870.8213;

BLE_HD double atm_h(int i) {
  const double t[8] = {-610.0, 17000.0, 21000.0, 32000.0, 47000.0, 51000.0, 71000.0, 85000.0};
  return t[i];
}
BLE_HD double atm_lapse(double alpha, int i) {
  const double lo[7] = {-0.007, 0.006, 0.001, 0.0028, 0.0, -0.0028, -0.002};
  const double hi[7] = {-0.0058, 0.005, 0.001, 0.0028, 0.0, -0.0028, -0.002};
  return (1 - alpha) * lo[i] + alpha * hi[i];                              // :83-84
}

// Full fp64 tables (reset path, generic fallback).  t_tr/p_tr: 8 entries each.
BLE_HD void atm_tables(double alpha, double* lapse, double* t_tr, double* p_tr) {
  t_tr[0] = kAtmT0;                                                        // :157
  p_tr[0] = kAtmP0;                                                        // :167
  for (int i = 0; i < 7; ++i) {
    lapse[i] = atm_lapse(alpha, i);
    const double dh = atm_h(i + 1) - atm_h(i);
    t_tr[i + 1] = t_tr[i] + lapse[i] * dh;                                 // :158-161
    if (lapse[i] == 0.0) {
      p_tr[i + 1] = p_tr[i] * exp(-(kGravity * dh) / (kRAir * t_tr[i + 1]));            // :186-192
    } else {
      p_tr[i + 1] = p_tr[i] * pow(t_tr[i + 1] / t_tr[i], -kGravity / (kRAir * lapse[i]));  // :194-202
    }
  }
}

// Generic at_pressure in fp64 (any layer).  Returns false if p is outside the atmosphere
// (the reference asserts, :126-127).
BLE_HD_NOINLINE bool atm_at_pressure_generic(double alpha, double p, double* height, double* temperature) {
  double lapse[7], t_tr[8], p_tr[8];
  atm_tables(alpha, lapse, t_tr, p_tr);
  if (!(p > p_tr[7]) || !(p <= p_tr[0])) return false;
  for (int i = 0; i < 7; ++i) {
    if (p > p_tr[i + 1]) {                                                 // :134
      double h;
      if (lapse[i] == 0.0) {
        h = (-kRAir * t_tr[i] / kGravity) * log(p / p_tr[i]) + atm_h(i);   // :137-140
      } else {
        h = (pow(p / p_tr[i], -kRAir * lapse[i] / kGravity) - 1) * t_tr[i] / lapse[i] + atm_h(i);  // :142-146
      }
      *height = h;
      *temperature = t_tr[i] + lapse[i] * (h - atm_h(i));                  // :148-149
      return true;
    }
  }
  return false;
}

BLE_HD bool atm_at_height_generic(double alpha, double h, double* pressure, double* temperature) {
  double lapse[7], t_tr[8], p_tr[8];
  atm_tables(alpha, lapse, t_tr, p_tr);
  if (!(h >= atm_h(0)) || !(h < atm_h(7))) return false;                   // :95-96
  for (int i = 0; i < 7; ++i) {
    if (h < atm_h(i + 1)) {                                                // :103
      const double t = t_tr[i] + lapse[i] * (h - atm_h(i));
      double p;
      if (lapse[i] == 0.0) {
        p = p_tr[i] * exp(-(kGravity * (h - atm_h(i))) / (kRAir * t));     // :109-111
      } else {
        p = p_tr[i] * pow(t / t_tr[i], -kGravity / (kRAir * lapse[i]));    // :113-115
      }
      *pressure = p;
      *temperature = t;
      return true;
    }
  }
  return false;
}

// Register-resident lower atmosphere (layers 0, 1, 2), always fp64: pressure <-> height feeds the
// stiff buoyancy loop of the Euler integrator (see the note above BalloonState).
struct Atmosphere {
  double alpha;
  double l0, l1, l2;        // lapse rates
  double t1, t2;            // temperature transitions (t0 = 300)
  double p1, p2, p3;        // pressure transitions (p0 = 108870.8213)
  bool ok = true;           // false once a query fell outside the atmosphere
  // X = (p/P_i)^k of the previous sub-step (production build only): the next X is
  // X_prev * (p/p_prev)^k with |p/p_prev - 1| ~ 1e-3, summed as a binomial series.
  // The cache starts EMPTY (lcache = -1) however the struct is built: the first query of an agent
  // step always takes the exp(k log) path.
  bool incremental = false;
  int lcache = -1;
  double pcache = 1.0, xcache = 1.0;

  // The one way the kernels (and the host replay of tests/hostemu) rebuild the struct from the nine
  // per-episode rows stored in HBM (DevState rows D_ALPHA .. D_P3).
  BLE_HD static Atmosphere from_rows(double alpha, double l0, double l1, double l2, double t1, double t2,
                                     double p1, double p2, double p3) {
    Atmosphere a;
    a.alpha = alpha;
    a.l0 = l0; a.l1 = l1; a.l2 = l2;
    a.t1 = t1; a.t2 = t2;
    a.p1 = p1; a.p2 = p2; a.p3 = p3;
    a.ok = true;
    a.incremental = false;
    a.lcache = -1; a.pcache = 1.0; a.xcache = 1.0;
    return a;
  }

  BLE_HD void init(double a) {
    double lapse[7], t_tr[8], p_tr[8];
    atm_tables(a, lapse, t_tr, p_tr);
    alpha = a;
    l0 = lapse[0]; l1 = lapse[1]; l2 = lapse[2];
    t1 = t_tr[1]; t2 = t_tr[2];
    p1 = p_tr[1]; p2 = p_tr[2]; p3 = p_tr[3];
    ok = true;
    incremental = false;
    lcache = -1; pcache = 1.0; xcache = 1.0;
  }

  // (1 + r)^k as a binomial series, |r| <= 0.02: the 10th term is below 1e-18.
  BLE_HD static double binomial_pow(double r, double k) {
    double acc = 1.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int n = 9; n >= 1; --n) acc = 1.0 + acc * r * (k - double(n - 1)) * (1.0 / double(n));
    return acc;
  }

  // What one Euler sub-step needs from the atmosphere, with ONE pow (as exp(k log x)):
  //   temperature at p (standard_atmosphere.py:148-149; T = T_i (p/P_i)^k exactly) and the
  //   secant h(p + dir) - h(p) of env/balloon/balloon.py:438-443.  With X = (p/P_i)^k,
  //   h(p) = (X - 1) T_i / L_i + H_i, so the secant is (T_i/L_i) X ((1 + dir/p)^k - 1); the last
  //   factor is summed as a binomial series (dir/p ~ 1e-4: five terms reach 1e-20), which is both
  //   cheaper and more accurate than subtracting two 17 km heights.  Falls back to two full
  //   evaluations when p and p + dir straddle a layer boundary or leave layers 0..2.
  BLE_HD void temperature_and_secant(double p, double dir, double* temperature, double* dh) {
    const double q = p + dir;
    double lapse, ti, pi;
    int layer = -1;
    if (p > p3 && p <= kAtmP0 && q > p3 && q <= kAtmP0) {
      if (p > p1) { if (q > p1) { layer = 0; lapse = l0; ti = kAtmT0; pi = kAtmP0; } }
      else if (p > p2) { if (q <= p1 && q > p2) { layer = 1; lapse = l1; ti = t1; pi = p1; } }
      else { if (q <= p2) { layer = 2; lapse = l2; ti = t2; pi = p2; } }
    }
    if (layer >= 0) {
      const double k = -kRAir * lapse / kGravity;
      double x;
      const double r = p / pcache - 1.0;
      if (incremental && layer == lcache && fabs(r) < 0.02) {
        x = xcache * binomial_pow(r, k);
      } else {
        x = exp(k * log(p / pi));
      }
      lcache = layer; pcache = p; xcache = x;
      const double e = dir / p;
      const double series = e * k * (1.0 + e * (k - 1.0) * (1.0 / 2.0) * (1.0 + e * (k - 2.0) * (1.0 / 3.0) *
                            (1.0 + e * (k - 3.0) * (1.0 / 4.0) * (1.0 + e * (k - 4.0) * (1.0 / 5.0)))));
      *temperature = ti * x;
      *dh = (ti / lapse) * x * series;
    } else {
      lcache = -1;
      double h0, h1, t1_unused;
      at_pressure(p, &h0, temperature);
      at_pressure(q, &h1, &t1_unused);
      *dh = h1 - h0;
    }
  }

  // height [m] and temperature [K] at pressure p (:122-154).
  BLE_HD void at_pressure(double p, double* height, double* temperature) {
    if (p > p3 && p <= kAtmP0) {
      double lapse, ti, pi, hi;
      if (p > p1) { lapse = l0; ti = kAtmT0; pi = kAtmP0; hi = -610.0; }
      else if (p > p2) { lapse = l1; ti = t1; pi = p1; hi = 17000.0; }
      else { lapse = l2; ti = t2; pi = p2; hi = 21000.0; }
      const double h = (pow(p / pi, -kRAir * lapse / kGravity) - 1) * ti / lapse + hi;   // :142-146
      *height = h;
      *temperature = ti + lapse * (h - hi);                                            // :148-149
    } else {
      double h, t;
      if (!atm_at_pressure_generic(alpha, p, &h, &t)) { ok = false; h = 0; t = 1; }
      *height = h;
      *temperature = t;
    }
  }
};

// ---- spherical offset (utils/spherical_geometry.py:44-76) -----------------------------------
template <typename Real>
BLE_HD void latlng_from_offset(Real lat0, Real lng0, Real x_m, Real y_m, Real* lat, Real* lng) {
  const Real heading = r_atan2(x_m / Real(1000), y_m / Real(1000));        // :61
  const Real angle = r_sqrt(x_m * x_m + y_m * y_m) / Real(kEarthRadiusM);  // :62
  const Real ca = r_cos(angle), sa = r_sin(angle);
  const Real sfl = r_sin(lat0), cfl = r_cos(lat0);
  const Real sin_lat = ca * sfl + sa * cfl * r_cos(heading);               // :69-70
  const Real d_lng = r_atan2(sa * cfl * r_sin(heading), ca - sfl * sin_lat);
  Real new_lat = r_asin(r_min(r_max(sin_lat, Real(-1)), Real(1)));
  new_lat = r_min(r_max(new_lat, Real(-kPi / 2)), Real(kPi / 2));
  Real new_lng = lng0 + d_lng;
  // s2 LatLng.normalized(): IEEE remainder(lng, 2 pi)
  if (new_lng > Real(kPi) || new_lng < Real(-kPi)) {
    new_lng = new_lng - Real(2 * kPi) * Real(rint(double(new_lng) / (2 * kPi)));
  }
  *lat = new_lat;
  *lng = new_lng;
}

// ---- NOAA solar calculator (env/balloon/solar.py:43-174) --------------------------------------
// Time-only part (shared by every balloon at the same timestamp, but computed per balloon
// here: start times differ per episode).  Julian arithmetic is fp64 in both modes.
template <typename Real>
struct SolarTime {
  double fod;               // fraction_of_day
  Real sin_decl, cos_decl;  // of the sun's declination
  double eq_time_deg;       // degrees(equation_of_time), as used at :115
  Real flux;
};

template <typename Real>
BLE_HD SolarTime<Real> solar_time(int64_t ts) {
  SolarTime<Real> o;
  int64_t days = ts / 86400;
  int64_t sod = ts - days * 86400;
  if (sod < 0) { sod += 86400; days -= 1; }
  o.fod = double(sod) / 86400.0;                                           // :66-68
  // Julian day number of 0h UT == 2440587.5 + days since the UNIX epoch (the calendar
  // formula at :71-75 evaluates to exactly this; tests/test_oracle_golden.py proves it).
  const double julian_time = (2440587.5 + double(days)) + o.fod;
  const double jc = (julian_time - 2451545.0) / 36525.0;                   // :78-79
  const double l0_deg = 280.46646 + jc * (36000.76983 + jc * 0.0003032);   // :82-83
  const double m0_deg = 357.52911 + jc * (35999.05029 - 0.0001537 * jc);   // :88-89
  const double om_deg = 125.04 - 1934.136 * jc;
  Real l0, m0, om;
  if (is_double<Real>::value) {
    l0 = Real(l0_deg * (kPi / 180.0)); m0 = Real(m0_deg * (kPi / 180.0)); om = Real(om_deg * (kPi / 180.0));
  } else {  // reduce the ~4800 degree angles in fp64 before fp32 trig
    l0 = Real(fmod(l0_deg, 360.0) * (kPi / 180.0));
    m0 = Real(fmod(m0_deg, 360.0) * (kPi / 180.0));
    om = Real(fmod(om_deg, 360.0) * (kPi / 180.0));
  }
  const Real jcr = Real(jc);
  const Real sin2l0 = r_sin(Real(2) * l0), cos2l0 = r_cos(Real(2) * l0), sin4l0 = r_sin(Real(4) * l0);
  const Real sinm0 = r_sin(m0), sin2m0 = r_sin(Real(2) * m0), sin3m0 = r_sin(Real(3) * m0);
  const Real mean_obl = r_rad(Real(23.0) + (Real(26.0) + ((Real(21.448) - jcr * (Real(46.815) + jcr *
                        (Real(0.00059) - jcr * Real(0.001813))))) / Real(60.0)) / Real(60.0));  // :94-97
  const Real obl = mean_obl + r_rad(Real(0.00256) * r_cos(om));            // :99-100
  const Real ty = r_tan(obl / Real(2));
  const Real var_y = ty * ty;                                              // :102
  const Real ecc = Real(0.016708634) - jcr * (Real(0.000042037) + Real(0.0000001267) * jcr);  // :104-105
  const Real eq_time = Real(4.0) * (var_y * sin2l0 - Real(2.0) * ecc * sinm0 +
                                    Real(4.0) * ecc * var_y * sinm0 * cos2l0 -
                                    Real(0.5) * var_y * var_y * sin4l0 -
                                    Real(1.25) * ecc * ecc * sin2m0);      // :107-111
  o.eq_time_deg = double(r_deg(eq_time));
  const Real eq_center = r_rad(sinm0 * (Real(1.914602) - jcr * (Real(0.004817) + Real(0.000014) * jcr)) +
                               sin2m0 * (Real(0.019993) - Real(0.000101) * jcr) +
                               sin3m0 * Real(0.000289));                   // :122-127
  const Real true_long = l0 + eq_center;
  const Real app_long = true_long - r_rad(Real(0.00569) - Real(0.00478) * r_sin(om));  // :129-131
  o.sin_decl = r_sin(obl) * r_sin(app_long);                               // sin(arcsin(.)) :132-133
  const Real decl = r_asin(o.sin_decl);
  o.cos_decl = r_cos(decl);
  const Real e1 = (Real(1) + ecc) / (Real(1) - ecc);
  o.flux = Real(1366.0) * (Real(1) + Real(0.5) * (e1 * e1 - Real(1)) * r_cos(m0));  // :170-172
  return o;
}

// Position-dependent part: cos(zenith angle) (:113-138).
template <typename Real>
BLE_HD Real solar_cos_zenith(const SolarTime<Real>& st, Real lat, Real lng) {
  const double lng_deg = double(lng) * (180.0 / kPi);
  const double ha_min = fmod(1440.0 * st.fod + st.eq_time_deg + 4.0 * lng_deg, 1440.0);  // :113-116
  double ha = (ha_min * (kPi / 180.0)) / 4.0;
  ha = (ha < 0) ? ha + kPi : ha - kPi;                                     // :117-120
  const Real hour_angle = Real(ha);
  const Real cz = r_sin(lat) * st.sin_decl + r_cos(lat) * st.cos_decl * r_cos(hour_angle);  // :135-138
  return r_min(r_max(cz, Real(-1)), Real(1));
}

// cos(zenith) -> refraction-corrected solar elevation in degrees (:135-157).
template <typename Real>
BLE_HD Real elevation_from_cos_zenith(Real cz) {
  const Real zenith = r_acos(cz);
  const Real el_unc = Real(90.0) - r_deg(zenith);                          // :141
  Real refraction;
  if (el_unc > Real(85.0)) {
    refraction = Real(0);
  } else if (el_unc > Real(5.0)) {
    const Real t = r_tan(r_rad(el_unc));
    refraction = Real(58.1) / t - Real(0.07) / (t * t * t) + Real(0.000086) / (t * t * t * t * t);
  } else if (el_unc > Real(-0.575)) {
    refraction = Real(1735.0) + el_unc * (Real(-518.2) + el_unc * (Real(103.4) + el_unc *
                 (Real(-12.79) + el_unc * Real(0.711))));
  } else {
    refraction = Real(-20.772) / r_tan(r_rad(el_unc));
  }
  return el_unc + refraction / Real(3600.0);                               // :157
}

// solar.py:212-236 with the two panel heights of solar_power folded to constants:
// degrees(atan2(sqrt(h (10.41603 + h)), 8.69275)) for h = 3.3 m and 2.7 m.
constexpr double kShadowEl33 = 37.738149050524044;
constexpr double kShadowEl27 = 34.39486500086289;

// Elevation plus its sine and cosine, which is all that attenuation (solar.py:204) and the panel
// projection (solar.py:532-534) need.  fp64: straightforward.  fp32: no trig beyond the acos --
// sin(el_unc) = cos(zenith), cos(el_unc) = sqrt(1 - cz^2) >= 0, tan(el_unc) = their ratio, and the
// refraction angle (<= 0.5 degree) is added with a 4th-order small-angle rotation (error < 1e-11).
template <typename Real>
struct SunAngles { Real el, sin_el, cos_el; };

template <typename Real>
BLE_HD SunAngles<Real> sun_angles_from_cos_zenith(Real cz) {
  SunAngles<Real> o;
  if (is_double<Real>::value) {
    o.el = elevation_from_cos_zenith<Real>(cz);
    o.sin_el = r_sin(r_rad(o.el));
    o.cos_el = r_cos(r_rad(o.el));
    return o;
  }
  const Real sz = r_sqrt(r_max(Real(1) - cz * cz, Real(0)));
  const Real el_unc = Real(90.0) - r_deg(r_acos(cz));                       // :141
  Real refraction;
  if (el_unc > Real(85.0)) {
    refraction = Real(0);
  } else if (el_unc > Real(5.0)) {
    const Real t = cz / sz;
    refraction = Real(58.1) / t - Real(0.07) / (t * t * t) + Real(0.000086) / (t * t * t * t * t);
  } else if (el_unc > Real(-0.575)) {
    refraction = Real(1735.0) + el_unc * (Real(-518.2) + el_unc * (Real(103.4) + el_unc *
                 (Real(-12.79) + el_unc * Real(0.711))));
  } else {
    refraction = Real(-20.772) * sz / cz;
  }
  const Real d_deg = refraction / Real(3600.0);
  const Real d = r_rad(d_deg), d2 = d * d;
  const Real sd = d * (Real(1) - d2 * Real(1.0 / 6.0));
  const Real cd = Real(1) - d2 * (Real(0.5) - d2 * Real(1.0 / 24.0));
  o.el = el_unc + d_deg;                                                   // :157
  o.sin_el = cz * cd + sz * sd;
  o.cos_el = sz * cd - cz * sd;
  return o;
}

// solar.py:177-209 with sin(el) supplied.
template <typename Real>
BLE_HD Real solar_attenuation_s(Real el_deg, Real sin_el, Real pressure) {
  if (el_deg < Real(kMinSolarElDeg)) return Real(0);
  const Real s = Real(614.0) * sin_el;
  const Real airmass = Real(0.34764) * (pressure / Real(101325.0)) * (r_sqrt(Real(1229.0) + s * s) - s);
  return Real(0.5) * (r_exp(Real(-0.65) * airmass) + r_exp(Real(-0.95) * airmass));
}

// solar.py:515-536 [W] with the attenuation and sin/cos(el) supplied:
// cos(el - a) = cos(el) cos(a) + sin(el) sin(a).
template <typename Real>
BLE_HD Real solar_power_sc(const SunAngles<Real>& a, Real att) {
  const Real sh33 = (a.el >= Real(kShadowEl33)) ? Real(0.4392) : Real(1);
  const Real sh27 = (a.el >= Real(kShadowEl27)) ? Real(0.4392) : Real(1);
  const Real c35 = a.cos_el * Real(0.81915204428899178969) + a.sin_el * Real(0.57357643635104609610);
  const Real c65 = a.cos_el * Real(0.42261826174069943619) + a.sin_el * Real(0.90630778703664996324);
  return Real(210.0) * att * (Real(4) * c35 * sh33 + Real(2) * c65 * sh27);
}

template <typename Real>
BLE_HD Real solar_elevation(const SolarTime<Real>& st, Real lat, Real lng) {
  return elevation_from_cos_zenith<Real>(solar_cos_zenith<Real>(st, lat, lng));
}

template <typename Real>
BLE_HD void solar_calculator(Real lat, Real lng, int64_t ts, Real* el_deg, Real* flux) {
  const SolarTime<Real> st = solar_time<Real>(ts);
  *el_deg = solar_elevation<Real>(st, lat, lng);
  *flux = st.flux;
}

// solar.py:177-209
template <typename Real>
BLE_HD Real solar_attenuation(Real el_deg, Real pressure) {
  if (el_deg < Real(kMinSolarElDeg)) return Real(0);
  const Real s = Real(614.0) * r_sin(r_rad(el_deg));
  const Real airmass = Real(0.34764) * (pressure / Real(101325.0)) * (r_sqrt(Real(1229.0) + s * s) - s);
  return Real(0.5) * (r_exp(Real(-0.65) * airmass) + r_exp(Real(-0.95) * airmass));
}


// solar.py:515-536 [W], with the attenuation factor (solar.py:177-209) supplied by the caller.
template <typename Real>
BLE_HD Real solar_power_att(Real el_deg, Real att) {
  const Real sh33 = (el_deg >= Real(kShadowEl33)) ? Real(0.4392) : Real(1);
  const Real sh27 = (el_deg >= Real(kShadowEl27)) ? Real(0.4392) : Real(1);
  return Real(210.0) * att * (Real(4) * r_cos(r_rad(el_deg - Real(35))) * sh33 +
                              Real(2) * r_cos(r_rad(el_deg - Real(65))) * sh27);
}

// solar.py:515-536 [W]
template <typename Real>
BLE_HD Real solar_power(Real el_deg, Real pressure) {
  const Real att = solar_attenuation<Real>(el_deg, pressure);
  const Real sh33 = (el_deg >= Real(kShadowEl33)) ? Real(0.4392) : Real(1);
  const Real sh27 = (el_deg >= Real(kShadowEl27)) ? Real(0.4392) : Real(1);
  return Real(210.0) * att * (Real(4) * r_cos(r_rad(el_deg - Real(35))) * sh33 +
                              Real(2) * r_cos(r_rad(el_deg - Real(65))) * sh27);
}

// ---- thermal model (env/balloon/thermal.py:28-230) ---------------------------------------------
template <typename Real> BLE_HD Real absorptivity_ir(Real t) {
  return Real(0.04587) + Real(0.000232) * (t - Real(210));                 // :77-91
}
template <typename Real> BLE_HD Real total_absorptivity(Real a) {
  const Real refl = Real(0.0291);
  return a * (Real(1) + (Real(1) - a - refl) / (Real(1) - refl));          // :138-147
}

// q_earth / balloon_area: depends on the per-episode upwelling IR only (thermal.py:213-217).
template <typename Real>
BLE_HD Real earth_heat_per_area(Real earth_flux) {
  const Real sigma = Real(0.000000056704);
  const Real t_earth = r_sqrt(r_sqrt(earth_flux / sigma));                 // (flux/sigma)^0.25 :66-75
  return earth_flux * Real(0.4605) * total_absorptivity<Real>(absorptivity_ir<Real>(t_earth));
}

// d_balloon_temperature_dt (thermal.py:175-230) with the attenuated solar flux (flux * att) and
// the earth term per unit area supplied by the caller.
template <typename Real>
BLE_HD Real d_temperature_dt_core(Real volume, Real mass, Real t_balloon, Real t_ambient, Real pressure,
                                  Real attenuated_flux, Real earth_per_area) {
  const Real sigma = Real(0.000000056704);
  const Real radius = r_cbrt(Real(3) * volume / Real(4 * kPi));            // :199
  const Real area = Real(4 * kPi) * radius * radius;
  const Real q_solar = attenuated_flux * Real(0.25) * area * total_absorptivity<Real>(Real(0.01435));
  const Real q_earth = earth_per_area * area;
  const Real tb2 = t_balloon * t_balloon;
  const Real q_emit = sigma * tb2 * tb2 * area * total_absorptivity<Real>(absorptivity_ir<Real>(t_balloon));
  // convective_heat_air_factor :150-172
  const Real visc = Real(1.458e-6) * (t_ambient * r_sqrt(t_ambient)) / (t_ambient + Real(110.4));
  const Real cond = Real(0.0241) * r_pow_fast(t_ambient / Real(273.15), Real(0.9));
  const Real prandtl = Real(0.804) - Real(3.25e-4) * t_ambient;
  const Real rho = pressure * Real(kMAir) / (Real(kR) * t_ambient);
  const Real d = Real(2) * radius;
  const Real grashof = (Real(9.80665) * (rho * rho) * (d * d * d) / (t_ambient * (visc * visc))) *
                       r_abs(t_ambient - t_balloon);
  const Real ra = prandtl * grashof;
  const Real nusselt = Real(2) + Real(0.457) * r_sqrt(r_sqrt(ra)) +
                       r_pow_fast(Real(1) + Real(2.69e-8) * ra, Real(1.0 / 12.0));
  const Real k_heat = nusselt * cond / d;
  const Real q_conv = area * (k_heat * (t_ambient - t_balloon));
  return (q_solar + q_earth + q_conv - q_emit) / (Real(1500) * mass);      // :229-230
}

template <typename Real>
BLE_HD Real d_balloon_temperature_dt(Real volume, Real mass, Real t_balloon, Real t_ambient,
                                     Real pressure, Real el_deg, Real flux, Real earth_flux) {
  const Real att = solar_attenuation<Real>(el_deg, pressure);
  return d_temperature_dt_core<Real>(volume, mass, t_balloon, t_ambient, pressure, flux * att,
                                     earth_heat_per_area<Real>(earth_flux));
}

// ---- ACS tables (env/balloon/acs.py:24-68) ----------------------------------------------------------
template <typename Real>
BLE_HD Real acs_most_efficient_power(Real pr) {
  // interp1d([1, 1.05, 1.2, 1.25, 1.35] -> [100, 100, 300, 400, 400]), linear, extrapolating
  // (both end segments are flat).
  if (pr <= Real(1.05)) return Real(100);
  if (pr <= Real(1.2)) return (Real(300.0 - 100.0) / Real(1.2 - 1.05)) * (pr - Real(1.05)) + Real(100);
  if (pr <= Real(1.25)) return (Real(400.0 - 300.0) / Real(1.25 - 1.2)) * (pr - Real(1.2)) + Real(300);
  return Real(400);
}

// [power 100,200,300,400][pressure ratio linspace(1.05, 1.35, 13)]  (acs.py:35-41).
// Kept in constant memory on the device (a function-local array would be rebuilt on the
// thread's stack at every call: 52 local stores per sub-step in the first profile).
#define BLE_ACS_EFF_TABLE                                                                \
  {{0.4, 0.4, 0.3, 0.2, 0.2, 0., 0., 0., 0., 0., 0., 0., 0.},                             \
   {0.4, 0.3, 0.3, 0.30, 0.25, 0.23, 0.20, 0.15, 0.12, 0.10, 0., 0., 0.},                 \
   {0., 0.3, 0.25, 0.25, 0.25, 0.20, 0.20, 0.20, 0.2, 0.15, 0.13, 0.12, 0.11},            \
   {0., 0.23, 0.23, 0.23, 0.23, 0.23, 0.20, 0.20, 0.20, 0.18, 0.16, 0.15, 0.13}}
static const double kAcsEffHost[4][13] = BLE_ACS_EFF_TABLE;
#if defined(__CUDACC__)
static __constant__ double kAcsEffDev[4][13] = BLE_ACS_EFF_TABLE;
#endif
BLE_HD double acs_eff_table(int j, int i) {
#if defined(__CUDA_ARCH__)
  return kAcsEffDev[j][i];
#else
  return kAcsEffHost[j][i];
#endif
}

template <typename Real>
BLE_HD Real acs_fan_efficiency(Real pr, Real power_w) {
  // bilinear, clamped outside the table (interp2d fill_value=None -> nearest)
  const double prc = fmin(fmax(double(pr), 1.05), 1.35);
  const double wc = fmin(fmax(double(power_w), 100.0), 400.0);
  const double fx = (prc - 1.05) / (0.3 / 12.0);
  const double fy = (wc - 100.0) / 100.0;
  int i = int(fx); if (i > 11) i = 11;
  int j = int(fy); if (j > 2) j = 2;
  const Real tx = Real(fx - i), ty = Real(fy - j);
  const Real e00 = Real(acs_eff_table(j, i)), e01 = Real(acs_eff_table(j, i + 1));
  const Real e10 = Real(acs_eff_table(j + 1, i)), e11 = Real(acs_eff_table(j + 1, i + 1));
  return (Real(1) - tx) * (Real(1) - ty) * e00 + tx * (Real(1) - ty) * e01 +
         (Real(1) - tx) * ty * e10 + tx * ty * e11;
}

// ---- envelope (env/balloon/balloon.py:552-609) --------------------------------------------------------
template <typename Real>
BLE_HD void superpressure_and_volume(Real mols_gas, Real mols_air, Real t_int, Real pressure,
                                     Real* volume, Real* superpressure) {
  const Real vu = (mols_gas + mols_air) * Real(kR) * t_int / pressure;     // :581-584
  if (vu <= Real(kVolBase)) {
    *volume = vu;
    *superpressure = Real(0);
    return;
  }
  const Real b = -(Real(kVolBase) - Real(kDvDp) * pressure);
  const Real c = -(Real(kDvDp) * vu * pressure);
  const Real v = Real(0.5) * (-b + r_sqrt(b * b - Real(4) * c));           // :604
  *volume = v;
  *superpressure = pressure * vu / v - pressure;                           // :605-607
}

// ---- safety layers ---------------------------------------------------------------------------------------
BLE_HD int paused_action(int a) { return a == kDown ? kStay : a; }         // power_safety.py:120-126

// env/balloon/envelope_safety.py:109-157
template <typename Real>
BLE_HD int envelope_safety(int action, Real sp, int* state) {
  int s = *state, ns;
  const bool low_keep = (s == kEnvLowCritical) || (s == kEnvLow);
  const bool high_keep = (s == kEnvHigh) || (s == kEnvHighCritical);
  if (sp < Real(150)) ns = kEnvLowCritical;
  else if (sp < Real(250)) ns = kEnvLow;
  else if (sp < Real(250 + 50)) ns = low_keep ? kEnvLow : kEnvNominal;
  else if (sp < Real(kMaxSuperpressure - 250 - 50)) ns = kEnvNominal;
  else if (sp < Real(kMaxSuperpressure - 250)) ns = high_keep ? kEnvHigh : kEnvNominal;
  else if (sp < Real(kMaxSuperpressure - 150)) ns = kEnvHigh;
  else ns = kEnvHighCritical;
  *state = ns;
  if (ns == kEnvLowCritical || ns == kEnvHighCritical) return kUp;
  if (ns == kEnvLow || ns == kEnvHigh) return paused_action(action);
  return action;
}

// env/balloon/altitude_safety.py:73-111 (50 kft floor, 500 ft buffer + 500 ft hysteresis)
constexpr double kAltMin = 50000.0 * 0.3048;
constexpr double kAltBuf = 500.0 * 0.3048;
template <typename Real>
BLE_HD int altitude_safety(int action, Real altitude_m, int* state) {
  const int s = *state;
  int ns;
  const bool was_low = (s == kAltVeryLow) || (s == kAltLow);
  if (altitude_m < Real(kAltMin)) ns = kAltVeryLow;
  else if (altitude_m < Real(kAltMin + kAltBuf)) ns = kAltLow;
  else if (altitude_m < Real(kAltMin + kAltBuf + kAltBuf)) ns = was_low ? kAltLow : kAltNominal;
  else ns = kAltNominal;
  *state = ns;
  if (ns == kAltVeryLow) return kUp;
  if (ns == kAltLow) return paused_action(action);
  return action;
}

// env/balloon/power_safety.py:52-118.  Times are integer UNIX seconds.
template <typename Real>
BLE_HD int power_safety(int action, int64_t now, Real charge_wh, int64_t* sunrise_h, int64_t* sunset,
                        int* paused) {
  if (now > *sunrise_h) *sunrise_h += ((now - *sunrise_h + 86399) / 86400) * 86400;  // :83-84
  if (now > *sunset) *sunset += ((now - *sunset + 86399) / 86400) * 86400;           // :85-86
  if (*sunset < *sunrise_h) {                                              // daytime :88
    const Real soc = charge_wh / Real(kBatteryCapacityWh);
    if (*paused && soc < Real(0.05)) return paused_action(action);
    *paused = 0;
    return action;
  }
  if (*paused) return paused_action(action);
  const Real hours = Real(double(*sunrise_h - now) / 3600.0);
  const Real floating = Real(kNightLoadW) * hours;                         // :107-109
  if ((charge_wh - floating) / Real(kBatteryCapacityWh) < Real(0.025)) {
    *paused = 1;
    return paused_action(action);
  }
  return action;
}

// ---- balloon state + Euler sub-step (env/balloon/balloon.py:356-549) ------------------------------------------
//
// Precision note.  The reference integrates with explicit Euler at h = 10 s, and its vertical
// dynamics  dp/dt = (dp/dh) * sign * sqrt(2 |rho V - m| g / (rho C_d V^(2/3)))  are stiff near
// equilibrium: measured on the oracle, a pressure perturbation entering one agent step is damped
// 1000x in the median but amplified up to ~70x (p99 ~4x) after stable-init/venting, and
// perturbations injected at EVERY sub-step by fp32 rounding of p, T_ambient or V (3e-5 kg of
// lift) grew to 8 Pa (8e-4 relative) within one step.  The variables of that loop -- x, y,
// pressure, ambient/internal temperature, volume, superpressure, mols_air, battery charge -- are
// therefore carried in fp64 (registers and HBM), and only the smooth right-hand sides (solar
// geometry, thermal model, ACS tables, solar power, reward) are evaluated in `Real`.
template <typename Real>
struct BalloonState {
  double x, y, pressure, t_ambient, t_internal, volume, superpressure, mols_air, charge;   // fp64 accumulators
  Real acs_power, acs_flow, solar_w, load_w;
  // per-episode constants
  Real lat0, lng0, ir, mols_gas;
  int64_t date_time;         // UNIX seconds
  int32_t time_elapsed;      // seconds
  int status;
};

// Sun along one agent step.  The 18 sub-steps need the solar elevation at (x_k, y_k, t_k) with
// x, y moving linearly (wind is held constant, env/balloon/balloon.py:321-325) and t_k = t_0 + 10 k.
//   Real = double (audit): evaluated exactly at every sub-step, as the reference does (:451-452).
//   Real = float: cos(zenith) is evaluated exactly at k = 0, 9, 18 and interpolated quadratically
//   (max error 2.5e-8 over 2e5 random tracks, below fp32 epsilon; elevation itself has a cusp at
//   the zenith and must not be interpolated); the acos, the refraction branches (:143-155) and
//   everything downstream stay per sub-step.  Flux is interpolated linearly (1e-6 relative / step).
template <typename Real>
struct SunTrack {
  Real c0, c1, c2, f0, f2;

  BLE_HD static void exact(const BalloonState<Real>& s, double x, double y, int64_t ts, Real* cz, Real* flux) {
    Real lat, lng;
    latlng_from_offset<Real>(s.lat0, s.lng0, Real(x), Real(y), &lat, &lng);
    const SolarTime<Real> st = solar_time<Real>(ts);
    *cz = solar_cos_zenith<Real>(st, lat, lng);
    *flux = st.flux;
  }

  BLE_HD void init(const BalloonState<Real>& s, double u, double v) {
    if (!is_double<Real>::value) {
      Real f1;
      exact(s, s.x, s.y, s.date_time, &c0, &f0);
      exact(s, s.x + u * 90.0, s.y + v * 90.0, s.date_time + 90, &c1, &f1);
      exact(s, s.x + u * 180.0, s.y + v * 180.0, s.date_time + 180, &c2, &f2);
    }
  }

  // Sun at the state `s` has reached after k sub-steps of this agent step.
  BLE_HD void at(const BalloonState<Real>& s, int k, SunAngles<Real>* sun, Real* flux) const {
    Real cz;
    if (is_double<Real>::value) {
      exact(s, s.x, s.y, s.date_time, &cz, flux);
    } else {
      const Real t = Real(k) * Real(1.0 / kSubSteps);
      cz = c0 + t * ((Real(-3) * c0 + Real(4) * c1 - c2) + t * (Real(2) * c0 - Real(4) * c1 + Real(2) * c2));
      cz = r_min(r_max(cz, Real(-1)), Real(1));
      *flux = f0 + t * (f2 - f0);
    }
    *sun = sun_angles_from_cos_zenith<Real>(cz);
  }
};

// ---- one explicit-Euler sub-step (:356-549) as four independent "roles" ---------------------------------
// Every right-hand side reads only the OLD state, so the sub-step splits into four pieces with no
// data dependence on each other.  The single-thread path (euler_substep) calls them in sequence;
// k_step_ws in ble_engine.cu gives each role its own warp and exchanges the results through shared
// memory once per sub-step, which cuts the dependent instruction chain per sub-step ~5x.

// Role P: buoyancy -> dh/dt -> dp/dt and the ambient temperature (:412-445, :457-458), fp64.
template <typename Real>
BLE_HD void role_pressure(Atmosphere& atm, double pressure, double t_ambient, double volume, double mols_air,
                          double mols_gas, double* new_pressure, double* new_t_ambient) {
  const double dt = double(kStrideS);
  const double rho = (pressure * kMAir) / (kR * t_ambient);
  const double cv = double(r_cbrt(Real(volume)));                          // drag enters multiplicatively
  const double drag = kCod * cv * cv;                                      // V^(2/3) :415
  const double mass = kMHe * mols_gas + kMAir * mols_air + kEnvelopeMass + kPayloadMass;
  const double lift = rho * volume;
  const double direction = (lift >= mass) ? 1.0 : -1.0;
  const double dh_dt = direction * sqrt(fabs(2 * (lift - mass) * kGravity / (rho * drag)));   // :424-427
  double dh;
  atm.temperature_and_secant(pressure, direction, new_t_ambient, &dh);     // :438-441
  const double dp_dh = direction / dh;                                     // :442
  *new_pressure = pressure + dp_dh * dh_dt * dt;                           // :443-445
}

// Role T: the sun-independent part of d_balloon_temperature_dt (thermal.py:175-230):
// (q_earth + q_convective - q_emitted) / (c m).
template <typename Real>
BLE_HD Real role_thermal_body(double volume, double t_internal, double t_ambient, double pressure, Real earth_per_area) {
  return d_temperature_dt_core<Real>(Real(volume), Real(kEnvelopeMass), Real(t_internal), Real(t_ambient),
                                     Real(pressure), Real(0), earth_per_area);
}

// Role E: envelope volume / superpressure (:470-482) and the ACS (:487-519).
template <typename Real>
BLE_HD void role_envelope_acs(double mols_gas, double mols_air, double t_internal, double pressure, double superpressure,
                              int action, double* new_volume, double* new_sp, double* new_mols_air,
                              Real* acs_power_out, Real* flow_out, int* status_env) {
  const double dt = double(kStrideS);
  superpressure_and_volume<double>(mols_gas, mols_air, t_internal, pressure, new_volume, new_sp);
  int status = kOk;
  if (*new_sp > kMaxSuperpressure) status = kBurst;
  if (*new_sp <= 0.0) status = kZeroPressure;
  *status_env = status;
  Real acs_power = Real(0), flow = Real(0);
  if (action == kUp) {
    const Real valve_area = Real(kPi * kValveDiameter * kValveDiameter / 4.0);
    const Real gas_density = Real((superpressure + pressure) * kMAir / (kR * t_internal));
    flow = Real(-kValveCd) * valve_area * r_sqrt(Real(2) * Real(superpressure) * gas_density);
  } else if (action == kDown) {
    const Real pr = Real((pressure + fmax(superpressure, 0.0)) / pressure);   // :247-250
    acs_power = acs_most_efficient_power<Real>(pr);
    flow = acs_fan_efficiency<Real>(pr, acs_power) * acs_power / Real(3600);   // acs.py:67-68
  }
  *new_mols_air = fmax(mols_air + (double(flow) / kMAir) * dt, 0.0);
  *acs_power_out = acs_power;
  *flow_out = flow;
}

// Role S: everything that needs the sun -- the solar part of dT/dt and the power system (:524-542).
// `acs_power` is the power role E chose for this sub-step (a pure function of the old state, so
// role S recomputes it instead of waiting for role E).
template <typename Real>
BLE_HD void role_sun_power(const SunAngles<Real>& sun, Real flux, double volume, double pressure, double superpressure,
                           double charge, int action, Real* d_t_solar, Real* solar_w_out, Real* load_w_out,
                           double* new_charge, int* out_of_power, Real* acs_power_out = nullptr) {
  const Real p_r = Real(pressure);
  const Real att = solar_attenuation_s<Real>(sun.el, sun.sin_el, p_r);
  const Real radius = r_cbrt(Real(3) * Real(volume) / Real(4 * kPi));
  const Real area = Real(4 * kPi) * radius * radius;
  *d_t_solar = (flux * att) * Real(0.25) * area * total_absorptivity<Real>(Real(0.01435)) /
               (Real(1500) * Real(kEnvelopeMass));
  Real acs_power = Real(0);
  if (action == kDown) acs_power = acs_most_efficient_power<Real>(Real((pressure + fmax(superpressure, 0.0)) / pressure));
  const bool is_day = sun.el > Real(kMinSolarElDeg);
  const Real solar_w = is_day ? solar_power_sc<Real>(sun, att) : Real(0);
  const Real load_w = (is_day ? Real(kDayLoadW) : Real(kNightLoadW)) + acs_power;
  double c = charge + double(solar_w - load_w) * (double(kStrideS) / 3600.0);
  c = fmin(fmax(c, 0.0), kBatteryCapacityWh);
  *new_charge = c;
  *out_of_power = (c <= 0.0) ? 1 : 0;
  *solar_w_out = solar_w;
  *load_w_out = load_w;
  if (acs_power_out != nullptr) *acs_power_out = acs_power;
}

// The four roles in sequence = one sub-step for one thread.
template <typename Real>
BLE_HD void euler_substep(BalloonState<Real>& s, Atmosphere& atm, double u, double v, int action,
                          const SunAngles<Real>& sun, Real flux, Real earth_per_area) {
  const double dt = double(kStrideS);
  double new_pressure, new_t_ambient, new_volume, new_sp, new_mols_air, new_charge;
  Real acs_power, flow, d_t_solar, solar_w, load_w;
  int status_env, out_of_power;
  role_pressure<Real>(atm, s.pressure, s.t_ambient, s.volume, s.mols_air, double(s.mols_gas), &new_pressure, &new_t_ambient);
  const Real d_t_body = role_thermal_body<Real>(s.volume, s.t_internal, s.t_ambient, s.pressure, earth_per_area);
  role_envelope_acs<Real>(double(s.mols_gas), s.mols_air, s.t_internal, s.pressure, s.superpressure, action,
                          &new_volume, &new_sp, &new_mols_air, &acs_power, &flow, &status_env);
  role_sun_power<Real>(sun, flux, s.volume, s.pressure, s.superpressure, s.charge, action, &d_t_solar, &solar_w,
                       &load_w, &new_charge, &out_of_power);
  int status = s.status;
  if (status_env != kOk) status = status_env;
  if (out_of_power) status = kOutOfPower;                                  // later assignment wins (:541-542)
  s.x += u * dt;                                                           // :394-395
  s.y += v * dt;
  s.pressure = new_pressure;
  s.t_ambient = new_t_ambient;
  s.t_internal = s.t_internal + double(d_t_body + d_t_solar) * dt;         // :462-467
  s.volume = new_volume;
  s.superpressure = new_sp;
  s.mols_air = new_mols_air;
  s.charge = new_charge;
  s.acs_power = acs_power;
  s.acs_flow = flow;
  s.solar_w = solar_w;
  s.load_w = load_w;
  s.status = status;
  s.date_time += kStrideS;                                                 // :546-547
  s.time_elapsed += kStrideS;
}

// env/balloon_env.py:44-102 evaluated on the post-step state; `el` = sun at that state.
template <typename Real>
BLE_HD Real perciatelli_reward(const BalloonState<Real>& s, int last_command, Real el) {
  const Real dist = Real(sqrt(s.x * s.x + s.y * s.y));
  const Real radius = Real(50000.0);
  Real reward = Real(1);
  if (!(dist <= radius)) {
    reward = Real(0.4) * r_exp(Real(-0.69314718056 / 100.0) * ((dist - radius) / Real(1000)));
  }
  if (last_command == kDown) {
    const bool excess = (solar_power<Real>(el, Real(s.pressure)) > Real(kDayLoadW)) &&
                        (s.charge / kBatteryCapacityWh > 0.99);            // balloon.py:231-238
    if (!excess) {
      const Real scale = r_min(r_max((s.acs_power - Real(100)) / Real(200), Real(0)), Real(1));
      reward *= Real(0.95) - Real(0.3) * scale;
    }
  }
  return reward;
}

}  // namespace ble

// =====================================================================================================
// Reset-time helpers (always fp64: not on the step path, and the sunrise/sunset search makes
// discrete choices from elevation comparisons).
// =====================================================================================================
namespace ble {

// _find_solar_elevation_binary_search (env/balloon/solar.py:296-375).
// kind 0: minimise el ('minimum'), 1: maximise el ('maximum'), 2: minimise |el - target|.
BLE_HD double search_objective(double lat, double lng, int64_t ts, int kind, double target) {
  double el, flux;
  solar_calculator<double>(lat, lng, ts, &el, &flux);
  return kind == 0 ? el : (kind == 1 ? -el : fabs(el - target));
}

BLE_HD int64_t find_solar_elevation(double lat, double lng, int64_t min_ts, int64_t max_ts, int kind,
                                    double target, bool* ok) {
  const int64_t delta = 180;                                               // _SEARCH_TIME_DELTA :39
  if (max_ts < min_ts) { *ok = false; return min_ts; }                     // :318-319
  const int64_t max_steps = (max_ts - min_ts) / delta;                     // :321
  if (max_steps <= 0) { *ok = false; return min_ts; }                      // :322
  int64_t low = 0, high = max_steps;
  while (high > low + 1) {                                                 // :358-364
    const double midpoint = double(low) + double(high - low) / 2.0;
    if (search_objective(lat, lng, min_ts + delta * low, kind, target) <
        search_objective(lat, lng, min_ts + delta * high, kind, target)) {
      high = int64_t(ceil(midpoint));
    } else {
      low = int64_t(floor(midpoint));
    }
  }
  const int64_t idx = (search_objective(lat, lng, min_ts + delta * low, kind, target) <
                       search_objective(lat, lng, min_ts + delta * high, kind, target)) ? low : high;
  return min_ts + delta * idx;                                             // :367-372
}

// get_next_sunrise_sunset (env/balloon/solar.py:432-483).
BLE_HD bool next_sunrise_sunset(double lat, double lng, int64_t ts, int64_t* sunrise, int64_t* sunset) {
  bool ok = fabs(lat * (180.0 / kPi)) < 60.0;                              // :449
  const int64_t h12 = 12 * 3600, h24 = 24 * 3600;
  double el0, el1, f;
  solar_calculator<double>(lat, lng, ts, &el0, &f);
  solar_calculator<double>(lat, lng, ts + 1, &el1, &f);
  const bool afternoon = el1 < el0;                                        // is_solar_afternoon :239-255
  const int64_t noon_lo = afternoon ? ts + h12 : ts;                       // get_next_solar_noon :405-429
  const int64_t next_noon = find_solar_elevation(lat, lng, noon_lo, noon_lo + h12, 1, 0.0, &ok);
  const int64_t mid_lo = afternoon ? ts : ts + h12;                        // get_next_solar_midnight :378-402
  const int64_t next_midnight = find_solar_elevation(lat, lng, mid_lo, mid_lo + h12, 0, 0.0, &ok);
  const int64_t sr_lo = afternoon ? next_midnight : next_midnight - h24;   // :458-475
  int64_t sr = find_solar_elevation(lat, lng, sr_lo, next_noon, 2, kMinSolarElDeg, &ok);
  const int64_t ss_lo = afternoon ? next_noon - h24 : next_noon;
  int64_t ss = find_solar_elevation(lat, lng, ss_lo, next_midnight, 2, kMinSolarElDeg, &ok);
  if (sr < ts) sr += h24;                                                  // :478-481
  if (ss < ts) ss += h24;
  *sunrise = sr;
  *sunset = ss;
  return ok;
}

// calculate_stable_params_for_pressure (env/balloon/stable_init.py:40-129).
struct StableParams { double t_ambient, t_internal, mols_air, volume, superpressure; bool ok; };

BLE_HD StableParams stable_params(double alpha, double pressure, double mols_gas, double lat, double lng,
                                  int64_t ts, double ir) {
  StableParams o;
  double h;
  o.ok = atm_at_pressure_generic(alpha, pressure, &h, &o.t_ambient);       // :76
  double mols_air = (pressure * kMAir * kVolBase / (kR * o.t_ambient) - kEnvelopeMass - kPayloadMass -
                     kMHe * mols_gas) / kMAir;                             // :92-96
  o.mols_air = mols_air < 0.0 ? 0.0 : mols_air;                            // :98
  double t_int = 206.0;                                                    // :101
  double el, flux;
  solar_calculator<double>(lat, lng, ts, &el, &flux);                      // :102
  const double delta = 0.01;
  for (int it = 0; it < 10; ++it) {                                        // :107-127
    const double d1 = d_balloon_temperature_dt<double>(kVolBase, kEnvelopeMass, t_int - delta / 2,
                                                       o.t_ambient, pressure, el, flux, ir);
    const double d2 = d_balloon_temperature_dt<double>(kVolBase, kEnvelopeMass, t_int + delta / 2,
                                                       o.t_ambient, pressure, el, flux, ir);
    const double d2t = (d2 - d1) / delta;
    const double mean_d = (d1 + d2) / 2.0;
    if (fabs(d2t) > 0.0) t_int -= mean_d / d2t;
    if (fabs(mean_d) < 1e-5) break;
  }
  o.t_internal = t_int;
  superpressure_and_volume<double>(mols_gas, o.mols_air, t_int, pressure, &o.volume, &o.superpressure);
  return o;
}

}  // namespace ble

// =====================================================================================================
// One agent step for one balloon: safety-layer chain, 18 Euler sub-steps, reward.
// Balloon.simulate_step (env/balloon/balloon.py:263-328) + BalloonEnv.step's reward/terminal
// (env/balloon_env.py:157-190).
// =====================================================================================================
namespace ble {

struct SafetyState {
  int64_t sunrise_h, sunset;
  int envelope_state, altitude_state, power_paused, power_safety_enabled, last_command;
};

template <typename Real>
BLE_HD Real agent_step(BalloonState<Real>& s, Atmosphere& atm, SafetyState& ss, int action,
                       double u, double v, int* effective_action) {
  ss.last_command = action;                                                // balloon.py:286
  int eff = action;
  if (ss.power_safety_enabled) {                                           // :305-309
    eff = power_safety<double>(eff, s.date_time, s.charge, &ss.sunrise_h, &ss.sunset, &ss.power_paused);
  }
  eff = envelope_safety<double>(eff, s.superpressure, &ss.envelope_state);  // :310-311
  double altitude, t_unused;
  atm.at_pressure(s.pressure, &altitude, &t_unused);
  eff = altitude_safety<double>(eff, altitude, &ss.altitude_state);         // :312-313
  *effective_action = eff;
  SunTrack<Real> sun;
  sun.init(s, u, v);
  const Real earth_per_area = earth_heat_per_area<Real>(s.ir);
  atm.incremental = !is_double<Real>::value;
  int k = 0;
  SunAngles<Real> ang;
  Real flux;
#pragma unroll 1
  for (; k < kSubSteps;) {                                                 // :321-328
    sun.at(s, k, &ang, &flux);
    euler_substep<Real>(s, atm, u, v, eff, ang, flux, earth_per_area);
    ++k;
    if (s.status != kOk) break;
  }
  ang.el = Real(0);
  if (action == kDown) sun.at(s, k, &ang, &flux);                          // excess_energy's sun (balloon.py:231-238)
  return perciatelli_reward<Real>(s, action, ang.el);
}

}  // namespace ble
