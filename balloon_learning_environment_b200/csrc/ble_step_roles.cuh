// Production (fp32 right-hand sides, fp64 stiff state) Euler sub-step, split into four ROLES that
// only read the OLD state (explicit Euler, env/balloon/balloon.py:356-549), written branch-light on
// the primitives of ble_fastmath.cuh.  k_step_fused (ble_step_fused.cuh) gives every role its own
// warp; tests/hostemu replays the same functions on the CPU in sequence.
//
// Same physics as ble_physics.cuh's role_* functions (which stay the fp64 audit path and cite the
// reference line by line); what changes is the arithmetic shape:
//   * no IEEE division / sqrt expansions and no libm slow paths in the loop (fm::rcp, fm::rsqrt, ...);
//   * rho = (p M)/(R T) is never formed: |2 (rho V - m) g / (rho C_d V^(2/3))| = 2 g |p M V - m R T| / (p M C_d V^(2/3));
//   * cbrt(V) is computed once per sub-step by the envelope role and shared (drag, balloon area);
//   * X = (p/P_i)^k is carried incrementally with a degree-8 binomial series in Estrin form
//     (4 dependent DFMA instead of 27) whose coefficients C(k, n) are per-step constants; the same
//     coefficients give the height secant h(p +- 1) - h(p).
#pragma once
#include "ble_fastmath.cuh"
#include "ble_physics.cuh"

namespace ble {
namespace roles {

// Rare paths are kept out of line: the step kernel is instruction-cache bound if they are inlined at every use.
BLE_HD_NOINLINE double pow_by_exp_log(double x, double k) { return exp(k * log(x)); }
BLE_HD_NOINLINE void secant_slow_path(Atmosphere atm, double p, double direction, double* t_new, double* dh, bool* ok) {
  atm.lcache = -1;
  atm.incremental = false;
  atm.temperature_and_secant(p, direction, t_new, dh);
  *ok = atm.ok;
}

// ---- role P: buoyancy -> dh/dt -> dp/dt, ambient temperature (balloon.py:412-445,457-458) -----------
// Holds only what the fast path reads, so that it lives in registers; the atmosphere tables themselves are fetched
// through `load_atm()` (a callable returning an Atmosphere) on the rare paths: first use, a change of layer, a
// pressure outside layers 0..2.
struct PressureRole {
  double mass0;              // kMHe * mols_gas + envelope + payload
  // cached atmosphere layer of the balloon (standard_atmosphere.py:122-154)
  int layer;                 // 0, 1, 2; -1 = take the generic path
  double p_lo, p_hi;         // the layer holds pressures in (p_lo, p_hi]
  double k, ti, pi, tl;      // k = -R_air L / g, T_i, P_i, T_i / L
  double c1, c2, c3, c4, c5, c6, c7, c8;   // binomial coefficients C(k, n)
  // X = (p / P_i)^k at the pressure of the previous evaluation
  bool have_x, atm_ok;
  double x_prev, p_prev, inv_p_prev;

  BLE_HD void select_layer(const Atmosphere& atm, double p) {
    double lapse = 0;
    layer = -1;
    if (p > atm.p3 && p <= kAtmP0) {
      if (p > atm.p1) { layer = 0; lapse = atm.l0; ti = kAtmT0; pi = kAtmP0; p_lo = atm.p1; p_hi = kAtmP0; }
      else if (p > atm.p2) { layer = 1; lapse = atm.l1; ti = atm.t1; pi = atm.p1; p_lo = atm.p2; p_hi = atm.p1; }
      else { layer = 2; lapse = atm.l2; ti = atm.t2; pi = atm.p2; p_lo = atm.p3; p_hi = atm.p2; }
    }
    have_x = false;
    if (layer < 0 || lapse == 0.0) { layer = -1; return; }
    k = -kRAir * lapse / kGravity;
    tl = ti / lapse;
    c1 = k;
    c2 = c1 * (k - 1.0) * (1.0 / 2.0);
    c3 = c2 * (k - 2.0) * (1.0 / 3.0);
    c4 = c3 * (k - 3.0) * (1.0 / 4.0);
    c5 = c4 * (k - 4.0) * (1.0 / 5.0);
    c6 = c5 * (k - 5.0) * (1.0 / 6.0);
    c7 = c6 * (k - 6.0) * (1.0 / 7.0);
    c8 = c7 * (k - 7.0) * (1.0 / 8.0);
  }

  BLE_HD void init(const Atmosphere& atm, double mols_gas, double p) {
    mass0 = kMHe * mols_gas + kEnvelopeMass + kPayloadMass;
    atm_ok = true;
    x_prev = 1.0; p_prev = p; inv_p_prev = 0.0;
    select_layer(atm, p);
  }

  // X = (p / P_i)^k from scratch (once per agent step; also what the altitude safety layer needs:
  // h = (X - 1) T_i / L + H_i).
  BLE_HD double x_from_scratch(double p) const { return pow_by_exp_log(p / pi, k); }

  BLE_HD void seed_x(double p, double x) { have_x = true; p_prev = p; x_prev = x; inv_p_prev = fm::rcp(p); }

  // (1 + r)^k, |r| <= 0.02, degree 8 in Estrin form: the neglected term is C(k, 9) r^9 < 1e-17.
  BLE_HD double binomial8(double r) const {
    const double r2 = r * r, r4 = r2 * r2;
    const double a = fma(c1, r, 1.0), b = fma(c3, r, c2), c = fma(c5, r, c4), d = fma(c7, r, c6);
    const double lo = fma(r2, b, a), hi = fma(r2, d, c);
    return fma(r4 * r4, c8, fma(r4, hi, lo));
  }

  // One sub-step: new pressure and the ambient temperature AT THE OLD pressure (balloon.py:457-458).
  // cv = cbrt(volume).
  template <typename LoadAtm>
  BLE_HD void step(double p, double t_ambient, double volume, double mols_air, float cv, LoadAtm load_atm,
                   double* new_pressure, double* new_t_ambient) {
    const double dt = double(kStrideS);
    const double a = p * kMAir, b = kR * t_ambient;                        // rho = a / b
    const double drag = kCod * double(cv) * double(cv);                    // C_d V^(2/3) :415
    const double mass = fma(kMAir, mols_air, mass0);
    const double num = fma(a, volume, -(mass * b));                        // (rho V - m) b
    const double direction = (num >= 0.0) ? 1.0 : -1.0;                    // lift >= mass :421-423
    const double dh_dt_abs = fm::sqrt_pos((2.0 * kGravity) * fabs(num) * fm::rcp(a * drag));   // :424-427
    const double q = p + direction;
    double t_new, dh;
    if (layer >= 0 && p > p_lo && p <= p_hi && q > p_lo && q <= p_hi) {
      const double inv_p = fm::rcp(p);
      double x;
      const double r = (p - p_prev) * inv_p_prev;
      if (have_x && fabs(r) < 0.02) x = x_prev * binomial8(r);
      else x = x_from_scratch(p);
      have_x = true; p_prev = p; x_prev = x; inv_p_prev = inv_p;
      const double e = direction * inv_p;
      const double series = e * fma(e, fma(e, fma(e, c4, c3), c2), c1);    // (1 + e)^k - 1, e ~ 1e-4
      t_new = ti * x;                                                      // standard_atmosphere.py:148-149
      dh = tl * x * series;                                                // h(p + dir) - h(p) :438-441
    } else {
      const Atmosphere atm = load_atm();
      bool ok;
      secant_slow_path(atm, p, direction, &t_new, &dh, &ok);
      atm_ok = atm_ok && ok;
      select_layer(atm, p);
    }
    *new_t_ambient = t_new;
    *new_pressure = fma(dh_dt_abs * dt, fm::rcp(dh), p);                   // p + (dir / dh) (dir |dh/dt|) dt :442-445
  }
};

// ---- role T: sun-independent part of d_balloon_temperature_dt (thermal.py:175-230) -------------------
// cv = cbrt(volume); returns (q_earth + q_convective - q_emitted) / (c m) in K/s.
BLE_HD float thermal_body(float cv, double t_internal, double t_ambient, double pressure, float earth_per_area) {
  const float sigma = 0.000000056704f;
  const float radius = cv * 0.62035049089940f;                             // (3 / 4 pi)^(1/3)  :199
  const float area = float(4 * kPi) * radius * radius;
  const float tb = float(t_internal), ta = float(t_ambient), pr = float(pressure);
  const float dtemp = float(t_ambient - t_internal);
  const float abs_ir = 0.04587f + 0.000232f * (tb - 210.0f);               // :77-91
  const float tot = abs_ir * (1.0f + (1.0f - abs_ir - 0.0291f) * (1.0f / (1.0f - 0.0291f)));   // :138-147
  const float tb2 = tb * tb;
  const float q_emit = sigma * tb2 * tb2 * tot;                            // per unit area
  // convective_heat_air_factor :150-172
  const float inv_ta = fm::rcpf(ta);
  const float visc = 1.458e-6f * (ta * fm::sqrtf_pos(ta)) * fm::rcpf(ta + 110.4f);
  const float cond = 0.0241f * fm::powf_pos(ta * (1.0f / 273.15f), 0.9f);
  const float prandtl = 0.804f - 3.25e-4f * ta;
  const float rho = pr * float(kMAir / kR) * inv_ta;
  const float d = 2.0f * radius;
  const float rv = rho * fm::rcpf(visc);
  const float grashof = 9.80665f * (rv * rv) * (d * d * d) * inv_ta * fabsf(dtemp);
  const float ra = prandtl * grashof;
  const float nusselt = 2.0f + 0.457f * fm::sqrtf_pos(fm::sqrtf_pos(ra)) +
                        fm::powf_pos(1.0f + 2.69e-8f * ra, 1.0f / 12.0f);
  const float k_heat = nusselt * cond * fm::rcpf(d);
  const float per_area = earth_per_area + k_heat * dtemp - q_emit;
  return area * per_area * float(1.0 / (1500.0 * kEnvelopeMass));          // :229-230
}

// ---- role E: envelope volume / superpressure (balloon.py:470-482,552-609) and the ACS (:487-519) ------
struct EnvelopeOut {
  double volume, superpressure, mols_air;
  float cv, acs_power, flow;
  int status;
};

BLE_HD float acs_fan_efficiency_f(float pr_minus_1, float power_w) {
  // bilinear over pr in linspace(1.05, 1.35, 13) x W in (100, 200, 300, 400), clamped outside (acs.py:24-68)
  const float fx = fminf(fmaxf((pr_minus_1 - 0.05f) * 40.0f, 0.0f), 12.0f);
  const float fy = fminf(fmaxf((power_w - 100.0f) * 0.01f, 0.0f), 3.0f);
  int i = int(fx); if (i > 11) i = 11;
  int j = int(fy); if (j > 2) j = 2;
  const float tx = fx - float(i), ty = fy - float(j);
  const float e00 = float(acs_eff_table(j, i)), e01 = float(acs_eff_table(j, i + 1));
  const float e10 = float(acs_eff_table(j + 1, i)), e11 = float(acs_eff_table(j + 1, i + 1));
  const float lo = e00 + tx * (e01 - e00), hi = e10 + tx * (e11 - e10);
  return lo + ty * (hi - lo);
}

BLE_HD float acs_power_f(float pr_minus_1) {
  // interp1d([1, 1.05, 1.2, 1.25, 1.35] -> [100, 100, 300, 400, 400]) (acs.py:24-29)
  const float a = 100.0f + (200.0f / 0.15f) * (pr_minus_1 - 0.05f);
  const float b = 300.0f + (100.0f / 0.05f) * (pr_minus_1 - 0.2f);
  const float w = pr_minus_1 <= 0.2f ? a : b;
  return fminf(fmaxf(w, 100.0f), 400.0f);
}

BLE_HD EnvelopeOut envelope_acs(double mols_gas, double mols_air, double t_internal, double pressure,
                                double superpressure, int action) {
  EnvelopeOut o;
  const double dt = double(kStrideS);
  const double inv_p = fm::rcp(pressure);
  const double vu = (mols_gas + mols_air) * kR * t_internal * inv_p;       // :581-584
  const double bq = fma(kDvDp, pressure, -kVolBase);                       // b of the quadratic :601
  const double disc = fma(bq, bq, 4.0 * kDvDp * vu * pressure);
  const double v = 0.5 * (fm::sqrt_pos(disc) - bq);                        // :604
  const bool slack = vu <= kVolBase;
  o.volume = slack ? vu : v;
  o.superpressure = slack ? 0.0 : fma(pressure * vu, fm::rcp(v), -pressure);   // :605-607
  int status = kOk;
  if (o.superpressure > kMaxSuperpressure) status = kBurst;                // :479-480
  if (o.superpressure <= 0.0) status = kZeroPressure;                      // :481-482
  o.status = status;
  o.cv = fm::cbrtf_pos(float(o.volume));
  float acs_power = 0.0f, flow = 0.0f;
  if (action == kUp) {                                                     // vent :487-499
    const float valve_area = float(kPi * kValveDiameter * kValveDiameter / 4.0);
    const float gas_density = float((superpressure + pressure) * (kMAir / kR)) * fm::rcpf(float(t_internal));
    flow = float(-kValveCd) * valve_area * fm::sqrtf_pos(fmaxf(2.0f * float(superpressure) * gas_density, 0.0f));
  } else if (action == kDown) {                                            // compressor :500-517
    const float prm1 = float(fmax(superpressure, 0.0) * inv_p);            // pressure_ratio - 1 :247-250
    acs_power = acs_power_f(prm1);
    flow = acs_fan_efficiency_f(prm1, acs_power) * acs_power * (1.0f / 3600.0f);   // acs.py:67-68
  }
  o.mols_air = fmax(fma(double(flow) * (1.0 / kMAir), dt, mols_air), 0.0); // :519
  o.acs_power = acs_power;
  o.flow = flow;
  return o;
}

// ---- role S: sun-dependent part of dT/dt and the power system (balloon.py:451-467,524-542) -----------
struct SunOut {
  float d_t_solar, solar_w, load_w, acs_power;
  double charge;
  int out_of_power;
};

// sun_angles_from_cos_zenith<float> (ble_physics.cuh) on the fast primitives.
BLE_HD SunAngles<float> sun_angles_fast(float cz) {
  SunAngles<float> o;
  const float sz = fm::sqrtf_pos(fmaxf(1.0f - cz * cz, 0.0f));
  const float el_unc = 90.0f - r_deg(acosf(cz));                           // solar.py:141
  float refraction;
  if (el_unc > 85.0f) {
    refraction = 0.0f;
  } else if (el_unc > 5.0f) {
    const float it = sz * fm::rcpf(cz);                                    // 1 / tan(el_unc)
    const float it2 = it * it;
    refraction = it * (58.1f + it2 * (-0.07f + it2 * 0.000086f));          // :145-147
  } else if (el_unc > -0.575f) {
    refraction = 1735.0f + el_unc * (-518.2f + el_unc * (103.4f + el_unc * (-12.79f + el_unc * 0.711f)));
  } else {
    refraction = -20.772f * sz * fm::rcpf(cz);
  }
  const float d_deg = refraction * (1.0f / 3600.0f);
  const float d = r_rad(d_deg), d2 = d * d;
  const float sd = d * (1.0f - d2 * (1.0f / 6.0f));
  const float cd = 1.0f - d2 * (0.5f - d2 * (1.0f / 24.0f));
  o.el = el_unc + d_deg;                                                   // :157
  o.sin_el = cz * cd + sz * sd;
  o.cos_el = sz * cd - cz * sd;
  return o;
}

BLE_HD SunOut sun_power(const SunAngles<float>& sun, float flux, float cv, double pressure, double superpressure,
                        double charge, int action) {
  SunOut o;
  const float p_r = float(pressure);
  float att = 0.0f;
  if (!(sun.el < float(kMinSolarElDeg))) {                                 // solar.py:177-209
    const float s = 614.0f * sun.sin_el;
    const float airmass = 0.34764f * (p_r * (1.0f / 101325.0f)) * (fm::sqrtf_pos(1229.0f + s * s) - s);
    att = 0.5f * (fm::expf_fast(-0.65f * airmass) + fm::expf_fast(-0.95f * airmass));
  }
  const float radius = cv * 0.62035049089940f;
  const float area = float(4 * kPi) * radius * radius;
  const float tot_solar = 0.01435f * (1.0f + (1.0f - 0.01435f - 0.0291f) / (1.0f - 0.0291f));
  o.d_t_solar = (flux * att) * 0.25f * area * tot_solar * float(1.0 / (1500.0 * kEnvelopeMass));
  float acs_power = 0.0f;
  if (action == kDown) acs_power = acs_power_f(float(fmax(superpressure, 0.0) * fm::rcp(pressure)));
  const bool is_day = sun.el > float(kMinSolarElDeg);                      // balloon.py:524-530
  const float solar_w = is_day ? solar_power_sc<float>(sun, att) : 0.0f;
  const float load_w = (is_day ? float(kDayLoadW) : float(kNightLoadW)) + acs_power;
  double c = fma(double(solar_w - load_w), double(kStrideS) / 3600.0, charge);
  c = fmin(fmax(c, 0.0), kBatteryCapacityWh);                              // :537-539
  o.charge = c;
  o.out_of_power = (c <= 0.0) ? 1 : 0;                                     // :541-542
  o.solar_w = solar_w;
  o.load_w = load_w;
  o.acs_power = acs_power;
  return o;
}

// cos(zenith) of the quadratic sun track at sub-step k (SunTrack<float>::at, ble_physics.cuh).
BLE_HD void sun_track_at(const SunTrack<float>& tr, int k, float* cz, float* flux) {
  const float t = float(k) * float(1.0 / kSubSteps);
  float c = tr.c0 + t * ((-3.0f * tr.c0 + 4.0f * tr.c1 - tr.c2) + t * (2.0f * tr.c0 - 4.0f * tr.c1 + 2.0f * tr.c2));
  *cz = fminf(fmaxf(c, -1.0f), 1.0f);
  *flux = tr.f0 + t * (tr.f2 - tr.f0);
}

// ---- the sun along one agent step (production build) ---------------------------------------------------------
// Same quantities as solar_time<float> / latlng_from_offset<float> / solar_cos_zenith<float> (ble_physics.cuh,
// following env/balloon/solar.py:43-174 and utils/spherical_geometry.py:44-76), reshaped so that one evaluation
// needs 6 sincos instead of 14 libm trig calls, 2 atan2, an asin and 4 fp64 fmod:
//   * double / triple angles by identities; cos(declination) = sqrt(1 - sin^2);
//   * heading = atan2(x, y) only ever enters through its sine and cosine, which are x / r and y / r;
//   * the hour angle only enters through its cosine, so the fmod(., 1440) and the +-pi wrap (solar.py:113-120) drop
//     out: cos(ha) = -cos(2 pi fod + eq_time pi / 720 + lng), and with lng = lng0 + atan2(Y, X) the atan2 is
//     replaced by the angle-addition formula on (X, Y) / hypot(X, Y).
struct SolarTimeFast {
  double phase;            // 2 pi fraction_of_day + equation_of_time [min] * pi / 720, in radians (fp64, unreduced)
  float sin_decl, cos_decl, flux;
};

BLE_HD SolarTimeFast solar_time_fast(int64_t ts) {
  SolarTimeFast o;
  int64_t days = ts / 86400;
  int64_t sod = ts - days * 86400;
  if (sod < 0) { sod += 86400; days -= 1; }
  const double fod = double(sod) * (1.0 / 86400.0);                        // solar.py:66-68
  const double jc = ((2440587.5 + double(days)) + fod - 2451545.0) * (1.0 / 36525.0);   // :71-79
  const double l0_deg = 280.46646 + jc * (36000.76983 + jc * 0.0003032);   // :82-83
  const double m0_deg = 357.52911 + jc * (35999.05029 - 0.0001537 * jc);   // :88-89
  const double om_deg = 125.04 - 1934.136 * jc;
  const float l0 = float(fm::mod_pos(l0_deg, 360.0, 1.0 / 360.0) * (kPi / 180.0));
  const float m0 = float(fm::mod_pos(m0_deg, 360.0, 1.0 / 360.0) * (kPi / 180.0));
  const float om = float(fm::mod_pos(om_deg, 360.0, 1.0 / 360.0) * (kPi / 180.0));
  const float jcr = float(jc);
  float sl, cl, sm, cm, so, co;
  fm::sincosf_fast(l0, &sl, &cl);
  fm::sincosf_fast(m0, &sm, &cm);
  fm::sincosf_fast(om, &so, &co);
  const float sin2l0 = 2.0f * sl * cl, cos2l0 = 1.0f - 2.0f * sl * sl, sin4l0 = 2.0f * sin2l0 * cos2l0;
  const float sin2m0 = 2.0f * sm * cm, sin3m0 = sm * (3.0f - 4.0f * sm * sm);
  const float mean_obl = r_rad(23.0f + (26.0f + ((21.448f - jcr * (46.815f + jcr * (0.00059f - jcr * 0.001813f)))) *
                               (1.0f / 60.0f)) * (1.0f / 60.0f));          // :94-97
  const float obl = mean_obl + r_rad(0.00256f * co);                       // :99-100
  float sh, ch;
  fm::sincosf_fast(0.5f * obl, &sh, &ch);
  const float ty = sh * fm::rcpf(ch);
  const float var_y = ty * ty;                                             // :102
  const float sin_obl = 2.0f * sh * ch;
  const float ecc = 0.016708634f - jcr * (0.000042037f + 0.0000001267f * jcr);   // :104-105
  const float eq_time = 4.0f * (var_y * sin2l0 - 2.0f * ecc * sm + 4.0f * ecc * var_y * sm * cos2l0 -
                                0.5f * var_y * var_y * sin4l0 - 1.25f * ecc * ecc * sin2m0);   // :107-111
  const float eq_center = r_rad(sm * (1.914602f - jcr * (0.004817f + 0.000014f * jcr)) +
                                sin2m0 * (0.019993f - 0.000101f * jcr) + sin3m0 * 0.000289f);   // :122-127
  const float app_long = l0 + eq_center - r_rad(0.00569f - 0.00478f * so);                    // :129-131
  float sa, ca;
  fm::sincosf_fast(app_long, &sa, &ca);
  o.sin_decl = sin_obl * sa;                                               // :132-133
  o.cos_decl = fm::sqrtf_pos(fmaxf(1.0f - o.sin_decl * o.sin_decl, 0.0f));
  const float e1 = (1.0f + ecc) * fm::rcpf(1.0f - ecc);
  o.flux = 1366.0f * (1.0f + 0.5f * (e1 * e1 - 1.0f) * cm);                // :170-172
  // degrees(eq_time) is what the reference adds, as minutes, to the hour angle (:115): minutes * pi / 720 radians
  o.phase = 2.0 * kPi * fod + double(r_deg(eq_time)) * (kPi / 720.0);
  return o;
}

// Per-step constants of a balloon's station: sin / cos of the centre latitude.
struct StationTrig { float sfl, cfl; double lng0; };
BLE_HD StationTrig station_trig(float lat0, float lng0) {
  StationTrig t;
  fm::sincosf_fast(lat0, &t.sfl, &t.cfl);
  t.lng0 = double(lng0);
  return t;
}

// cos(zenith) at offset (x, y) [m] from the station (utils/spherical_geometry.py:61-76 + solar.py:113-138).
BLE_HD float cos_zenith_fast(const StationTrig& st, double phase, float sin_decl, float cos_decl, float x, float y) {
  const float r2 = x * x + y * y;
  float ch = 1.0f, sh = 0.0f, angle = 0.0f;                                // heading = atan2(0, 0) = 0 at the station
  if (r2 > 0.0f) {
    const float inv_r = fm::rsqrtf_pos(r2);
    ch = y * inv_r; sh = x * inv_r;                                        // cos / sin of atan2(x, y) :61
    angle = (r2 * inv_r) * float(1.0 / kEarthRadiusM);                     // :62
  }
  float sa, ca;
  fm::sincosf_fast(angle, &sa, &ca);
  float sin_lat = ca * st.sfl + sa * st.cfl * ch;                          // :69-70
  sin_lat = fminf(fmaxf(sin_lat, -1.0f), 1.0f);
  const float X = ca - st.sfl * sin_lat, Y = sa * st.cfl * sh;             // d_lng = atan2(Y, X)
  const float h2 = X * X + Y * Y;
  float cosd = 1.0f, sind = 0.0f;
  if (h2 > 1e-30f) { const float inv_h = fm::rsqrtf_pos(h2); cosd = X * inv_h; sind = Y * inv_h; }
  const double a = phase + st.lng0;
  const float ar = float(fma(-2.0 * kPi, rint(a * (0.5 / kPi)), a));       // reduced to [-pi, pi] in fp64
  float sA, cA;
  fm::sincosf_fast(ar, &sA, &cA);
  const float cos_ha = -(cA * cosd - sA * sind);                           // cos(hour angle), solar.py:113-120
  const float cos_lat = fm::sqrtf_pos(fmaxf(1.0f - sin_lat * sin_lat, 0.0f));
  const float cz = sin_lat * sin_decl + cos_lat * cos_decl * cos_ha;       // :135-138
  return fminf(fmaxf(cz, -1.0f), 1.0f);
}

// The quadratic sun track of one agent step: cos(zenith) at t0, t0 + 90 s, t0 + 180 s (SunTrack<float>, ble_physics.cuh).
// The time-only terms are evaluated at the two ends; at the mid point they are the mean of the ends (declination and
// equation of time move by < 1.5e-5 rad in 180 s: the interpolation error is below 1e-10).
BLE_HD SunTrack<float> sun_track_fast(const SolarTimeFast& a, const SolarTimeFast& b, float lat0, float lng0, double x,
                                      double y, double u, double v) {
  SunTrack<float> tr;
  const StationTrig st = station_trig(lat0, lng0);
  // fraction_of_day wraps at midnight: keep the phase continuous across the step
  double pb = b.phase;
  if (pb < a.phase - kPi) pb += 2.0 * kPi;
  tr.c0 = cos_zenith_fast(st, a.phase, a.sin_decl, a.cos_decl, float(x), float(y));
  tr.c1 = cos_zenith_fast(st, 0.5 * (a.phase + pb), 0.5f * (a.sin_decl + b.sin_decl), 0.5f * (a.cos_decl + b.cos_decl),
                          float(x + u * 90.0), float(y + v * 90.0));
  tr.c2 = cos_zenith_fast(st, pb, b.sin_decl, b.cos_decl, float(x + u * 180.0), float(y + v * 180.0));
  tr.f0 = a.flux; tr.f2 = b.flux;
  return tr;
}

// ---- the four roles in sequence: one agent step for one balloon (host replay + reference for the kernel) ----
// Same contract as agent_step<float> (ble_physics.cuh).
BLE_HD float agent_step_roles(BalloonState<float>& s, Atmosphere& atm, SafetyState& ss, int action,
                              double u, double v, int* effective_action) {
  ss.last_command = action;                                                // balloon.py:286
  int eff = action;
  if (ss.power_safety_enabled) {
    eff = power_safety<double>(eff, s.date_time, s.charge, &ss.sunrise_h, &ss.sunset, &ss.power_paused);
  }
  eff = envelope_safety<double>(eff, s.superpressure, &ss.envelope_state);
  PressureRole pr;
  pr.init(atm, double(s.mols_gas), s.pressure);
  auto load_atm = [&atm]() { return atm; };
  double altitude;
  if (pr.layer >= 0) {                                                     // altitude from the same X the loop starts with
    const double x0 = pr.x_from_scratch(s.pressure);
    pr.seed_x(s.pressure, x0);
    altitude = (x0 - 1.0) * pr.tl + atm_h(pr.layer);
  } else {
    double t_unused;
    atm.at_pressure(s.pressure, &altitude, &t_unused);
  }
  eff = altitude_safety<double>(eff, altitude, &ss.altitude_state);
  *effective_action = eff;
  const SunTrack<float> sun = sun_track_fast(solar_time_fast(s.date_time), solar_time_fast(s.date_time + 180), s.lat0,
                                             s.lng0, s.x, s.y, u, v);
  const float earth_per_area = earth_heat_per_area<float>(s.ir);
  float cv = fm::cbrtf_pos(float(s.volume));
  int k = 0;
  for (; k < kSubSteps;) {
    float cz, flux;
    sun_track_at(sun, k, &cz, &flux);
    const SunAngles<float> ang = sun_angles_fast(cz);
    double np, nt;
    pr.step(s.pressure, s.t_ambient, s.volume, s.mols_air, cv, load_atm, &np, &nt);
    const float dtb = thermal_body(cv, s.t_internal, s.t_ambient, s.pressure, earth_per_area);
    const EnvelopeOut eo = envelope_acs(double(s.mols_gas), s.mols_air, s.t_internal, s.pressure, s.superpressure, eff);
    const SunOut so = sun_power(ang, flux, cv, s.pressure, s.superpressure, s.charge, eff);
    int status = s.status;
    if (eo.status != kOk) status = eo.status;
    if (so.out_of_power) status = kOutOfPower;                             // later assignment wins (:541-542)
    s.x += u * double(kStrideS);
    s.y += v * double(kStrideS);
    s.pressure = np; s.t_ambient = nt;
    s.t_internal = s.t_internal + double(dtb + so.d_t_solar) * double(kStrideS);
    s.volume = eo.volume; s.superpressure = eo.superpressure; s.mols_air = eo.mols_air; cv = eo.cv;
    s.charge = so.charge;
    s.acs_power = eo.acs_power; s.acs_flow = eo.flow; s.solar_w = so.solar_w; s.load_w = so.load_w;
    s.status = status;
    s.date_time += kStrideS;
    s.time_elapsed += kStrideS;
    ++k;
    if (s.status != kOk) break;
  }
  atm.ok = atm.ok && pr.atm_ok;
  float el = 0.0f;
  if (action == kDown) {                                                   // excess_energy's sun (balloon.py:231-238)
    float cz, flux;
    sun_track_at(sun, k, &cz, &flux);
    el = sun_angles_fast(cz).el;
  }
  return perciatelli_reward<float>(s, action, el);
}

}  // namespace roles
}  // namespace ble
