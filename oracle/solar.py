"""Batched NOAA solar model (ORACLE / test infrastructure).

Follows env/balloon/solar.py:38-536.  Times are int64 UNIX seconds (UTC); every
time the reference manipulates on this path is a whole number of seconds.
"""
import numpy as np

from oracle import constants as C


def civil_from_unix(ts):
  """int64 unix seconds -> (year, month, day, seconds_of_day), proleptic Gregorian, UTC."""
  ts = np.asarray(ts, np.int64)
  days = np.floor_divide(ts, 86400)
  sod = ts - days * 86400
  z = days + 719468
  era = np.floor_divide(z, 146097)
  doe = z - era * 146097
  yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
  y = yoe + era * 400
  doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
  mp = (5 * doy + 2) // 153
  d = doy - (153 * mp + 2) // 5 + 1
  m = np.where(mp < 10, mp + 3, mp - 9)
  y = np.where(m <= 2, y + 1, y)
  return y, m, d, sod


def solar_calculator(lat_rad, lng_rad, ts):
  """-> (el_deg, az_deg, flux) following solar.py:43-174."""
  lat = np.asarray(lat_rad, np.float64)
  lng_deg = np.degrees(np.asarray(lng_rad, np.float64))
  ts = np.asarray(ts, np.int64)
  if np.any(np.abs(lat) > np.pi / 2) or np.any(np.abs(lng_deg) > 180.0):
    raise ValueError('solar_calculator: latlng is invalid')               # :60-61
  year, month, day, sod = civil_from_unix(ts)
  year = year.astype(np.float64); month = month.astype(np.float64); day = day.astype(np.float64)
  fraction_of_day = sod / C.NUM_SECONDS_PER_DAY                           # :66-68
  jdn = (367.0 * year - np.floor(7.0 * (year + np.floor((month + 9.0) / 12.0)) / 4.0)
         - np.floor(3.0 * (np.floor((year + (month - 9.0) / 7.0) / 100.0) + 1.0) / 4.0)
         + np.floor(275.0 * month / 9.0) + day + 1721028.5)               # :71-75
  julian_time = jdn + fraction_of_day
  jc = (julian_time - 2451545.0) / 36525.0                                # :78-79

  l0 = np.radians(280.46646 + jc * (36000.76983 + jc * 0.0003032))        # :82-83
  sin2l0, cos2l0, sin4l0 = np.sin(2.0 * l0), np.cos(2.0 * l0), np.sin(4.0 * l0)
  m0 = np.radians(357.52911 + jc * (35999.05029 - 0.0001537 * jc))        # :88-89
  sinm0, sin2m0, sin3m0 = np.sin(m0), np.sin(2.0 * m0), np.sin(3.0 * m0)
  mean_obliquity = np.radians(
      23.0 + (26.0 + ((21.448 - jc * (46.815 + jc * (0.00059 - jc * 0.001813)))) / 60.0) / 60.0)
  obliquity = mean_obliquity + np.radians(
      0.00256 * np.cos(np.radians(125.04 - 1934.136 * jc)))               # :99-100
  var_y = np.tan(obliquity / 2.0) ** 2
  ecc = 0.016708634 - jc * (0.000042037 + 0.0000001267 * jc)              # :104-105
  eq_time = (4.0 * (var_y * sin2l0 - 2.0 * ecc * sinm0 + 4.0 * ecc * var_y * sinm0 * cos2l0
                    - 0.5 * var_y * var_y * sin4l0 - 1.25 * ecc * ecc * sin2m0))  # :107-111
  hour_angle = np.radians(
      np.fmod(1440.0 * fraction_of_day + np.degrees(eq_time) + 4.0 * lng_deg, 1440.0)) / 4.0
  hour_angle = np.where(hour_angle < 0, hour_angle + np.pi, hour_angle - np.pi)  # :117-120
  eq_center = np.radians(sinm0 * (1.914602 - jc * (0.004817 + 0.000014 * jc))
                         + sin2m0 * (0.019993 - 0.000101 * jc) + sin3m0 * 0.000289)
  true_long = l0 + eq_center
  apparent_long = true_long - np.radians(
      0.00569 - 0.00478 * np.sin(np.radians(125.04 - 1934.136 * jc)))     # :129-131
  declination = np.arcsin(np.sin(obliquity) * np.sin(apparent_long))
  zenith = np.arccos(np.sin(lat) * np.sin(declination)
                     + np.cos(lat) * np.cos(declination) * np.cos(hour_angle))  # :135-138
  el_unc = 90.0 - np.degrees(zenith)

  with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
    tan_seu = np.tan(np.radians(el_unc))
    r_hi = 58.1 / tan_seu - 0.07 / (tan_seu ** 3) + 0.000086 / (tan_seu ** 5)
    r_mid = (1735.0 + el_unc * (-518.2 + el_unc * (103.4 + el_unc * (-12.79 + el_unc * 0.711))))
    r_lo = -20.772 / tan_seu
  refraction = np.where(el_unc > 85.0, 0.0,
                        np.where(el_unc > 5.0, r_hi,
                                 np.where(el_unc > -0.575, r_mid, r_lo)))  # :143-155
  el_deg = el_unc + refraction / 3600.0

  with np.errstate(divide='ignore', invalid='ignore'):
    cos_az = ((np.sin(lat) * np.cos(zenith) - np.sin(declination)) /
              (np.cos(lat) * np.sin(zenith)))                             # :160-162
  az_unwrapped = np.arccos(np.clip(cos_az, -1.0, 1.0))
  az_deg = np.where(hour_angle > 0, np.degrees(az_unwrapped) + 180.0,
                    180.0 - np.degrees(az_unwrapped))                     # :164-167
  flux = 1366.0 * (1 + 0.5 * (((1 + ecc) / (1 - ecc)) ** 2 - 1) * np.cos(m0))  # :170-172
  return el_deg, az_deg, flux


def solar_atmospheric_attenuation(el_deg, pressure_pa):
  """solar.py:177-209."""
  el_deg = np.asarray(el_deg, np.float64); pressure_pa = np.asarray(pressure_pa, np.float64)
  if np.any(el_deg > 90.0) or np.any(el_deg < -90.0):
    raise ValueError('solar_atmospheric_attenuation: elevation out of range')
  if np.any(pressure_pa > 101325.0) or np.any(pressure_pa < 0.0):
    raise ValueError('solar_atmospheric_attenuation: pressure out of range')
  tmp = 614.0 * np.sin(np.radians(el_deg))
  airmass = 0.34764 * (pressure_pa / 101325.0) * (np.sqrt(1229.0 + tmp * tmp) - tmp)
  att = 0.5 * (np.exp(-0.65 * airmass) + np.exp(-0.95 * airmass))
  return np.where(el_deg < C.MIN_SOLAR_EL_DEG, 0.0, att)


def balloon_shadow(el_deg, panel_height_below_balloon_m):
  """solar.py:212-236."""
  h = panel_height_below_balloon_m
  shadow_el = np.degrees(np.arctan2(np.sqrt(h * (10.41603 + h)), 8.69275))
  return np.where(np.asarray(el_deg) >= shadow_el, 0.4392, 1.0)


def solar_power(el_deg, pressure_pa):
  """Watts; solar.py:515-536."""
  el_deg = np.asarray(el_deg, np.float64)
  att = solar_atmospheric_attenuation(el_deg, pressure_pa)
  return 210.0 * att * (
      4 * np.cos(np.radians(el_deg - 35)) * balloon_shadow(el_deg, 3.3) +
      2 * np.cos(np.radians(el_deg - 65)) * balloon_shadow(el_deg, 2.7))


# ---- sunrise / sunset search (solar.py:239-483); reset-time + features ----------------

_MIN, _MAX, _TARGET = 0, 1, 2


def _objective(lat, lng, ts, kind, target):
  el, _, _ = solar_calculator(lat, lng, ts)
  if kind == _MIN:
    return el
  if kind == _MAX:
    return -el
  return np.abs(el - target)


def _find_solar_elevation(lat, lng, min_ts, max_ts, kind, target=0.0,
                          delta=C.SOLAR_SEARCH_DELTA_S):
  """Vectorised _find_solar_elevation_binary_search (solar.py:296-375) -> time (int64)."""
  lat = np.asarray(lat, np.float64); lng = np.asarray(lng, np.float64)
  min_ts = np.asarray(min_ts, np.int64); max_ts = np.asarray(max_ts, np.int64)
  if np.any(max_ts < min_ts):
    raise ValueError('Time interval must have positive extent.')
  # int((max - min) / delta): timedelta / timedelta is true division, int() truncates.
  max_steps = ((max_ts - min_ts) / float(delta)).astype(np.int64)
  assert np.all(max_steps > 0)
  low = np.zeros_like(max_steps)
  high = max_steps.copy()
  obj = lambda idx: _objective(lat, lng, min_ts + delta * idx, kind, target)
  while True:
    active = high > low + 1
    if not active.any():
      break
    midpoint = low + (high - low) / 2
    lt = obj(low) < obj(high)
    new_high = np.where(lt, np.ceil(midpoint).astype(np.int64), high)
    new_low = np.where(lt, low, np.floor(midpoint).astype(np.int64))
    high = np.where(active, new_high, high)
    low = np.where(active, new_low, low)
  min_index = np.where(obj(low) < obj(high), low, high)
  return min_ts + delta * min_index


def is_solar_afternoon(lat, lng, ts):
  """solar.py:239-255."""
  now_el, _, _ = solar_calculator(lat, lng, ts)
  then_el, _, _ = solar_calculator(lat, lng, np.asarray(ts, np.int64) + 1)
  return then_el < now_el


def get_next_sunrise_sunset(lat, lng, ts, delta=C.SOLAR_SEARCH_DELTA_S):
  """-> (sunrise_ts, sunset_ts) int64; solar.py:378-483."""
  lat = np.atleast_1d(np.asarray(lat, np.float64)); lng = np.atleast_1d(np.asarray(lng, np.float64))
  ts = np.atleast_1d(np.asarray(ts, np.int64))
  assert np.all(np.abs(np.degrees(lat)) < 60.0), 'High latitudes not supported.'
  h12, h24 = 12 * 3600, 24 * 3600
  aft = is_solar_afternoon(lat, lng, ts)
  # get_next_solar_noon :405-429 / get_next_solar_midnight :378-402
  noon_lo = np.where(aft, ts + h12, ts)
  next_noon = _find_solar_elevation(lat, lng, noon_lo, noon_lo + h12, _MAX, delta=delta)
  mid_lo = np.where(aft, ts, ts + h12)
  next_midnight = _find_solar_elevation(lat, lng, mid_lo, mid_lo + h12, _MIN, delta=delta)
  # :458-475
  sr_lo = np.where(aft, next_midnight, next_midnight - h24)
  sunrise = _find_solar_elevation(lat, lng, sr_lo, next_noon, _TARGET, C.MIN_SOLAR_EL_DEG, delta)
  ss_lo = np.where(aft, next_noon - h24, next_noon)
  sunset = _find_solar_elevation(lat, lng, ss_lo, next_midnight, _TARGET, C.MIN_SOLAR_EL_DEG, delta)
  sunrise = np.where(sunrise < ts, sunrise + h24, sunrise)                # :478-481
  sunset = np.where(sunset < ts, sunset + h24, sunset)
  return sunrise, sunset
