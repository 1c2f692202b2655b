"""Perciatelli observation surface (ORACLE / test infrastructure).

Follows env/features.py:56-103,269-581 (PerciatelliFeatureConstructor), env/wind_gp.py:33-241
(WindGP on scikit-learn's GaussianProcessRegressor), env/balloon/pressure_range_builder.py:43-275,
env/balloon/power_table.py:21-38 and utils/transforms.py:45-94, batched over N balloons.

scikit-learn's GPR (requirements.txt:59) is restated as the closed form it implements:
  K = 12.96 exp(-||(a - b) / l||_2) + 0.05 I,  L = chol(K),  alpha = K^-1 y,
  mean = k*^T alpha,  var = max(12.96 - ||L^-1 k*||^2, 0)
(kernel 12.96 * Matern(nu=0.5, length_scale=(357 km, 357 km, 326 Pa, 34560 s)), alpha=0.05,
optimizer=None, normalize_y=False; env/wind_gp.py:33-38,66-79).
"""
import bisect

import numpy as np
import scipy.linalg

from oracle import atmosphere as atmosphere_lib  # noqa: F401
from oracle import balloon as balloon_lib
from oracle import constants as C
from oracle import solar
from oracle import stable_init
from oracle import wind

NUM_LEVELS = 181
NUM_FEATURES = 3 * (NUM_LEVELS * 2 - 1) + 16          # 1099 (features.py:291)
PRESSURE_LEVELS = np.linspace(C.PERCIATELLI_PRESSURE_RANGE_MIN, C.PERCIATELLI_PRESSURE_RANGE_MAX, NUM_LEVELS)
TOLERANCE_M = 1e-5                                    # features.py:53

GP_LENGTH_SCALE = np.array([357000.0, 357000.0, 326.0, 34560.0])   # wind_gp.py:33-35
GP_SIGMA2 = 3.6 ** 2
GP_NOISE = 0.05
GP_HORIZON_S = 6 * 3600


# ----------------------------------------------------------------------------- power table

_PR_INTERVALS = [1.08, 1.11, 1.14, 1.17, 1.2, 1.23, 1.26]
_SOC_MAPPINGS = [([0.3, 0.4, 0.5], [0, 150, 175, 200]), ([0.3, 0.4, 0.7], [0, 200, 200, 225]),
                 ([0.3, 0.4, 0.6], [0, 225, 225, 250]), ([0.3, 0.4, 0.5], [0, 200, 225, 250]),
                 ([0.3, 0.4, 0.5], [0, 225, 250, 275]), ([0.4, 0.5], [0, 275, 300]),
                 ([0.5, 0.6], [0, 300, 325]), ([0.5, 0.6], [0, 325, 350])]


def power_table_lookup(pressure_ratio: float, soc: float) -> float:
  """env/balloon/power_table.py:21-38."""
  assert 0.99 <= pressure_ratio <= 5
  pr_id = bisect.bisect(_PR_INTERVALS, pressure_ratio)
  soc_id = bisect.bisect(_SOC_MAPPINGS[pr_id][0], soc)
  return float(_SOC_MAPPINGS[pr_id][1][soc_id])


# ----------------------------------------------------------------------------- sunrise time

def compute_sunrise_time(lat, lng, ts):
  """features.py:72-103 -> normalised solar cycle time in [0, 2 pi]."""
  sunrise, sunset = solar.get_next_sunrise_sunset(lat, lng, ts)
  ts = np.asarray(ts, np.int64)
  day = 86400
  assert np.all((sunrise - day <= ts) & (ts <= sunrise)) and np.all((sunset - day <= ts) & (ts <= sunset))
  is_day = sunset < sunrise
  prev_sunrise = sunrise - day
  prev_sunset = sunset - day
  day_val = np.pi * (ts - prev_sunrise) / np.where(is_day, sunset - prev_sunrise, 1)
  night_val = np.pi + np.pi * (ts - prev_sunset) / np.where(is_day, 1, sunrise - prev_sunset)
  return np.where(is_day, day_val, night_val)


# ----------------------------------------------------------------------------- pressure range

def _x_crossing(x1, y1, x2, y2, y_star):
  """pressure_range_builder.py:43-71."""
  if y_star < min(y1, y2) or y_star > max(y1, y2):
    raise ValueError('y_star must be in [y1, y2].')
  if x1 >= x2:
    raise ValueError('x2 must be greater than x1.')
  if y1 == y2:
    raise ValueError('y1 may not be equal to y2.')
  alpha = abs((y_star - y1) / (y2 - y1))
  return alpha * (x2 - x1) + x1


def _safe_pressure(p1, sp1, p2, sp2, min_sp, max_sp):
  """pressure_range_builder.py:74-102."""
  if p1 >= p2:
    raise ValueError('pressure2 must be greater than pressure1.')
  if sp1 == sp2:
    raise ValueError('sp1 and sp2 may not be equal.')
  if (sp1 < min_sp and sp2 >= min_sp) or (sp1 >= min_sp and sp2 < min_sp):
    return _x_crossing(p1, sp1, p2, sp2, min_sp)
  if (sp1 > max_sp and sp2 <= max_sp) or (sp1 <= max_sp and sp2 > max_sp):
    return _x_crossing(p1, sp1, p2, sp2, max_sp)
  raise ValueError('Unable to find valid superpressure crossing for input params.')


def _search(levels, sp_levels, significant, sp_significant, min_sp, max_sp, direction):
  """pressure_range_builder.py:105-182 for one balloon, with superpressures precomputed."""
  if min_sp <= sp_significant <= max_sp:
    return significant
  last = (significant, sp_significant)
  order = range(len(levels) - 1, -1, -1) if direction == 'min' else range(len(levels))
  for j in order:
    pressure = levels[j]
    if (direction == 'min' and pressure > significant) or (direction == 'max' and pressure < significant):
      continue
    sp = sp_levels[j]
    if sp > max_sp or sp < min_sp:
      last = (pressure, sp)
      continue
    if direction == 'min':
      return _safe_pressure(pressure, sp, last[0], last[1], min_sp, max_sp)
    return _safe_pressure(last[0], last[1], pressure, sp, min_sp, max_sp)
  raise ValueError('Unable to find safe pressure for balloon.')


def get_pressure_range(b: balloon_lib.BalloonBatch, atm):
  """pressure_range_builder.py:203-275 -> (min_pressure[N], max_pressure[N])."""
  n = b.n
  min_sp = C.ENV_BUFFER
  max_sp = C.ENVELOPE_MAX_SUPERPRESSURE - C.ENV_BUFFER
  search_max, _ = atm.at_height(np.full(n, C.ALT_MIN_ALTITUDE_M))          # :230
  levels = np.linspace(1000.0, search_max, 20, axis=1)                     # [N, 20] :231
  lat, lng = b.latlng()
  p_over_t = np.empty((n, 20)); sp_levels = np.empty((n, 20))
  for j in range(20):
    _, t = atm.at_pressure(levels[:, j])
    p_over_t[:, j] = levels[:, j] / t                                      # :241
    sp_levels[:, j] = stable_init.calculate_stable_params_for_pressure(
        levels[:, j], b.mols_lift_gas, lat, lng, b.date_time, b.upwelling_infrared, atm)['superpressure']
  assert np.all(np.diff(p_over_t, axis=1) > 0)                             # :242
  total_empty_mass = C.PAYLOAD_MASS + C.ENVELOPE_MASS + b.mols_lift_gas * C.HE_MOLAR_MASS
  target = total_empty_mass * C.UNIVERSAL_GAS_CONSTANT / (C.DRY_AIR_MOLAR_MASS * C.ENVELOPE_VOLUME_BASE)
  # interp1d(p_over_t_column, pressure_levels, linear, extrapolate)(target)  :243-245
  idx = np.clip(np.array([np.searchsorted(p_over_t[e], target[e], side='left') for e in range(n)]) - 1, 0, 18)
  r = np.arange(n)
  x0, x1 = p_over_t[r, idx], p_over_t[r, idx + 1]
  y0, y1 = levels[r, idx], levels[r, idx + 1]
  min_pressure = (y1 - y0) / (x1 - x0) * (target - x0) + y0
  sp_min_sig = stable_init.calculate_stable_params_for_pressure(
      min_pressure, b.mols_lift_gas, lat, lng, b.date_time, b.upwelling_infrared, atm)['superpressure']
  sp_max_sig = sp_levels[:, -1]                                             # max_pressure = search_range_max :247
  out_min = np.empty(n); out_max = np.empty(n)
  for e in range(n):
    out_min[e] = _search(levels[e], sp_levels[e], min_pressure[e], sp_min_sig[e], min_sp, max_sp, 'max')
    out_max[e] = _search(levels[e], sp_levels[e], levels[e, -1], sp_max_sig[e], min_sp, max_sp, 'min')
  return out_min, out_max


# ----------------------------------------------------------------------------- wind GP

class WindGP:
  """env/wind_gp.py:41-241 for N balloons.  History is kept forever (as the reference does) and
  filtered to the last 6 h at query time."""

  def __init__(self, n):
    self.n = n
    self.locations = [[] for _ in range(n)]     # per balloon: list of [x, y, p, t]
    self.errors = [[] for _ in range(n)]        # per balloon: list of [du, dv]

  def observe(self, x, y, pressure, elapsed_s, wind_u, wind_v, forecast_u, forecast_v, mask=None):
    """wind_gp.py:98-119."""
    for e in range(self.n):
      if mask is not None and not mask[e]:
        continue
      self.locations[e].append([x[e], y[e], pressure[e], float(elapsed_s[e])])
      self.errors[e].append([wind_u[e] - forecast_u[e], wind_v[e] - forecast_v[e]])

  def window(self, e, now_s):
    loc = np.asarray(self.locations[e], np.float64).reshape(-1, 4)
    err = np.asarray(self.errors[e], np.float64).reshape(-1, 2)
    fresh = np.abs(loc[:, 3] - now_s) < GP_HORIZON_S                         # :172-178
    return loc[fresh], err[fresh]

  def query_column(self, e, x, y, elapsed_s):
    """-> (error means [181, 2], normalised variance [181]) at the 181 pressure levels (:143-207)."""
    loc, err = self.window(e, elapsed_s)
    if len(self.locations[e]) == 0:
      return np.zeros((NUM_LEVELS, 2)), np.zeros(NUM_LEVELS)
    q = np.stack([np.full(NUM_LEVELS, x), np.full(NUM_LEVELS, y), PRESSURE_LEVELS,
                  np.full(NUM_LEVELS, float(elapsed_s))], axis=1)
    a = loc / GP_LENGTH_SCALE
    k = GP_SIGMA2 * np.exp(-np.sqrt(((a[:, None, :] - a[None, :, :]) ** 2).sum(-1)))
    k[np.diag_indices_from(k)] += GP_NOISE
    chol = scipy.linalg.cholesky(k, lower=True)
    alpha = scipy.linalg.cho_solve((chol, True), err)
    qa = q / GP_LENGTH_SCALE
    k_star = GP_SIGMA2 * np.exp(-np.sqrt(((qa[:, None, :] - a[None, :, :]) ** 2).sum(-1)))   # [181, m]
    means = k_star @ alpha
    v = scipy.linalg.solve_triangular(chol, k_star.T, lower=True)
    var = np.maximum(GP_SIGMA2 - (v ** 2).sum(0), 0.0)
    return means, var / GP_SIGMA2                                            # :193


# ----------------------------------------------------------------------------- feature vector

class PerciatelliFeatures:
  """PerciatelliFeatureConstructor for N balloons riding on an oracle.env.OracleArena."""

  def __init__(self, arena):
    self.arena = arena
    self.gp = WindGP(arena.state.n)

  def observe(self, mask=None):
    """features.py:299-306: measurement = ground truth wind at the CURRENT state."""
    s = self.arena.state
    wu, wv = self.arena.ground_truth_at_balloon()
    fu, fv = self.arena.forecast(s.x, s.y, s.pressure, s.time_elapsed)
    self.gp.observe(s.x, s.y, s.pressure, s.time_elapsed, wu, wv, fu, fv, mask)

  def get_features(self):
    s, atm = self.arena.state, self.arena.atmosphere
    n = s.n
    out = np.zeros((n, NUM_FEATURES), np.float32)
    lat, lng = s.latlng()
    el, _, _ = solar.solar_calculator(lat, lng, s.date_time)
    sunrise_time = compute_sunrise_time(lat, lng, s.date_time)
    soc = s.battery_soc()
    pr = s.pressure_ratio()
    dist_m = np.sqrt(s.x * s.x + s.y * s.y)
    heading = np.arctan2(-s.x / 1000.0, -s.y / 1000.0)
    paused = s.navigation_is_paused()
    excess = s.excess_energy()
    # ambient features (features.py:382-455)
    out[:, 0] = np.clip((s.pressure - 5000.0) / 9000.0, 0.0, 1.0)
    out[:, 1] = soc
    out[:, 2] = np.clip((el + 90.0) / 180.0, 0.0, 1.0)
    out[:, 3] = np.sin(sunrise_time)
    out[:, 4] = np.cos(sunrise_time)
    out[:, 5] = np.sin(heading)
    out[:, 6] = np.cos(heading)
    out[:, 7] = (dist_m / 1000.0) / (dist_m / 1000.0 + 250.0)
    out[:, 8] = s.last_command == C.UP
    out[:, 9] = s.last_command == C.STAY
    out[:, 10] = s.last_command == C.DOWN
    out[:, 11] = paused
    out[:, 12] = ~paused
    out[:, 13] = excess
    out[:, 14] = [np.clip((power_table_lookup(pr[e], soc[e]) - 100.0) / 200.0, 0.0, 1.0) for e in range(n)]
    out[:, 15] = pr
    # wind features (features.py:457-556)
    pmin, pmax = get_pressure_range(s, atm)
    for e in range(n):
      err_means, dev = self.gp.query_column(e, s.x[e], s.y[e], s.time_elapsed[e])
      fu, fv = self.arena_forecast_column(e)
      means = err_means + np.stack([fu, fv], axis=1)                         # wind_gp.py:217-241
      p = min(max(s.pressure[e], 5000.0), 14000.0)
      level = int(round((p - 5000.0) / (PRESSURE_LEVELS[1] - PRESSURE_LEVELS[0])))   # banker's round, :376-377
      lower = NUM_LEVELS - level - 1
      station = -np.array([s.x[e], s.y[e]]) / (dist_m[e] + TOLERANCE_M)
      mag = np.sqrt((means ** 2).sum(1))
      unit = means / (mag + TOLERANCE_M)[:, None]
      if dist_m[e] < TOLERANCE_M:
        angle = np.zeros(NUM_LEVELS)
      else:
        angle = np.arccos(np.clip(unit @ station, -1.0, 1.0))
        angle = np.where(mag < TOLERANCE_M, np.pi, angle)
      col = np.empty((2 * NUM_LEVELS - 1, 3), np.float32)
      col[:] = (0.0, 1.0, 1.0)                                               # unreachable (features.py:558-581)
      reachable = (PRESSURE_LEVELS >= pmin[e]) & (PRESSURE_LEVELS <= pmax[e])
      rows = np.stack([dev, angle / np.pi, mag / (mag + 30.0)], axis=1).astype(np.float32)
      seg = col[lower:lower + NUM_LEVELS]
      seg[reachable] = rows[reachable]
      out[e, 16:] = col.reshape(-1)
    return out

  def arena_forecast_column(self, e):
    s = self.arena.state
    a = self.arena
    one = lambda v: np.full(NUM_LEVELS, v[e])
    static = np.broadcast_to(np.asarray(a.static_wind, bool), (s.n,))[e]
    if static:
      p = PRESSURE_LEVELS
      u = np.select([p < 8000.0, p < 10000.0, p < 12000.0], [10.0, 0.0, -10.0], default=0.0)
      v = np.select([p < 8000.0, p < 10000.0, p < 12000.0], [0.0, 10.0, 0.0], default=-10.0)
      return u, v
    return wind.get_forecast(a.fields, np.full(NUM_LEVELS, np.asarray(a.field_idx)[e]), one(s.x), one(s.y),
                             PRESSURE_LEVELS, one(s.time_elapsed))


# ---------------------------------------------------------------------------------------------------------
# The step-to-step update of the WindGP Cholesky factor used by csrc/ble_gp_kernels.cuh (k_gp_update),
# restated for the CPU tests: dropping the OLDEST measurement is a rank-1 update of the trailing factor with
# the factor's first column, appending the NEWEST is one forward substitution.  Mathematically identical to
# sklearn's refit on the new window (wind_gp.py:172-190).
# ---------------------------------------------------------------------------------------------------------
def cholesky_drop_first(l_factor):
  """Lower factor of K[1:, 1:] from the lower factor of K (K22 = L22 L22^T + l21 l21^T)."""
  m = l_factor.shape[0]
  x = l_factor[1:, 0].copy()
  out = l_factor[1:, 1:].copy()
  for k in range(m - 1):
    r = np.hypot(out[k, k], x[k])
    c, s = r / out[k, k], x[k] / out[k, k]
    out[k, k] = r
    out[k + 1:, k] = (out[k + 1:, k] + s * x[k + 1:]) / c
    x[k + 1:] = c * x[k + 1:] - s * out[k + 1:, k]
  return out


def cholesky_append(l_factor, k_new, k_diag):
  """Lower factor of [[K, k_new], [k_new^T, k_diag]] from the lower factor of K."""
  m = l_factor.shape[0]
  r = np.zeros(m)
  for j in range(m):
    r[j] = (k_new[j] - l_factor[j, :j] @ r[:j]) / l_factor[j, j]
  out = np.zeros((m + 1, m + 1))
  out[:m, :m] = l_factor
  out[m, :m] = r
  out[m, m] = np.sqrt(k_diag - r @ r)
  return out
