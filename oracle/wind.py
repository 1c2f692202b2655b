"""Wind field lookup: grid interpolation + simplex noise (ORACLE / test infrastructure).

Follows env/grid_based_wind_field.py:70-187 (+ generative/vae.py:26-93 for the grid),
env/wind_field.py:125-218, env/simplex_wind_noise.py:50-211.
"""
import numpy as np

from oracle import constants as C
from oracle import opensimplex4

GRID_XY = np.linspace(-C.FIELD_DISPLACEMENT_KM, C.FIELD_DISPLACEMENT_KM, C.FIELD_XY)   # vae.py:77-81
GRID_P = np.linspace(C.FIELD_MIN_PRESSURE, C.FIELD_MAX_PRESSURE, C.FIELD_P)            # vae.py:83-87
GRID_T = np.linspace(0, C.FIELD_TIME_HORIZON_H, C.FIELD_T).astype(np.int32).astype(np.float64)
FIELD_SHAPE = (C.FIELD_XY, C.FIELD_XY, C.FIELD_P, C.FIELD_T, 2)

# weight, x, y, pressure, time spacings (simplex_wind_noise.py:50-64)
U_HARMONICS = np.array([
    [0.1445, 702.269, 2116.987, 2587.802, 245.0],
    [0.2766, 1483.570, 752.124, 646.208, 16.39],
    [0.2627, 276.810, 147.040, 587.702, 3.836],
    [0.2137, 10214.525, 1512.216, 965.629, 41.780],
    [0.1025, 181.286, 420.942, 8500.0, 245.0]])
V_HARMONICS = np.array([
    [0.2716, 1974.228, 2028.814, 713.697, 26.435],
    [0.2684, 699.738, 541.845, 632.116, 9.530],
    [0.2348, 217.750, 196.522, 686.825, 3.546],
    [0.1186, 47.500, 43.048, 66.553, 8.424],
    [0.1066, 3663.291, 232.023, 7499.741, 225.0]])
HARMONICS = np.stack([U_HARMONICS, V_HARMONICS])    # [2, 5, 5]


def boomerang(t, max_val):
  """grid_based_wind_field.py:134-143 (vectorised)."""
  cycle = (t / max_val).astype(np.int64) % 2
  rem = np.mod(t, max_val)
  return np.where(cycle % 2 == 0, rem, max_val - rem)


def prepare_points(x_m, y_m, pressure, elapsed_s):
  """-> fp32 [N, 4] = (x_km, y_km, p, t_h) as in _prepare_get_forecast_inputs (:145-187)."""
  x_km = np.clip(np.asarray(x_m, np.float64) / 1000.0, -C.FIELD_DISPLACEMENT_KM, C.FIELD_DISPLACEMENT_KM)
  y_km = np.clip(np.asarray(y_m, np.float64) / 1000.0, -C.FIELD_DISPLACEMENT_KM, C.FIELD_DISPLACEMENT_KM)
  p = np.clip(np.asarray(pressure, np.float64), C.FIELD_MIN_PRESSURE, C.FIELD_MAX_PRESSURE)
  hours = np.asarray(elapsed_s, np.float64) / 3600.0
  t = np.where(hours < C.FIELD_TIME_HORIZON_H, hours, boomerang(hours, float(C.FIELD_TIME_HORIZON_H)))
  x_km, y_km, p, t = np.broadcast_arrays(x_km, y_km, p, t)
  return np.stack([x_km, y_km, p, t], axis=-1).astype(np.float32)


def _axis(grid, v):
  n = grid.shape[0]
  i = np.clip(np.searchsorted(grid, v, side='right') - 1, 0, n - 2)
  w = (v - grid[i]) / (grid[i + 1] - grid[i])
  return i, w


def interpolate(fields, field_idx, points):
  """16-corner multilinear interpolation == scipy.interpolate.interpn(linear) (:91).

  fields: float32 [F, 21, 21, 10, 9, 2]; field_idx: int [N]; points: fp32 [N, 4].
  Coordinates are rounded to fp32 first (the reference packs them into an fp32 array,
  :181), weights and accumulation are fp64.
  """
  pts = np.asarray(points, np.float32).astype(np.float64)
  fi = np.asarray(field_idx, np.int64)
  ix, wx = _axis(GRID_XY, pts[:, 0])
  iy, wy = _axis(GRID_XY, pts[:, 1])
  ip, wp = _axis(GRID_P, pts[:, 2])
  it, wt = _axis(GRID_T, pts[:, 3])
  out = np.zeros((pts.shape[0], 2))
  for a in (0, 1):
    for b in (0, 1):
      for c in (0, 1):
        for d in (0, 1):
          w = ((wx if a else 1 - wx) * (wy if b else 1 - wy) *
               (wp if c else 1 - wp) * (wt if d else 1 - wt))
          out += w[:, None] * fields[fi, ix + a, iy + b, ip + c, it + d, :].astype(np.float64)
  return out


def get_forecast(fields, field_idx, x_m, y_m, pressure, elapsed_s):
  """(u, v) m/s; GridBasedWindField.get_forecast (:70-94)."""
  uv = interpolate(fields, field_idx, prepare_points(x_m, y_m, pressure, elapsed_s))
  return uv[:, 0], uv[:, 1]


class SimplexWindNoise:
  """N independent noise models (wind_field.py:188-218).

  seeds: int [N, 2, 5] (component u/v, harmonic); offsets: float [N, 2, 5, 4].
  The reference draws seed = jax.random.choice(key, 1634753849) and
  offsets = uniform(key, (4,)) * 2 - 1 as fp32 (simplex_wind_noise.py:98-114).
  """

  def __init__(self, seeds, offsets):
    self.seeds = np.asarray(seeds, np.int64)
    self.offsets = np.asarray(offsets, np.float64)
    assert self.seeds.shape[1:] == (2, 5) and self.offsets.shape[1:] == (2, 5, 4)
    self.perms = opensimplex4.make_perms(self.seeds)         # uint8 [N, 2, 5, 256]

  def get_wind_noise(self, x_m, y_m, pressure, elapsed_s):
    """-> (du, dv); NoisyWindComponent.get_noise (:180-211) for u and v."""
    x_km = np.asarray(x_m, np.float64) / 1000.0
    y_km = np.asarray(y_m, np.float64) / 1000.0
    p = np.asarray(pressure, np.float64)
    t_h = np.asarray(elapsed_s, np.float64) / 3600.0
    out = []
    for comp in (0, 1):
      weighted = 0.0; total_w = 0.0; total_w2 = 0.0
      for h in range(5):
        wgt, sx, sy, sp, st = HARMONICS[comp, h]
        off = self.offsets[:, comp, h, :]
        noise = C.NOISE_MAGNITUDE * opensimplex4.noise4d(
            self.perms[:, comp, h, :], x_km / sx + off[:, 0], y_km / sy + off[:, 1],
            p / sp + off[:, 2], t_h / st + off[:, 3])                        # :142-146
        weighted = weighted + noise * wgt
        total_w += wgt
        total_w2 += wgt ** 2
      weighted = weighted / total_w
      weighted = weighted * np.sqrt(total_w / total_w2)                      # :205-207
      out.append(weighted)
    return out[0], out[1]
