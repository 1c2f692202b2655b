"""Batched standard atmosphere (ORACLE / test infrastructure).

Follows env/balloon/standard_atmosphere.py:60-202.  One atmosphere per balloon,
parameterised by alpha in [0, 1] (the reference draws alpha = jax.random.uniform(key),
:82); lapse = (1 - alpha) * LOW + alpha * HIGH (:83-84).
"""
import numpy as np

from oracle import constants as C

HEIGHT_TRANSITIONS = np.array(
    [-610.0, 17000.0, 21000.0, 32000.0, 47000.0, 51000.0, 71000.0, 85000.0])  # :66-67
LAPSE_RATES_LOW = np.array([-0.007, 0.006, 0.001, 0.0028, 0.0, -0.0028, -0.002])   # :68-69
LAPSE_RATES_HIGH = np.array([-0.0058, 0.005, 0.001, 0.0028, 0.0, -0.0028, -0.002])  # :70-71
BASE_TEMPERATURE = 300.0      # :157
BASE_PRESSURE = 108870.8213   # :167
_R = C.DRY_AIR_SPECIFIC_GAS_CONSTANT
_G = C.GRAVITY


class Atmosphere:
  """N independent atmospheres (alpha[N])."""

  def __init__(self, alpha):
    alpha = np.atleast_1d(np.asarray(alpha, np.float64))
    self.alpha = alpha
    n = alpha.shape[0]
    self.lapse = ((1 - alpha)[:, None] * LAPSE_RATES_LOW[None, :] +
                  alpha[:, None] * LAPSE_RATES_HIGH[None, :])            # :83-84
    self.t_tr = np.empty((n, 8))
    self.t_tr[:, 0] = BASE_TEMPERATURE
    for i in range(7):                                                    # :156-162
      self.t_tr[:, i + 1] = self.t_tr[:, i] + self.lapse[:, i] * (
          HEIGHT_TRANSITIONS[i + 1] - HEIGHT_TRANSITIONS[i])
    self.p_tr = np.empty((n, 8))
    self.p_tr[:, 0] = BASE_PRESSURE
    for i in range(7):                                                    # :164-183
      lapse = self.lapse[:, i]
      zero = lapse == 0.0
      safe = np.where(zero, 1.0, lapse)
      const_t = self.p_tr[:, i] * np.exp(
          -(_G * (HEIGHT_TRANSITIONS[i + 1] - HEIGHT_TRANSITIONS[i])) /
          (_R * self.t_tr[:, i + 1]))                                     # :186-192
      lin_t = self.p_tr[:, i] * (
          (self.t_tr[:, i + 1] / self.t_tr[:, i]) ** (-_G / (_R * safe)))  # :194-202
      self.p_tr[:, i + 1] = np.where(zero, const_t, lin_t)

  def subset(self, idx):
    return Atmosphere(self.alpha[idx])

  def _layer_for_pressure(self, pressure):
    # First i with pressure > p_tr[i+1]  (:132-134).
    gt = pressure[:, None] > self.p_tr[:, 1:]
    if not np.all(gt.any(axis=1)) or not np.all(pressure <= self.p_tr[:, 0]):
      raise AssertionError('pressure outside atmosphere range')           # :126-127
    return np.argmax(gt, axis=1)

  def at_pressure(self, pressure):
    """-> (height [m], temperature [K]) following :122-154."""
    pressure = np.broadcast_to(np.asarray(pressure, np.float64), self.alpha.shape)
    i = self._layer_for_pressure(pressure)
    r = np.arange(pressure.shape[0])
    lapse, t_i, p_i, h_i = self.lapse[r, i], self.t_tr[r, i], self.p_tr[r, i], HEIGHT_TRANSITIONS[i]
    zero = lapse == 0.0
    safe = np.where(zero, 1.0, lapse)
    h_const = (-_R * t_i / _G) * np.log(pressure / p_i) + h_i            # :137-140
    h_lin = ((pressure / p_i) ** (-_R * safe / _G) - 1) * t_i / safe + h_i  # :142-146
    height = np.where(zero, h_const, h_lin)
    temperature = t_i + lapse * (height - h_i)                           # :148-149
    return height, temperature

  def at_height(self, height):
    """-> (pressure [Pa], temperature [K]) following :89-120."""
    height = np.broadcast_to(np.asarray(height, np.float64), self.alpha.shape)
    assert np.all(height >= HEIGHT_TRANSITIONS[0]) and np.all(height < HEIGHT_TRANSITIONS[-1])
    i = np.argmax(height[:, None] < HEIGHT_TRANSITIONS[None, 1:], axis=1)  # :103
    r = np.arange(height.shape[0])
    lapse, t_i, p_i, h_i = self.lapse[r, i], self.t_tr[r, i], self.p_tr[r, i], HEIGHT_TRANSITIONS[i]
    temperature = t_i + lapse * (height - h_i)                           # :105-106
    zero = lapse == 0.0
    safe = np.where(zero, 1.0, lapse)
    p_const = p_i * np.exp(-(_G * (height - h_i)) / (_R * temperature))   # :109-111
    p_lin = p_i * ((temperature / t_i) ** (-_G / (_R * safe)))            # :113-115
    return np.where(zero, p_const, p_lin), temperature
