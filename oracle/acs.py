"""ACS compressor tables (ORACLE / test infrastructure).  Follows env/balloon/acs.py:24-68.

The reference builds scipy `interp1d(..., fill_value='extrapolate')` and the (removed)
`interp2d(..., fill_value=None)`; both are plain piecewise-linear interpolants, restated
here: 1-D linear with linear extrapolation from the end segments, 2-D bilinear with
nearest (clamped) extrapolation.
"""
import numpy as np

from oracle import constants as C

_PR_KNOTS = np.array([1.0, 1.05, 1.2, 1.25, 1.35])            # :26
_POWER_KNOTS = np.array([100.0, 100.0, 300.0, 400.0, 400.0])  # :27
_EFF_PR = np.linspace(1.05, 1.35, 13)                         # :33
_EFF_W = np.linspace(100.0, 400.0, 4)                         # :34
_EFF = np.array([0.4, 0.4, 0.3, 0.2, 0.2, 0.00000, 0.00000, 0.00000, 0.00000,
                 0.00000, 0.00000, 0.00000, 0.00000, 0.4, 0.3, 0.3, 0.30, 0.25,
                 0.23, 0.20, 0.15, 0.12, 0.10, 0.00000, 0.00000, 0.00000,
                 0.00000, 0.3, 0.25, 0.25, 0.25, 0.20, 0.20, 0.20, 0.2, 0.15,
                 0.13, 0.12, 0.11, 0.00000, 0.23, 0.23, 0.23, 0.23, 0.23, 0.20,
                 0.20, 0.20, 0.18, 0.16, 0.15, 0.13]).reshape(4, 13)   # [W][pr] :35-41


def get_most_efficient_power(pressure_ratio):
  """Watts; acs.py:44-60 (interp1d linear, extrapolating)."""
  pr = np.asarray(pressure_ratio, np.float64)
  i = np.clip(np.searchsorted(_PR_KNOTS, pr, side='left') - 1, 0, len(_PR_KNOTS) - 2)
  x0, x1 = _PR_KNOTS[i], _PR_KNOTS[i + 1]
  y0, y1 = _POWER_KNOTS[i], _POWER_KNOTS[i + 1]
  slope = (y1 - y0) / (x1 - x0)
  return slope * (pr - x0) + y0


def get_fan_efficiency(pressure_ratio, power_w):
  """acs.py:63-66 (bilinear, clamped outside the table)."""
  pr = np.clip(np.asarray(pressure_ratio, np.float64), _EFF_PR[0], _EFF_PR[-1])
  w = np.clip(np.asarray(power_w, np.float64), _EFF_W[0], _EFF_W[-1])
  i = np.clip(np.searchsorted(_EFF_PR, pr, side='right') - 1, 0, 11)
  j = np.clip(np.searchsorted(_EFF_W, w, side='right') - 1, 0, 2)
  tx = (pr - _EFF_PR[i]) / (_EFF_PR[i + 1] - _EFF_PR[i])
  ty = (w - _EFF_W[j]) / (_EFF_W[j + 1] - _EFF_W[j])
  return ((1 - tx) * (1 - ty) * _EFF[j, i] + tx * (1 - ty) * _EFF[j, i + 1]
          + (1 - tx) * ty * _EFF[j + 1, i] + tx * ty * _EFF[j + 1, i + 1])


def get_mass_flow(power_w, efficiency):
  """kg/s; acs.py:67-68."""
  return efficiency * power_w / C.NUM_SECONDS_PER_HOUR
