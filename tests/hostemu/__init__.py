"""TEST-ONLY: builds and loads the host replay of the CUDA device headers (see hostemu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libhostemu.so')
_SRC = os.path.join(_HERE, 'hostemu.cpp')
_HDRS = [os.path.join(_HERE, '..', '..', 'balloon_learning_environment_b200', 'csrc', f)
         for f in ('ble_physics.cuh', 'ble_wind.cuh', 'ble_features.cuh', 'ble_agents.cuh', 'ble_step_roles.cuh', 'ble_fastmath.cuh')] + [os.path.join(_HERE, '..', '..', 'include', 'ble_b200.h')]


def build(force=False):
  newest = max(os.path.getmtime(p) for p in [_SRC] + _HDRS)
  if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off',
                           '-x', 'c++', _SRC, '-o', _SO])
  return _SO


def load():
  lib = ctypes.CDLL(build())
  return lib


def ptr(a):
  return a.ctypes.data_as(ctypes.c_void_p)


F_ROWS = ['x', 'y', 'pressure', 'ambient_temperature', 'internal_temperature', 'envelope_volume',
          'superpressure', 'mols_air', 'mols_lift_gas', 'battery_charge', 'acs_power',
          'acs_mass_flow', 'solar_charging', 'power_load', 'center_lat', 'center_lng',
          'upwelling_infrared', 'alpha']
I_ROWS = ['date_time', 'time_elapsed', 'last_command', 'status', 'envelope_state', 'altitude_state',
          'power_paused', 'sunrise_h', 'sunset', 'power_safety_enabled']


def pack_state(batch, alpha, power_safety_enabled=True):
  """oracle BalloonBatch -> (f64 [NF,N], i64 [NI,N]) in the C-ABI row order (include/ble_b200.h)."""
  n = batch.n
  f = np.empty((len(F_ROWS), n), np.float64)
  i = np.empty((len(I_ROWS), n), np.int64)
  for r, k in enumerate(F_ROWS):
    f[r] = alpha if k == 'alpha' else getattr(batch, k)
  for r, k in enumerate(I_ROWS):
    i[r] = (np.broadcast_to(np.asarray(power_safety_enabled, np.int64), (n,))
            if k == 'power_safety_enabled' else getattr(batch, k))
  return f, i


def unpack_state(f, i, batch):
  for r, k in enumerate(F_ROWS):
    if k != 'alpha':
      setattr(batch, k, f[r].copy())
  for r, k in enumerate(I_ROWS):
    if k != 'power_safety_enabled':
      setattr(batch, k, i[r].copy())
  return batch


def emu_step(lib, precision, f, i, actions, wind):
  n = f.shape[1]
  reward = np.zeros(n); eff = np.zeros(n, np.int32)
  actions = np.ascontiguousarray(actions, np.int32); wind = np.ascontiguousarray(wind, np.float64)
  lib.emu_step(ctypes.c_int(precision), ctypes.c_int64(n), ptr(f), ptr(i), ptr(actions), ptr(wind),
               ptr(reward), ptr(eff))
  return reward, eff
