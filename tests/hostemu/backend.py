"""TEST-ONLY stand-in for BatchedBalloonArena on a machine without a GPU.

`CudaBalloonArena` (the N = 1 adaptor behind the reference's BalloonArenaInterface) talks to a
`BatchedBalloonArena`; this class offers the same methods on the CPU so that the ADAPTOR's Python surface -- the
objects it hands to the reference's BalloonEnv, reward function and info dict -- can be executed here, under the
reference itself (tests/test_adaptor_under_reference.py).  The transition runs the device headers' production
arithmetic through tests/hostemu (emu_step, precision 2); winds, derived properties and the atmosphere come from the
oracle.  Never imported by the product package.
"""
import numpy as np
import torch

from oracle import atmosphere as atmosphere_lib
from oracle import balloon as balloon_lib
from oracle import constants as C
from oracle import solar
from oracle import wind as wind_lib
from tests import hostemu


class HostBackend:
  enable_features = False
  device = torch.device('cpu')

  def __init__(self, num_envs=1, precision=2):
    self.num_envs = num_envs
    self._lib = hostemu.load()
    self._prec = precision
    self._f = np.zeros((len(hostemu.F_ROWS), num_envs))
    self._i = np.zeros((len(hostemu.I_ROWS), num_envs), np.int64)
    self._fields, self._fidx, self._noise = None, np.zeros(num_envs, np.int64), None
    self._last = None

  # -- wind ---------------------------------------------------------------------------------------------------------
  def set_wind_fields(self, fields, env_to_field=None):
    self._fields = np.ascontiguousarray(fields.numpy(), np.float32)
    if env_to_field is not None:
      self._fidx = env_to_field.numpy().astype(np.int64)

  def set_wind_noise(self, seeds, offsets):
    self._noise = wind_lib.SimplexWindNoise(seeds.numpy(), offsets.numpy().astype(np.float64))

  def _wind(self, x, y, p, t, idx, with_noise):
    u, v = wind_lib.get_forecast(self._fields, self._fidx[idx], x, y, p, t)
    if with_noise and self._noise is not None:
      sub = wind_lib.SimplexWindNoise(self._noise.seeds[idx], self._noise.offsets[idx])
      du, dv = sub.get_wind_noise(x, y, p, t)
      u, v = u + du, v + dv
    return u, v

  def wind_query(self, xyzt, env_idx, with_noise):
    q, idx = xyzt.numpy(), env_idx.numpy().astype(np.int64)
    u, v = self._wind(q[:, 0], q[:, 1], q[:, 2], q[:, 3], idx, with_noise)
    return torch.from_numpy(np.stack([u, v], 1).astype(np.float32))

  def wind_at_balloon(self):
    r = hostemu.F_ROWS.index
    idx = np.arange(self.num_envs)
    u, v = self._wind(self._f[r('x')], self._f[r('y')], self._f[r('pressure')],
                      self._i[hostemu.I_ROWS.index('time_elapsed')].astype(np.float64), idx, True)
    return torch.from_numpy(np.stack([u, v], 1).astype(np.float32))

  def atmosphere_query(self, which, q, env_idx):
    alpha = self._f[hostemu.F_ROWS.index('alpha')][env_idx.numpy().astype(np.int64)]
    atm = atmosphere_lib.Atmosphere(alpha)
    out = np.full((q.shape[0], 4), np.nan)
    v = q.numpy()
    try:
      if which == 'pressure':
        h, t = atm.at_pressure(v)
        out[:, 0], out[:, 1], out[:, 2] = h, t, v
      else:
        p, t = atm.at_height(v)
        out[:, 0], out[:, 1], out[:, 2] = v, t, p
      out[:, 3] = out[:, 2] * C.DRY_AIR_MOLAR_MASS / (C.UNIVERSAL_GAS_CONSTANT * out[:, 1])
    except AssertionError:
      pass
    return torch.from_numpy(out)

  def reset(self, seeds):
    """Any valid state: the tests inject the state they compare from right after construction."""
    del seeds
    b = balloon_lib.make_batch(self.num_envs, center_lat=0.0, center_lng=0.0, date_time=1364203532, pressure=9000.0)
    self._f, self._i = hostemu.pack_state(b, 0.5, True)

  # -- state --------------------------------------------------------------------------------------------------------
  def set_state(self, f64, i64):
    self._f = np.ascontiguousarray(f64.numpy(), np.float64).copy()
    self._i = np.ascontiguousarray(i64.numpy(), np.int64).copy()

  def get_state(self):
    return torch.from_numpy(self._f.copy()), torch.from_numpy(self._i.copy())

  def _batch(self):
    b = balloon_lib.make_batch(self.num_envs, center_lat=0.0, center_lng=0.0, date_time=0)
    return hostemu.unpack_state(self._f, self._i, b)

  def get_derived(self):
    b = self._batch()
    lat, lng = b.latlng()
    el, _, flux = solar.solar_calculator(lat, lng, b.date_time)
    atm = atmosphere_lib.Atmosphere(self._f[hostemu.F_ROWS.index('alpha')])
    h, _ = atm.at_pressure(b.pressure)
    vals = dict(lat=lat, lng=lng, solar_elevation=el, solar_flux=flux, excess_energy=b.excess_energy().astype(np.float64),
                navigation_is_paused=b.navigation_is_paused().astype(np.float64), pressure_ratio=b.pressure_ratio(),
                battery_soc=b.battery_soc(), altitude=h)
    return {k: torch.from_numpy(np.asarray(v, np.float64)) for k, v in vals.items()}

  # -- step ---------------------------------------------------------------------------------------------------------
  def step(self, actions):
    acts = actions.numpy().astype(np.int32)
    wind = self.wind_at_balloon().numpy().astype(np.float64)
    live = self._i[hostemu.I_ROWS.index('status')] == 0
    reward, _ = hostemu.emu_step(self._lib, self._prec, self._f, self._i, acts, wind)
    done = (self._i[hostemu.I_ROWS.index('status')] != 0).astype(np.uint8)
    self._last = (torch.from_numpy(np.where(live, reward, 0.0).astype(np.float32)), torch.from_numpy(done),
                  torch.from_numpy(wind.astype(np.float32)))
    return self._last

  def step_info(self):
    st = torch.from_numpy(self._i[hostemu.I_ROWS.index('status')].copy())
    return {'out_of_power': st == 1, 'envelope_burst': st == 2, 'zeropressure': st == 3,
            'time_elapsed': torch.from_numpy(self._i[hostemu.I_ROWS.index('time_elapsed')].astype(np.int32)),
            'sim_error': torch.zeros(self.num_envs, dtype=torch.bool)}

  def features_clear(self): pass
  def features_observe(self): pass
  def close(self): pass
