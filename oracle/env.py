"""Batched arena / env step (ORACLE / test infrastructure).

Follows env/balloon_arena.py:184-202,222-226,270-275 (arena step and measurements) and
env/balloon_env.py:44-102,157-190,280-290 (reward, terminal, info).
"""
import numpy as np

from oracle import balloon as balloon_lib
from oracle import constants as C
from oracle import wind


def perciatelli_reward(b, *, station_keeping_radius_km=C.REWARD_RADIUS_KM,
                       reward_dropoff=C.REWARD_DROPOFF, reward_halflife=C.REWARD_HALFLIFE):
  """env/balloon_env.py:44-102 on a BalloonBatch (post-step state)."""
  distance_m = np.sqrt(b.x * b.x + b.y * b.y)                              # units.relative_distance
  radius_m = station_keeping_radius_km * 1000.0
  outside = reward_dropoff * np.exp(
      -0.69314718056 / reward_halflife * ((distance_m - radius_m) / 1000.0))
  reward = np.where(distance_m <= radius_m, 1.0, outside)                  # :82-86
  penalise = (b.last_command == C.DOWN) & ~b.excess_energy()               # :88-89
  scale = np.clip((b.acs_power - 100.0) / (300.0 - 100.0), 0.0, 1.0)       # transforms.py:63-66
  multiplier = 0.95 - 0.3 * scale
  return np.where(penalise, reward * multiplier, reward)


class OracleArena:
  """BalloonArena.step restated for N balloons with per-balloon fields / noise / atmosphere.

  fields: float32 [F,21,21,10,9,2] + field_idx int[N]  (GridBasedWindField), or None with
  `static_wind=True` for SimpleStaticWindField (env/wind_field.py:149-184).
  noise: wind.SimplexWindNoise or None (forecast == ground truth).
  """

  def __init__(self, state, atmosphere, fields=None, field_idx=None, noise=None,
               static_wind=False, power_safety_layer_enabled=True):
    self.state = state
    self.atmosphere = atmosphere
    self.fields = fields
    self.field_idx = field_idx
    self.noise = noise
    self.static_wind = static_wind     # bool or bool[N]
    self.power_safety_layer_enabled = power_safety_layer_enabled
    self.last_effective_action = None
    self.last_wind = None

  def forecast(self, x, y, pressure, elapsed_s):
    static = np.broadcast_to(np.asarray(self.static_wind, bool), np.shape(pressure))
    p = np.asarray(pressure)
    us = np.select([p < 8000.0, p < 10000.0, p < 12000.0], [10.0, 0.0, -10.0], default=0.0)
    vs = np.select([p < 8000.0, p < 10000.0, p < 12000.0], [0.0, 10.0, 0.0], default=-10.0)
    if static.all():
      return us, vs
    fi = np.where(static, 0, np.asarray(self.field_idx))
    u, v = wind.get_forecast(self.fields, fi, x, y, pressure, elapsed_s)
    return np.where(static, us, u), np.where(static, vs, v)

  def ground_truth_at_balloon(self):
    """env/balloon_arena.py:270-275 -> wind_field.py:125-145."""
    s = self.state
    u, v = self.forecast(s.x, s.y, s.pressure, s.time_elapsed)
    if self.noise is not None:
      du, dv = self.noise.get_wind_noise(s.x, s.y, s.pressure, s.time_elapsed)
      u, v = u + du, v + dv
    return u, v

  def step(self, action):
    """env/balloon_arena.py:184-202 (without feature construction)."""
    u, v = self.ground_truth_at_balloon()                                  # PRE-step lookup
    self.last_wind = (u, v)
    self.last_effective_action = balloon_lib.simulate_step(
        self.state, u, v, self.atmosphere, action,
        power_safety_layer_enabled=self.power_safety_layer_enabled)
    return self.state


class OracleEnv:
  """BalloonEnv.step restated for N balloons: (reward, done, info)."""

  def __init__(self, arena: OracleArena):
    self.arena = arena

  def step(self, action):
    was_live = self.arena.state.status == C.STATUS_OK
    s = self.arena.step(action)
    reward = perciatelli_reward(s)
    info = dict(out_of_power=s.status == C.STATUS_OUT_OF_POWER,
                envelope_burst=s.status == C.STATUS_BURST,
                zeropressure=s.status == C.STATUS_ZEROPRESSURE,
                time_elapsed=s.time_elapsed.copy())
    done = info['out_of_power'] | info['envelope_burst'] | info['zeropressure']
    return np.where(was_live, reward, 0.0), done, info
