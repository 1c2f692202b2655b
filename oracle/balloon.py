"""Batched balloon state + Euler integrator (ORACLE / test infrastructure).

Follows env/balloon/balloon.py:73-250 (BalloonState), :263-328 (simulate_step),
:356-549 (_simulate_step_internal) and :552-609 (calculate_superpressure_and_volume).
All arrays are fp64 / int64 of shape [N].
"""
import copy
import dataclasses

import numpy as np

from oracle import acs
from oracle import constants as C
from oracle import geometry
from oracle import safety
from oracle import solar
from oracle import thermal

FLOAT_FIELDS = ('x', 'y', 'pressure', 'ambient_temperature', 'internal_temperature',
                'envelope_volume', 'superpressure', 'mols_air', 'mols_lift_gas',
                'battery_charge', 'acs_power', 'acs_mass_flow', 'solar_charging', 'power_load',
                'center_lat', 'center_lng', 'upwelling_infrared')
INT_FIELDS = ('date_time', 'time_elapsed', 'last_command', 'status', 'envelope_state',
              'altitude_state', 'power_paused', 'sunrise_h', 'sunset')


@dataclasses.dataclass
class BalloonBatch:
  """Struct-of-arrays BalloonState (env/balloon/balloon.py:73-208) for N balloons.

  x, y metres; pressure Pa; battery_charge Wh; powers W; date_time / sunrise_h / sunset
  int64 unix seconds; time_elapsed int64 seconds; center_lat/lng radians.
  """
  x: np.ndarray
  y: np.ndarray
  pressure: np.ndarray
  ambient_temperature: np.ndarray
  internal_temperature: np.ndarray
  envelope_volume: np.ndarray
  superpressure: np.ndarray
  mols_air: np.ndarray
  mols_lift_gas: np.ndarray
  battery_charge: np.ndarray
  acs_power: np.ndarray
  acs_mass_flow: np.ndarray
  solar_charging: np.ndarray
  power_load: np.ndarray
  center_lat: np.ndarray
  center_lng: np.ndarray
  upwelling_infrared: np.ndarray
  date_time: np.ndarray
  time_elapsed: np.ndarray
  last_command: np.ndarray
  status: np.ndarray
  envelope_state: np.ndarray
  altitude_state: np.ndarray
  power_paused: np.ndarray
  sunrise_h: np.ndarray      # PowerSafetyLayer._sunrise_with_hysteresis
  sunset: np.ndarray         # PowerSafetyLayer._sunset

  @property
  def n(self):
    return self.x.shape[0]

  def copy(self):
    return copy.deepcopy(self)

  def latlng(self):                                                        # :217-220
    return geometry.latlng_from_offset(self.center_lat, self.center_lng, self.x, self.y)

  def battery_soc(self):                                                   # :222-229
    return self.battery_charge / C.BATTERY_CAPACITY_WH

  def excess_energy(self):                                                 # :231-238
    lat, lng = self.latlng()
    el, _, _ = solar.solar_calculator(lat, lng, self.date_time)
    return (solar.solar_power(el, self.pressure) > C.DAYTIME_POWER_LOAD_W) & (
        self.battery_soc() > 0.99)

  def navigation_is_paused(self):                                          # :240-245
    return (self.power_paused.astype(bool) | (self.envelope_state != C.ENV_NOMINAL) |
            (self.altitude_state != C.ALT_NOMINAL))

  def pressure_ratio(self):                                                # :247-250
    return (self.pressure + np.maximum(self.superpressure, 0.0)) / self.pressure

  def select(self, idx):
    return BalloonBatch(**{f.name: getattr(self, f.name)[idx].copy()
                           for f in dataclasses.fields(self)})

  def as_dict(self):
    return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}


def make_batch(n, *, center_lat, center_lng, date_time, x=0.0, y=0.0, pressure=6000.0,
               upwelling_infrared=250.0, battery_charge=C.DEFAULT_BATTERY_CHARGE_WH,
               time_elapsed=0):
  """Defaults of BalloonState (:175-208) + the PowerSafetyLayer constructor (:210-215)."""
  f = lambda v: np.broadcast_to(np.asarray(v, np.float64), (n,)).copy()
  i = lambda v: np.broadcast_to(np.asarray(v, np.int64), (n,)).copy()
  b = BalloonBatch(
      x=f(x), y=f(y), pressure=f(pressure), ambient_temperature=f(206.0),
      internal_temperature=f(206.0), envelope_volume=f(1804.0), superpressure=f(0.0),
      mols_air=f(0.0), mols_lift_gas=f(C.DEFAULT_MOLS_LIFT_GAS), battery_charge=f(battery_charge),
      acs_power=f(0.0), acs_mass_flow=f(0.0), solar_charging=f(0.0), power_load=f(0.0),
      center_lat=f(center_lat), center_lng=f(center_lng), upwelling_infrared=f(upwelling_infrared),
      date_time=i(date_time), time_elapsed=i(time_elapsed), last_command=i(C.STAY),
      status=i(C.STATUS_OK), envelope_state=i(C.ENV_NOMINAL), altitude_state=i(C.ALT_NOMINAL),
      power_paused=i(0), sunrise_h=i(0), sunset=i(0))
  init_power_safety(b)
  return b


def init_power_safety(b: BalloonBatch):
  """PowerSafetyLayer.__init__ (env/balloon/power_safety.py:33-50)."""
  lat, lng = b.latlng()
  sunrise, sunset = solar.get_next_sunrise_sunset(lat, lng, b.date_time)
  b.sunrise_h = sunrise + C.POWER_TIME_HYSTERESIS_S
  b.sunset = sunset
  b.power_paused = np.zeros(b.n, np.int64)


def calculate_superpressure_and_volume(mols_lift_gas, mols_air, internal_temperature, pressure,
                                       volume_base=C.ENVELOPE_VOLUME_BASE,
                                       dv_pressure=C.ENVELOPE_VOLUME_DV_PRESSURE):
  """-> (envelope_volume, superpressure); :552-609."""
  unconstrained = ((mols_lift_gas + mols_air) * C.UNIVERSAL_GAS_CONSTANT *
                   internal_temperature / pressure)                        # :581-584
  b = -(volume_base - dv_pressure * pressure)                              # :597-599
  c = -(dv_pressure * unconstrained * pressure)                            # :600-602
  vol_full = 0.5 * (-b + np.sqrt(b * b - 4 * c))                           # :604
  sp_full = pressure * unconstrained / vol_full - pressure                 # :605-607
  slack = unconstrained <= volume_base                                     # :586
  return np.where(slack, unconstrained, vol_full), np.where(slack, 0.0, sp_full)


def simulate_step_internal(s: BalloonBatch, u, v, atmosphere, action, stride=C.PHYSICS_STRIDE_S):
  """One explicit-Euler sub-step; returns a dict of changed fields (:356-549).

  Every right-hand side reads the OLD state `s`.
  """
  ch = {}
  ch['x'] = s.x + u * stride                                               # :394-395
  ch['y'] = s.y + v * stride

  rho_air = (s.pressure * C.DRY_AIR_MOLAR_MASS) / (
      C.UNIVERSAL_GAS_CONSTANT * s.ambient_temperature)                    # :412-413
  drag = C.ENVELOPE_COD * s.envelope_volume ** (2.0 / 3.0)                 # :415
  total_mass = (C.HE_MOLAR_MASS * s.mols_lift_gas + C.DRY_AIR_MOLAR_MASS * s.mols_air +
                C.ENVELOPE_MASS + C.PAYLOAD_MASS)                          # :417-420
  direction = np.where(rho_air * s.envelope_volume >= total_mass, 1.0, -1.0)
  dh_dt = direction * np.sqrt(np.abs(
      2 * (rho_air * s.envelope_volume - total_mass) * C.GRAVITY / (rho_air * drag)))  # :424-427
  dp = 1.0
  height0, _ = atmosphere.at_pressure(s.pressure)                          # :439
  height1, _ = atmosphere.at_pressure(s.pressure + direction * dp)         # :440-441
  dp_dh = direction * dp / (height1 - height0)
  dp_dt = dp_dh * dh_dt
  ch['pressure'] = s.pressure + dp_dt * stride                             # :445

  lat, lng = s.latlng()
  el, _, flux = solar.solar_calculator(lat, lng, s.date_time)              # :451-452
  _, ch['ambient_temperature'] = atmosphere.at_pressure(s.pressure)        # :457-458
  d_temp = thermal.d_balloon_temperature_dt(
      s.envelope_volume, C.ENVELOPE_MASS, s.internal_temperature, s.ambient_temperature,
      s.pressure, el, flux, s.upwelling_infrared)                          # :462-465
  ch['internal_temperature'] = s.internal_temperature + d_temp * stride    # :466-467

  ch['envelope_volume'], ch['superpressure'] = calculate_superpressure_and_volume(
      s.mols_lift_gas, s.mols_air, s.internal_temperature, s.pressure)     # :470-477
  status = s.status.copy()
  status = np.where(ch['superpressure'] > C.ENVELOPE_MAX_SUPERPRESSURE, C.STATUS_BURST, status)
  status = np.where(ch['superpressure'] <= 0.0, C.STATUS_ZEROPRESSURE, status)  # :479-482

  # ACS (:487-513).  UP branch takes sqrt of the OLD superpressure.
  valve_area = np.pi * C.ACS_VALVE_HOLE_DIAMETER_M ** 2 / 4.0
  gas_density = (s.superpressure + s.pressure) * C.DRY_AIR_MOLAR_MASS / (
      C.UNIVERSAL_GAS_CONSTANT * s.internal_temperature)
  with np.errstate(invalid='ignore'):
    flow_up = -1 * C.DEFAULT_VALVE_HOLE_CD * valve_area * np.sqrt(
        2.0 * s.superpressure * gas_density)
  pr = s.pressure_ratio()
  power_down = acs.get_most_efficient_power(pr)
  flow_down = acs.get_mass_flow(power_down, acs.get_fan_efficiency(pr, power_down))
  ch['acs_power'] = np.where(action == C.DOWN, power_down, 0.0)
  ch['acs_mass_flow'] = np.where(action == C.UP, flow_up,
                                 np.where(action == C.DOWN, flow_down, 0.0))
  ch['mols_air'] = np.maximum(
      s.mols_air + (ch['acs_mass_flow'] / C.DRY_AIR_MOLAR_MASS) * stride, 0.0)  # :515-519

  is_day = el > C.MIN_SOLAR_EL_DEG                                         # :524
  ch['solar_charging'] = np.where(is_day, solar.solar_power(el, s.pressure), 0.0)
  ch['power_load'] = np.where(is_day, C.DAYTIME_POWER_LOAD_W,
                              C.NIGHTTIME_POWER_LOAD_W) + ch['acs_power']  # :529-531
  charge = s.battery_charge + (ch['solar_charging'] - ch['power_load']) * (stride / 3600.0)
  ch['battery_charge'] = np.minimum(np.maximum(charge, 0.0), C.BATTERY_CAPACITY_WH)  # :535-539
  status = np.where(ch['battery_charge'] <= 0.0, C.STATUS_OUT_OF_POWER, status)  # :541-542
  ch['status'] = status
  ch['date_time'] = s.date_time + stride                                   # :546-547
  ch['time_elapsed'] = s.time_elapsed + stride
  return ch


def effective_action(s: BalloonBatch, atmosphere, action, power_safety_layer_enabled=True):
  """Safety-layer chain, evaluated once per agent step (:304-313).  Mutates layer state.

  power_safety_layer_enabled: bool or bool[N] (BalloonState.power_safety_layer_enabled, :200).
  """
  eff = np.asarray(action, np.int64).copy()
  psl = np.broadcast_to(np.asarray(power_safety_layer_enabled, bool), eff.shape)
  if psl.any():
    e2, sr, ss, paused = safety.power_safety_get_action(
        eff, s.date_time, s.battery_charge, s.sunrise_h, s.sunset, s.power_paused.astype(bool))
    eff = np.where(psl, e2, eff)
    s.sunrise_h = np.where(psl, sr, s.sunrise_h)
    s.sunset = np.where(psl, ss, s.sunset)
    s.power_paused = np.where(psl, paused, s.power_paused.astype(bool)).astype(np.int64)
  eff, s.envelope_state = safety.envelope_safety_get_action(eff, s.superpressure, s.envelope_state)
  altitude, _ = atmosphere.at_pressure(s.pressure)
  eff, s.altitude_state = safety.altitude_safety_get_action(eff, altitude, s.altitude_state)
  return eff


def simulate_step(s: BalloonBatch, u, v, atmosphere, action,
                  time_delta=C.AGENT_TIME_STEP_S, stride=C.PHYSICS_STRIDE_S,
                  power_safety_layer_enabled=True):
  """In-place agent step (:263-328).  Returns the effective action.

  Balloons whose status is not OK are left untouched (the reference asserts, :288; the
  batched contract makes stepping a finished balloon a no-op).
  """
  action = np.broadcast_to(np.asarray(action, np.int64), (s.n,))
  u = np.broadcast_to(np.asarray(u, np.float64), (s.n,))
  v = np.broadcast_to(np.asarray(v, np.float64), (s.n,))
  live = np.nonzero(s.status == C.STATUS_OK)[0]
  eff_all = np.full(s.n, C.STAY, np.int64)
  if live.size == 0:
    return eff_all
  sub = s.select(live)
  atm = atmosphere.subset(live)
  sub.last_command = action[live].copy()                                   # :286
  psl = np.broadcast_to(np.asarray(power_safety_layer_enabled, bool), (s.n,))[live]
  eff = effective_action(sub, atm, action[live], psl)
  assert time_delta % stride == 0
  running = np.ones(live.size, bool)
  for _ in range(time_delta // stride):                                    # :321-328
    if not running.any():
      break
    idx = np.nonzero(running)[0]
    part = sub.select(idx)
    ch = simulate_step_internal(part, u[live][idx], v[live][idx], atm.subset(idx), eff[idx], stride)
    for k, val in ch.items():
      getattr(sub, k)[idx] = val
    running[idx] = ch['status'] == C.STATUS_OK
  for f in dataclasses.fields(s):
    getattr(s, f.name)[live] = getattr(sub, f.name)
  eff_all[live] = eff
  return eff_all
