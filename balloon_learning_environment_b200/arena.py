"""`CudaBalloonArena`: the N = 1 adaptor behind the reference's `BalloonArenaInterface`.

The reference lets a caller inject the simulator: `BalloonEnv(arena=...)`
(env/balloon_env.py:113,144-148) with the interface of env/balloon_arena.py:42-120
(reset / step / get_simulator_state / set_simulator_state / get_balloon_state /
set_balloon_state / get_measurements).  This class implements that interface on top of a
one-balloon `BatchedBalloonArena`, returning objects whose attribute names match the reference's
`BalloonState`, `SimulatorState`, `SimulatorObservation` and `WindVector`
(env/balloon/balloon.py:73-250, env/simulator_data.py:25-46, env/wind_field.py:38-51), so the
reference's `BalloonEnv.step`, reward function and info dict run unchanged on top of it.
It is duck-typed (this package never imports the reference); INTEGRATION.md shows the two-line
subclass a maintainer adds to make `isinstance(arena, BalloonArenaInterface)` hold.
"""
import dataclasses
import datetime as dt
import enum
import time
from typing import Any, Callable, Optional

import numpy as np
import torch

from balloon_learning_environment_b200 import _lib
from balloon_learning_environment_b200 import batched_env
from balloon_learning_environment_b200 import units


class AltitudeControlCommand(enum.IntEnum):   # env/balloon/control.py:21-25
  DOWN = 0
  STAY = 1
  UP = 2


class BalloonStatus(enum.Enum):               # env/balloon/balloon.py:66-70
  OK = 0
  OUT_OF_POWER = 1
  BURST = 2
  ZEROPRESSURE = 3


@dataclasses.dataclass
class LatLng:
  """Stand-in for s2sphere.LatLng: `.lat().radians/.degrees`, `.lng()...`."""
  lat_radians: float
  lng_radians: float

  @dataclasses.dataclass
  class _Angle:
    radians: float

    @property
    def degrees(self):
      return float(np.degrees(self.radians))

  def lat(self):
    return LatLng._Angle(self.lat_radians)

  def lng(self):
    return LatLng._Angle(self.lng_radians)


@dataclasses.dataclass
class WindVector:                             # env/wind_field.py:38-51
  u: units.Velocity
  v: units.Velocity


@dataclasses.dataclass
class BalloonState:
  """Host snapshot of one balloon, attribute-compatible with env/balloon/balloon.py:73-250."""
  center_latlng: LatLng
  date_time: dt.datetime
  time_elapsed: dt.timedelta
  x: units.Distance
  y: units.Distance
  pressure: float
  ambient_temperature: float
  mols_lift_gas: float
  mols_air: float
  internal_temperature: float
  envelope_volume: float
  superpressure: float
  acs_power: units.Power
  acs_mass_flow: float
  solar_charging: units.Power
  power_load: units.Power
  battery_charge: units.Energy
  last_command: AltitudeControlCommand
  status: BalloonStatus
  power_safety_layer_enabled: bool
  upwelling_infrared: float
  # derived properties (balloon.py:217-250), evaluated on the device at snapshot time
  latlng: LatLng
  battery_soc: float
  excess_energy: bool
  navigation_is_paused: bool
  pressure_ratio: float
  # safety-layer internals (needed to restore a checkpoint exactly)
  envelope_state: int = 0
  altitude_state: int = 0
  power_paused: bool = False
  sunrise_with_hysteresis: Optional[dt.datetime] = None
  sunset: Optional[dt.datetime] = None
  atmosphere_alpha: float = 0.5
  battery_capacity: units.Energy = units.Energy(3058.56)
  daytime_power_load: units.Power = units.Power(120.4)
  nighttime_power_load: units.Power = units.Power(183.7)


@dataclasses.dataclass
class SimulatorState:                         # env/simulator_data.py:25-34
  balloon_state: BalloonState
  wind_field: Any
  atmosphere: Any


@dataclasses.dataclass
class SimulatorObservation:                   # env/simulator_data.py:38-46
  balloon_observation: BalloonState
  wind_at_balloon: WindVector


class _NullFeatureConstructor:
  """Observes nothing: the reference's hot path with a null feature constructor."""
  observation_space = None

  def __init__(self, arena):
    del arena

  def observe(self, observation):
    del observation

  def get_features(self) -> np.ndarray:
    return np.zeros((0,), np.float32)


class CudaPerciatelliFeatureConstructor:
  """FeatureConstructor interface (env/features.py:106-144) over the device-side implementation.

  The measurement history (WindGP) lives in the handle and is updated by reset()/step(), so
  observe() has nothing left to do; get_features() runs the feature kernels for this balloon.
  """

  def __init__(self, arena: 'CudaBalloonArena'):
    self._arena = arena
    self.observation_space = batched_env.perciatelli_observation_space()

  def observe(self, observation):
    del observation

  def get_features(self) -> np.ndarray:
    return self._arena._arena.features().cpu().numpy()[0]


def _utc(ts: int) -> dt.datetime:
  return dt.datetime.fromtimestamp(int(ts), tz=dt.timezone.utc)


class CudaBalloonArena:
  """One balloon flying on the GPU, behind the reference's arena interface."""

  def __init__(self, feature_constructor_factory: Optional[Callable[[Any], Any]] = None,
               wind_field: Optional[np.ndarray] = None, seed: Optional[int] = None, *,
               device: str = 'cuda:0', precision: str = 'fp32', wind_model: str = 'grid',
               enable_noise: bool = True, observation: Optional[str] = 'perciatelli'):
    """wind_field: float32 [21,21,10,9,2] grid (GridWindFieldSampler.sample_field layout) for the
    'grid' model; `wind_model='simple_static'` reproduces SimpleStaticWindField.
    observation='perciatelli' (default, as in the reference) makes reset()/step() return the
    1099-feature observation computed on the device; None returns an empty vector.
    feature_constructor_factory(arena) may supply any object with observe/get_features/
    observation_space instead."""
    self._arena = batched_env.BatchedBalloonArena(1, device=device, precision=precision,
                                                  wind_model=wind_model, enable_noise=enable_noise,
                                                  enable_features=observation == 'perciatelli')
    default_factory = CudaPerciatelliFeatureConstructor if observation == 'perciatelli' else _NullFeatureConstructor
    self._factory = feature_constructor_factory or default_factory
    if wind_model == 'grid':
      if wind_field is None:
        raise ValueError("wind_model='grid' needs a [21,21,10,9,2] wind_field")
      self.set_wind_field(wind_field)
    self.feature_constructor = None
    self.reset(seed)

  # -- BalloonArenaInterface ---------------------------------------------------------------------
  def reset(self, seed: Optional[int] = None) -> np.ndarray:
    if seed is None:
      seed = int(time.time() * 1e6)                          # env/balloon_arena.py:168-169
    if isinstance(seed, np.ndarray):                          # a jax PRNG key (uint32[2])
      seed = int(np.asarray(seed, np.uint64).ravel()[-1]) + (int(np.asarray(seed, np.uint64).ravel()[0]) << 32)
    self._arena.reset(torch.tensor([int(seed) & (2**63 - 1)], dtype=torch.int64))
    self.feature_constructor = self._factory(self)
    self.feature_constructor.observe(self.get_measurements())
    return self.feature_constructor.get_features()

  def step(self, action) -> np.ndarray:
    self._arena.step(torch.tensor([int(action)], dtype=torch.int32))
    self.feature_constructor.observe(self.get_measurements())
    return self.feature_constructor.get_features()

  def get_simulator_state(self) -> SimulatorState:
    state = self.get_balloon_state()
    return SimulatorState(state, self, state.atmosphere_alpha)

  def set_simulator_state(self, new_state: SimulatorState) -> None:
    self.set_balloon_state(new_state.balloon_state)

  def get_balloon_state(self) -> BalloonState:
    f, i = self._arena.get_state()
    d = self._arena.get_derived()
    f = {k: float(f[r, 0]) for r, k in enumerate(_lib.F_ROWS)}
    i = {k: int(i[r, 0]) for r, k in enumerate(_lib.I_ROWS)}
    d = {k: float(v[0]) for k, v in d.items()}
    return BalloonState(
        center_latlng=LatLng(f['center_lat'], f['center_lng']), date_time=_utc(i['date_time']),
        time_elapsed=dt.timedelta(seconds=i['time_elapsed']), x=units.Distance(f['x']), y=units.Distance(f['y']),
        pressure=f['pressure'], ambient_temperature=f['ambient_temperature'], mols_lift_gas=f['mols_lift_gas'],
        mols_air=f['mols_air'], internal_temperature=f['internal_temperature'],
        envelope_volume=f['envelope_volume'], superpressure=f['superpressure'],
        acs_power=units.Power(f['acs_power']), acs_mass_flow=f['acs_mass_flow'],
        solar_charging=units.Power(f['solar_charging']), power_load=units.Power(f['power_load']),
        battery_charge=units.Energy(f['battery_charge']), last_command=AltitudeControlCommand(i['last_command']),
        status=BalloonStatus(i['status']), power_safety_layer_enabled=bool(i['power_safety_enabled']),
        upwelling_infrared=f['upwelling_infrared'], latlng=LatLng(d['lat'], d['lng']),
        battery_soc=d['battery_soc'], excess_energy=bool(d['excess_energy']),
        navigation_is_paused=bool(d['navigation_is_paused']), pressure_ratio=d['pressure_ratio'],
        envelope_state=i['envelope_state'], altitude_state=i['altitude_state'],
        power_paused=bool(i['power_paused']), sunrise_with_hysteresis=_utc(i['sunrise_h']),
        sunset=_utc(i['sunset']), atmosphere_alpha=f['atmosphere_alpha'])

  def set_balloon_state(self, s: BalloonState) -> None:
    f = {'x': s.x.m, 'y': s.y.m, 'pressure': s.pressure, 'ambient_temperature': s.ambient_temperature,
         'internal_temperature': s.internal_temperature, 'envelope_volume': s.envelope_volume,
         'superpressure': s.superpressure, 'mols_air': s.mols_air, 'mols_lift_gas': s.mols_lift_gas,
         'battery_charge': s.battery_charge.watt_hours, 'acs_power': s.acs_power.watts,
         'acs_mass_flow': s.acs_mass_flow, 'solar_charging': s.solar_charging.watts,
         'power_load': s.power_load.watts, 'center_lat': s.center_latlng.lat().radians,
         'center_lng': s.center_latlng.lng().radians, 'upwelling_infrared': s.upwelling_infrared,
         'atmosphere_alpha': s.atmosphere_alpha}
    i = {'date_time': int(s.date_time.timestamp()), 'time_elapsed': int(s.time_elapsed.total_seconds()),
         'last_command': int(s.last_command), 'status': int(s.status.value),
         'envelope_state': s.envelope_state, 'altitude_state': s.altitude_state,
         'power_paused': int(s.power_paused), 'sunrise_h': int(s.sunrise_with_hysteresis.timestamp()),
         'sunset': int(s.sunset.timestamp()), 'power_safety_enabled': int(s.power_safety_layer_enabled)}
    fm = torch.tensor([[f[k]] for k in _lib.F_ROWS], dtype=torch.float64)
    im = torch.tensor([[i[k]] for k in _lib.I_ROWS], dtype=torch.int64)
    self._arena.set_state(fm, im)

  def get_measurements(self) -> SimulatorObservation:
    uv = self._arena.wind_at_balloon().cpu().numpy()
    return SimulatorObservation(self.get_balloon_state(),
                                WindVector(units.Velocity(float(uv[0, 0])), units.Velocity(float(uv[0, 1]))))

  # -- extras --------------------------------------------------------------------------------------
  def set_wind_field(self, field: np.ndarray) -> None:
    field = torch.as_tensor(np.asarray(field, np.float32)).reshape(1, *batched_env.FIELD_SHAPE)
    self._arena.set_wind_fields(field)

  def set_wind_noise(self, seeds: np.ndarray, offsets: np.ndarray) -> None:
    self._arena.set_wind_noise(torch.as_tensor(np.asarray(seeds, np.int64)).reshape(1, 2, 5),
                               torch.as_tensor(np.asarray(offsets, np.float32)).reshape(1, 2, 5, 4))

  def reset_feature_history(self) -> None:
    """A fresh FeatureConstructor that has observed only the current state (what BalloonArena.reset
    does at env/balloon_arena.py:179-182); use after set_balloon_state when starting a new episode."""
    if self._arena.enable_features:
      self._arena.features_clear()
      self._arena.features_observe()

  def close(self):
    self._arena.close()
