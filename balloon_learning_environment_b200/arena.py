"""`CudaBalloonArena`: the N = 1 adaptor behind the reference's `BalloonArenaInterface`.

The reference lets a caller inject the simulator: `BalloonEnv(arena=...)`
(env/balloon_env.py:113,144-148) with the interface of env/balloon_arena.py:42-120
(reset / step / get_simulator_state / set_simulator_state / get_balloon_state /
set_balloon_state / get_measurements).  This class implements that interface on top of a
one-balloon `BatchedBalloonArena`, returning objects whose attribute names match the reference's
`BalloonState`, `SimulatorState`, `SimulatorObservation` and `WindVector`
(env/balloon/balloon.py:73-250, env/simulator_data.py:25-46, env/wind_field.py:38-51), so the
reference's `BalloonEnv.step`, reward function and info dict run unchanged on top of it.

Types.  When the reference package is importable the adaptor returns the REFERENCE'S OWN value types --
`units.Distance/Power/Energy/Velocity`, `control.AltitudeControlCommand`, `balloon.BalloonStatus`,
`wind_field.WindVector`, `standard_atmosphere.AtmosphericValues`, `simulator_data.SimulatorState /
SimulatorObservation` -- so that the reference's comparisons (`status == BalloonStatus.BURST`, env/balloon_env.py:281-283)
and arithmetic (`units.relative_distance`, :80) work on them; otherwise it uses the attribute-compatible stand-ins
defined here.  `SimulatorState.wind_field` is a `CudaWindField` (get_forecast / get_forecast_column /
get_ground_truth, env/wind_field.py:69-145) and `.atmosphere` a `CudaAtmosphere` (at_pressure / at_height,
env/balloon/standard_atmosphere.py:89-154), both evaluated on the device for this balloon.
INTEGRATION.md shows the two-line subclass a maintainer adds to make `isinstance(arena, BalloonArenaInterface)` hold.
"""
import dataclasses
import datetime as dt
import enum
import inspect
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

  @property
  def is_valid(self):                         # read by env/balloon/solar.py:60
    return abs(self.lat_radians) <= np.pi / 2 and abs(self.lng_radians) <= np.pi

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
  # flight-vehicle constants (balloon.py:155-172), read by env/balloon/pressure_range_builder.py:134-245
  envelope_volume_base: float = 1804.0
  envelope_volume_dv_pressure: float = 0.0199
  envelope_mass: float = 68.5
  envelope_max_superpressure: float = 2380.0
  envelope_cod: float = 0.25
  payload_mass: float = 92.5
  acs_valve_hole_diameter: units.Distance = units.Distance(0.04)


@dataclasses.dataclass
class SimulatorState:                         # env/simulator_data.py:25-34
  balloon_state: BalloonState
  wind_field: Any
  atmosphere: Any


@dataclasses.dataclass
class SimulatorObservation:                   # env/simulator_data.py:38-46
  balloon_observation: BalloonState
  wind_at_balloon: WindVector


@dataclasses.dataclass
class AtmosphericValues:                      # env/balloon/standard_atmosphere.py:48-54
  height: units.Distance
  temperature: float
  pressure: float
  density: float


class _Types:
  """The value types that cross the boundary: the reference's own when it is importable, else the stand-ins above."""

  def __init__(self):
    self.reference = False
    self.Distance = lambda m: units.Distance(m)
    self.Velocity = lambda mps: units.Velocity(mps)
    self.Power = lambda watts: units.Power(watts)
    self.Energy = lambda wh: units.Energy(wh)
    self.Command, self.Status = AltitudeControlCommand, BalloonStatus
    self.WindVector, self.AtmosphericValues = WindVector, AtmosphericValues
    self.SimulatorState, self.SimulatorObservation = SimulatorState, SimulatorObservation
    self.LatLng = LatLng
    try:
      import s2sphere
      from balloon_learning_environment.env import simulator_data as ref_sim
      from balloon_learning_environment.env import wind_field as ref_wind
      from balloon_learning_environment.env.balloon import balloon as ref_balloon
      from balloon_learning_environment.env.balloon import control as ref_control
      from balloon_learning_environment.env.balloon import standard_atmosphere as ref_atm
      from balloon_learning_environment.utils import units as ref_units
    except Exception:  # pylint: disable=broad-except
      return
    self.reference = True
    self.Distance = lambda m: ref_units.Distance(m=m)
    self.Velocity = lambda mps: ref_units.Velocity(mps=mps)
    self.Power = lambda watts: ref_units.Power(watts=watts)
    self.Energy = lambda wh: ref_units.Energy(watt_hours=wh)
    self.Command, self.Status = ref_control.AltitudeControlCommand, ref_balloon.BalloonStatus
    self.WindVector, self.AtmosphericValues = ref_wind.WindVector, ref_atm.AtmosphericValues
    self.SimulatorState, self.SimulatorObservation = ref_sim.SimulatorState, ref_sim.SimulatorObservation
    self.LatLng = s2sphere.LatLng.from_radians


_TYPES = None


def types() -> _Types:
  global _TYPES
  if _TYPES is None:
    _TYPES = _Types()
  return _TYPES


class CudaWindField:
  """`SimulatorState.wind_field` of a CudaBalloonArena: the WindField queries of env/wind_field.py:55-145 for THIS
  balloon's wind grid and noise generators, evaluated by ble_wind_query."""

  def __init__(self, backend, env_index: int = 0):
    self._backend, self._e = backend, int(env_index)

  def _query(self, x, y, pressures, elapsed_time, with_noise):
    t = elapsed_time.total_seconds()
    pts = torch.tensor([[x.m, y.m, float(p), t] for p in pressures], dtype=torch.float64)
    idx = torch.full((len(pressures),), self._e, dtype=torch.int32)
    uv = self._backend.wind_query(pts, idx, with_noise).cpu().numpy()
    ty = types()
    return [ty.WindVector(ty.Velocity(float(u)), ty.Velocity(float(v))) for u, v in uv]

  def get_forecast(self, x, y, pressure: float, elapsed_time: dt.timedelta):
    return self._query(x, y, [pressure], elapsed_time, False)[0]

  def get_forecast_column(self, x, y, pressures, elapsed_time: dt.timedelta):
    return self._query(x, y, list(pressures), elapsed_time, False)

  def get_ground_truth(self, x, y, pressure: float, elapsed_time: dt.timedelta):
    return self._query(x, y, [pressure], elapsed_time, True)[0]


class CudaAtmosphere:
  """`SimulatorState.atmosphere`: Atmosphere.at_pressure / at_height (env/balloon/standard_atmosphere.py:89-154) of this
  balloon's atmosphere, evaluated by ble_atmosphere_query.  Out-of-range queries raise AssertionError as the reference."""

  def __init__(self, backend, env_index: int = 0):
    self._backend, self._e = backend, int(env_index)

  def _query(self, which, value):
    out = self._backend.atmosphere_query(which, torch.tensor([float(value)], dtype=torch.float64),
                                         torch.tensor([self._e], dtype=torch.int32)).cpu().numpy()[0]
    assert not np.isnan(out[0]), f'atmosphere query outside the table: {which} = {value}'
    ty = types()
    return ty.AtmosphericValues(ty.Distance(float(out[0])), float(out[1]), float(out[2]), float(out[3]))

  def at_pressure(self, pressure: float):
    return self._query('pressure', pressure)

  def at_height(self, height):
    return self._query('height', height.m)


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
    self.feature_constructor = self._make_feature_constructor()
    self.feature_constructor.observe(self.get_measurements())
    return self.feature_constructor.get_features()

  def _make_feature_constructor(self):
    """The reference's factories take (forecast: WindField, atmosphere: Atmosphere) (env/balloon_arena.py:179-181,
    env/features.py:122-124); the device-side ones here take the arena.  Told apart by arity."""
    try:
      params = [p for p in inspect.signature(self._factory).parameters.values()
                if p.default is p.empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    except (TypeError, ValueError):
      params = [None]
    if len(params) == 2:
      return self._factory(CudaWindField(self._arena), CudaAtmosphere(self._arena))
    return self._factory(self)

  def step(self, action) -> np.ndarray:
    self._arena.step(torch.tensor([int(action)], dtype=torch.int32))
    self.feature_constructor.observe(self.get_measurements())
    return self.feature_constructor.get_features()

  def get_simulator_state(self):
    return types().SimulatorState(self.get_balloon_state(), CudaWindField(self._arena), CudaAtmosphere(self._arena))

  def set_simulator_state(self, new_state: SimulatorState) -> None:
    self.set_balloon_state(new_state.balloon_state)

  def get_balloon_state(self) -> BalloonState:
    f, i = self._arena.get_state()
    d = self._arena.get_derived()
    f = {k: float(f[r, 0]) for r, k in enumerate(_lib.F_ROWS)}
    i = {k: int(i[r, 0]) for r, k in enumerate(_lib.I_ROWS)}
    d = {k: float(v[0]) for k, v in d.items()}
    ty = types()
    return BalloonState(
        center_latlng=ty.LatLng(f['center_lat'], f['center_lng']), date_time=_utc(i['date_time']),
        time_elapsed=dt.timedelta(seconds=i['time_elapsed']), x=ty.Distance(f['x']), y=ty.Distance(f['y']),
        pressure=f['pressure'], ambient_temperature=f['ambient_temperature'], mols_lift_gas=f['mols_lift_gas'],
        mols_air=f['mols_air'], internal_temperature=f['internal_temperature'],
        envelope_volume=f['envelope_volume'], superpressure=f['superpressure'],
        acs_power=ty.Power(f['acs_power']), acs_mass_flow=f['acs_mass_flow'],
        solar_charging=ty.Power(f['solar_charging']), power_load=ty.Power(f['power_load']),
        battery_charge=ty.Energy(f['battery_charge']), last_command=ty.Command(i['last_command']),
        status=ty.Status(i['status']), power_safety_layer_enabled=bool(i['power_safety_enabled']),
        upwelling_infrared=f['upwelling_infrared'], latlng=ty.LatLng(d['lat'], d['lng']),
        battery_soc=d['battery_soc'], excess_energy=bool(d['excess_energy']),
        navigation_is_paused=bool(d['navigation_is_paused']), pressure_ratio=d['pressure_ratio'],
        envelope_state=i['envelope_state'], altitude_state=i['altitude_state'],
        power_paused=bool(i['power_paused']), sunrise_with_hysteresis=_utc(i['sunrise_h']),
        sunset=_utc(i['sunset']), atmosphere_alpha=f['atmosphere_alpha'],
        battery_capacity=ty.Energy(3058.56), daytime_power_load=ty.Power(120.4), nighttime_power_load=ty.Power(183.7),
        acs_valve_hole_diameter=ty.Distance(0.04))

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
    ty = types()
    return ty.SimulatorObservation(self.get_balloon_state(),
                                   ty.WindVector(ty.Velocity(float(uv[0, 0])), ty.Velocity(float(uv[0, 1]))))

  # -- extras --------------------------------------------------------------------------------------
  def get_info(self):
    """BalloonEnv._get_info (env/balloon_env.py:280-290) for the step that just ran, as written by the step kernel
    (no state download): out_of_power / envelope_burst / zeropressure bool, time_elapsed timedelta."""
    info = self._arena.step_info()
    return {'out_of_power': bool(info['out_of_power'][0]), 'envelope_burst': bool(info['envelope_burst'][0]),
            'zeropressure': bool(info['zeropressure'][0]),
            'time_elapsed': dt.timedelta(seconds=int(info['time_elapsed'][0]))}

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
