"""Dumps golden vectors by running the UNMODIFIED reference (this container only).

    python -m tests.golden.tier0.make_golden

Writes tests/golden/kat.json (per-function known answers) and tests/golden/traj.npz
(multi-step trajectories through the reference's own BalloonEnv.step with injected
initial state, wind field and noise parameters).  The reference is imported from
/root/reference under the Tier-0 stubs (tests/golden/tier0/boot.py); nothing from it is
copied.  The committed outputs are what travels to the GPU box.
"""
import datetime as dt
import json
import os

import tests.golden.tier0.boot as boot  # noqa: F401  (first: makes the reference importable)

import jax
import numpy as np
import s2sphere as s2
from balloon_learning_environment.env import balloon_arena
from balloon_learning_environment.env import balloon_env
from balloon_learning_environment.env import grid_based_wind_field
from balloon_learning_environment.env import grid_wind_field_sampler
from balloon_learning_environment.env import simplex_wind_noise
from balloon_learning_environment.env import simulator_data
from balloon_learning_environment.env import wind_field
from balloon_learning_environment.env.balloon import acs
from balloon_learning_environment.env.balloon import balloon
from balloon_learning_environment.env.balloon import control
from balloon_learning_environment.env.balloon import solar
from balloon_learning_environment.env.balloon import stable_init
from balloon_learning_environment.env.balloon import standard_atmosphere
from balloon_learning_environment.env.balloon import thermal
from balloon_learning_environment.generative import vae
from balloon_learning_environment.utils import spherical_geometry
from balloon_learning_environment.utils import test_helpers
from balloon_learning_environment.utils import units
import opensimplex

from tests.golden import fields as golden_fields

OUT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UTC = dt.timezone.utc


def ts_of(d: dt.datetime) -> int:
  return int(d.timestamp())


def make_atmosphere(alpha: float) -> standard_atmosphere.Atmosphere:
  atm = standard_atmosphere.Atmosphere(jax.random.PRNGKey(0))
  atm._lapse_rates = ((1 - alpha) * atm._LAPSE_RATES_LOW + alpha * atm._LAPSE_RATES_HIGH)
  atm._initialize_temperature_transitions()
  atm._initialize_pressure_transitions()
  return atm


# ----------------------------------------------------------------------------- KATs

def kat_atmosphere():
  out = []
  for alpha in (0.0, 0.11815023, 0.41845703, 0.5, 0.99589407, 1.0):
    atm = make_atmosphere(alpha)
    rows = []
    for p in (14000.0, 12027.7, 11000.0, 9380.0, 9000.0, 8806.4, 8806.3, 8229.0, 7000.0,
              6000.0, 5000.0, 4400.0, 3000.0, 700.0, 80.0, 40.0, 2.0):
      v = atm.at_pressure(p)
      rows.append([p, v.height.meters, v.temperature, v.density])
    hrows = []
    for h in (0.0, 15240.0, 16999.0, 17000.0, 18000.0, 20999.9, 25000.0, 40000.0, 48000.0, 60000.0):
      v = atm.at_height(units.Distance(m=h))
      hrows.append([h, v.pressure, v.temperature, v.density])
    out.append(dict(alpha=alpha, lapse=list(map(float, atm._lapse_rates)),
                    t_tr=list(map(float, atm._temperature_transitions)),
                    p_tr=list(map(float, atm._pressure_transitions)),
                    at_pressure=rows, at_height=hrows))
  return out


def kat_solar():
  rng = np.random.default_rng(7)
  rows = []
  t0 = ts_of(units.datetime(2011, 1, 1)); t1 = ts_of(units.datetime(2014, 12, 31))
  cases = [(0.0, 0.0, ts_of(test_helpers.START_DATE_TIME)),
           (0.0, 0.0, ts_of(units.datetime(2013, 9, 21, 18, 0, 0))),
           (9.99, -174.9, ts_of(units.datetime(2012, 2, 29, 23, 59, 59))),
           (-10.0, 175.0, ts_of(units.datetime(2012, 3, 1, 0, 0, 0))),
           (45.0, 10.0, ts_of(units.datetime(2020, 6, 21, 11, 20, 0)))]
  for _ in range(120):
    cases.append((float(rng.uniform(-12, 12)), float(rng.uniform(-180, 180)),
                  int(rng.integers(t0, t1))))
  for lat, lng, ts in cases:
    ll = s2.LatLng.from_degrees(lat, lng)
    d = units.datetime_from_timestamp(ts)
    el, az, flux = solar.solar_calculator(ll, d)
    rows.append([ll.lat().radians, ll.lng().radians, ts, float(el), float(az), float(flux)])
  power = []
  for el in (-90.0, -10.0, -4.243, -4.242, -4.0, 0.0, 5.0, 20.0, 29.9, 35.0, 42.0, 60.0, 85.0, 90.0):
    for p in (5000.0, 9000.0, 14000.0, 101325.0):
      power.append([el, p, float(solar.solar_atmospheric_attenuation(el, p)),
                    float(solar.solar_power(el, p).watts)])
  sun = []
  for lat, lng, ts in cases[:40]:
    ll = s2.LatLng.from_degrees(lat, lng)
    d = units.datetime_from_timestamp(ts)
    sr, ss = solar.get_next_sunrise_sunset(ll, d)
    sun.append([ll.lat().radians, ll.lng().radians, ts, ts_of(sr), ts_of(ss)])
  return dict(calculator=rows, power=power, sunrise_sunset=sun)


def kat_thermal_acs_geometry():
  rng = np.random.default_rng(11)
  th = []
  for _ in range(60):
    a = dict(balloon_volume=float(rng.uniform(900, 1900)), balloon_mass=68.5,
             balloon_temperature_k=float(rng.uniform(180, 260)),
             ambient_temperature_k=float(rng.uniform(185, 230)),
             pressure_altitude_pa=float(rng.uniform(4000, 15000)),
             solar_elevation_deg=float(rng.uniform(-60, 89)),
             solar_flux=float(rng.uniform(1320, 1420)), earth_flux=float(rng.uniform(225, 315)))
    th.append(list(a.values()) + [float(thermal.d_balloon_temperature_dt(**a))])
  ac = []
  for pr in list(np.linspace(0.98, 1.42, 45)) + [1.05, 1.2, 1.25, 1.35, 1.0750000001]:
    w = acs.get_most_efficient_power(float(pr))
    eff = acs.get_fan_efficiency(float(pr), w)
    ac.append([float(pr), float(w.watts), float(eff), float(acs.get_mass_flow(w, eff))])
  eff_grid = []
  for pr in (1.0, 1.05, 1.06, 1.19, 1.3, 1.35, 1.4):
    for w in (50.0, 100.0, 150.0, 299.0, 300.0, 400.0, 500.0):
      eff_grid.append([pr, w, float(acs.get_fan_efficiency(pr, units.Power(watts=w)))])
  geo = []
  for _ in range(60):
    lat, lng = float(rng.uniform(-12, 12)), float(rng.uniform(-179.9, 179.9))
    x, y = float(rng.normal(0, 3e5)), float(rng.normal(0, 3e5))
    c = s2.LatLng.from_degrees(lat, lng)
    o = spherical_geometry.calculate_latlng_from_offset(c, units.Distance(m=x), units.Distance(m=y))
    geo.append([c.lat().radians, c.lng().radians, x, y, o.lat().radians, o.lng().radians])
  c = s2.LatLng.from_degrees(0.0, 0.0)
  o = spherical_geometry.calculate_latlng_from_offset(c, units.Distance(m=0.0), units.Distance(m=0.0))
  geo.append([0.0, 0.0, 0.0, 0.0, o.lat().radians, o.lng().radians])
  sp = []
  for _ in range(60):
    a = (6830.0, float(rng.uniform(0, 9000)), float(rng.uniform(180, 260)), float(rng.uniform(4500, 14500)))
    vol, s = balloon.calculate_superpressure_and_volume(*a, 1804.0, 0.0199)
    sp.append(list(a) + [float(vol), float(s)])
  return dict(thermal=th, acs=ac, eff_grid=eff_grid, geometry=geo, superpressure=sp)


def kat_stable_init():
  rng = np.random.default_rng(13)
  rows = []
  for _ in range(40):
    alpha = float(rng.uniform(0, 1))
    atm = make_atmosphere(alpha)
    p = float(rng.uniform(6500, 12000))
    lat, lng = float(rng.uniform(-10, 10)), float(rng.uniform(-175, 175))
    ts = int(rng.integers(ts_of(units.datetime(2011, 1, 1)), ts_of(units.datetime(2014, 12, 31))))
    ir = float(rng.uniform(225, 315))
    sp = stable_init.calculate_stable_params_for_pressure(
        p, 1804.0, 0.0199, 68.5, 92.5, 6830.0, s2.LatLng.from_degrees(lat, lng),
        units.datetime_from_timestamp(ts), ir, atm)
    ll = s2.LatLng.from_degrees(lat, lng)
    rows.append([alpha, p, ll.lat().radians, ll.lng().radians, ts, ir, float(sp.ambient_temperature),
                 float(sp.internal_temperature), float(sp.mols_air), float(sp.envelope_volume),
                 float(sp.superpressure)])
  return rows


def kat_safety():
  """Random input sequences through the reference's three safety layers (state machines)."""
  from balloon_learning_environment.env.balloon import altitude_safety, envelope_safety, power_safety
  rng = np.random.default_rng(23)
  env_rows = []
  layer = envelope_safety.EnvelopeSafetyLayer(2380.0)
  sp = 1000.0
  for _ in range(600):
    sp = float(np.clip(sp + rng.normal(0, 120), -50, 2500))
    if rng.uniform() < 0.05:
      sp = float(rng.choice([149.9, 150.0, 249.9, 250.0, 299.9, 300.0, 2079.9, 2080.0, 2129.9,
                             2130.0, 2229.9, 2230.0]))
    a = int(rng.integers(0, 3))
    out = layer.get_action(control.AltitudeControlCommand(a), sp)
    env_rows.append([sp, a, int(out), _ENV_STATE[layer._state_machine.state.name],
                     int(layer.navigation_is_paused)])
  alt_rows = []
  alpha = 0.37
  atm = make_atmosphere(alpha)
  layer = altitude_safety.AltitudeSafetyLayer()
  h = 15800.0
  for _ in range(600):
    h = float(np.clip(h + rng.normal(0, 60), 15000, 16500))
    p = float(atm.at_height(units.Distance(m=h)).pressure)
    a = int(rng.integers(0, 3))
    out = layer.get_action(control.AltitudeControlCommand(a), atm, p)
    alt_rows.append([p, a, int(out), _ALT_STATE[layer._state_machine.state.name]])
  pow_rows = []
  lat, lng, ts0 = 3.0, 77.0, ts_of(units.datetime(2012, 6, 1, 14, 0, 0))
  layer = power_safety.PowerSafetyLayer(s2.LatLng.from_degrees(lat, lng),
                                        units.datetime_from_timestamp(ts0))
  init = [ts_of(layer._sunrise_with_hysteresis), ts_of(layer._sunset)]
  ts, charge = ts0, 1500.0
  for _ in range(900):
    ts += 180 * int(rng.integers(1, 4))
    charge = float(np.clip(charge + rng.normal(-4, 25), 0.0, 3058.56))
    if rng.uniform() < 0.03:
      charge = float(rng.choice([60.0, 100.0, 152.0, 153.5, 400.0, 2900.0]))
    a = int(rng.integers(0, 3))
    out = layer.get_action(control.AltitudeControlCommand(a), units.datetime_from_timestamp(ts),
                           units.Power(watts=183.7), units.Energy(watt_hours=charge),
                           units.Energy(watt_hours=3058.56))
    pow_rows.append([ts, charge, a, int(out), int(layer.navigation_is_paused),
                     ts_of(layer._sunrise_with_hysteresis), ts_of(layer._sunset)])
  return dict(envelope=env_rows, altitude=dict(alpha=alpha, rows=alt_rows),
              power=dict(init=init, rows=pow_rows))


class _BankSampler(grid_wind_field_sampler.GridWindFieldSampler):
  def __init__(self, field):
    self._field = field
  @property
  def field_shape(self):
    return vae.FieldShape()
  def sample_field(self, key, date_time):
    return self._field


def kat_interp():
  bank = golden_fields.field_bank()
  rng = np.random.default_rng(17)
  rows = []
  for f in range(bank.shape[0]):
    wf = grid_based_wind_field.GridBasedWindField(_BankSampler(bank[f]))
    wf.reset_forecast(None, None)
    pts = [(123.4e3, -321.0e3, 8765.4, 13.7 * 3600), (600e3, 0.0, 4000.0, 50 * 3600),
           (12.5e3, 487.5e3, 13999.0, 100 * 3600), (0.0, 0.0, 9500.0, 172764),
           (-500e3, 500e3, 5000.0, 0), (500e3, -500e3, 14000.0, 48 * 3600),
           (0.0, 0.0, 9000.0, 96 * 3600), (-1e6, 1e6, 20000.0, 144 * 3600 + 10)]
    for _ in range(40):
      pts.append((float(rng.uniform(-6e5, 6e5)), float(rng.uniform(-6e5, 6e5)),
                  float(rng.uniform(4000, 15000)), int(rng.integers(0, 8 * 86400 // 10)) * 10))
    for (x, y, p, t) in pts:
      w = wf.get_forecast(units.Distance(m=x), units.Distance(m=y), p, dt.timedelta(seconds=int(t)))
      rows.append([f, x, y, p, int(t), float(w.u.mps), float(w.v.mps)])
  return rows


# ----------------------------------------------------------------------------- trajectories

class _NullFeatures:
  """A FeatureConstructor that observes nothing (keeps BalloonEnv.step on the hot path only)."""

  def __init__(self, forecast, atmosphere):
    del forecast, atmosphere
  def observe(self, observation):
    pass
  def get_features(self):
    return np.zeros(1, np.float32)
  @property
  def observation_space(self):
    return None


def _inject_noise(wf, seeds, offsets):
  for c, comp in enumerate((wf._noise_model.noise_u, wf._noise_model.noise_v)):
    for h, harm in enumerate(comp._harmonics):
      harm._simplex_generator = opensimplex.OpenSimplex(seed=int(seeds[c, h]))
      harm._offsets = simplex_wind_noise.SimplexOffset(*[float(o) for o in offsets[c, h]])


_ENV_STATE = {s: i for i, s in enumerate(['NOMINAL', 'LOW_CRITICAL', 'LOW', 'HIGH', 'HIGH_CRITICAL'])}
_ALT_STATE = {s: i for i, s in enumerate(['NOMINAL', 'LOW', 'VERY_LOW'])}

FLOATS = ('x', 'y', 'pressure', 'ambient_temperature', 'internal_temperature', 'envelope_volume',
          'superpressure', 'mols_air', 'mols_lift_gas', 'battery_charge', 'acs_power',
          'acs_mass_flow', 'solar_charging', 'power_load', 'center_lat', 'center_lng',
          'upwelling_infrared')
INTS = ('date_time', 'time_elapsed', 'last_command', 'status', 'envelope_state', 'altitude_state',
        'power_paused', 'sunrise_h', 'sunset')


def snapshot(st: balloon.BalloonState):
  g = lambda v: float(getattr(v, 'watts', getattr(v, 'watt_hours', getattr(v, 'm', v))))
  f = dict(x=st.x.m, y=st.y.m, pressure=st.pressure, ambient_temperature=st.ambient_temperature,
           internal_temperature=st.internal_temperature, envelope_volume=st.envelope_volume,
           superpressure=st.superpressure, mols_air=st.mols_air, mols_lift_gas=st.mols_lift_gas,
           battery_charge=st.battery_charge.watt_hours, acs_power=g(st.acs_power),
           acs_mass_flow=st.acs_mass_flow, solar_charging=g(st.solar_charging),
           power_load=g(st.power_load), center_lat=st.center_latlng.lat().radians,
           center_lng=st.center_latlng.lng().radians, upwelling_infrared=st.upwelling_infrared)
  psl = st.power_safety_layer
  i = dict(date_time=ts_of(st.date_time), time_elapsed=int(st.time_elapsed.total_seconds()),
           last_command=int(st.last_command), status=int(st.status.value),
           envelope_state=_ENV_STATE[st.envelope_safety_layer._state_machine.state.name],
           altitude_state=_ALT_STATE[st.altitude_safety_layer._state_machine.state.name],
           power_paused=int(psl.navigation_is_paused),
           sunrise_h=ts_of(psl._sunrise_with_hysteresis), sunset=ts_of(psl._sunset))
  return ([float(f[k]) for k in FLOATS], [int(i[k]) for k in INTS])


SCENARIOS = [
    # name, dict(...)
    dict(name='default_static', alpha=0.5, lat=0.0, lng=0.0, date=(2013, 3, 25, 9, 25, 32),
         pressure=9000.0, x=0.0, y=0.0, ir=250.0, power=0.95, field=-1, policy='random', steps=960),
    dict(name='grid_random', alpha=0.11815023, lat=5.3, lng=120.7, date=(2011, 7, 4, 18, 0, 0),
         pressure=7500.0, x=50e3, y=-80e3, ir=300.0, power=0.95, field=0, policy='random', steps=960),
    dict(name='grid_down_night_lowbatt', alpha=0.9, lat=-8.0, lng=-60.0, date=(2012, 12, 1, 23, 30, 0),
         pressure=8200.0, x=-120e3, y=30e3, ir=240.0, power=0.30, field=1, policy='down', steps=600),
    dict(name='grid_up', alpha=0.3, lat=2.0, lng=10.0, date=(2014, 5, 5, 5, 0, 0),
         pressure=11000.0, x=10e3, y=10e3, ir=280.0, power=0.6, field=2, policy='up', steps=400),
    dict(name='grid_down', alpha=0.7, lat=-3.0, lng=170.0, date=(2013, 10, 10, 12, 0, 0),
         pressure=7000.0, x=0.0, y=0.0, ir=260.0, power=0.99, field=3, policy='down', steps=600),
    dict(name='nopsl_down_outofpower', alpha=0.5, lat=0.0, lng=30.0, date=(2013, 1, 10, 17, 0, 0),
         pressure=9000.0, x=0.0, y=0.0, ir=250.0, power=0.12, field=0, policy='down', steps=400,
         power_safety=False),
    dict(name='sticky_random', alpha=0.41845703, lat=9.5, lng=-174.0, date=(2011, 1, 1, 0, 0, 7),
         pressure=10500.0, x=190e3, y=20e3, ir=226.0, power=0.5, field=1, policy='sticky', steps=960),
    dict(name='zeropressure', alpha=0.5, lat=0.0, lng=0.0, date=(2013, 3, 25, 9, 25, 32),
         pressure=13500.0, x=0.0, y=0.0, ir=250.0, power=0.95, field=0, policy='random', steps=3,
         stable=False),
    dict(name='burst', alpha=0.5, lat=0.0, lng=0.0, date=(2013, 3, 25, 9, 25, 32),
         pressure=6000.0, x=0.0, y=0.0, ir=250.0, power=0.95, field=0, policy='random', steps=3,
         stable=False, mols_air=3.0e4),
    dict(name='all_terminal', alpha=0.5, lat=0.0, lng=0.0, date=(2013, 3, 25, 22, 25, 32),
         pressure=13500.0, x=0.0, y=0.0, ir=250.0, power=1e-5, field=0, policy='random', steps=3,
         stable=False, power_safety=False),
]


def run_scenario(sc, bank, rng):
  date = units.datetime(*sc['date'])
  atm = make_atmosphere(sc['alpha'])
  b = test_helpers.create_balloon(
      x=units.Distance(m=sc['x']), y=units.Distance(m=sc['y']), center_lat=sc['lat'],
      center_lng=sc['lng'], pressure=sc['pressure'], power_percent=sc['power'], date_time=date,
      power_safety_layer_enabled=sc.get('power_safety', True),
      use_stable_init=sc.get('stable', True), upwelling_infrared=sc['ir'], atmosphere=atm)
  if 'mols_air' in sc:
    b.state.mols_air = sc['mols_air']
  if sc['field'] < 0:
    wf = wind_field.SimpleStaticWindField()
  else:
    wf = grid_based_wind_field.GridBasedWindField(_BankSampler(bank[sc['field']]))
  wf.reset(jax.random.PRNGKey(1), date)
  seeds = rng.integers(0, 1634753849, size=(2, 5))
  offsets = (rng.uniform(0, 1, size=(2, 5, 4)).astype(np.float32) * np.float32(2.0)
             - np.float32(1.0)).astype(np.float64)
  _inject_noise(wf, seeds, offsets)

  # BalloonArena.__init__ and BalloonEnv.__init__ both reset (re-sampling atmosphere, balloon,
  # wind field); build them first, then inject the scenario's objects.
  arena = balloon_arena.BalloonArena(_NullFeatures, wf, seed=0)
  env = balloon_env.BalloonEnv(arena=arena, seed=0)
  atm = make_atmosphere(sc['alpha'])
  _inject_noise(wf, seeds, offsets)
  if sc['field'] >= 0:
    wf.field = bank[sc['field']]
  arena._balloon = b
  arena._atmosphere = atm
  arena.feature_constructor = _NullFeatures(None, None)

  n = sc['steps']
  if sc['policy'] == 'random':
    actions = rng.integers(0, 3, size=n)
  elif sc['policy'] == 'sticky':
    actions = np.repeat(rng.integers(0, 3, size=n // 20 + 1), 20)[:n]
  else:
    actions = np.full(n, dict(down=0, stay=1, up=2)[sc['policy']])
  f0, i0 = snapshot(b.state)
  fl, il, rew, done, winds = [], [], [], [], []
  for a in actions:
    w = arena._get_wind_ground_truth_at_balloon()
    _, r, d, info = env.step(int(a))
    f, i = snapshot(arena.get_balloon_state())
    fl.append(f); il.append(i); rew.append(float(r)); done.append(bool(d))
    winds.append([w.u.mps, w.v.mps])
    assert int(info['time_elapsed'].total_seconds()) == i[1]
    if d:
      break
  k = len(fl)
  return dict(alpha=sc['alpha'], field=sc['field'], power_safety=int(sc.get('power_safety', True)),
              seeds=seeds, offsets=offsets, actions=np.asarray(actions[:k], np.int64),
              f0=np.asarray(f0), i0=np.asarray(i0, np.int64), f=np.asarray(fl),
              i=np.asarray(il, np.int64), reward=np.asarray(rew), done=np.asarray(done),
              wind=np.asarray(winds, np.float64))


def run_feature_scenario(sc, bank, rng, steps):
  """The reference's PerciatelliFeatureConstructor (1099 features, WindGP) along an episode."""
  from balloon_learning_environment.env import features as features_lib
  date = units.datetime(*sc['date'])
  atm = make_atmosphere(sc['alpha'])
  b = test_helpers.create_balloon(
      x=units.Distance(m=sc['x']), y=units.Distance(m=sc['y']), center_lat=sc['lat'], center_lng=sc['lng'],
      pressure=sc['pressure'], power_percent=sc['power'], date_time=date, upwelling_infrared=sc['ir'],
      atmosphere=atm)
  wf = (wind_field.SimpleStaticWindField() if sc['field'] < 0 else
        grid_based_wind_field.GridBasedWindField(_BankSampler(bank[sc['field']])))
  wf.reset(jax.random.PRNGKey(1), date)
  seeds = rng.integers(0, 1634753849, size=(2, 5))
  offsets = (rng.uniform(0, 1, size=(2, 5, 4)).astype(np.float32) * np.float32(2.0) - np.float32(1.0)).astype(np.float64)
  arena = balloon_arena.BalloonArena(features_lib.PerciatelliFeatureConstructor, wf, seed=0)
  env = balloon_env.BalloonEnv(arena=arena, seed=0)
  atm = make_atmosphere(sc['alpha'])
  _inject_noise(wf, seeds, offsets)
  if sc['field'] >= 0:
    wf.field = bank[sc['field']]
  arena._balloon = b
  arena._atmosphere = atm
  arena.feature_constructor = features_lib.PerciatelliFeatureConstructor(wf, atm)
  arena.feature_constructor.observe(arena.get_measurements())
  obs0 = arena.feature_constructor.get_features()
  actions = rng.integers(0, 3, size=steps)
  f0, i0 = snapshot(b.state)
  obs, fl, il, rew = [obs0], [], [], []
  for a in actions:
    o, r, d, _ = env.step(int(a))
    f, i = snapshot(arena.get_balloon_state())
    obs.append(o); fl.append(f); il.append(i); rew.append(float(r))
    assert not d
  pr = arena.feature_constructor  # final pressure range for a direct KAT
  from balloon_learning_environment.env.balloon import pressure_range_builder
  rng_final = pressure_range_builder.get_pressure_range(arena.get_balloon_state(), atm)
  return dict(alpha=sc['alpha'], field=sc['field'], power_safety=1, seeds=seeds, offsets=offsets,
              actions=np.asarray(actions, np.int64), f0=np.asarray(f0), i0=np.asarray(i0, np.int64),
              f=np.asarray(fl), i=np.asarray(il, np.int64), reward=np.asarray(rew),
              obs=np.asarray(obs, np.float32),
              final_pressure_range=np.array([rng_final.min_pressure, rng_final.max_pressure]))


def main():
  kat = dict(atmosphere=kat_atmosphere(), solar=kat_solar(), stable_init=kat_stable_init(),
             interp=kat_interp(), safety=kat_safety(), float_fields=FLOATS, int_fields=INTS, **kat_thermal_acs_geometry())
  # opensimplex restatement self-KAT (pins oracle <-> CUDA, not the third-party package)
  gen = opensimplex.OpenSimplex(seed=1234567)
  rng = np.random.default_rng(3)
  pts = rng.uniform(-40, 40, (64, 4))
  kat['opensimplex_port'] = dict(seed=1234567, perm=[int(v) for v in gen._perm],
                                 points=pts.tolist(), values=[gen.noise4d(*p) for p in pts])
  # VAE decoder: the reference's flax module (under the Tier-0 flax/jax stand-ins) on seeded
  # random-init weights of its own architecture; a few slices are enough to pin the restatement.
  from oracle import vae as vae_oracle
  params = vae_oracle.synthetic_params(11)
  zs = np.random.default_rng(12).standard_normal((2, 64)).astype(np.float32)
  dec = [np.asarray(vae.Decoder().apply({'params': params}, z)) for z in zs]
  kat['vae'] = dict(params_seed=11, latents=zs.tolist(),
                    slices=[[d[:, :, 3, 4, :].tolist(), d[10, 5, :, :, :].tolist()] for d in dec],
                    mean_abs=[float(np.abs(d).mean()) for d in dec])
  with open(os.path.join(OUT, 'kat.json'), 'w') as f:
    json.dump(kat, f)
  bank = golden_fields.field_bank()
  rng = np.random.default_rng(2024)
  flat = {}
  names = []
  for sc in SCENARIOS:
    res = run_scenario(sc, bank, rng)
    names.append(sc['name'])
    for k, v in res.items():
      flat[f"{sc['name']}/{k}"] = np.asarray(v)
    st = res['i'][:, INTS.index('status')]
    print(f"{sc['name']:28s} steps={len(res['reward']):4d} final_status={st[-1]} "
          f"env_states={sorted(set(res['i'][:, 4]))} alt_states={sorted(set(res['i'][:, 5]))} "
          f"paused_steps={int(res['i'][:, 6].sum())} mean_reward={res['reward'].mean():.3f}")
  flat['names'] = np.asarray(names)
  np.savez_compressed(os.path.join(OUT, 'traj.npz'), **flat)
  # observation surface: two episodes through the reference's PerciatelliFeatureConstructor
  rng = np.random.default_rng(77)
  feat = {}
  for sc, steps in ((SCENARIOS[1], 140), (SCENARIOS[0], 40), (SCENARIOS[6], 60)):
    res = run_feature_scenario(sc, bank, rng, steps)
    for k, v in res.items():
      feat[f"{sc['name']}/{k}"] = np.asarray(v)
    print(f"features {sc['name']:20s} steps={steps} obs {res['obs'].shape} range {res['final_pressure_range']}")
  feat['names'] = np.asarray([SCENARIOS[1]['name'], SCENARIOS[0]['name'], SCENARIOS[6]['name']])
  np.savez_compressed(os.path.join(OUT, 'features.npz'), **feat)
  print('wrote', os.path.join(OUT, 'kat.json'), os.path.join(OUT, 'traj.npz'))


if __name__ == '__main__':
  main()
