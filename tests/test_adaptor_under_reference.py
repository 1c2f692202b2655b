"""The reference's REAL `BalloonEnv` (and its real PerciatelliFeatureConstructor) run over the adaptor's Python surface.

`CudaBalloonArena` is what a user passes as `BalloonEnv(arena=...)` (env/balloon_env.py:113,144-148).  Everything the
reference does with an arena afterwards -- `arena.step(command)`, `get_simulator_state()`, the reward function on
`simulator_state.balloon_state`, `_get_info`, a FeatureConstructor built from `(wind_field, atmosphere)` and fed
`get_measurements()` -- is executed here by the reference's own code, twice: once over the reference's BalloonArena and
once over `HostReplayArena`, a test-side subclass that swaps ONLY the adaptor's backend (`BatchedBalloonArena`, CUDA)
for tests/hostemu/backend.py (the same device headers compiled for the host).  The product class stays CUDA-only; the
GPU test below it runs the same comparison of the wind_field / atmosphere views on the real backend against the oracle.

Tier-0 only: needs /root/reference (skipped elsewhere; the GPU box runs the `gpu`-marked test, which needs no reference).
"""
import datetime as dt
import os

import numpy as np
import pytest
import torch

HAVE_REFERENCE = os.path.isdir(os.environ.get('BLE_REFERENCE_ROOT', '/root/reference'))
needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason='Tier-0 harness: needs the reference checkout')


def _host_replay_arena_class():
  from balloon_learning_environment_b200 import arena as arena_lib
  from tests.hostemu import backend as backend_lib

  class HostReplayArena(arena_lib.CudaBalloonArena):
    """CudaBalloonArena with the device backend replaced by the host replay (tests only)."""

    def __init__(self, feature_constructor_factory, wind_field, precision=1):
      self._arena = backend_lib.HostBackend(1, precision=precision)
      self._factory = feature_constructor_factory
      self.set_wind_field(wind_field)
      self.feature_constructor = None
      self.reset(0)

  return HostReplayArena


def _scenario(mg, name):
  return next(sc for sc in mg.SCENARIOS if sc['name'] == name)


def _build_pair(mg, sc, bank, factory, precision):
  """-> (reference env over its own arena, reference env over the adaptor), same balloon / field / noise / atmosphere."""
  import jax
  from balloon_learning_environment.env import balloon_arena, balloon_env, grid_based_wind_field
  from balloon_learning_environment.utils import test_helpers, units
  from tests import hostemu

  date = units.datetime(*sc['date'])
  rng = np.random.default_rng(5)
  seeds = rng.integers(0, 1634753849, size=(2, 5))
  offsets = (rng.uniform(0, 1, size=(2, 5, 4)).astype(np.float32) * np.float32(2.0) - np.float32(1.0)).astype(np.float64)

  def balloon():
    b = test_helpers.create_balloon(
        x=units.Distance(m=sc['x']), y=units.Distance(m=sc['y']), center_lat=sc['lat'], center_lng=sc['lng'],
        pressure=sc['pressure'], power_percent=sc['power'], date_time=date,
        power_safety_layer_enabled=sc.get('power_safety', True), use_stable_init=sc.get('stable', True),
        upwelling_infrared=sc['ir'], atmosphere=mg.make_atmosphere(sc['alpha']))
    if 'mols_air' in sc:
      b.state.mols_air = sc['mols_air']
    return b

  # the reference arm (as tests/golden/tier0/make_golden.py:run_scenario)
  wf = grid_based_wind_field.GridBasedWindField(mg._BankSampler(bank[sc['field']]))
  wf.reset(jax.random.PRNGKey(1), date)
  ref_arena = balloon_arena.BalloonArena(factory, wf, seed=0)
  ref_env = balloon_env.BalloonEnv(arena=ref_arena, seed=0)
  mg._inject_noise(wf, seeds, offsets)
  wf.field = bank[sc['field']]
  ref_arena._balloon = balloon()
  ref_arena._atmosphere = mg.make_atmosphere(sc['alpha'])
  ref_arena.feature_constructor = factory(wf, ref_arena._atmosphere)
  ref_arena.feature_constructor.observe(ref_arena.get_measurements())

  # the adaptor arm: the reference's BalloonEnv over HostReplayArena
  ours = _host_replay_arena_class()(factory, bank[sc['field']], precision=precision)
  our_env = balloon_env.BalloonEnv(arena=ours, seed=0)
  ours.set_wind_noise(seeds, offsets)
  f0, i0 = mg.snapshot(balloon().state)
  f = np.array(f0 + [sc['alpha']])[:, None]
  i = np.array(i0 + [int(sc.get('power_safety', True))], np.int64)[:, None]
  assert len(hostemu.F_ROWS) == f.shape[0] and len(hostemu.I_ROWS) == i.shape[0]
  ours._arena.set_state(torch.from_numpy(f), torch.from_numpy(i))
  ours.feature_constructor = ours._make_feature_constructor()
  ours.feature_constructor.observe(ours.get_measurements())
  return ref_env, our_env, wf


def _mps(w):
  return np.array([w.u.meters_per_second, w.v.meters_per_second])


@needs_reference
@pytest.mark.parametrize('name,steps,precision,tol', [
    ('grid_random', 60, 1, 1e-9), ('grid_random', 60, 2, 2e-4), ('grid_down_night_lowbatt', 40, 1, 1e-9),
    ('burst', 3, 1, 1e-9), ('all_terminal', 3, 2, 2e-4)])
def test_reference_balloon_env_runs_over_the_adaptor(name, steps, precision, tol):
  import tests.golden.tier0.boot  # noqa: F401
  from balloon_learning_environment.env import simulator_data, wind_field as ref_wind
  from balloon_learning_environment.env.balloon import balloon as ref_balloon, control, standard_atmosphere
  from balloon_learning_environment.utils import units
  from tests.golden import fields as golden_fields
  from tests.golden.tier0 import make_golden as mg

  sc = _scenario(mg, name)
  ref_env, our_env, wf = _build_pair(mg, sc, golden_fields.field_bank(), mg._NullFeatures, precision)
  rng = np.random.default_rng(3)
  probe_p = [5000.0, 8765.4, 13999.0]
  for t in range(steps):
    a = int(rng.integers(0, 3)) if sc['policy'] in ('random', 'sticky') else dict(down=0, up=2)[sc['policy']]
    _, r_ref, d_ref, info_ref = ref_env.step(a)
    obs, r, d, info = our_env.step(a)
    assert isinstance(obs, np.ndarray)
    assert abs(float(r) - float(r_ref)) <= max(tol, 1e-9) * 10, (t, r, r_ref)
    assert bool(d) == bool(d_ref), t
    assert set(info) == set(info_ref) == {'out_of_power', 'envelope_burst', 'zeropressure', 'time_elapsed'}
    for k in info_ref:
      assert info[k] == info_ref[k], (t, k)
    assert our_env.arena.get_info() == {k: (bool(v) if k != 'time_elapsed' else v) for k, v in info_ref.items()}

    ss, ss_ref = our_env.get_simulator_state(), ref_env.get_simulator_state()
    assert isinstance(ss, simulator_data.SimulatorState)
    b, b_ref = ss.balloon_state, ss_ref.balloon_state
    # the reference's own types cross the boundary
    assert isinstance(b.x, units.Distance) and isinstance(b.acs_power, units.Power)
    assert isinstance(b.battery_charge, units.Energy) and isinstance(b.status, ref_balloon.BalloonStatus)
    assert isinstance(b.last_command, control.AltitudeControlCommand)
    assert b.status == b_ref.status and b.last_command == b_ref.last_command
    assert b.date_time == b_ref.date_time and b.time_elapsed == b_ref.time_elapsed
    assert b.navigation_is_paused == b_ref.navigation_is_paused and b.excess_energy == b_ref.excess_energy
    for k, floor in (('pressure', 0), ('ambient_temperature', 0), ('internal_temperature', 0), ('envelope_volume', 0),
                     ('superpressure', 100.0), ('mols_air', 100.0), ('battery_soc', 0), ('pressure_ratio', 0)):
      want = float(getattr(b_ref, k))
      assert abs(float(getattr(b, k)) - want) <= tol * 10 * max(abs(want), floor) + 1e-12, (t, k)
    # the adaptor's winds are float32 (|err| ~1e-6 m/s x 180 s per step)
    assert abs(b.x.m - b_ref.x.m) <= tol * 1e4 + 0.05 and abs(b.y.m - b_ref.y.m) <= tol * 1e4 + 0.05
    assert abs(b.latlng.lat().radians - b_ref.latlng.lat().radians) < 1e-9 + tol * 1e-2
    assert abs(b.battery_charge.watt_hours - b_ref.battery_charge.watt_hours) <= tol * 10 * 3058.56

    if d_ref:
      break
    if t % 10 == 0:
      # SimulatorState.wind_field: the three WindField queries on THIS balloon's field and noise
      x, y, el = b_ref.x, b_ref.y, b_ref.time_elapsed
      for p in probe_p:
        got, want = ss.wind_field.get_forecast(x, y, p, el), wf.get_forecast(x, y, p, el)
        assert isinstance(got, ref_wind.WindVector)
        np.testing.assert_allclose(_mps(got), _mps(want), atol=2e-5)
        got, want = ss.wind_field.get_ground_truth(x, y, p, el), wf.get_ground_truth(x, y, p, el)
        np.testing.assert_allclose(_mps(got), _mps(want), atol=2e-5)
      col = ss.wind_field.get_forecast_column(x, y, probe_p, el)
      want = wf.get_forecast_column(x, y, probe_p, el)
      np.testing.assert_allclose([_mps(w) for w in col], [_mps(w) for w in want], atol=2e-5)
      # SimulatorState.atmosphere
      for p in (14000.0, 8806.3, 5000.0):
        got, want = ss.atmosphere.at_pressure(p), ss_ref.atmosphere.at_pressure(p)
        assert isinstance(got, standard_atmosphere.AtmosphericValues)
        np.testing.assert_allclose([got.height.m, got.temperature, got.pressure, got.density],
                                   [want.height.m, want.temperature, want.pressure, want.density], rtol=1e-12)
      for h in (15240.0, 17000.0, 20999.9):
        got, want = ss.atmosphere.at_height(units.Distance(m=h)), ss_ref.atmosphere.at_height(units.Distance(m=h))
        np.testing.assert_allclose([got.height.m, got.temperature, got.pressure, got.density],
                                   [want.height.m, want.temperature, want.pressure, want.density], rtol=1e-12)
  with pytest.raises(AssertionError):
    our_env.get_simulator_state().atmosphere.at_pressure(1.0e6)


@needs_reference
def test_reference_feature_constructor_runs_over_the_adaptor():
  """env/features.py's PerciatelliFeatureConstructor (WindGP included) built from the ADAPTOR'S wind_field / atmosphere and
  fed the adaptor's get_measurements(): its 1099 features equal the ones it computes over the reference arena."""
  import tests.golden.tier0.boot  # noqa: F401
  from balloon_learning_environment.env import features as features_lib
  from tests.golden import fields as golden_fields
  from tests.golden.tier0 import make_golden as mg

  sc = _scenario(mg, 'grid_random')
  ref_env, our_env, _ = _build_pair(mg, sc, golden_fields.field_bank(), features_lib.PerciatelliFeatureConstructor, 1)
  assert isinstance(our_env.arena.feature_constructor, features_lib.PerciatelliFeatureConstructor)
  rng = np.random.default_rng(4)
  worst = 0.0
  for t in range(12):
    a = int(rng.integers(0, 3))
    o_ref, r_ref, _, _ = ref_env.step(a)
    o, r, _, _ = our_env.step(a)
    assert o.shape == o_ref.shape == (1099,) and o.dtype == np.float32
    worst = max(worst, float(np.abs(o - o_ref).max()))
    assert abs(r - r_ref) < 1e-8
  assert worst < 5e-5, worst     # the adaptor's winds are float32


@pytest.mark.gpu
def test_wind_field_and_atmosphere_views_on_the_device():
  """CudaWindField / CudaAtmosphere of a live CudaBalloonArena against the oracle (no reference needed)."""
  from balloon_learning_environment_b200 import arena as arena_lib, units
  from oracle import atmosphere as atmosphere_lib, wind as wind_lib
  from tests.golden import fields as golden_fields

  bank = golden_fields.field_bank()
  rng = np.random.default_rng(9)
  seeds = rng.integers(0, 1634753849, size=(2, 5))
  offsets = (rng.uniform(0, 1, size=(2, 5, 4)).astype(np.float32) * np.float32(2.0) - np.float32(1.0)).astype(np.float64)
  a = arena_lib.CudaBalloonArena(wind_field=bank[1], seed=11, observation=None)
  a.set_wind_noise(seeds, offsets)
  for _ in range(3):
    a.step(2)
  ss = a.get_simulator_state()
  alpha = ss.balloon_state.atmosphere_alpha
  atm = atmosphere_lib.Atmosphere(alpha)
  noise = wind_lib.SimplexWindNoise(seeds[None], offsets[None])
  fields = np.ascontiguousarray(bank[1:2], np.float32)
  for _ in range(20):
    x, y = rng.uniform(-6e5, 6e5, 2)
    t = int(rng.integers(0, 4 * 86400))
    ps = rng.uniform(4000, 15000, 7)
    col = ss.wind_field.get_forecast_column(units.Distance(x), units.Distance(y), ps, dt.timedelta(seconds=t))
    u, v = wind_lib.get_forecast(fields, np.zeros(7, np.int64), np.full(7, x), np.full(7, y), ps, np.full(7, t))
    np.testing.assert_allclose([[w.u.mps, w.v.mps] for w in col], np.stack([u, v], 1), atol=2e-5)
    one = ss.wind_field.get_forecast(units.Distance(x), units.Distance(y), ps[0], dt.timedelta(seconds=t))
    assert abs(one.u.mps - u[0]) < 2e-5 and abs(one.v.mps - v[0]) < 2e-5
    gt = ss.wind_field.get_ground_truth(units.Distance(x), units.Distance(y), ps[0], dt.timedelta(seconds=t))
    du, dv = noise.get_wind_noise(np.array([x]), np.array([y]), ps[:1], np.array([t]))
    assert abs(gt.u.mps - (u[0] + du[0])) < 5e-5 and abs(gt.v.mps - (v[0] + dv[0])) < 5e-5
  for p in (14000.0, 12027.7, 8806.3, 5000.0, 700.0):
    got = ss.atmosphere.at_pressure(p)
    h, temp = atm.at_pressure(np.array([p]))
    np.testing.assert_allclose([got.height.m, got.temperature, got.pressure], [h[0], temp[0], p], rtol=1e-12)
  for h in (0.0, 15240.0, 17000.0, 25000.0):
    got = ss.atmosphere.at_height(units.Distance(h))
    p, temp = atm.at_height(np.array([h]))
    np.testing.assert_allclose([got.height.m, got.temperature, got.pressure], [h, temp[0], p[0]], rtol=1e-12)
  with pytest.raises(AssertionError):
    ss.atmosphere.at_pressure(2.0e5)
  info = a.get_info()
  assert set(info) == {'out_of_power', 'envelope_burst', 'zeropressure', 'time_elapsed'}
  assert info['time_elapsed'] == dt.timedelta(seconds=540)
  a.close()
