"""Known answers and properties from the reference's own unit tests, replayed on the oracle.

Reference: env/balloon_env_test.py:87-206 (reward), env/wind_gp_test.py:44-65 (GP variance at a measured
point), env/balloon/balloon_test.py:93-212 (dynamics properties), env/balloon/stable_init_test.py:36-83
(stable initialisation holds pressure and temperature).
"""
import numpy as np

from oracle import balloon, constants as C, env as oenv, features as F

# env/balloon/balloon_test / balloon_env_test create_balloon defaults: (0, 0), 2013-03-25 09:25:32 UTC
_T0 = 1364203532


def _state(x_m, y_m, **kw):
  x_m = np.atleast_1d(np.asarray(x_m, np.float64))
  return balloon.make_batch(len(x_m), center_lat=0.0, center_lng=0.0, date_time=_T0, x=x_m,
                            y=np.atleast_1d(np.asarray(y_m, np.float64)), **kw)


def test_reward_in_radius_is_exactly_one():
  # env/balloon_env_test.py:87-110: radius 50 km (and 10 km), dropoff 0.0 -> reward == 1.0
  x = np.array([1.0, 49.99, 0.0, -35.355]) * 1000.0
  y = np.array([-1.0, 0.0, -49.99, 35.3]) * 1000.0
  r = oenv.perciatelli_reward(_state(x, y), station_keeping_radius_km=50.0, reward_dropoff=0.0)
  assert (r == 1.0).all()
  r = oenv.perciatelli_reward(_state([-9990.0], [0.0]), station_keeping_radius_km=10.0, reward_dropoff=0.0)
  assert r[0] == 1.0
  # the boundary itself is inside (<=, env/balloon_env.py:82)
  assert oenv.perciatelli_reward(_state([50000.0], [0.0]))[0] == 1.0


def test_reward_equals_dropoff_just_outside_radius():
  # env/balloon_env_test.py:112-137: 0.1 km outside the radius, |reward - dropoff| < 1e-3
  for radius_km, angle, dropoff in [(50.0, 0.6, 0.0), (50.0, 1.3, 0.4), (10.0, 2.1, 0.0)]:
    d = (radius_km + 0.1) * 1000.0
    r = oenv.perciatelli_reward(_state([d * np.cos(angle)], [d * np.sin(angle)]),
                                station_keeping_radius_km=radius_km, reward_dropoff=dropoff)
    assert abs(r[0] - dropoff) < 1e-3


def test_reward_halves_after_the_halflife_distance():
  # env/balloon_env_test.py:139-171: 51 km vs 101 km from the origin, halflife 50 km, dropoff 1.0
  r = oenv.perciatelli_reward(_state([47_548.69, 94_165.06], [18_442.39, 36_523.16]),
                              station_keeping_radius_km=50.0, reward_dropoff=1.0, reward_halflife=50.0)
  assert abs(r[0] * 0.5 - r[1]) < 1e-3


def test_power_regulariser_applies_only_to_down_without_excess_energy():
  # env/balloon_env_test.py:173-206 (expected 1.0 / 1.0 / 0.95 / 1.0 to two places).  Excess energy needs
  # daylight AND a full battery (balloon.py:231-238): 09:25 UTC at (0, 0) is daylight, 12:00 local night is not.
  day = _state([0.0, 0.0], [0.0, 0.0], battery_charge=C.BATTERY_CAPACITY_WH)
  assert day.excess_energy().all()
  day.last_command[:] = [C.DOWN, C.STAY]
  np.testing.assert_array_equal(oenv.perciatelli_reward(day), [1.0, 1.0])
  low = _state([0.0, 0.0], [0.0, 0.0])                       # 95 % battery -> no excess energy
  assert not low.excess_energy().any()
  low.last_command[:] = [C.DOWN, C.STAY]
  np.testing.assert_allclose(oenv.perciatelli_reward(low), [0.95, 1.0], atol=5e-3)
  # env/balloon_env.py:90-100 + utils/transforms.py:63-66: the multiplier falls linearly from 0.95 at
  # <= 100 W of ACS power to 0.65 at >= 300 W
  ramp = _state(np.zeros(5), np.zeros(5))
  ramp.last_command[:] = C.DOWN
  ramp.acs_power[:] = [0.0, 100.0, 200.0, 300.0, 1000.0]
  np.testing.assert_allclose(oenv.perciatelli_reward(ramp), [0.95, 0.95, 0.80, 0.65, 0.65], rtol=1e-12)


def test_gp_variance_at_a_measured_point():
  # env/wind_gp_test.py:44-54: SIGMA_NOISE^2 / (SIGMA_NOISE^2 + SIGMA_EXP^2) = 0.003843 to three places
  gp = F.WindGP(1)
  level = 77
  p = F.PRESSURE_LEVELS[level]
  gp.observe([0.0], [0.0], [p], [0], [1.0], [1.0], [0.0], [0.0])
  means, var = gp.query_column(0, 0.0, 0.0, 0)
  assert abs(var[level] - 0.003843) < 5e-4
  assert abs(var[level] - F.GP_NOISE / (F.GP_NOISE + F.GP_SIGMA2)) < 1e-12
  # the posterior mean at the measured point is the measurement shrunk by the same factor
  np.testing.assert_allclose(means[level], np.array([1.0, 1.0]) * F.GP_SIGMA2 / (F.GP_SIGMA2 + F.GP_NOISE), rtol=1e-12)
  # env/wind_gp_test.py:56-65: a nearby query moves continuously away from the prior (0 error, variance 1)
  far = gp.query_column(0, 50.0, 0.0, 0)
  assert (far[0][level] != 0.0).all() and far[1][level] < 1.0
  # 6 h horizon (wind_gp.py:172-178): an old measurement no longer informs the query
  stale_means, stale_var = gp.query_column(0, 0.0, 0.0, F.GP_HORIZON_S)
  assert (stale_means == 0.0).all() and (stale_var == 1.0).all()


# ------------------------------------------------------------------ dynamics properties of the reference's unit tests

def _sub_step(b, atm, action, u=10.0, v=12.0, n=1):
  """n calls of Balloon.simulate_step(..., time_delta = 10 s) (one physics sub-step each)."""
  for _ in range(n):
    balloon.simulate_step(b, u, v, atm, np.full(b.n, action), time_delta=10, stride=10)


def _default_balloon(atm, pressure=9000.0, stable=True, **kw):
  from oracle import stable_init
  b = balloon.make_batch(1, center_lat=0.0, center_lng=0.0, date_time=_T0, pressure=pressure, **kw)
  if stable:
    stable_init.cold_start_to_stable_params(b, atm)
  return b


def test_balloon_dynamics_properties():
  """env/balloon/balloon_test.py:93-212 on the oracle (the solar-calculator mocks become a day / night date)."""
  from oracle import atmosphere
  atm = atmosphere.Atmosphere([0.5])
  b = _default_balloon(atm)
  _sub_step(b, atm, C.STAY)
  assert b.x[0] == 100.0 and b.y[0] == 120.0                              # :93-105 goes in the wind direction
  for p0, sign in [(20123.0, -1), (2345.0, +1)]:                          # :107-131 up when low, down when high
    b = _default_balloon(atm, pressure=p0, stable=False)
    _sub_step(b, atm, C.STAY, u=3.0, v=-4.0)
    assert np.sign(b.pressure[0] - p0) == sign
  day, night = _T0, _T0 + 12 * 3600                                       # 09:25 / 21:25 UTC at (0, 0)
  half = 0.5 * C.BATTERY_CAPACITY_WH
  b = balloon.make_batch(1, center_lat=0.0, center_lng=0.0, date_time=day, pressure=9000.0, battery_charge=half)
  _sub_step(b, atm, C.STAY)
  assert b.battery_charge[0] > half and b.power_load[0] == C.DAYTIME_POWER_LOAD_W      # :139-152, :184-197
  b = balloon.make_batch(1, center_lat=0.0, center_lng=0.0, date_time=night, pressure=9000.0, battery_charge=half)
  _sub_step(b, atm, C.STAY)
  assert b.battery_charge[0] < half and b.power_load[0] == C.NIGHTTIME_POWER_LOAD_W    # :154-182
  assert b.solar_charging[0] == 0.0
  b = _default_balloon(atm)
  _sub_step(b, atm, C.DOWN)
  assert b.power_load[0] > C.DAYTIME_POWER_LOAD_W and b.acs_power[0] > 0.0             # :199-212


def test_stable_init_holds_pressure_and_temperature():
  """env/balloon/stable_init_test.py:36-83: after cold_start_to_stable_params the balloon stays within 100 Pa over
  100 sub-steps and dT/dt < 1e-3 K/s.  Atmosphere(PRNGKey(38)) draws alpha = 0.99589407."""
  from oracle import atmosphere, solar, stable_init, thermal
  atm = atmosphere.Atmosphere([0.99589407])
  midnight = 1590969600                                                    # 2020-06-01 00:00:00 UTC
  for p0 in (9500.0, 11500.0, 6500.0):
    b = balloon.make_batch(1, center_lat=0.0, center_lng=0.0, date_time=midnight, pressure=p0)
    stable_init.cold_start_to_stable_params(b, atm)
    _sub_step(b, atm, C.STAY, u=3.0, v=-4.0, n=100)
    assert b.status[0] == C.STATUS_OK and abs(b.pressure[0] - p0) < 100.0, (p0, b.pressure[0])
  for p0 in (9500.0, 11500.0, 5000.0):
    b = _default_balloon(atm, pressure=p0)
    lat, lng = b.latlng()
    el, _, flux = solar.solar_calculator(lat, lng, b.date_time)
    d_temp = thermal.d_balloon_temperature_dt(b.envelope_volume, C.ENVELOPE_MASS, b.internal_temperature,
                                              b.ambient_temperature, b.pressure, el, flux, b.upwelling_infrared)
    assert d_temp[0] < 1e-3
