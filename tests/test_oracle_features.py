"""Observation surface: the oracle against the reference's recorded Perciatelli observations, and the
scalar device helpers (host replay) against the oracle."""
import ctypes
import os

import numpy as np

from oracle import atmosphere, balloon, constants as C, features as F
from tests import golden_io, hostemu

KAT = golden_io.load_kat()
FF, IF = KAT['float_fields'], KAT['int_fields']
FEAT = np.load(os.path.join(golden_io.GOLDEN_DIR, 'features.npz'))
P = hostemu.ptr


def _scenario(name):
  return {k.split('/', 1)[1]: FEAT[k] for k in FEAT.files if k.startswith(name + '/')}


def test_oracle_reproduces_reference_observations():
  """Three episodes through the reference's PerciatelliFeatureConstructor (WindGP, pressure range,
  sunrise time, power table): 1099 float32 features, compared at 1e-6 absolute."""
  for name, stride in (('default_static', 8), ('sticky_random', 12), ('grid_random', 20)):
    sc = _scenario(name)
    env = golden_io.oracle_env_for_scenario(sc, FF, IF)
    feat = F.PerciatelliFeatures(env.arena)
    feat.observe()
    steps = len(sc['actions'])
    for t in range(steps + 1):
      if t > 0:
        env.step(np.array([sc['actions'][t - 1]]))
        feat.observe()
      if t % stride and t not in (1, steps):
        continue
      obs = feat.get_features()[0]
      assert obs.shape == (1099,) and obs.dtype == np.float32
      np.testing.assert_allclose(obs, sc['obs'][t], rtol=0, atol=1e-6, err_msg=f'{name} t={t}')
    pmin, pmax = F.get_pressure_range(env.arena.state, env.arena.atmosphere)
    np.testing.assert_allclose([pmin[0], pmax[0]], sc['final_pressure_range'], rtol=1e-10)


def test_feature_layout_properties():
  """env/features_test.py:92-478: shape, one-hot order UP/STAY/DOWN, padding triple (0, 1, 1)."""
  sc = _scenario('grid_random')
  obs = sc['obs'][-1]
  assert obs.shape[0] == F.NUM_FEATURES == 1099
  assert obs[8:11].sum() == 1.0 and obs[11] + obs[12] == 1.0
  col = obs[16:].reshape(361, 3)
  pad = np.all(col == np.array([0.0, 1.0, 1.0], np.float32), axis=1)
  assert pad[0] and pad[-1] and not pad[180]           # the balloon's own level sits at index 180
  # survey KAT: get_pressure_range for alpha = 0.5, 9000 Pa default balloon
  atm = atmosphere.Atmosphere([0.5])
  b = balloon.make_batch(1, center_lat=0.0, center_lng=0.0, date_time=1364203532, pressure=9000.0)
  b.battery_charge[:] = 0.95 * C.BATTERY_CAPACITY_WH
  from oracle import stable_init
  stable_init.cold_start_to_stable_params(b, atm)
  pmin, pmax = F.get_pressure_range(b, atm)
  np.testing.assert_allclose([pmin[0], pmax[0]], [5973.567203433708, 12027.731351001139], rtol=1e-11)


def test_power_table_reference_rows():
  # env/balloon/power_table_test.py:32-73 (a sample of its 37 exact rows)
  for pr, soc, want in [(1.0, 0.2, 0), (1.0, 0.35, 150), (1.0, 0.45, 175), (1.0, 0.9, 200), (1.09, 0.5, 200),
                        (1.09, 0.75, 225), (1.12, 0.65, 250), (1.15, 0.45, 225), (1.19, 0.55, 275),
                        (1.21, 0.45, 275), (1.21, 0.55, 300), (1.25, 0.55, 300), (1.25, 0.65, 325),
                        (1.3, 0.3, 0), (1.3, 0.55, 325), (1.3, 0.7, 350)]:
    assert F.power_table_lookup(pr, soc) == want


def test_device_scalar_helpers_match_oracle():
  lib = hostemu.load()
  lib.emu_power_table.restype = ctypes.c_double
  lib.emu_power_table.argtypes = [ctypes.c_double, ctypes.c_double]
  lib.emu_nearest_level.argtypes = [ctypes.c_double]
  rng = np.random.default_rng(4)
  for pr, soc in zip(rng.uniform(0.99, 1.4, 400), rng.uniform(0, 1, 400)):
    assert lib.emu_power_table(pr, soc) == F.power_table_lookup(pr, soc)
  for p in list(rng.uniform(4000, 15000, 400)) + [5025.0, 5075.0, 5125.0, 13975.0, 4000.0, 14001.0]:
    q = min(max(p, 5000.0), 14000.0)
    assert lib.emu_nearest_level(p) == int(round((q - 5000.0) / 50.0))
  n = 64
  alpha = rng.uniform(0, 1, n); atm = atmosphere.Atmosphere(alpha)
  pmax, _ = atm.at_height(np.full(n, 15240.0))
  b = balloon.make_batch(n, center_lat=np.radians(rng.uniform(-10, 10, n)), center_lng=np.radians(rng.uniform(-175, 175, n)),
                         date_time=1293840000 + rng.integers(0, 126144000, n), pressure=rng.uniform(6500, pmax),
                         upwelling_infrared=rng.uniform(225, 315, n), x=rng.normal(0, 1e5, n), y=rng.normal(0, 1e5, n))
  lat, lng = b.latlng()
  want_min, want_max = F.get_pressure_range(b, atm)
  out = np.zeros((n, 2)); ok = np.zeros(n, np.int32)
  lib.emu_pressure_range(ctypes.c_int64(n), P(alpha), P(b.mols_lift_gas), P(lat), P(lng), P(b.date_time),
                         P(b.upwelling_infrared), P(out), P(ok))
  assert ok.all()
  np.testing.assert_allclose(out[:, 0], want_min, rtol=1e-9)
  np.testing.assert_allclose(out[:, 1], want_max, rtol=1e-9)
  st = np.zeros(n)
  lib.emu_sunrise_time(ctypes.c_int64(n), P(lat), P(lng), P(b.date_time), P(st))
  np.testing.assert_allclose(st, F.compute_sunrise_time(lat, lng, b.date_time), rtol=1e-12)


def test_incremental_cholesky_matches_refit():
  """The drop-oldest / append-newest factor update (k_gp_update's algorithm) against a fresh Cholesky of the
  Matern-1/2 Gram matrix of the shifted window, over 300 consecutive shifts of a 120-point window."""
  from oracle import features as features_lib
  rng = np.random.default_rng(7)
  n_pts, m = 420, 120
  pts = np.cumsum(rng.normal(0, 0.05, (n_pts, 4)), axis=0)          # a wandering balloon in scaled coordinates
  pts[:, 3] = np.arange(n_pts) * (180.0 / 34560.0)

  def gram(p):
    d = np.sqrt(((p[:, None, :] - p[None, :, :]) ** 2).sum(-1))
    return 3.6 ** 2 * np.exp(-d) + 0.05 * np.eye(len(p))

  factor = np.linalg.cholesky(gram(pts[:m]))
  worst = 0.0
  for first in range(1, 301):
    window = pts[first:first + m]
    factor = features_lib.cholesky_drop_first(factor)
    k_new = 3.6 ** 2 * np.exp(-np.sqrt(((window[:-1] - window[-1]) ** 2).sum(-1)))
    factor = features_lib.cholesky_append(factor, k_new, 3.6 ** 2 + 0.05)
    if first % 50 == 0 or first < 3:
      want = np.linalg.cholesky(gram(window))
      worst = max(worst, np.abs(factor - want).max())
  assert worst < 1e-11, worst                                        # no drift over 300 carried steps
