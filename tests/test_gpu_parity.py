"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference
goldens.  Run on the B200 box with `pytest -m gpu`.

Tolerances (BASELINE.json north_star): fp32 state within 1e-4 relative of the reference after a
step; discrete decisions (effective action masks -> safety-layer states, status, last command,
time) bit-exact.  The fp64 audit build of the same kernels is held to ~1e-8.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import atmosphere as atmosphere_lib
from oracle import balloon as balloon_lib
from oracle import constants as C
from oracle import env as env_lib
from oracle import solar as solar_lib
from oracle import stable_init as stable_init_lib
from oracle import wind as wind_lib
from tests import golden_io
from tests.golden import fields as golden_fields

pytestmark = pytest.mark.gpu

KAT = golden_io.load_kat()
TRAJ = golden_io.load_traj()
FF, IF = KAT['float_fields'], KAT['int_fields']

# relative-error floors for fields that pass through zero (fraction of typical magnitude)
FLOORS = {'x': 1e4, 'y': 1e4, 'acs_mass_flow': 1e-2, 'superpressure': 100.0, 'solar_charging': 50.0,
          'acs_power': 100.0, 'mols_air': 100.0}


@pytest.fixture(scope='module')
def ble():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  import balloon_learning_environment_b200 as pkg
  from balloon_learning_environment_b200 import _lib, batched_env
  _lib.load()          # raises if the CUDA library is missing: GPU tests must not silently skip it
  return batched_env


def pack(batch, alpha, psl):
  from balloon_learning_environment_b200 import _lib
  n = batch.n
  f = np.empty((len(_lib.F_ROWS), n)); i = np.empty((len(_lib.I_ROWS), n), np.int64)
  for r, k in enumerate(_lib.F_ROWS):
    f[r] = alpha if k == 'atmosphere_alpha' else getattr(batch, k)
  for r, k in enumerate(_lib.I_ROWS):
    i[r] = np.broadcast_to(np.asarray(psl, np.int64), (n,)) if k == 'power_safety_enabled' else getattr(batch, k)
  return torch.from_numpy(f), torch.from_numpy(i)


def state_np(arena):
  from balloon_learning_environment_b200 import _lib
  f, i = arena.get_state()
  torch.cuda.synchronize()
  f, i = f.cpu().numpy(), i.cpu().numpy()
  d = {k: f[r] for r, k in enumerate(_lib.F_ROWS)}
  d.update({k: i[r] for r, k in enumerate(_lib.I_ROWS)})
  return d


def rel_err(got, ref, key):
  scale = np.maximum(np.abs(ref), FLOORS.get(key, 1e-30))
  return float((np.abs(got - ref) / scale).max())


# ------------------------------------------------------------------------------ wind gather

@pytest.mark.parametrize('precision,tol', [('fp64', 1e-11), ('fp32', 1e-5)])
def test_wind_gather_matches_oracle(ble, precision, tol):
  rng = np.random.default_rng(9)
  bank = golden_fields.field_bank()
  arena = ble.BatchedBalloonArena(8, precision=precision, enable_noise=False)
  arena.set_wind_fields(torch.from_numpy(bank), torch.zeros(8, dtype=torch.int32))
  m = 20000
  xyzt = np.stack([rng.uniform(-600, 600, m), rng.uniform(-600, 600, m), rng.uniform(4000, 15000, m),
                   rng.uniform(0, 200, m)], 1).astype(np.float32)
  xyzt[:10, 3] = [0, 6, 47.999, 48, 48.001, 96, 144, 95.99, 191.9, 0.001]
  xyzt[10:14, 0] = [-500, 500, -500.001, 499.999]
  xyzt[14:18, 2] = [5000, 14000, 4999.9, 14000.1]
  fidx = rng.integers(0, 4, m).astype(np.int32)
  uv = arena.wind_forecast(torch.from_numpy(xyzt), torch.from_numpy(fidx)).cpu().numpy()
  pts = wind_lib.prepare_points(xyzt[:, 0].astype(np.float64) * 1000, xyzt[:, 1].astype(np.float64) * 1000,
                                xyzt[:, 2].astype(np.float64), xyzt[:, 3].astype(np.float64) * 3600.0)
  want = wind_lib.interpolate(bank, fidx, pts)
  err = float(np.abs(uv - want).max())
  print(f'wind gather {precision}: worst |err| = {err:.2e} m/s over {m} lookups (max |wind| {np.abs(want).max():.1f} m/s)')
  # fp64 engine: the only error is the float32 cast of the output; fp32 engine: fp32 weights and a 16-term fp32 blend
  assert err < (3e-6 if precision == 'fp64' else tol)
  arena.close()


def test_wind_gather_reference_kat_and_empty(ble):
  rows = np.array(KAT['interp'])
  bank = golden_fields.field_bank()
  arena = ble.BatchedBalloonArena(4, precision='fp64', enable_noise=False)
  arena.set_wind_fields(torch.from_numpy(bank))
  xyzt = np.stack([rows[:, 1] / 1000.0, rows[:, 2] / 1000.0, rows[:, 3], rows[:, 4] / 3600.0], 1)
  keep = np.all(xyzt.astype(np.float32).astype(np.float64) == xyzt, axis=1)
  uv = arena.wind_forecast(torch.from_numpy(xyzt[keep].astype(np.float32)),
                           torch.from_numpy(rows[keep, 0].astype(np.int32))).cpu().numpy()
  np.testing.assert_allclose(uv, rows[keep][:, 5:7], rtol=1e-6, atol=2e-6)      # float32 output
  empty = arena.wind_forecast(torch.zeros(0, 4), torch.zeros(0, dtype=torch.int32))
  assert tuple(empty.shape) == (0, 2)
  arena.close()


# ------------------------------------------------------------------------------ noise

@pytest.mark.parametrize('precision,tol', [('fp64', 1e-6), ('fp32', 3e-4)])
def test_noise_matches_oracle(ble, precision, tol):
  rng = np.random.default_rng(3)
  n = 1000                                      # not a multiple of the 128-balloon noise block
  arena = ble.BatchedBalloonArena(n, precision=precision, enable_noise=True)
  arena.set_wind_fields(torch.zeros(1, *ble.FIELD_SHAPE))          # forecast == 0 -> wind == noise
  seeds = rng.integers(0, 1634753849, (n, 2, 5))
  offsets = (rng.uniform(0, 1, (n, 2, 5, 4)).astype(np.float32) * 2 - 1)
  arena.set_wind_noise(torch.from_numpy(seeds), torch.from_numpy(offsets))
  b = balloon_lib.make_batch(n, center_lat=0.0, center_lng=0.0, date_time=1364203532)
  b.x = rng.uniform(-8e5, 8e5, n); b.y = rng.uniform(-8e5, 8e5, n)
  b.pressure = rng.uniform(5000, 14000, n); b.time_elapsed = rng.integers(0, 400000, n) // 10 * 10
  f, i = pack(b, 0.5, 1)
  arena.set_state(f, i)
  got = arena.wind_at_balloon().cpu().numpy().astype(np.float64)
  st = state_np(arena)                          # fp32 build rounds x, y, p on upload
  noise = wind_lib.SimplexWindNoise(seeds, offsets.astype(np.float64))
  du, dv = noise.get_wind_noise(st['x'], st['y'], st['pressure'], st['time_elapsed'])
  assert np.abs(got[:, 0] - du).max() < tol and np.abs(got[:, 1] - dv).max() < tol
  assert np.abs(du).max() > 0.5                 # the comparison is not vacuous
  arena.close()


# ------------------------------------------------------------------------------ one step vs the reference

def _all_recorded_states():
  """Every recorded (state_t, action_{t+1}) -> (state_{t+1}, reward, wind) pair of the goldens."""
  names = sorted(TRAJ)
  fs, is_, acts, want_f, want_i, want_r, want_w, alpha, psl, field, seeds, offs = ([] for _ in range(12))
  for name in names:
    sc = TRAJ[name]
    n = len(sc['actions'])
    idx = np.arange(0, n - 1)
    if idx.size == 0:
      continue
    fs.append(sc['f'][idx]); is_.append(sc['i'][idx]); acts.append(sc['actions'][idx + 1])
    want_f.append(sc['f'][idx + 1]); want_i.append(sc['i'][idx + 1]); want_r.append(sc['reward'][idx + 1])
    want_w.append(sc['wind'][idx + 1])
    alpha.append(np.full(idx.size, float(sc['alpha']))); psl.append(np.full(idx.size, int(sc['power_safety'])))
    field.append(np.full(idx.size, int(sc['field'])))
    seeds.append(np.broadcast_to(sc['seeds'], (idx.size, 2, 5))); offs.append(np.broadcast_to(sc['offsets'], (idx.size, 2, 5, 4)))
  saf = golden_io.load_safety()          # single reference steps from every safety band / prior state / terminal status
  fs.append(saf['f']); is_.append(saf['i']); acts.append(saf['action']); want_f.append(saf['want_f'])
  want_i.append(saf['want_i']); want_r.append(saf['reward']); want_w.append(saf['wind']); alpha.append(saf['alpha'])
  psl.append(saf['psl']); field.append(saf['field']); seeds.append(saf['seeds']); offs.append(saf['offsets'])
  cat = np.concatenate
  return dict(f=cat(fs), i=cat(is_), actions=cat(acts), want_f=cat(want_f), want_i=cat(want_i),
              want_r=cat(want_r), want_w=cat(want_w), alpha=cat(alpha), psl=cat(psl), field=cat(field),
              seeds=cat(seeds), offsets=cat(offs))


@pytest.mark.parametrize('precision,tol,wtol,kernel,warps', [
    ('fp64', 1e-8, 2e-6, 'thread', None), ('fp32', 1e-4, 5e-4, 'fused', 0), ('fp32', 1e-4, 5e-4, 'fused', 4),
    ('fp32', 1e-4, 5e-4, 'fused', 8), ('fp32', 1e-4, 5e-4, 'fused', 14), ('fp32', 1e-4, 5e-4, 'thread', None),
    ('fp32', 1e-4, 5e-4, 'ws', None)])
def test_single_step_matches_reference(ble, monkeypatch, precision, tol, wtol, kernel, warps):
  """~6,400 reference-recorded (state, action) -> (state', reward, wind) pairs -- ten episodes plus 1,977 single
  steps started in every envelope / altitude / power safety band, prior machine state and terminal status
  (tests/golden/safety.npz) -- advanced by ONE BalloonEnv.step on the GPU, through every step kernel: the production
  kernels k_step_warp (shape 0) and k_step_roles<4 | 8 | 14>, the first-generation k_step / k_step_ws, and the fp64
  audit build."""
  monkeypatch.setenv('BLE_STEP_KERNEL', kernel)
  if warps is not None:
    monkeypatch.setenv('BLE_STEP_WARPS', str(warps))
  rec = _all_recorded_states()
  ie = IF.index('envelope_state'); ia = IF.index('altitude_state'); ip = IF.index('power_paused')
  for col, values in ((ie, range(5)), (ia, range(3)), (ip, range(2))):        # coverage of the discrete machinery
    for val in values:
      assert (rec['i'][:, col] == val).sum() >= 50 and (rec['want_i'][:, col] == val).sum() >= 50
  grid = rec['field'] >= 0                      # SimpleStaticWindField scenario runs in its own arena
  for model, sel in (('grid', grid), ('simple_static', ~grid)):
    n = int(sel.sum())
    arena = ble.BatchedBalloonArena(n, precision=precision, wind_model=model, enable_noise=True)
    if model == 'grid':
      arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()),
                            torch.from_numpy(rec['field'][sel].astype(np.int32)))
    arena.set_wind_noise(torch.from_numpy(rec['seeds'][sel].copy()),
                         torch.from_numpy(rec['offsets'][sel].astype(np.float32)))
    b = golden_io.batch_from_rows(FF, IF, rec['f'][sel], rec['i'][sel])
    arena.set_state(*pack(b, rec['alpha'][sel], rec['psl'][sel]))
    reward, done, wind = arena.step(torch.from_numpy(rec['actions'][sel].astype(np.int32)))
    torch.cuda.synchronize()
    st = state_np(arena)
    assert np.abs(wind.cpu().numpy() - rec['want_w'][sel]).max() < wtol
    worst = {k: rel_err(st[k], rec['want_f'][sel][:, j], k) for j, k in enumerate(FF)}
    assert max(worst.values()) < tol, worst
    for j, k in enumerate(IF):
      mism = st[k] != rec['want_i'][sel][:, j]
      if k in ('sunrise_h', 'sunset'):
        mism = mism & (rec['psl'][sel] == 1)
      assert not mism.any(), (k, int(mism.sum()))
    assert np.abs(reward.cpu().numpy() - rec['want_r'][sel]).max() < max(tol, 2e-7)
    np.testing.assert_array_equal(done.cpu().numpy() != 0, rec['want_i'][sel][:, IF.index('status')] != 0)
    # info of BalloonEnv.step (env/balloon_env.py:280-290), written by the step kernel
    info = arena.step_info()
    status = rec['want_i'][sel][:, IF.index('status')]
    np.testing.assert_array_equal(info['out_of_power'].cpu().numpy(), status == 1)
    np.testing.assert_array_equal(info['envelope_burst'].cpu().numpy(), status == 2)
    np.testing.assert_array_equal(info['zeropressure'].cpu().numpy(), status == 3)
    np.testing.assert_array_equal(info['time_elapsed'].cpu().numpy(), rec['want_i'][sel][:, IF.index('time_elapsed')])
    assert not info['sim_error'].any()
    arena.close()


def _free_run(ble, precision, kernel, use_rollout=False):
  """The ten recorded reference episodes (up to 960 steps) rolled out on the GPU from their initial states with the
  reference's actions and NO re-synchronisation: the device follows its own wind lookups.  Returns the worst relative
  drift per field (with the time and scenario it happened at), the per-field median over all steps, and the number
  of discrete mismatches."""
  names = sorted(TRAJ)
  scs = [TRAJ[n] for n in names]
  oenv = golden_io.oracle_env_for_scenarios(scs, FF, IF)
  field = np.array([int(sc['field']) for sc in scs])
  alpha = np.array([float(sc['alpha']) for sc in scs]); psl = np.array([int(sc['power_safety']) for sc in scs])
  worst = {k: (0.0, '', 0) for k in FF}
  errs = {k: [] for k in FF}
  mism, decisions = 0, 0
  for model, sel in (('grid', field >= 0), ('simple_static', field < 0)):
    idx = np.nonzero(sel)[0]
    arena = ble.BatchedBalloonArena(len(idx), precision=precision, wind_model=model, enable_noise=True)
    if model == 'grid':
      arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()), torch.from_numpy(field[idx].astype(np.int32)))
    arena.set_wind_noise(torch.from_numpy(np.stack([scs[e]['seeds'] for e in idx])),
                         torch.from_numpy(np.stack([scs[e]['offsets'] for e in idx]).astype(np.float32)))
    arena.set_state(*pack(oenv.arena.state.select(idx), alpha[idx], psl[idx]))
    horizon = max(len(scs[e]['actions']) for e in idx)
    acts_all = np.array([[scs[e]['actions'][t] if t < len(scs[e]['actions']) else 1 for e in idx] for t in range(horizon)], np.int32)
    chunk = 32 if use_rollout else 1
    for t0 in range(0, horizon, chunk):
      if use_rollout:
        arena.rollout(torch.from_numpy(acts_all[t0:t0 + chunk]))
      else:
        arena.step(torch.from_numpy(acts_all[t0]))
      t = min(t0 + chunk, horizon) - 1
      st = state_np(arena)
      for col, e in enumerate(idx):
        sc = scs[e]
        if t >= len(sc['actions']):
          continue
        for j, k in enumerate(FF):
          ref = sc['f'][t, j]
          err = abs(st[k][col] - ref) / max(abs(ref), FLOORS.get(k, 1e-30))
          errs[k].append(err)
          if err > worst[k][0]:
            worst[k] = (float(err), names[e], t)
        for j, k in enumerate(IF):
          if k in ('sunrise_h', 'sunset') and not sc['power_safety']:
            continue
          decisions += 1
          mism += int(st[k][col] != sc['i'][t, j])
    arena.close()
  median = {k: float(np.median(v)) for k, v in errs.items()}
  p99 = {k: float(np.quantile(v, 0.99)) for k, v in errs.items()}
  return worst, median, p99, mism, decisions


@pytest.mark.parametrize('precision,kernel,rollout', [('fp64', 'thread', False), ('fp32', 'fused', False), ('fp32', 'fused', True)])
def test_free_running_episodes_vs_reference(ble, monkeypatch, precision, kernel, rollout):
  """free-running rollouts of the recorded 960-step reference episodes (no per-step
  re-synchronisation), drift per field and discrete-mismatch count reported.

  What bounds the drift is the reference's own dynamics, not the arithmetic: dh/dt = +-sqrt(2 |lift - mass| g / ...)
  (env/balloon/balloon.py:424-427) has infinite slope where the balloon crosses buoyancy equilibrium, and a sub-step
  that lands close to the crossing amplifies whatever error has accumulated by up to 1e6 (episode 'sticky_random',
  step ~371: the fp64 audit build jumps from 1e-13 to 1e-7 there, the fp32 build from 1e-8 to 1e-3 on pressure; the
  perturbation then decays again, but x / y integrate the wind difference and keep an offset).  Away from such events
  the fp32 build tracks the reference to ~1e-8.  Hence: median and 99th-percentile drift are held to 1e-4, the worst
  case only to 10 % of a field's scale, and the discrete state (status, safety-layer states, time) must never differ."""
  monkeypatch.setenv('BLE_STEP_KERNEL', kernel)
  worst, median, p99, mism, decisions = _free_run(ble, precision, kernel, rollout)
  print(f'free-running {precision}/{kernel}{"/rollout" if rollout else ""}: discrete mismatches {mism} of {decisions}')
  for k in FF:
    print(f'  {k:22s} worst {worst[k][0]:.2e} ({worst[k][1]} @ step {worst[k][2]})  median {median[k]:.1e}  p99 {p99[k]:.1e}')
  assert mism == 0, (mism, decisions)
  tol_med = 1e-9 if precision == 'fp64' else 1e-5
  assert max(median.values()) < tol_med, median
  if not rollout:                       # the rollout variant samples the state every 32 steps only
    assert max(p99.values()) < (1e-7 if precision == 'fp64' else 1e-3), p99    # 1 % of the samples sit in the wake of an event
  assert max(w[0] for w in worst.values()) < (1e-5 if precision == 'fp64' else 0.2), worst


def test_rollout_equals_single_steps_and_shapes_agree(ble, monkeypatch):
  """ble_rollout (K steps in one launch) lands bit for bit where K ble_step calls do, for every shape of the production
  step kernel (0 = k_step_warp, 4 / 8 / 14 = k_step_roles), and all shapes agree with each other (they run the same
  role functions in a different warp layout)."""
  n, k = 1000, 12                                  # not a multiple of 32: the last CTA is ragged
  rng = np.random.default_rng(17)
  bank = golden_fields.field_bank()
  fidx = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32))
  acts = torch.from_numpy(rng.integers(0, 3, (k, n)).astype(np.int32))
  results = []
  for warps, use_rollout in ((0, False), (0, True), (4, True), (8, False), (14, True), (14, False)):
    monkeypatch.setenv('BLE_STEP_WARPS', str(warps))
    a = ble.BatchedBalloonArena(n, precision='fp32', enable_noise=True)
    a.set_wind_fields(torch.from_numpy(bank), fidx)
    a.reset(torch.arange(n, dtype=torch.int64) + 41)
    f, i = a.get_state()
    i[3, :7] = 2                                   # a few finished balloons: frozen, reward 0, done 1
    a.set_state(f, i)
    if use_rollout:
      reward, done = a.rollout(acts)
      reward, done = reward.cpu().numpy(), done.cpu().numpy()
    else:
      rs, ds = [], []
      for t in range(k):
        r, dn, _ = a.step(acts[t])
        rs.append(r.cpu().numpy().copy()); ds.append(dn.cpu().numpy().copy())
      reward, done = np.stack(rs), np.stack(ds)
    info = {kk: v.cpu().numpy() for kk, v in a.step_info().items()}
    results.append((warps, use_rollout, reward, done, state_np(a), info))
    a.close()
  ref = results[0]
  assert (ref[3][:, :7] == 1).all() and (ref[2][:, :7] == 0).all()
  np.testing.assert_array_equal(ref[5]['time_elapsed'], ref[4]['time_elapsed'])
  by_shape = {}
  for warps, use_rollout, reward, done, st, info in results:
    # across shapes: the same role functions, but the compiler contracts them differently -> last-ulp differences
    np.testing.assert_allclose(reward, ref[2], rtol=2e-6, atol=1e-7, err_msg=f'{warps} {use_rollout}')
    np.testing.assert_array_equal(done, ref[3])
    for kk in st:
      if st[kk].dtype.kind == 'f':      # x, y integrate the wind: one float ulp of 10 m/s over 12 x 180 s is 2 mm
        np.testing.assert_allclose(st[kk], ref[4][kk], rtol=1e-6, atol=1e-2 if kk in ('x', 'y') else 1e-6,
                                   err_msg=f'{kk} warps={warps} rollout={use_rollout}')
      else:
        np.testing.assert_array_equal(st[kk], ref[4][kk], err_msg=f'{kk} warps={warps} rollout={use_rollout}')
    if warps in by_shape:                      # same shape, rollout vs single steps: bit for bit
      other = by_shape[warps]
      np.testing.assert_array_equal(reward, other[2], err_msg=f'rollout vs steps, shape {warps}')
      for kk in st:
        np.testing.assert_array_equal(st[kk], other[4][kk], err_msg=f'{kk} rollout vs steps, shape {warps}')
      for kk in info:
        np.testing.assert_array_equal(info[kk], other[5][kk])
    by_shape[warps] = (warps, use_rollout, reward, done, st, info)


def test_fp64_trajectories_match_reference(ble):
  """The ten golden scenarios rolled out on the GPU (fp64 audit kernels), own wind lookups."""
  names = sorted(TRAJ)
  scs = [TRAJ[n] for n in names]
  oenv = golden_io.oracle_env_for_scenarios(scs, FF, IF)
  field = np.array([int(sc['field']) for sc in scs])
  alpha = np.array([float(sc['alpha']) for sc in scs]); psl = np.array([int(sc['power_safety']) for sc in scs])
  grid = field >= 0
  for model, sel in (('grid', grid), ('simple_static', ~grid)):
    idx = np.nonzero(sel)[0]
    arena = ble.BatchedBalloonArena(len(idx), precision='fp64', wind_model=model, enable_noise=True)
    if model == 'grid':
      arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()), torch.from_numpy(field[idx].astype(np.int32)))
    arena.set_wind_noise(torch.from_numpy(np.stack([scs[e]['seeds'] for e in idx])),
                         torch.from_numpy(np.stack([scs[e]['offsets'] for e in idx]).astype(np.float32)))
    arena.set_state(*pack(oenv.arena.state.select(idx), alpha[idx], psl[idx]))
    horizon = max(len(scs[e]['actions']) for e in idx)
    for t in range(horizon):
      acts = np.array([scs[e]['actions'][t] if t < len(scs[e]['actions']) else 1 for e in idx], np.int32)
      reward, done, wind = arena.step(torch.from_numpy(acts))
      if t % 16 and t != horizon - 1 and t > 4:
        continue
      st = state_np(arena)
      rw = reward.cpu().numpy()
      for col, e in enumerate(idx):
        sc = scs[e]
        if t >= len(sc['actions']):
          continue
        for j, k in enumerate(FF):
          atol = 1e-2 if k in ('x', 'y') else 1e-6
          np.testing.assert_allclose(st[k][col], sc['f'][t, j], rtol=1e-5, atol=atol, err_msg=f'{names[e]} t={t} {k}')
        for j, k in enumerate(IF):
          if k in ('sunrise_h', 'sunset') and not sc['power_safety']:
            continue
          assert st[k][col] == sc['i'][t, j], (names[e], t, k)
        np.testing.assert_allclose(rw[col], sc['reward'][t], rtol=1e-5, atol=1e-6)
    arena.close()


# ------------------------------------------------------------------------------ reset

def test_reset_derived_state_matches_oracle_and_distributions(ble):
  n = 4096
  arena = ble.BatchedBalloonArena(n, precision='fp64', enable_noise=True)
  arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()[:1]))
  seeds = torch.arange(n, dtype=torch.int64) * 7919 + 12345
  arena.reset(seeds)
  st = state_np(arena)
  # distributions (utils/sampling.py:37-152, env/balloon_arena.py:228-268)
  a = st['atmosphere_alpha']
  assert 0 <= a.min() and a.max() < 1 and abs(a.mean() - 0.5) < 0.03
  r_km = np.hypot(st['x'], st['y']) / 1000.0
  assert r_km.max() <= 200.0 and abs(r_km.mean() - 200 * 1.2 / 3.2) < 4.0          # Beta(1.2, 2) mean
  assert abs(np.degrees(st['center_lat'])).max() <= 10.0 and abs(np.degrees(st['center_lng'])).max() <= 175.0
  t0, t1 = 1293840000, 1293840000 + 126144000
  assert st['date_time'].min() >= t0 and st['date_time'].max() < t1
  assert st['upwelling_infrared'].min() >= 225.0 and st['upwelling_infrared'].max() <= 315.0
  atm = atmosphere_lib.Atmosphere(a)
  pmax, _ = atm.at_height(np.full(n, C.ALT_MIN_ALTITUDE_M))
  assert (st['pressure'] >= 6500.0).all() and (st['pressure'] <= pmax).all()
  assert (st['status'] == 0).all() and (st['battery_charge'] == 2905.6).all() and (st['time_elapsed'] == 0).all()
  # deterministic part: stable init + power-safety sunrise/sunset == oracle on the same draws
  b = balloon_lib.make_batch(n, center_lat=st['center_lat'], center_lng=st['center_lng'], date_time=st['date_time'],
                             x=st['x'], y=st['y'], pressure=st['pressure'], upwelling_infrared=st['upwelling_infrared'])
  stable_init_lib.cold_start_to_stable_params(b, atm)
  for k in ('ambient_temperature', 'internal_temperature', 'mols_air', 'envelope_volume', 'superpressure'):
    np.testing.assert_allclose(st[k], getattr(b, k), rtol=1e-9, atol=1e-7, err_msg=k)
  np.testing.assert_array_equal(st['sunrise_h'], b.sunrise_h)
  np.testing.assert_array_equal(st['sunset'], b.sunset)
  # same seeds -> same state; masked reset leaves unmasked balloons alone
  arena.reset(seeds)
  st2 = state_np(arena)
  for k in st:
    np.testing.assert_array_equal(st[k], st2[k])
  mask = torch.zeros(n, dtype=torch.uint8); mask[::2] = 1
  arena.reset(seeds + 1, mask)
  st3 = state_np(arena)
  assert (st3['x'][1::2] == st['x'][1::2]).all() and (st3['x'][::2] != st['x'][::2]).any()
  arena.close()


def test_reset_distributions_ks_against_reference_samples(ble):
  """Every quantity BalloonArena.reset samples (utils/sampling.py:37-152, env/balloon_arena.py:228-268,
  standard_atmosphere.py:76-87), device reset (Philox) vs 3,000 resets of the UNMODIFIED reference recorded by
  tests/golden/tier0/make_reset_samples.py: two-sample Kolmogorov-Smirnov, plus the joint structure the marginals do not
  see (pressure is uniform on [6500, p(MIN_ALTITUDE | alpha)], the angle of (x, y) is uniform and independent of r)."""
  import os
  from scipy import stats
  g = np.load(os.path.join(golden_io.GOLDEN_DIR, 'reset_samples.npz'))
  ref = {k: g['samples'][:, i] for i, k in enumerate(g['columns'])}
  n = 16384
  arena = ble.BatchedBalloonArena(n, precision='fp64', enable_noise=True)
  arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()[:1]))
  arena.reset(torch.arange(n, dtype=torch.int64) * 104729 + 5)
  st = state_np(arena)
  atm = atmosphere_lib.Atmosphere(st['atmosphere_alpha'])
  pmax, _ = atm.at_height(np.full(n, C.ALT_MIN_ALTITUDE_M))
  ours = {'alpha': st['atmosphere_alpha'], 'date_time': st['date_time'].astype(np.float64), 'x': st['x'], 'y': st['y'],
          'lat_deg': np.degrees(st['center_lat']), 'lng_deg': np.degrees(st['center_lng']), 'pressure': st['pressure'],
          'upwelling_infrared': st['upwelling_infrared'], 'max_pressure': pmax}
  pvalues = {}
  for k, v in ours.items():
    pvalues[k] = float(stats.ks_2samp(v, ref[k]).pvalue)
  derived = {
      'radius': (np.hypot(ours['x'], ours['y']), np.hypot(ref['x'], ref['y'])),
      'angle': (np.arctan2(ours['y'], ours['x']), np.arctan2(ref['y'], ref['x'])),
      'pressure_quantile': ((ours['pressure'] - 6500.0) / (pmax - 6500.0), (ref['pressure'] - 6500.0) / (ref['max_pressure'] - 6500.0)),
  }
  for k, (a, b) in derived.items():
    pvalues[k] = float(stats.ks_2samp(a, b).pvalue)
  print('reset KS p-values:', {k: round(v, 4) for k, v in pvalues.items()})
  # 12 tests: a correct sampler fails p > 1e-3 on any of them with probability ~1 %
  assert min(pvalues.values()) > 1e-3, pvalues
  assert stats.kstest(derived['pressure_quantile'][0], 'uniform').pvalue > 1e-3
  assert stats.kstest(derived['radius'][0] / 200.0e3, stats.beta(1.2, 2.0).cdf).pvalue > 1e-3   # balloon_arena.py:243-245
  assert abs(stats.spearmanr(derived['radius'][0], derived['angle'][0])[0]) < 0.03
  assert (st['battery_charge'] == ref['battery_charge'][0]).all()
  arena.close()


def test_auto_reset_restarts_finished_balloons_inside_the_step(ble):
  """ble_config.auto_reset: a balloon whose step returns done = 1 starts a new episode before the call returns control of
  the stream -- exactly the masked ble_reset a caller would issue, with the seed chain splitmix64(previous seed); the
  others fly on untouched.  reward / done / info describe the step that ended the episode."""
  from balloon_learning_environment_b200 import _lib

  def splitmix64(x):
    m = (1 << 64) - 1
    z = (x + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)

  n = 96
  seeds = torch.arange(n, dtype=torch.int64) * 1009 + 77
  rng = np.random.default_rng(2)
  auto = ble.BatchedBalloonArena(n, precision='fp32', wind_model='simple_static', enable_noise=True, auto_reset=True)
  manual = ble.BatchedBalloonArena(n, precision='fp32', wind_model='simple_static', enable_noise=True)
  ended = np.zeros(n, bool); ended[[3, 17, 40, 41, 95]] = True
  for a in (auto, manual):
    a.reset(seeds)
    f, i = a.get_state()
    i[_lib.I_ROWS.index('status'), torch.from_numpy(ended)] = 2          # BURST: the next step returns done = 1 for these
    a.set_state(f, i)
  acts = torch.from_numpy(rng.integers(0, 3, n).astype(np.int32))
  r_auto, d_auto, _ = auto.step(acts); r_auto, d_auto = r_auto.clone(), d_auto.clone()
  info = auto.step_info()
  r_man, d_man, _ = manual.step(acts)
  np.testing.assert_array_equal(d_auto.cpu().numpy() != 0, ended)
  np.testing.assert_array_equal(d_auto.cpu().numpy(), d_man.cpu().numpy())
  np.testing.assert_array_equal(r_auto.cpu().numpy(), r_man.cpu().numpy())
  np.testing.assert_array_equal(info['envelope_burst'].cpu().numpy(), ended)           # info of the step that ended them
  next_seeds = np.array([splitmix64(int(s)) for s in seeds.numpy()], np.uint64).view(np.int64)
  manual.reset(torch.from_numpy(next_seeds), torch.from_numpy(ended.astype(np.uint8)))
  sa, sm = state_np(auto), state_np(manual)
  for k in sa:
    np.testing.assert_array_equal(sa[k], sm[k], err_msg=k)
  assert (sa['status'] == 0).all() and (sa['time_elapsed'][ended] == 0).all() and (sa['time_elapsed'][~ended] == 180).all()
  # the chain continues: a second episode end of balloon 3 uses splitmix64 of the second seed
  for a in (auto, manual):
    f, i = a.get_state()
    i[_lib.I_ROWS.index('status'), 3] = 1
    a.set_state(f, i)
  auto.step(acts); manual.step(acts)
  again = np.zeros(n, np.uint8); again[3] = 1
  third = next_seeds.copy(); third[3] = np.array([splitmix64(int(next_seeds.view(np.uint64)[3]))], np.uint64).view(np.int64)[0]
  manual.reset(torch.from_numpy(third), torch.from_numpy(again))
  sa, sm = state_np(auto), state_np(manual)
  for k in sa:
    np.testing.assert_array_equal(sa[k], sm[k], err_msg=k)
  with pytest.raises(_lib.BleError):
    auto.rollout(acts[None].repeat(2, 1))
  auto.close(); manual.close()


# ------------------------------------------------------------------------------ fp32 rollout vs oracle

def test_fp32_rollout_tracks_oracle(ble):
  """256 balloons from a device reset, 40 agent steps: fp32 kernels vs the fp64 oracle fed the same
  actions.  State must stay within 1e-4 relative per step when re-synchronised each step, and
  decisions must agree."""
  n, steps = 256, 40
  rng = np.random.default_rng(11)
  arena = ble.BatchedBalloonArena(n, precision='fp32', enable_noise=True)
  bank = golden_fields.field_bank()
  fidx = rng.integers(0, 2, n).astype(np.int32)
  arena.set_wind_fields(torch.from_numpy(bank), torch.from_numpy(fidx))
  arena.reset(torch.arange(n, dtype=torch.int64) + 99)
  seeds = rng.integers(0, 1634753849, (n, 2, 5)); offsets = rng.uniform(-1, 1, (n, 2, 5, 4)).astype(np.float32)
  arena.set_wind_noise(torch.from_numpy(seeds), torch.from_numpy(offsets))
  noise = wind_lib.SimplexWindNoise(seeds, offsets.astype(np.float64))
  mism_total = 0
  worst = {}
  for t in range(steps):
    st = state_np(arena)
    b = balloon_lib.BalloonBatch(**{k: st[k].copy() for k in balloon_lib.FLOAT_FIELDS + balloon_lib.INT_FIELDS})
    atm = atmosphere_lib.Atmosphere(st['atmosphere_alpha'])
    oenv = env_lib.OracleEnv(env_lib.OracleArena(b, atm, fields=bank, field_idx=fidx, noise=noise))
    acts = rng.integers(0, 3, n).astype(np.int32)
    want_r, want_done, _ = oenv.step(acts)
    reward, done, wind = arena.step(torch.from_numpy(acts))
    got = state_np(arena)
    for k in balloon_lib.FLOAT_FIELDS:
      worst[k] = max(worst.get(k, 0.0), rel_err(got[k], getattr(b, k), k))
    for k in ('status', 'last_command', 'envelope_state', 'altitude_state', 'power_paused', 'time_elapsed',
              'date_time', 'sunrise_h', 'sunset'):
      mism_total += int((got[k] != getattr(b, k)).sum())
    assert np.abs(reward.cpu().numpy() - want_r).max() < 2e-4
  assert max(worst.values()) < 1e-4, worst
  assert mism_total == 0, mism_total
  arena.close()


# ------------------------------------------------------------------------------ API behaviour

def test_errors_and_host_step(ble):
  from balloon_learning_environment_b200 import _lib
  arena = ble.BatchedBalloonArena(64, precision='fp32', enable_noise=False)
  with pytest.raises(_lib.BleError):          # stepping before reset/fields (grid_based_wind_field.py:86-87)
    arena.step(torch.zeros(64, dtype=torch.int32))
  arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()[:1]))
  with pytest.raises(_lib.BleError):
    arena.step(torch.zeros(64, dtype=torch.int32))
  arena.reset(torch.arange(64, dtype=torch.int64))
  twin = ble.BatchedBalloonArena(64, precision='fp32', enable_noise=False)
  twin.set_wind_fields(torch.from_numpy(golden_fields.field_bank()[:1]))
  twin.reset(torch.arange(64, dtype=torch.int64))
  acts = np.random.default_rng(0).integers(0, 3, 64).astype(np.int32)
  r_host = np.zeros(64, np.float32); d_host = np.zeros(64, np.uint8)
  arena.step_host(acts, r_host, d_host)
  r_dev, d_dev, _ = twin.step(torch.from_numpy(acts))
  np.testing.assert_array_equal(r_host, r_dev.cpu().numpy())
  np.testing.assert_array_equal(d_host, d_dev.cpu().numpy())
  assert arena.launch_count > 0
  arena.close(); twin.close()


def test_host_step_sequence_matches_device_steps(ble):
  """ble_step_host queues the next step's noise kernel while the host copies results out; a sequence of host
  steps, with a state upload in the middle (which must invalidate the queued noise), lands bit for bit where
  the same sequence of device-resident ble_step calls does."""
  n = 1000
  rng = np.random.default_rng(21)
  bank = golden_fields.field_bank()
  fidx = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32))
  arenas = []
  for _ in range(2):
    a = ble.BatchedBalloonArena(n, precision='fp32', enable_noise=True)
    a.set_wind_fields(torch.from_numpy(bank), fidx)
    a.reset(torch.arange(n, dtype=torch.int64) + 5)
    arenas.append(a)
  host, dev = arenas
  r_host = np.zeros(n, np.float32); d_host = np.zeros(n, np.uint8)
  # page-locked caller buffers take the no-staging path (the kernel reads / writes them in place): steps 4 and 5
  pinned = [torch.zeros(n, dtype=dt).pin_memory().numpy() for dt in (torch.int32, torch.float32, torch.uint8)]
  for t in range(6):
    if t == 3:                                   # move every balloon: the wind queued for the old position is stale
      for a in arenas:
        f, i = a.get_state()
        f = f.clone(); f[0] += 25000.0; f[1] -= 40000.0
        a.set_state(f, i)
    acts = rng.integers(0, 3, n).astype(np.int32)
    if t >= 4:
      pinned[0][:] = acts
      pinned[1][:] = -1.0; pinned[2][:] = 7
      host.step_host(pinned[0], pinned[1], pinned[2])
      r_host, d_host = pinned[1], pinned[2]
    else:
      host.step_host(acts, r_host, d_host)
    r_dev, d_dev, _ = dev.step(torch.from_numpy(acts))
    np.testing.assert_array_equal(r_host, r_dev.cpu().numpy())
    np.testing.assert_array_equal(d_host, d_dev.cpu().numpy())
    sh, sd = state_np(host), state_np(dev)
    for k in sh:
      np.testing.assert_array_equal(sh[k], sd[k], err_msg=f'{k} at step {t}')
  # device path: one fused launch per step; host path: the same launch reading the queued noise + one k_noise queued
  # behind the copies per step (the one the upload made stale is recomputed inside the step)
  assert host.launch_count <= dev.launch_count + 8
  host.close(); dev.close()


def test_full_size_properties(ble):
  """BASELINE.json size (65,536 balloons): size-independent invariants of the transition."""
  n = 65536
  rng = np.random.default_rng(5)
  env = ble.BatchedBalloonEnv(n, precision='fp32', enable_noise=True, seed=3)
  bank = golden_fields.field_bank()
  env.arena.set_wind_fields(torch.from_numpy(bank), torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)))
  env.reset(seed=3)
  noise_seeds = rng.integers(0, 1634753849, (n, 2, 5)); noise_offsets = rng.uniform(-1, 1, (n, 2, 5, 4)).astype(np.float32)
  env.arena.set_wind_noise(torch.from_numpy(noise_seeds), torch.from_numpy(noise_offsets))   # known to the oracle
  s0 = state_np(env.arena)
  acts = torch.from_numpy(rng.integers(0, 3, n).astype(np.int32)).cuda()
  _, reward, done, _ = env.step(acts)
  wind = env.arena._wind.cpu().numpy().astype(np.float64)
  s1 = state_np(env.arena)
  ok = s1['status'] == 0
  assert ok.mean() > 0.99
  np.testing.assert_array_equal(s1['time_elapsed'][ok], 180)
  np.testing.assert_array_equal(s1['date_time'][ok] - s0['date_time'][ok], 180)
  np.testing.assert_allclose((s1['x'] - s0['x'])[ok], wind[ok, 0] * 180.0, rtol=2e-5, atol=0.2)   # x += u dt
  np.testing.assert_allclose((s1['y'] - s0['y'])[ok], wind[ok, 1] * 180.0, rtol=2e-5, atol=0.2)
  assert (s1['battery_charge'] >= 0).all() and (s1['battery_charge'] <= 3058.56 + 1e-3).all()
  r = reward.cpu().numpy()
  assert (r >= 0).all() and (r <= 1).all()
  np.testing.assert_array_equal(s1['last_command'][ok], acts.cpu().numpy()[ok])
  # the same launch against the oracle on a sample of the batch (this size runs k_step, one thread per balloon)
  idx = np.sort(rng.choice(n, 2048, replace=False))
  b = balloon_lib.BalloonBatch(**{k: s0[k][idx].copy() for k in balloon_lib.FLOAT_FIELDS + balloon_lib.INT_FIELDS})
  fidx_all = env.arena._keepalive[1].cpu().numpy()
  oenv = env_lib.OracleEnv(env_lib.OracleArena(
      b, atmosphere_lib.Atmosphere(s0['atmosphere_alpha'][idx]), fields=bank, field_idx=fidx_all[idx],
      noise=wind_lib.SimplexWindNoise(noise_seeds[idx], noise_offsets[idx].astype(np.float64))))
  want_r, want_done, _ = oenv.step(acts.cpu().numpy()[idx])
  for k in balloon_lib.FLOAT_FIELDS:
    assert rel_err(s1[k][idx], getattr(b, k), k) < 1e-4, k
  for k in ('status', 'last_command', 'envelope_state', 'altitude_state', 'power_paused', 'time_elapsed',
            'date_time', 'sunrise_h', 'sunset'):
    np.testing.assert_array_equal(s1[k][idx], getattr(b, k), err_msg=k)
  assert np.abs(r[idx] - want_r).max() < 2e-4
  np.testing.assert_array_equal(done.cpu().numpy()[idx] != 0, want_done)
  # determinism: a twin env with the same seeds and actions lands in the same state, bit for bit
  twin = ble.BatchedBalloonEnv(n, precision='fp32', enable_noise=True, seed=3)
  twin.arena.set_wind_fields(torch.from_numpy(bank), env.arena._keepalive[1])
  twin.reset(seed=3)
  twin.arena.set_wind_noise(torch.from_numpy(noise_seeds), torch.from_numpy(noise_offsets))
  twin.step(acts)
  s1b = state_np(twin.arena)
  for k in s1:
    np.testing.assert_array_equal(s1[k], s1b[k])
  # finished balloons are frozen: force a terminal status and step again
  f, i = env.arena.get_state()
  from balloon_learning_environment_b200 import _lib
  i[_lib.I_ROWS.index('status'), :100] = 2
  env.arena.set_state(f, i)
  before = state_np(env.arena)
  _, reward, done, _ = env.step(acts)
  after = state_np(env.arena)
  for k in before:
    np.testing.assert_array_equal(before[k][:100], after[k][:100])
  assert (done.cpu().numpy()[:100] == 1).all() and (reward.cpu().numpy()[:100] == 0).all()
  env.close(); twin.close()


# ------------------------------------------------------------------------------ N = 1 adaptor (drop-in boundary)

def test_cuda_balloon_arena_follows_reference_episode(ble):
  """The reference's own calling sequence -- obs = arena.step(action); reward_fn(arena.get_simulator_state()) --
  on CudaBalloonArena, against the recorded reference episode 'grid_random' (state, reward AND the
  1099-feature observation, first 60 steps)."""
  import datetime as dt
  import math
  import os
  from balloon_learning_environment_b200 import arena as arena_lib, units
  feat = np.load(os.path.join(golden_io.GOLDEN_DIR, 'features.npz'))
  sc = {k.split('/', 1)[1]: feat[k] for k in feat.files if k.startswith('grid_random/')}
  bank = golden_fields.field_bank()
  a = arena_lib.CudaBalloonArena(wind_field=bank[int(sc['field'])], seed=0, precision='fp64')
  assert a.feature_constructor.observation_space.shape == (1099,)
  a.set_wind_noise(sc['seeds'], sc['offsets'])
  s = a.get_balloon_state()
  f0 = dict(zip(FF, sc['f0'])); i0 = dict(zip(IF, sc['i0']))
  utc = lambda ts: dt.datetime.fromtimestamp(int(ts), tz=dt.timezone.utc)
  s.center_latlng = arena_lib.LatLng(f0['center_lat'], f0['center_lng'])
  s.x, s.y = units.Distance(f0['x']), units.Distance(f0['y'])
  for k in ('pressure', 'ambient_temperature', 'internal_temperature', 'envelope_volume', 'superpressure',
            'mols_air', 'mols_lift_gas', 'upwelling_infrared', 'acs_mass_flow'):
    setattr(s, k, f0[k])
  s.battery_charge = units.Energy(f0['battery_charge']); s.acs_power = units.Power(f0['acs_power'])
  s.solar_charging = units.Power(f0['solar_charging']); s.power_load = units.Power(f0['power_load'])
  s.date_time = utc(i0['date_time']); s.time_elapsed = dt.timedelta(seconds=int(i0['time_elapsed']))
  s.last_command = arena_lib.AltitudeControlCommand(int(i0['last_command']))
  s.status = arena_lib.BalloonStatus(int(i0['status']))
  s.envelope_state, s.altitude_state, s.power_paused = int(i0['envelope_state']), int(i0['altitude_state']), bool(i0['power_paused'])
  s.sunrise_with_hysteresis, s.sunset = utc(i0['sunrise_h']), utc(i0['sunset'])
  s.atmosphere_alpha = float(sc['alpha'])
  a.set_balloon_state(s)
  a.reset_feature_history()
  np.testing.assert_allclose(a.feature_constructor.get_features(), sc['obs'][0], rtol=0, atol=1e-5)

  def reward_fn(sim_state):                     # env/balloon_env.py:44-102 written against the state view
    b = sim_state.balloon_state
    distance_km = math.hypot(b.x.m, b.y.m) / 1000.0
    reward = 1.0 if distance_km <= 50.0 else 0.4 * math.exp(-0.69314718056 / 100.0 * (distance_km - 50.0))
    if b.last_command == arena_lib.AltitudeControlCommand.DOWN and not b.excess_energy:
      reward *= 0.95 - 0.3 * min(max((b.acs_power.watts - 100.0) / 200.0, 0.0), 1.0)
    return reward

  for t in range(60):
    obs = a.step(arena_lib.AltitudeControlCommand(int(sc['actions'][t])))
    assert isinstance(obs, np.ndarray) and obs.shape == (1099,) and obs.dtype == np.float32
    np.testing.assert_allclose(obs, sc['obs'][t + 1], rtol=0, atol=1e-5)
    sim = a.get_simulator_state()
    b = sim.balloon_state
    want = dict(zip(FF, sc['f'][t])); wi = dict(zip(IF, sc['i'][t]))
    assert abs(b.x.km - want['x'] / 1000.0) < 1e-5 and abs(b.pressure - want['pressure']) < 1e-4
    assert abs(b.battery_charge.watt_hours - want['battery_charge']) < 1e-6
    assert b.status.value == wi['status'] and int(b.last_command) == wi['last_command']
    assert int(b.time_elapsed.total_seconds()) == wi['time_elapsed']
    assert abs(reward_fn(sim) - sc['reward'][t]) < 1e-6
  a.close()


# ------------------------------------------------------------------------------ observation surface

@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_perciatelli_features_match_reference(ble, precision):
  """Three recorded reference episodes through PerciatelliFeatureConstructor (1099 float32 features,
  WindGP window filling up to 120 measurements) replayed on the GPU.  Tolerance 1e-4 absolute on
  every feature (north_star / SURVEY.md f-1); the (0, 1, 1) padding pattern must match exactly."""
  import os
  feat = np.load(os.path.join(golden_io.GOLDEN_DIR, 'features.npz'))
  scen = lambda name: {k.split('/', 1)[1]: feat[k] for k in feat.files if k.startswith(name + '/')}
  for model, names in (('grid', ['grid_random', 'sticky_random']), ('simple_static', ['default_static'])):
    scs = [scen(n) for n in names]
    n = len(scs)
    arena = ble.BatchedBalloonArena(n, precision=precision, wind_model=model, enable_noise=True, enable_features=True)
    if model == 'grid':
      arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()),
                            torch.tensor([int(sc['field']) for sc in scs], dtype=torch.int32))
    arena.set_wind_noise(torch.from_numpy(np.stack([sc['seeds'] for sc in scs])),
                         torch.from_numpy(np.stack([sc['offsets'] for sc in scs]).astype(np.float32)))
    b = golden_io.batch_from_rows(FF, IF, np.stack([sc['f0'] for sc in scs]), np.stack([sc['i0'] for sc in scs]))
    arena.set_state(*pack(b, np.array([float(sc['alpha']) for sc in scs]), 1))
    arena.features_clear()
    arena.features_observe()                    # a fresh constructor observes the initial state
    horizon = max(len(sc['actions']) for sc in scs)
    worst = 0.0
    for t in range(horizon + 1):
      if t > 0:
        acts = np.array([sc['actions'][t - 1] if t - 1 < len(sc['actions']) else 1 for sc in scs], np.int32)
        arena.step(torch.from_numpy(acts))
      if t % 10 and t not in (1, 2, 3) and t < horizon - 2:
        continue
      obs = arena.features().cpu().numpy()
      for e, sc in enumerate(scs):
        if t > len(sc['actions']):
          continue
        ref = sc['obs'][t]
        pad_ref = np.all(ref[16:].reshape(361, 3) == np.array([0, 1, 1], np.float32), axis=1)
        pad_got = np.all(obs[e, 16:].reshape(361, 3) == np.array([0, 1, 1], np.float32), axis=1)
        np.testing.assert_array_equal(pad_got, pad_ref, err_msg=f'{names[e]} t={t} padding')
        err = np.abs(obs[e] - ref)
        worst = max(worst, float(err.max()))
        assert err.max() < 1e-4, (names[e], t, int(err.argmax()), float(obs[e][err.argmax()]), float(ref[err.argmax()]))
    print(precision, model, 'worst feature error', worst)
    arena.close()


@pytest.mark.parametrize('steps,every_step', [(12, False), (131, True)])
def test_features_batch_matches_oracle_after_reset(ble, steps, every_step):
  """Device reset -> `steps` steps -> features for 48 balloons, against the oracle fed the same states
  and the same measurement history.  every_step: features() after every step, so that the Cholesky factor
  is carried through 120 appends and then 11 drop + append pairs before the final comparison."""
  from oracle import features as features_lib
  n = 48
  rng = np.random.default_rng(21)
  bank = golden_fields.field_bank()
  fidx = rng.integers(0, 4, n).astype(np.int32)
  arena = ble.BatchedBalloonArena(n, precision='fp32', enable_noise=True, enable_features=True)
  arena.set_wind_fields(torch.from_numpy(bank), torch.from_numpy(fidx))
  arena.reset(torch.arange(n, dtype=torch.int64) * 13 + 5)
  seeds = rng.integers(0, 1634753849, (n, 2, 5)); offsets = rng.uniform(-1, 1, (n, 2, 5, 4)).astype(np.float32)
  arena.set_wind_noise(torch.from_numpy(seeds), torch.from_numpy(offsets))
  arena.features_clear(); arena.features_observe()
  noise = wind_lib.SimplexWindNoise(seeds, offsets.astype(np.float64))
  st = state_np(arena)
  ob = balloon_lib.BalloonBatch(**{k: st[k].copy() for k in balloon_lib.FLOAT_FIELDS + balloon_lib.INT_FIELDS})
  oarena = env_lib.OracleArena(ob, atmosphere_lib.Atmosphere(st['atmosphere_alpha']), fields=bank, field_idx=fidx, noise=noise)
  ofeat = features_lib.PerciatelliFeatures(oarena)
  ofeat.observe()
  for t in range(steps):
    acts = rng.integers(0, 3, n).astype(np.int32)
    arena.step(torch.from_numpy(acts))
    # keep the oracle on the GPU's trajectory: copy the state, then let it observe
    st = state_np(arena)
    for k in balloon_lib.FLOAT_FIELDS + balloon_lib.INT_FIELDS:
      setattr(ob, k, st[k].copy())
    ofeat.observe()
    if every_step:
      arena.features()
  got = arena.features().cpu().numpy()
  want = ofeat.get_features()
  pad = lambda o: np.all(o[:, 16:].reshape(-1, 361, 3) == np.array([0, 1, 1], np.float32), axis=2)
  np.testing.assert_array_equal(pad(got), pad(want))
  err = np.abs(got - want)
  print(f'features vs fp64 oracle after {steps} steps: worst {err.max():.2e} (ambient {err[:, :16].max():.2e}, '
        f'uncertainty {err[:, 16::3].max():.2e}, angle {err[:, 17::3].max():.2e}, magnitude {err[:, 18::3].max():.2e})')
  assert err.max() < 1e-4, float(err.max())
  arena.close()


@pytest.mark.parametrize('layout', ['x64', 'x128'])
def test_column_tile_is_bit_identical_to_gathers(ble, monkeypatch, layout):
  """The forecast column of the observation comes from ONE TMA tensor copy of the balloon's 9 lookup windows (the 5-D
  tensor view of the field bank, `cp.async.bulk.tensor.5d`); with BLE_COLUMN_TILE=0 the same kernel gathers one
  128-byte window per level.  Both feed the same interp_window, so the 1099 features must agree bit for bit -- across
  cells (balloons drift over cell borders during 30 steps), both layouts, balloons at the clipped edges of the grid and
  past the 48 h boomerang of the time axis."""
  n = 64
  bank = golden_fields.field_bank()
  rng = np.random.default_rng(44)
  fidx = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32))
  seeds = torch.arange(n, dtype=torch.int64) * 11 + 5
  arenas = []
  for tile in ('1', '0'):
    monkeypatch.setenv('BLE_COLUMN_TILE', tile)
    a = ble.BatchedBalloonArena(n, precision='fp32', enable_noise=True, enable_features=True, field_layout=layout)
    a.set_wind_fields(torch.from_numpy(bank), fidx)
    a.reset(seeds)
    from balloon_learning_environment_b200 import _lib
    f, i = a.get_state()
    x, y, te = _lib.F_ROWS.index('x'), _lib.F_ROWS.index('y'), _lib.I_ROWS.index('time_elapsed')
    f[x, :8] = torch.tensor([-499.9e3, 499.9e3, -650e3, 650e3, 0.0, 25e3, -25e3, 449.99e3], dtype=torch.float64)
    f[y, :8] = torch.tensor([499.9e3, -499.9e3, 650e3, -650e3, 0.0, -475e3, 475e3, 0.01e3], dtype=torch.float64)
    i[te, 8:12] = torch.tensor([47 * 3600 + 3000, 48 * 3600, 50 * 3600, 143 * 3600], dtype=torch.int64)
    a.set_state(f, i)
    a.features_clear(); a.features_observe()
    arenas.append(a)
  monkeypatch.delenv('BLE_COLUMN_TILE')
  for t in range(30):
    acts = torch.from_numpy(rng.integers(0, 3, n).astype(np.int32))
    got, want = [], []
    for a, out in zip(arenas, (got, want)):
      a.step(acts)
      out.append(a.features().cpu().numpy())
    np.testing.assert_array_equal(got[0], want[0], err_msg=f'step {t}')
  assert np.abs(got[0][:, 16:]).max() > 0.1
  for a in arenas:
    a.close()


def test_incremental_gp_matches_full_refit(ble, monkeypatch):
  """k_gp_posterior (kernel matrix kept in ring-slot order in HBM, updated one row per observe; blocked fp64 Cholesky +
  3 x TF32 column sweep per call) against the first-generation kernels that rebuild K from the measurement ring at
  every call (BLE_GP_REFIT=1), over 150 steps: window filling, then sliding (the ring overwrites the oldest slot),
  features skipped for a few steps, a masked reset in the middle."""
  n, steps = 40, 150
  bank = golden_fields.field_bank()
  rng = np.random.default_rng(33)
  fidx = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32))
  seeds = torch.arange(n, dtype=torch.int64) * 7 + 3
  arenas = []
  for refit in ('0', '1'):
    monkeypatch.setenv('BLE_GP_REFIT', refit)
    a = ble.BatchedBalloonArena(n, precision='fp32', enable_noise=True, enable_features=True)
    a.set_wind_fields(torch.from_numpy(bank), fidx)
    a.reset(seeds)
    arenas.append(a)
  monkeypatch.delenv('BLE_GP_REFIT')
  worst = 0.0
  skip = set(range(60, 66)) | set(range(131, 141))          # 6 and 10 steps without a features() call
  for t in range(steps):
    acts = torch.from_numpy(rng.integers(0, 3, n).astype(np.int32))
    for a in arenas:
      a.step(acts)
    if t == 100:                                             # a third of the balloons start a new episode
      mask = torch.from_numpy((np.arange(n) % 3 == 0).astype(np.uint8))
      for a in arenas:
        a.reset(seeds + 1000, mask)
    if t in skip:
      continue
    got, want = (a.features().cpu().numpy() for a in arenas)
    err = float(np.abs(got - want).max())
    worst = max(worst, err)
    # two fp32-class column solves (first generation: V in fp32; k_gp_posterior: 3 x TF32), each held to 1e-4 against
    # the reference / the fp64 oracle by the tests above
    assert err < 1.5e-4, (t, err, int(np.abs(got - want).max(axis=1).argmax()))
  st = arenas[0].get_state_dict()
  assert int((st['status'] == 0).sum()) > n // 2             # most balloons flew the whole test
  print('incremental vs refit: worst feature difference', worst)
  for a in arenas:
    a.close()


def test_full_size_one_field_per_balloon(ble):
  """BASELINE configs[2] as the bench flies it: 65,536 balloons, ONE generated wind field per balloon (127 GB of X64
  windows), simplex noise on.  One step of the production kernel; a 512-balloon sample is checked against the oracle
  flying the native-layout view of the same fields (ble_decode_fields(ble_sample_latents(seed)))."""
  from oracle import vae as vae_oracle
  n = 65536
  free, _ = torch.cuda.mem_get_info()
  if free < 150e9:
    pytest.skip('needs 150 GB of free HBM')
  rng = np.random.default_rng(21)
  env = ble.BatchedBalloonEnv(n, precision='fp32', enable_noise=True, seed=5, decoder_params=vae_oracle.synthetic_params(6))
  seeds = torch.arange(n, dtype=torch.int64) * 7919 + 3
  env.reset(seeds=seeds)                                     # generates one field per balloon from its seed
  noise_seeds = rng.integers(0, 1634753849, (n, 2, 5)); noise_offsets = rng.uniform(-1, 1, (n, 2, 5, 4)).astype(np.float32)
  env.arena.set_wind_noise(torch.from_numpy(noise_seeds), torch.from_numpy(noise_offsets))
  s0 = state_np(env.arena)
  acts = torch.from_numpy(rng.integers(0, 3, n).astype(np.int32)).cuda()
  _, reward, done, info = env.step(acts)
  s1 = state_np(env.arena)
  idx = np.sort(rng.choice(n, 512, replace=False))
  fields = env.arena.decode_wind_fields(env.arena.sample_latents(seeds[idx])).cpu().numpy()
  assert np.abs(fields).max() > 1.0
  b = balloon_lib.BalloonBatch(**{k: s0[k][idx].copy() for k in balloon_lib.FLOAT_FIELDS + balloon_lib.INT_FIELDS})
  oenv = env_lib.OracleEnv(env_lib.OracleArena(
      b, atmosphere_lib.Atmosphere(s0['atmosphere_alpha'][idx]), fields=fields, field_idx=np.arange(512),
      noise=wind_lib.SimplexWindNoise(noise_seeds[idx], noise_offsets[idx].astype(np.float64))))
  want_r, want_done, _ = oenv.step(acts.cpu().numpy()[idx])
  for k in balloon_lib.FLOAT_FIELDS:
    assert rel_err(s1[k][idx], getattr(b, k), k) < 1e-4, k
  for k in ('status', 'last_command', 'envelope_state', 'altitude_state', 'power_paused', 'time_elapsed', 'date_time'):
    np.testing.assert_array_equal(s1[k][idx], getattr(b, k), err_msg=k)
  assert np.abs(reward.cpu().numpy()[idx] - want_r).max() < 2e-4
  np.testing.assert_array_equal(done.cpu().numpy()[idx] != 0, want_done)
  assert int(info['time_elapsed'][idx[0]]) == 180
  # every balloon flies its OWN field: the winds of two balloons at the same point differ
  q = torch.tensor([[0.0, 0.0, 9000.0, 1.0]] * 4, dtype=torch.float32)
  uv = env.arena.wind_forecast(q, torch.tensor([0, 1, 40000, 65535], dtype=torch.int32)).cpu().numpy()
  assert len({tuple(np.round(r, 4)) for r in uv}) == 4
  env.close()


# ------------------------------------------------------------------------------ VAE decoder (reset path)

@pytest.mark.parametrize('decoder_precision,tol', [('fp32', 2e-4), ('tf32', 4e-3)])
def test_decoder_matches_oracle(ble, decoder_precision, tol):
  """vae.Decoder (generative/vae.py:134-186) as cuBLASLt GEMMs + resize/curl kernel, against the
  NumPy restatement (oracle/vae.py, itself checked against the reference's flax module under the
  Tier-0 stubs) on seeded random-init weights of the reference architecture.  'fp32' is what the reference computes on
  a CPU (tolerance: GEMM summation order); 'tf32' (the default) is jax's default matmul precision on a GPU: inputs
  rounded to 10 mantissa bits, four layers deep."""
  from oracle import vae as vae_oracle
  params = vae_oracle.synthetic_params(3)
  arena = ble.BatchedBalloonArena(4, precision='fp32', enable_noise=False, decoder_precision=decoder_precision)
  arena.set_decoder(params)
  rng = np.random.default_rng(8)
  z = rng.standard_normal((5000, 64)).astype(np.float32)           # spans two 4096-field chunks
  got = arena.decode_wind_fields(torch.from_numpy(z)).cpu().numpy()
  assert got.shape == (5000, 21, 21, 10, 9, 2)
  idx = np.array([0, 1, 2, 4095, 4096, 4999])
  want = vae_oracle.decode(params, z[idx])
  scale = np.abs(want).max()
  assert scale > 1.0
  err = float(np.abs(got[idx] - want).max() / scale)
  print(f'decoder {decoder_precision}: worst |err| / max|field| = {err:.2e}')
  assert err < tol
  # the decoded field is a discrete curl of a stream function: central-difference divergence vanishes
  # u = D_x Psi, v = -D_y Psi with commuting central differences => D_y u + D_x v == 0
  u, v = got[idx][..., 0].astype(np.float64), got[idx][..., 1].astype(np.float64)
  div = (u[:, 1:-1, 2:] - u[:, 1:-1, :-2]) / 2 + (v[:, 2:, 1:-1] - v[:, :-2, 1:-1]) / 2
  assert np.abs(div).max() < 1e-4 * scale
  # generated bank feeds the gather: decode -> bank -> lookup at a grid node returns the node value
  arena.generate_wind_fields(16, seed=5, chunk=8)
  g = torch.Generator(device='cuda'); g.manual_seed(5)
  z0 = torch.randn(8, 64, generator=g, device='cuda')
  f0 = arena.decode_wind_fields(z0).cpu().numpy()
  q = torch.tensor([[-500.0 + 50.0 * 3, -500.0 + 50.0 * 7, 5000.0 + 1000.0 * 4, 6.0 * 2]], dtype=torch.float32)
  uv = arena.wind_forecast(q, torch.tensor([2], dtype=torch.int32)).cpu().numpy()[0]
  np.testing.assert_allclose(uv, f0[2, 3, 7, 4, 2], rtol=1e-5, atol=1e-5)
  arena.close()


@pytest.mark.parametrize('decoder_precision,tol', [('fp32', 2e-5), ('tf32', 4e-3)])
def test_decoder_on_the_real_checkpoint(ble, decoder_precision, tol):
  """The reference's offlineskies22 weights, read by the product's msgpack loader from oracle/_ref/ (staged by
  __graft_entry__.build() where the reference is present), against fields the reference's own flax module decoded
  (tests/golden/decoder_real.npz, make_decoder_golden.py) -- and the generation path on the same weights."""
  import os
  from balloon_learning_environment_b200 import models
  path = os.path.join(os.path.dirname(golden_io.GOLDEN_DIR.rstrip('/')), '..', 'oracle', '_ref', 'offlineskies22_decoder.msgpack')
  path = os.path.abspath(path)
  if not os.path.exists(path):
    pytest.skip('oracle/_ref/offlineskies22_decoder.msgpack not staged (run __graft_entry__.build() where the reference is)')
  gold = np.load(os.path.join(golden_io.GOLDEN_DIR, 'decoder_real.npz'))
  assert os.path.getsize(path) == int(gold['checkpoint_bytes'])
  params = models.load_decoder(path)
  arena = ble.BatchedBalloonArena(4, precision='fp32', enable_noise=False, decoder_precision=decoder_precision)
  arena.set_decoder(params)
  got = arena.decode_wind_fields(torch.from_numpy(gold['latents'])).cpu().numpy()
  scale = float(np.abs(gold['fields']).max())
  err = float(np.abs(got - gold['fields']).max() / scale)
  print(f'real checkpoint, {decoder_precision}: worst |err| / max|wind| = {err:.2e} (max |wind| {scale:.1f} m/s)')
  assert scale > 20.0 and err < tol
  arena.close()


@pytest.mark.parametrize('layout', ['x64', 'x128'])
def test_fused_field_generation_fills_the_same_windows(ble, layout):
  """ble_generate_fields writes the gather's windows straight from the decoder output (k_flow_to_windows).  Reading
  the bank back at every grid node gives the native field; it must equal ble_decode_fields(ble_sample_latents(seeds))
  (the native-layout decoder) up to the GEMM batch shape, loading THAT field through the two-pass path
  (ble_write_fields) must give bit-identical lookups everywhere, and the field must be a discrete curl."""
  from oracle import vae as vae_oracle
  params = vae_oracle.synthetic_params(4)
  n = 6
  arena = ble.BatchedBalloonArena(n, precision='fp32', enable_noise=False, field_layout=layout, decoder_precision='fp32')
  arena.set_decoder(params)
  arena.alloc_wind_fields(n)
  seeds = torch.tensor([11, 22, 33, 44, 55, 66], dtype=torch.int64)
  arena.sample_wind_fields(seeds)
  # scattered regeneration of two fields leaves the others untouched and is reproducible per seed
  ix, iy, ip, it = np.meshgrid(np.arange(21), np.arange(21), np.arange(10), np.arange(9), indexing='ij')
  nodes = np.stack([-500.0 + 50.0 * ix, -500.0 + 50.0 * iy, 5000.0 + 1000.0 * ip, 6.0 * it], -1).reshape(-1, 4).astype(np.float32)
  q = torch.from_numpy(nodes)

  def read_back(a, f):
    return a.wind_forecast(q, torch.full((len(nodes),), f, dtype=torch.int32)).cpu().numpy().reshape(21, 21, 10, 9, 2)

  fields = np.stack([read_back(arena, f) for f in range(n)])
  assert np.abs(fields).max() > 0.5
  z = arena.sample_latents(seeds)
  native = arena.decode_wind_fields(z).cpu().numpy()
  assert np.abs(fields - native).max() < 2e-5 * np.abs(native).max()          # 6-row vs 2048-row GEMM launch
  want = vae_oracle.decode(params, z.cpu().numpy())
  assert np.abs(fields - want).max() < 2e-4 * np.abs(want).max()
  u, v = fields[..., 0].astype(np.float64), fields[..., 1].astype(np.float64)
  div = (u[:, 1:-1, 2:] - u[:, 1:-1, :-2]) / 2 + (v[:, 2:, 1:-1] - v[:, :-2, 1:-1]) / 2
  assert np.abs(div).max() < 1e-4 * np.abs(fields).max()
  other = ble.BatchedBalloonArena(n, precision='fp32', enable_noise=False, field_layout=layout)
  other.set_wind_fields(torch.from_numpy(fields))
  rng = np.random.default_rng(3)
  pts = np.stack([rng.uniform(-520, 520, 20000), rng.uniform(-520, 520, 20000), rng.uniform(4000, 15000, 20000),
                  rng.uniform(0, 100, 20000)], -1).astype(np.float32)
  fidx = torch.from_numpy(rng.integers(0, n, 20000).astype(np.int32))
  got = arena.wind_forecast(torch.from_numpy(pts), fidx).cpu().numpy()
  want = other.wind_forecast(torch.from_numpy(pts), fidx).cpu().numpy()
  np.testing.assert_array_equal(got, want)
  arena.sample_wind_fields_at(torch.tensor([33, 11], dtype=torch.int64), torch.tensor([4, 1], dtype=torch.int32))
  np.testing.assert_array_equal(read_back(arena, 4), fields[2])      # seed 33 again, now in slot 4
  np.testing.assert_array_equal(read_back(arena, 1), fields[0])
  np.testing.assert_array_equal(read_back(arena, 3), fields[3])      # untouched
  arena.close(); other.close()


def test_latents_are_standard_normal_and_keyed_by_seed(ble):
  """z ~ N(0, I_64) per seed (env/generative_wind_field.py:57-58): Kolmogorov-Smirnov against the normal CDF over
  4,096 seeds x 64 latents, per-dimension moments, no correlation between neighbouring seeds, reproducible per seed."""
  from scipy import stats
  arena = ble.BatchedBalloonArena(4, precision='fp32', enable_noise=False)
  seeds = torch.arange(1000, 1000 + 4096, dtype=torch.int64)
  z = arena.sample_latents(seeds).cpu().numpy().astype(np.float64)
  assert z.shape == (4096, 64)
  assert stats.kstest(z.ravel(), 'norm').pvalue > 1e-3
  for dim in (0, 1, 31, 63):
    assert stats.kstest(z[:, dim], 'norm').pvalue > 1e-4
  assert np.abs(z.mean(0)).max() < 0.08 and np.abs(z.std(0) - 1).max() < 0.06
  assert abs(np.corrcoef(z[:-1].ravel(), z[1:].ravel())[0, 1]) < 0.01        # seed s vs seed s + 1
  assert abs(np.corrcoef(z[:, :-1].ravel(), z[:, 1:].ravel())[0, 1]) < 0.01  # latent k vs latent k + 1
  again = arena.sample_latents(seeds[100:103]).cpu().numpy()
  np.testing.assert_array_equal(again, z[100:103].astype(np.float32))
  arena.close()


# ---- evaluation surface ---------------------------------------------------
def _agent_gold():
  import os
  return np.load(os.path.join(golden_io.GOLDEN_DIR, 'agents.npz'))


def test_station_seeker_kernel_matches_reference(ble):
  """k_agent_station_seeker on the 499 golden observations (242 recorded by the reference's feature
  constructor, 257 synthetic with exact ties): action and chosen level bit-exact against the reference's
  StationSeekerAgent; RandomWalk rule against the oracle."""
  from oracle import agents as oracle_agents
  gold = _agent_gold()
  obs = torch.from_numpy(np.ascontiguousarray(gold['open_obs'], np.float32))
  arena = ble.BatchedBalloonArena(len(obs), precision='fp32', wind_model='simple_static', enable_noise=False)
  actions, best = arena.station_seeker_actions(obs, with_level=True)
  np.testing.assert_array_equal(actions.cpu().numpy(), gold['open_actions'])
  np.testing.assert_array_equal(best.cpu().numpy(), gold['open_best'])
  # an all-invalid column: the reference asserts, the kernel reports level -1 and holds altitude
  blank = obs.clone(); blank[:, 16:] = torch.tensor([0.0, 1.0, 1.0]).repeat(361)
  a2, b2 = arena.station_seeker_actions(blank, with_level=True)
  assert (b2 == -1).all() and (a2 == 1).all()
  # random walk: step 0 draws targets in [6500, 11400]; the action is the band rule around the target
  seeds = torch.arange(len(obs), dtype=torch.int64) + 7
  a0 = arena.random_walk_actions(obs, seeds, 0).cpu().numpy()
  a0_again = arena.random_walk_actions(obs, seeds, 0).cpu().numpy()
  np.testing.assert_array_equal(a0, a0_again)                      # Philox(seed, step): stateless and repeatable
  p = oracle_agents.balloon_pressure_from_obs(gold['open_obs'])
  assert np.all(a0[p - 100 > 11400] == 2) and np.all(a0[p + 100 < 6500] == 0)
  hist = [np.bincount(arena.random_walk_actions(obs, seeds, k).cpu().numpy(), minlength=3) for k in range(1, 40)]
  assert np.sum(hist, axis=0).min() > 0                             # the walk visits all three commands
  with pytest.raises(ValueError):
    arena.station_seeker_actions(obs[:5])
  arena.close()


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_eval_agent_closed_loop_matches_reference(ble, precision):
  """The reference's eval_lib.eval_agent flying its StationSeekerAgent for 100 steps (two injected episodes),
  against the CUDA closed loop: ble_step -> ble_features_perciatelli -> ble_agent_station_seeker ->
  ble_eval_accumulate.  Actions bit-exact; cumulative reward 1e-5 relative (float32 step rewards), time
  within radius / final step exact, flight path 1e-4 relative (north_star state tolerance)."""
  from balloon_learning_environment_b200 import agents as agents_lib
  gold = _agent_gold()
  names = [str(n) for n in gold['names']]
  scs = [{k.split('/', 1)[1]: gold[k] for k in gold.files if k.startswith(n + '/')} for n in names]
  n = len(scs)
  arena = ble.BatchedBalloonArena(n, precision=precision, wind_model='grid', enable_noise=True, enable_features=True)
  arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()),
                        torch.tensor([int(sc['field']) for sc in scs], dtype=torch.int32))
  arena.set_wind_noise(torch.from_numpy(np.stack([sc['seeds'] for sc in scs])),
                       torch.from_numpy(np.stack([sc['offsets'] for sc in scs]).astype(np.float32)))
  b = golden_io.batch_from_rows(FF, IF, np.stack([sc['f0'] for sc in scs]), np.stack([sc['i0'] for sc in scs]))
  arena.set_state(*pack(b, np.array([float(sc['alpha']) for sc in scs]), 1))
  arena.features_clear(); arena.features_observe()
  agent = agents_lib.StationSeekerAgent(3, (1099,), arena)
  steps = len(scs[0]['actions'])
  path = torch.empty(steps, 6, n, dtype=torch.float32, device=arena.device)
  arena.eval_begin()
  action = agent.begin_episode(arena.features())
  taken = []
  for t in range(steps):
    taken.append(action.cpu().numpy().copy())
    reward, done, _ = arena.step(action)
    arena.eval_accumulate(reward, path[t])
    action = agent.step(reward, arena.features())
  taken = np.asarray(taken)
  res = {k: v.cpu().numpy() for k, v in arena.eval_results().items()}
  path = path.cpu().numpy()
  for e, sc in enumerate(scs):
    np.testing.assert_array_equal(taken[:, e], sc['actions'], err_msg=names[e])
    np.testing.assert_allclose(res['cumulative_reward'][e], sc['cumulative_reward'], rtol=1e-5)
    assert res['time_within_radius'][e] == sc['time_within_radius']
    assert res['final_timestep'][e] == sc['final_timestep'] and res['active'][e] == 1
    assert (res['out_of_power'][e], res['envelope_burst'][e], res['zeropressure'][e]) == tuple(sc['flags'])
    np.testing.assert_allclose(path[:, :, e], sc['flight_path'], rtol=1e-4, atol=2e-3, err_msg=names[e])
  arena.close()


def test_vectorised_eval_driver(ble):
  """eval_lib.eval_agent over a 24-seed suite with decoder-generated wind fields: per-seed results do not depend
  on how the suite is sharded (seed -> field, state and noise are functions of the seed alone), the JSON has
  the reference's schema (tests/golden/eval.json), terminated flights stop accumulating."""
  import json
  import os
  from balloon_learning_environment_b200 import agents as agents_lib, eval_lib, suites
  from oracle import vae as vae_oracle
  params = vae_oracle.synthetic_params(3)
  suite = suites.EvaluationSuite(list(range(100, 124)), 40)

  def fly(sub):
    env = ble.BatchedBalloonEnv(len(sub.seeds), observation='perciatelli', decoder_params=params, field_layout='x128')
    agent = agents_lib.create_agent('station_seeker', 3, (1099,), env.arena)
    out = eval_lib.eval_agent(agent, env, sub)
    env.close()
    return out

  whole = fly(suite)
  parts = fly(suites.shard(suite, 0, 2)) + fly(suites.shard(suite, 1, 2))
  assert [r.seed for r in whole] == list(suite.seeds) == [r.seed for r in parts]
  for a, b in zip(whole, parts):
    assert a.cumulative_reward == b.cumulative_reward and a.time_within_radius == b.time_within_radius
    assert a.final_timestep == b.final_timestep == 40 and len(a.flight_path) == 40
  assert 0.0 < np.mean([r.cumulative_reward for r in whole]) <= 40.0
  assert all(0.0 <= r.time_within_radius <= 1.0 for r in whole)
  with open(os.path.join(golden_io.GOLDEN_DIR, 'eval.json')) as f:
    ref = json.load(f)
  mine = json.loads(eval_lib.results_to_json(whole))
  assert list(mine[0].keys()) == list(ref[0].keys())
  assert list(mine[0]['flight_path'][0].keys()) == list(ref[0]['flight_path'][0].keys())
  merged = json.loads(eval_lib.combine_shards([eval_lib.results_to_json(parts[12:]), eval_lib.results_to_json(parts[:12])]))
  assert [r['seed'] for r in merged] == list(suite.seeds)
