"""GPU parity of the QR-DQN learner kernels (csrc/ble_learner.cu, through the C ABI) against
oracle/qrdqn.py.  Tolerance: fp32 kernels vs the fp64 oracle, 2e-5 relative on losses / gradients;
index work (greedy actions, replay picks, exploration phases) bit-exact.
"""
import copy

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import qrdqn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def lrn():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from balloon_learning_environment_b200 import _lib, learner
  _lib.load()
  return learner


def dev(a, dtype=None):
  t = torch.as_tensor(np.asarray(a))
  return (t.to(dtype) if dtype is not None else t).cuda()


def test_greedy_target_loss_match_oracle(lrn):
  rng = np.random.default_rng(0)
  b = 777                                           # not a multiple of the 4 samples per block
  logits = rng.normal(0, 2, (b, 3, 51)).astype(np.float32)
  logits[5, 0] = logits[5, 2] = logits[5, 1] + 1.0  # exact tie -> first maximum
  next_logits = rng.normal(0, 2, (b, 3, 51)).astype(np.float32)
  reward = rng.uniform(0, 1, b).astype(np.float32)
  discount = np.where(rng.uniform(size=b) < 0.2, 0.0, 0.993 ** 5).astype(np.float32)
  actions = rng.integers(0, 3, b).astype(np.int32)

  got_a, got_q = lrn.greedy_actions(dev(logits), with_q=True)
  np.testing.assert_array_equal(got_a.cpu().numpy(), qrdqn.greedy_actions(logits.astype(np.float64)))
  np.testing.assert_allclose(got_q.cpu().numpy(), logits.astype(np.float64).mean(2), rtol=1e-5, atol=1e-6)

  want_t = qrdqn.target_distribution(next_logits.astype(np.float64), reward, discount)
  got_t = lrn.target_distribution(dev(next_logits), dev(reward), dev(discount))
  np.testing.assert_allclose(got_t.cpu().numpy(), want_t, rtol=1e-6, atol=1e-6)

  want_l, want_g = qrdqn.quantile_huber_loss(logits.astype(np.float64), actions, want_t)
  tl = dev(logits).requires_grad_(True)
  mean, per = lrn.quantile_huber_loss(tl, dev(actions), got_t)
  mean.backward()
  np.testing.assert_allclose(per.cpu().numpy(), want_l, rtol=2e-5, atol=1e-6)
  assert abs(float(mean.detach()) - want_l.mean()) < 2e-5 * want_l.mean()
  np.testing.assert_allclose(tl.grad.cpu().numpy(), want_g, rtol=2e-5, atol=1e-9)

  # weights: invalid samples contribute neither loss nor gradient
  weight = (rng.uniform(size=b) < 0.7).astype(np.float32)
  tl2 = dev(logits).requires_grad_(True)
  mean2, _ = lrn.quantile_huber_loss(tl2, dev(actions), got_t, dev(weight))
  mean2.backward()
  np.testing.assert_allclose(tl2.grad.cpu().numpy(), want_g * weight[:, None, None], rtol=2e-5, atol=1e-9)
  assert abs(float(mean2.detach()) - (want_l * weight).mean()) < 2e-5


def test_loss_other_shapes(lrn):
  rng = np.random.default_rng(1)
  for a, n, kappa in ((2, 7, 0.5), (5, 32, 1.0), (3, 64, 2.0), (1, 33, 1.0)):
    logits = rng.normal(0, 1, (9, a, n)).astype(np.float32)
    target = rng.normal(0, 1, (9, n)).astype(np.float32)
    actions = rng.integers(0, a, 9).astype(np.int32)
    want_l, want_g = qrdqn.quantile_huber_loss(logits.astype(np.float64), actions, target.astype(np.float64), kappa)
    tl = dev(logits).requires_grad_(True)
    mean, per = lrn.quantile_huber_loss(tl, dev(actions), dev(target), None, kappa)
    mean.backward()
    np.testing.assert_allclose(per.cpu().numpy(), want_l, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(tl.grad.cpu().numpy(), want_g, rtol=2e-5, atol=1e-8)
  with pytest.raises(Exception):                        # more atoms than a warp holds two per lane
    lrn.quantile_huber_loss(torch.zeros(2, 3, 65, device='cuda'), torch.zeros(2, dtype=torch.int32, device='cuda'),
                            torch.zeros(2, 65, device='cuda'))


def _filled_replay(lrn, rng, envs=6, capacity=16, steps=27, features=23):
  rep = lrn.DeviceReplay(envs, capacity, num_features=features, n_step=5, gamma=0.993, seed=3)
  hist = []
  for t in range(steps):
    obs = rng.normal(0, 1, (envs, features)).astype(np.float32)
    act = rng.integers(0, 3, envs).astype(np.int32)
    rew = rng.uniform(0, 1, envs).astype(np.float32)
    term = rng.uniform(size=envs) < 0.08
    trunc = (rng.uniform(size=envs) < 0.08) & ~term
    rep.add(dev(obs), dev(act), dev(rew), dev(term.astype(np.uint8)), dev(trunc.astype(np.uint8)))
    hist.append((obs, act, rew, term, trunc))
  return rep, hist


def test_replay_sample_matches_oracle(lrn):
  rng = np.random.default_rng(2)
  envs, cap, steps, feat = 6, 16, 27, 23
  rep, hist = _filled_replay(lrn, rng, envs, cap, steps, feat)
  reward = np.zeros((cap, envs)); term = np.zeros((cap, envs), bool); trunc = np.zeros((cap, envs), bool)
  for t in range(steps - cap, steps):
    reward[t % cap], term[t % cap], trunc[t % cap] = hist[t][2], hist[t][3], hist[t][4]
  # every (t, e), including stale and future steps
  idx = np.array([(t, e) for t in range(0, steps + 2) for e in range(envs)], np.int64)
  out = rep.sample(len(idx), indices=dev(idx))
  valid = out['valid'].cpu().numpy()
  n_valid = 0
  for k, (t, e) in enumerate(idx):
    want = qrdqn.nstep_transition(reward, term, trunc, steps, cap, int(t), int(e))
    assert bool(valid[k]) == (want is not None), (t, e)
    if want is None:
      continue
    n_valid += 1
    _, ret, disc, t_next = want
    assert abs(float(out['return'][k]) - ret) < 1e-5
    assert abs(float(out['discount'][k]) - disc) < 1e-6
    assert int(out['action'][k]) == int(hist[t][1][e])
    np.testing.assert_array_equal(out['state'][k].cpu().numpy(), hist[t][0][e])
    np.testing.assert_array_equal(out['next_state'][k].cpu().numpy(), hist[t_next][0][e])
  assert n_valid > 20
  # random draws: all valid, reproducible per draw counter, spread over steps and balloons
  out = rep.sample(4096)
  assert int(out['valid'].sum()) == 4096
  picked = out['indices'].cpu().numpy()
  for t, e in picked[:200]:
    assert qrdqn.nstep_transition(reward, term, trunc, steps, cap, int(t), int(e)) is not None
  assert len(np.unique(picked[:, 1])) == envs and len(np.unique(picked[:, 0])) >= 5
  # empty ring: nothing to sample
  empty = lrn.DeviceReplay(4, 8, num_features=feat)
  assert int(empty.sample(16)['valid'].sum()) == 0
  # sampling into the learner's own row-padded input buffers (out_pitch of the replay view): same rows, padding untouched
  cfg = lrn.QrDqnConfig(num_layers=2, hidden_units=16, num_features=feat)
  learner = lrn.QrDqnLearner(cfg, seed=0)
  sel = dev(idx[valid.astype(bool)][:40])
  plain = rep.sample(len(sel), indices=sel)
  bufs = learner.sample_buffers(len(sel))
  assert bufs['state'].stride(0) == 32 and bufs['state'].shape == (len(sel), feat)
  placed = rep.sample(len(sel), indices=sel, out=bufs)
  for k in ('state', 'next_state', 'action', 'return', 'discount', 'valid'):
    assert placed[k].data_ptr() == bufs[k].data_ptr()
    assert torch.equal(placed[k], plain[k]), k
  loss = learner.step(placed)                         # consumed in place (no copy): a finite loss and a parameter update
  assert np.isfinite(float(loss))


def test_adam_kernel_matches_oracle(lrn):
  rng = np.random.default_rng(3)
  n = 100_003
  p = rng.normal(0, 1, n).astype(np.float32).astype(np.float64); m = np.zeros(n); v = np.zeros(n)
  start = p.copy()
  tp, tm, tv = dev(p, torch.float32), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
  for step in range(1, 5):
    g = rng.normal(0, 1e-2, n).astype(np.float32).astype(np.float64)
    p, m, v = qrdqn.adam_update(p, g / 2.0, m, v, step, 1e-3, eps=2e-5)
    lrn.adam_step(tp, dev(g, torch.float32), tm, tv, step, 1e-3, eps=2e-5, grad_scale=0.5)
  # the parameters move by ~4e-3 in total; fp32 storage of p costs up to 2.4e-7 per step
  assert np.abs(p - start).mean() > 1e-3
  np.testing.assert_allclose(tp.cpu().numpy(), p, rtol=0, atol=2e-6)
  np.testing.assert_allclose(tm.cpu().numpy(), m, rtol=1e-5, atol=1e-9)
  np.testing.assert_allclose(tv.cpu().numpy(), v, rtol=1e-5, atol=1e-12)


def test_marco_polo_matches_oracle(lrn):
  rng = np.random.default_rng(4)
  n = 512
  mp = lrn.MarcoPoloExploration(n, exploratory_episode_probability=0.8, seed=9)
  oracles = [qrdqn.MarcoPolo(0.8) for _ in range(n)]
  begin = np.ones(n, np.uint8)
  explored = 0
  for k in range(260):
    obs = np.zeros((n, 1099), np.float32)
    obs[:, 0] = rng.uniform(0.1, 0.9, n)
    rl = rng.integers(0, 3, n).astype(np.int32)
    if k == 130:
      begin = (rng.uniform(size=n) < 0.3).astype(np.uint8)      # some balloons start a new episode mid-way
    prev_target = mp.walk_target.cpu().numpy().copy()
    got = mp.step(dev(obs), dev(rl), dev(begin)).cpu().numpy()
    st = mp.state.cpu().numpy()
    target = mp.walk_target.cpu().numpy()
    for e in range(n):
      o = oracles[e]
      if begin[e]:
        # the kernel's own Philox draws, recovered from what it stored (the first step of an episode is
        # always in the RL phase, so the walk target still holds the begin_episode draw)
        o.begin_episode(u_episode=0.0 if st[0, e] else 1.0, u_target=(target[e] - 6500.0) / 4900.0)
      probe = copy.copy(o)
      probe.step(obs[e, 0], int(rl[e]), 0.0)
      z = (target[e] - o.target) / (probe.walk_elapsed * 0.1666) if probe.exploratory_phase else 0.0
      want = o.step(obs[e, 0], int(rl[e]), z)
      assert want == got[e], (k, e)
      assert (int(o.exploratory_episode), int(o.exploratory_phase), o.phase_elapsed, o.walk_elapsed) == tuple(st[:, e]), (k, e)
      assert abs(o.target - target[e]) < 1e-6
      if not o.exploratory_phase:
        assert got[e] == rl[e] and (begin[e] or target[e] == prev_target[e])
      explored += int(o.exploratory_phase)
    if k == 0:
      frac = st[0].mean()
      assert 0.72 < frac < 0.88                                   # exploratory_episode_probability = 0.8
      assert target.min() >= 6500.0 and target.max() <= 11400.0    # sample_pressure without atmosphere
    begin = np.zeros(n, np.uint8)
  assert explored > 0


def test_learner_step_matches_torch_fp64(lrn):
  """One full learner update (target net -> target distribution -> loss -> backward -> Adam) against the
  same computation in fp64 torch on the CPU with the oracle's loss gradient."""
  cfg = lrn.QrDqnConfig(num_layers=3, hidden_units=32, num_features=19, learning_rate=1e-3, target_update_period=2,
                        tf32_matmul=False)
  learner = lrn.QrDqnLearner(cfg, seed=1)
  ref = lrn.QuantileNetwork(cfg).double()
  ref_t = lrn.QuantileNetwork(cfg).double()
  with torch.no_grad():
    for dst, src in zip(ref.parameters(), learner.online.parameters()):
      dst.copy_(src.detach().cpu().double())
    for dst, src in zip(ref_t.parameters(), learner.target.parameters()):
      dst.copy_(src.detach().cpu().double())
  opt = torch.optim.Adam(ref.parameters(), lr=1e-3, eps=cfg.adam_eps)
  rng = np.random.default_rng(5)
  for step in range(3):
    b = 24
    batch = {'state': rng.normal(0, 1, (b, 19)).astype(np.float32), 'next_state': rng.normal(0, 1, (b, 19)).astype(np.float32),
             'action': rng.integers(0, 3, b).astype(np.int32), 'return': rng.uniform(0, 2, b).astype(np.float32),
             'discount': np.full(b, 0.993 ** 5, np.float32), 'valid': np.ones(b, np.uint8)}
    loss = learner.step({k: dev(v) for k, v in batch.items()})
    with torch.no_grad():
      nl = ref_t(torch.tensor(batch['next_state']).double()).numpy()
    tgt = qrdqn.target_distribution(nl, batch['return'], batch['discount'])
    logits = ref(torch.tensor(batch['state']).double())
    want_l, want_g = qrdqn.quantile_huber_loss(logits.detach().numpy(), batch['action'], tgt)
    opt.zero_grad()
    logits.backward(torch.tensor(want_g))
    opt.step()
    if (step + 1) % 2 == 0:
      ref_t.load_state_dict(ref.state_dict())
    assert abs(float(loss) - want_l.mean()) < 1e-4 * max(1.0, want_l.mean())
    for got, want in zip(learner.online.parameters(), ref.parameters()):
      np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=0, atol=2e-5)
    for got, want in zip(learner.target.parameters(), ref_t.parameters()):
      np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=0, atol=2e-5)


def test_training_loop_smoke(lrn):
  """run_training on 96 balloons: bookkeeping of the vectorised EnvironmentLoop."""
  from balloon_learning_environment_b200 import batched_env
  from tests.golden import fields as golden_fields
  n = 96
  env = batched_env.BatchedBalloonEnv(n, observation='perciatelli', field_layout='x128', seed=3)
  env.arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()),
                            torch.arange(n, dtype=torch.int32) % 4)
  cfg = lrn.QrDqnConfig(num_layers=3, hidden_units=64, batch_size=64, min_replay_size=200, max_episode_length=7,
                        learning_rate=1e-4)
  learner = lrn.QrDqnLearner(cfg, seed=0)
  before = learner.flat.clone()
  explore = lrn.MarcoPoloExploration(n, seed=1)
  replay = lrn.DeviceReplay(n, 12, n_step=cfg.n_step, gamma=cfg.discount)
  stats = lrn.run_training(env, learner, num_iterations=20, replay=replay, exploration=explore,
                           learner_steps_per_iteration=2, seed=5)
  assert stats['env_steps'] == 20 * n and replay.count == 20
  assert stats['episodes'] >= 2 * n                       # max_episode_length 7 -> two truncations per balloon
  assert stats['learner_steps'] >= 20 and np.isfinite(stats['last_loss']) and stats['last_loss'] > 0
  assert 0.0 <= stats['mean_reward'] <= 1.0
  assert float((learner.flat - before).abs().max()) > 0
  assert int(replay.truncated.sum()) >= n                 # the step-limit truncations were recorded
  env.close()


def test_quantile_agent_interface(lrn, tmp_path):
  """QuantileAgent behind the batched Agent interface: greedy in EVAL mode through eval_lib.eval_agent, replay +
  SGD + exploration in TRAIN mode, checkpoint round trip."""
  from balloon_learning_environment_b200 import agents, batched_env, eval_lib, suites
  from tests.golden import fields as golden_fields
  n = 32
  env = batched_env.BatchedBalloonEnv(n, observation='perciatelli', field_layout='x128', seed=1)
  env.arena.set_wind_fields(torch.from_numpy(golden_fields.field_bank()), torch.arange(n, dtype=torch.int32) % 4)
  cfg = lrn.QrDqnConfig(num_layers=3, hidden_units=64, batch_size=64, min_replay_size=128, learning_rate=1e-4)
  agent = lrn.QuantileAgent(3, (1099,), env.arena, config=cfg, seed=4, replay_steps=32, learner_steps_per_step=2)
  assert isinstance(agents.create_agent('quantile', 3, (1099,), env.arena), lrn.QuantileAgent)
  suite = suites.EvaluationSuite(seeds=list(range(100, 100 + n)), max_episode_length=12)
  results = eval_lib.eval_agent(agent, env, suite, calculate_flight_path=False)
  assert len(results) == n and all(r.final_timestep == 12 for r in results)
  assert agent.replay.count == 0 and agent.learner.steps == 0             # EVAL mode neither stores nor learns
  # greedy and deterministic: the same observation gives the same action
  obs = env.reset(seeds=torch.arange(n, dtype=torch.int64))
  a1 = agent.begin_episode(obs).clone(); a2 = agent.begin_episode(obs)
  assert torch.equal(a1, a2) and int(a1.min()) >= 0 and int(a1.max()) <= 2
  # TRAIN mode: one lockstep episode of 10 steps
  agent.set_mode(agents.AgentMode.TRAIN)
  before = agent.learner.flat.clone()
  action = agent.begin_episode(obs)
  for t in range(10):
    obs, reward, done, _ = env.step(action)
    action = agent.step(reward, obs, done)
  agent.end_episode(reward, done)
  assert agent.replay.count == 11 and int(agent.replay.truncated[10].sum()) == n - int(done.ne(0).sum())
  assert agent.learner.steps > 0 and float((agent.learner.flat - before).abs().max()) > 0
  assert np.isfinite(float(agent.last_loss))
  agent.save_checkpoint(str(tmp_path), 3)
  saved = agent.learner.flat.clone()
  agent.learner.flat.zero_()
  assert agent.reload_latest_checkpoint(str(tmp_path)) == 3
  assert torch.equal(agent.learner.flat, saved)
  assert agent.reload_latest_checkpoint(str(tmp_path / 'missing')) == -1
  env.close()
