"""CPU checks of the QR-DQN oracle (oracle/qrdqn.py) and of the learner's host logic.

The loss arithmetic belongs to dopamine-rl / rlax, which are not vendored under the reference, so it
is pinned here by (a) finite differences, (b) an independent PyTorch-autograd restatement of the
published formula in fp32, (c) hand-computed small cases.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from balloon_learning_environment_b200 import learner as learner_lib
from oracle import qrdqn


def torch_reference_loss(logits, actions, target, kappa=1.0):
  """The published QR-DQN loss written with torch ops (dopamine quantile_agent.train)."""
  b, _, n = logits.shape
  theta = logits[torch.arange(b), actions.long()]
  d = target[:, None, :] - theta[:, :, None]
  huber = torch.where(d.abs() <= kappa, 0.5 * d ** 2, kappa * (d.abs() - 0.5 * kappa))
  tau = (torch.arange(n, dtype=logits.dtype) + 0.5) / n
  w = (tau[None, :, None] - (d.detach() < 0).to(logits.dtype)).abs()
  return (w * huber).sum(2).mean(1)


def test_loss_hand_case():
  # one sample, one action, two atoms: theta = [0, 1], target = [0.5, 3]
  logits = np.array([[[0.0, 1.0]]])
  target = np.array([[0.5, 3.0]])
  loss, grad = qrdqn.quantile_huber_loss(logits, np.array([0]), target, kappa=1.0)
  # i=0 (tau .25): d = .5 -> .25*.125 ; d = 3 -> .25*2.5 ; i=1 (tau .75): d = -.5 -> |.75-1|*.125 ; d = 2 -> .75*1.5
  want = ((0.25 * 0.125 + 0.25 * 2.5) + (0.25 * 0.125 + 0.75 * 1.5)) / 2
  assert abs(loss[0] - want) < 1e-12
  # d loss / d theta_0 = -(.25*.5 + .25*1)/2 ; d theta_1 = -(.25*(-.5) + .75*1)/2
  np.testing.assert_allclose(grad[0, 0], [-(0.125 + 0.25) / 2, -(-0.125 + 0.75) / 2], atol=1e-12)


def test_loss_gradient_matches_finite_differences():
  rng = np.random.default_rng(0)
  logits = rng.normal(0, 2, (5, 3, 51))
  target = rng.normal(0, 2, (5, 51))
  actions = rng.integers(0, 3, 5)
  loss, grad = qrdqn.quantile_huber_loss(logits, actions, target)
  eps = 1e-6
  for _ in range(40):
    b, a, i = rng.integers(0, 5), rng.integers(0, 3), rng.integers(0, 51)
    lp, lm = logits.copy(), logits.copy()
    lp[b, a, i] += eps; lm[b, a, i] -= eps
    fd = (qrdqn.quantile_huber_loss(lp, actions, target)[0].mean() -
          qrdqn.quantile_huber_loss(lm, actions, target)[0].mean()) / (2 * eps)
    assert abs(fd - grad[b, a, i]) < 1e-7
  assert loss.shape == (5,)


def test_loss_matches_torch_autograd_fp32():
  rng = np.random.default_rng(1)
  logits = rng.normal(0, 1.5, (16, 3, 51)).astype(np.float32)
  target = rng.normal(0, 1.5, (16, 51)).astype(np.float32)
  actions = rng.integers(0, 3, 16)
  loss, grad = qrdqn.quantile_huber_loss(logits.astype(np.float64), actions, target.astype(np.float64))
  tl = torch.tensor(logits, requires_grad=True)
  per = torch_reference_loss(tl, torch.tensor(actions), torch.tensor(target))
  per.mean().backward()
  np.testing.assert_allclose(per.detach().numpy(), loss, rtol=2e-5, atol=1e-5)   # fp32 vs fp64
  np.testing.assert_allclose(tl.grad.numpy(), grad, rtol=1e-4, atol=2e-7)


def test_target_distribution_and_greedy_ties():
  logits = np.zeros((3, 3, 4))
  logits[0, 1] = 1.0                       # action 1 wins
  logits[1, 0] = logits[1, 2] = 2.0        # tie 0/2 -> first maximum
  logits[2, 2] = [4, 0, 0, 0]              # mean 1 beats zeros
  np.testing.assert_array_equal(qrdqn.greedy_actions(logits), [1, 0, 2])
  tgt = qrdqn.target_distribution(logits, np.array([1.0, 2.0, 3.0]), np.array([0.5, 0.0, 1.0]))
  np.testing.assert_allclose(tgt[0], 1.0 + 0.5 * 1.0)
  np.testing.assert_allclose(tgt[1], 2.0)
  np.testing.assert_allclose(tgt[2], [7, 3, 3, 3])


def test_network_matches_oracle_and_reference_shapes():
  cfg = learner_lib.QrDqnConfig()
  assert (cfg.num_layers, cfg.hidden_units, cfg.num_atoms, cfg.num_actions) == (8, 600, 51, 3)   # acme_utils.py:234-237
  small = learner_lib.QrDqnConfig(num_layers=4, hidden_units=48, num_features=37)
  rng = np.random.default_rng(2)
  params = qrdqn.init_params(rng, num_inputs=37, num_layers=4, hidden=48)
  for (w, _), fan_in in zip(params, [37, 48, 48, 48]):
    limit = np.sqrt(np.sqrt(3.0) / fan_in)
    assert np.abs(w).max() <= limit and np.abs(w).max() > 0.9 * limit
  net = learner_lib.QuantileNetwork(small)
  net.load_flax_params({f'Dense_{i}': {'kernel': w.astype(np.float32), 'bias': (b + 0.01 * i).astype(np.float32)}
                        for i, (w, b) in enumerate(params)})
  params = [(w, b + 0.01 * i) for i, (w, b) in enumerate(params)]
  x = rng.normal(0, 1, (9, 37)).astype(np.float32)
  want, want_q = qrdqn.forward(params, x)
  got = net(torch.tensor(x)).detach().numpy()
  assert got.shape == (9, 3, 51)
  np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
  np.testing.assert_allclose(got.mean(2), want_q, rtol=1e-4, atol=1e-5)
  # default initialiser of the torch module follows networks.py:79-82
  for layer in learner_lib.QuantileNetwork(small).layers:
    limit = np.sqrt(np.sqrt(3.0) / layer.in_features)
    assert float(layer.weight.detach().abs().max()) <= limit + 1e-7 and float(layer.bias.detach().abs().max()) == 0.0


def _ring(capacity, envs, count, rng):
  reward = rng.uniform(0, 1, (capacity, envs))
  terminal = np.zeros((capacity, envs), bool)
  truncated = np.zeros((capacity, envs), bool)
  return reward, terminal, truncated


def test_nstep_transitions():
  rng = np.random.default_rng(3)
  cap, envs, count = 16, 2, 12
  reward, terminal, truncated = _ring(cap, envs, count, rng)
  g = qrdqn.GAMMA
  n, ret, disc, t_next = qrdqn.nstep_transition(reward, terminal, truncated, count, cap, 2, 1)
  assert (n, t_next) == (5, 7)
  assert abs(ret - sum(g ** k * reward[2 + k, 1] for k in range(5))) < 1e-12 and abs(disc - g ** 5) < 1e-12
  # the next observation (t + 5) must already be written
  assert qrdqn.nstep_transition(reward, terminal, truncated, count, cap, 7, 0) is None
  assert qrdqn.nstep_transition(reward, terminal, truncated, count, cap, 6, 0) is not None
  # terminal inside the window: cut, discount 0
  terminal[4, 1] = True
  n, ret, disc, t_next = qrdqn.nstep_transition(reward, terminal, truncated, count, cap, 2, 1)
  assert (n, disc, t_next) == (3, 0.0, 5)
  assert abs(ret - sum(g ** k * reward[2 + k, 1] for k in range(3))) < 1e-12
  # a terminal last step is usable even at the write cursor
  terminal[11, 0] = True
  assert qrdqn.nstep_transition(reward, terminal, truncated, count, cap, 11, 0) == (1, reward[11, 0], 0.0, 11)
  # truncation inside the window: not sampled
  truncated[8, 0] = True
  assert qrdqn.nstep_transition(reward, terminal, truncated, count, cap, 5, 0) is None
  assert qrdqn.nstep_transition(reward, terminal, truncated, count, cap, 9, 1) is None    # window passes the cursor
  # wrapped ring: step 20 lives in slot 4; steps older than count - capacity are gone
  count = 30
  assert qrdqn.nstep_transition(reward, np.zeros_like(terminal), np.zeros_like(truncated), count, cap, 13, 0) is None
  n, ret, _, t_next = qrdqn.nstep_transition(reward, np.zeros_like(terminal), np.zeros_like(truncated), count, cap, 14, 0)
  assert t_next == 19 and abs(ret - sum(g ** k * reward[(14 + k) % cap, 0] for k in range(5))) < 1e-12


def test_adam_matches_torch_adam():
  rng = np.random.default_rng(4)
  p = rng.normal(0, 1, 50); m = np.zeros(50); v = np.zeros(50)
  tp = torch.tensor(p.copy(), requires_grad=True)
  opt = torch.optim.Adam([tp], lr=2e-3, eps=2e-5)
  for step in range(1, 6):
    g = rng.normal(0, 1, 50)
    p, m, v = qrdqn.adam_update(p, g, m, v, step, 2e-3)
    tp.grad = torch.tensor(g.copy())
    opt.step()
  np.testing.assert_allclose(tp.detach().numpy(), p, rtol=1e-10, atol=1e-12)


def test_marco_polo_phase_schedule():
  mp_ = qrdqn.MarcoPolo(0.8)
  mp_.begin_episode(u_episode=0.5, u_target=0.5)
  assert mp_.exploratory_episode and abs(mp_.target - 8950.0) < 1e-9
  phases = []
  for k in range(1, 250):
    a = mp_.step(pressure_feature=0.5, rl_action=1, z=0.0)
    phases.append(mp_.exploratory_phase)
    if not mp_.exploratory_phase:
      assert a == 1
  # 4 h of RL (calls 1..79), 2 h of exploration (calls 80..119), 4 h of RL (120..199), ...
  assert not any(phases[:79]) and all(phases[79:119]) and not any(phases[119:199]) and all(phases[199:239])
  # pressure feature 0.5 -> 9,500 Pa > target + 100 -> UP while exploring
  mp_.begin_episode(0.5, 0.5)
  for k in range(80):
    a = mp_.step(0.5, 1, 0.0)
  assert mp_.exploratory_phase and a == 2
  # a non-exploratory episode never leaves the RL phase
  mp_.begin_episode(u_episode=0.81, u_target=0.1)
  assert all(mp_.step(0.5, 0, 1.0) == 0 and not mp_.exploratory_phase for _ in range(300))


def test_flatten_parameters_keeps_views():
  net = learner_lib.QuantileNetwork(learner_lib.QrDqnConfig(num_layers=3, hidden_units=8, num_features=5))
  before = [p.detach().clone() for p in net.parameters()]
  flat = learner_lib.flatten_parameters(net, 'cpu')
  assert flat.numel() == sum(p.numel() for p in net.parameters())
  for p, b in zip(net.parameters(), before):
    assert torch.equal(p, b)
  out = net(torch.randn(4, 5))
  net.flat_grad.zero_()
  out.sum().backward()
  assert float(net.flat_grad.abs().sum()) > 0                   # autograd accumulated into the flat buffer
  flat.zero_()
  assert all(float(p.abs().sum()) == 0 for p in net.parameters())


def _dp_worker(rank, world, port, out):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  cfg = learner_lib.QrDqnConfig(num_layers=3, hidden_units=16, num_features=11)
  torch.manual_seed(0)
  net = learner_lib.QuantileNetwork(cfg)
  learner_lib.flatten_parameters(net, 'cpu')
  g = torch.Generator().manual_seed(5)
  x = torch.randn(8, 11, generator=g); tgt = torch.randn(8, 51, generator=g); act = torch.randint(0, 3, (8,), generator=g)
  lo, hi = rank * 4, rank * 4 + 4                               # each rank sees half of the batch
  net.flat_grad.zero_()
  torch_reference_loss(net(x[lo:hi]), act[lo:hi], tgt[lo:hi]).mean().backward()
  world_size = learner_lib.allreduce_sum_(net.flat_grad)
  out[rank] = (net.flat_grad / world_size).clone()
  dist.destroy_process_group()


def test_data_parallel_gradient_equals_full_batch_gradient():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]
  manager = mp.Manager()
  out = manager.dict()
  mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
  cfg = learner_lib.QrDqnConfig(num_layers=3, hidden_units=16, num_features=11)
  torch.manual_seed(0)
  net = learner_lib.QuantileNetwork(cfg)
  learner_lib.flatten_parameters(net, 'cpu')
  g = torch.Generator().manual_seed(5)
  x = torch.randn(8, 11, generator=g); tgt = torch.randn(8, 51, generator=g); act = torch.randint(0, 3, (8,), generator=g)
  net.flat_grad.zero_()
  torch_reference_loss(net(x), act, tgt).mean().backward()
  for rank in range(2):
    torch.testing.assert_close(out[rank], net.flat_grad, rtol=1e-5, atol=1e-7)


def test_kernels_refuse_cpu_tensors():
  with pytest.raises(Exception):
    learner_lib.greedy_actions(torch.zeros(2, 3, 51))
  with pytest.raises(Exception):
    learner_lib.DeviceReplay(4, 8, device='cpu')
