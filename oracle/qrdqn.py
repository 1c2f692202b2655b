"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the QR-DQN learner surface.

Restates, in NumPy fp64 / plain Python loops:

* the quantile network, `agents/networks.py:63-98` (8 Dense layers, 600 hidden units, ReLU,
  `num_actions * num_atoms` outputs reshaped to `[num_actions, num_atoms]`, q = mean over atoms;
  initialiser `variance_scaling(1/sqrt(3), 'fan_in', 'uniform')`), as configured by
  `acme_utils.py:235-238` and `agents/configs/quantile.gin:16-17`;
* the QR-DQN target and quantile-Huber loss.  The arithmetic lives in THIRD-PARTY code that is not
  vendored under /root/reference: `dopamine-rl==4.0.0` (`requirements.txt:14`;
  `dopamine/jax/agents/quantile/quantile_agent.py`, `target_distribution` and `train`, called from
  `agents/quantile_agent.py:122-139`) and `rlax.quantile_q_learning` behind `acme.agents.jax.dqn.QrDqn`
  (`acme_utils.py:245`, `huber_param=1`).  Both published forms reduce to
      loss_b = (1/N) * sum_i sum_j |tau_i - 1{d_ij < 0}| * huber_kappa(d_ij),   d_ij = target_j - theta_i,
      tau_i = (i + 0.5) / N,  target_j = r + discount * theta'_{a*, j},  a* = argmax_a mean_j theta'_{a, j}
  (rlax divides the Huber term by kappa, which is 1 here).  PARITY UNPINNED at this boundary: the
  reference holds no golden vectors for the loss; the restatement is checked against finite
  differences and a PyTorch-autograd fp32 restatement in tests/test_oracle_qrdqn.py;
* n-step transitions (`config.n_step = 5`, `config.discount = 0.993`, `acme_utils.py:223-224`;
  `JaxQuantileAgent.update_horizon = 5`, `gamma = 0.993`, `quantile.gin:18-19`): a window that
  reaches a terminal step is cut there with discount 0; a window that crosses a step-limit
  truncation is not sampled (dopamine's circular replay buffer rule);
* Adam as `optax.adam(learning_rate, eps=adam_eps)` (`acme_utils.py:225,233`; optax 0.0.9,
  `requirements.txt:45`): bias-corrected moments, update = m_hat / (sqrt(v_hat) + eps);
* MarcoPoloExploration (`agents/marco_polo_exploration.py:36-93`) driven the way
  `acme_utils.CombinedActor` drives it (`acme_utils.py:161-183`).
"""
import numpy as np

NUM_ACTIONS = 3
NUM_ATOMS = 51
NUM_LAYERS = 8
HIDDEN_UNITS = 600
NUM_FEATURES = 1099
GAMMA = 0.993
N_STEP = 5
KAPPA = 1.0
RL_PHASE_S = 4 * 3600          # marco_polo_exploration.py:35
EXPLORATORY_PHASE_S = 2 * 3600  # marco_polo_exploration.py:36
AGENT_STEP_S = 180             # utils/constants.py:35


def init_params(rng, num_inputs=NUM_FEATURES, num_layers=NUM_LAYERS, hidden=HIDDEN_UNITS,
                num_actions=NUM_ACTIONS, num_atoms=NUM_ATOMS):
  """networks.py:79-92: uniform(-l, l), l = sqrt(3 * scale / fan_in), scale = 1/sqrt(3); zero bias."""
  dims = [num_inputs] + [hidden] * (num_layers - 1) + [num_actions * num_atoms]
  params = []
  for fan_in, fan_out in zip(dims[:-1], dims[1:]):
    limit = np.sqrt(3.0 * (1.0 / np.sqrt(3.0)) / fan_in)
    params.append((rng.uniform(-limit, limit, (fan_in, fan_out)), np.zeros(fan_out)))
  return params


def forward(params, x, num_actions=NUM_ACTIONS, num_atoms=NUM_ATOMS):
  """networks.py:84-97 for a batch: returns (logits [B, A, N], q_values [B, A])."""
  h = np.asarray(x, np.float64)
  for i, (w, b) in enumerate(params):
    h = h @ w + b
    if i + 1 < len(params):
      h = np.maximum(h, 0.0)
  logits = h.reshape(-1, num_actions, num_atoms)
  return logits, logits.mean(axis=2)


def greedy_actions(logits):
  """argmax_a of the mean over atoms; first maximum wins (jnp.argmax)."""
  return np.argmax(logits.mean(axis=2), axis=1).astype(np.int32)


def target_distribution(next_logits, reward, discount):
  """dopamine quantile_agent.target_distribution: r + discount * theta'[argmax_a mean theta'] -> [B, N]."""
  a = greedy_actions(next_logits)
  chosen = next_logits[np.arange(len(a)), a]
  return np.asarray(reward, np.float64)[:, None] + np.asarray(discount, np.float64)[:, None] * chosen


def quantile_huber_loss(logits, actions, target, kappa=KAPPA):
  """Per-sample loss [B] and d(mean_b loss_b)/d logits [B, A, N] (zero for the actions not taken)."""
  b, a, n = logits.shape
  theta = logits[np.arange(b), actions]                          # [B, N]
  d = target[:, None, :] - theta[:, :, None]                     # [B, i, j]
  absd = np.abs(d)
  huber = np.where(absd <= kappa, 0.5 * d * d, kappa * (absd - 0.5 * kappa))
  tau = (np.arange(n) + 0.5) / n
  weight = np.abs(tau[None, :, None] - (d < 0))
  loss = (weight * huber).sum(axis=2).mean(axis=1)
  dhuber = np.where(absd <= kappa, d, kappa * np.sign(d))
  gtheta = -(weight * dhuber).sum(axis=2) / n / b                # d mean_b(loss_b) / d theta_i
  grad = np.zeros_like(logits)
  grad[np.arange(b), actions] = gtheta
  return loss, grad


def nstep_transition(reward, terminal, truncated, count, capacity, t, e, n=N_STEP, gamma=GAMMA):
  """One n-step transition starting at absolute step t of balloon e, from ring arrays [capacity, E].

  Returns None when (t, e) cannot be sampled, else (n_used, R, discount, t_next) with
  R = sum_{k<n_used} gamma^k r_{t+k}, discount = gamma^n_used * (1 - terminal), next observation = obs[t_next].
  """
  if t < max(0, count - capacity) or t >= count:
    return None
  ret, g = 0.0, 1.0
  for k in range(n):
    if t + k >= count:
      return None                                   # the window runs past what has been written
    slot = (t + k) % capacity
    ret += g * float(reward[slot, e])
    g *= gamma
    if terminal[slot, e]:
      t_next = t + k + 1 if t + k + 1 < count else t
      return k + 1, ret, 0.0, t_next
    if truncated[slot, e]:
      return None                                   # final observation of a truncated episode is not stored
  if t + n >= count:
    return None
  return n, ret, g, t + n


def adam_update(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=2e-5):
  """optax.adam: step counts from 1.  Returns (p, m, v)."""
  m = b1 * m + (1.0 - b1) * g
  v = b2 * v + (1.0 - b2) * g * g
  m_hat = m / (1.0 - b1 ** step)
  v_hat = v / (1.0 - b2 ** step)
  return p - lr * m_hat / (np.sqrt(v_hat) + eps), m, v


class MarcoPolo:
  """One balloon's MarcoPoloExploration + RandomWalkAgent, with the random draws supplied by the caller
  (the reference draws them from jax.random; streams are unpinned by design)."""

  def __init__(self, probability=0.8):
    self.probability = probability
    self.exploratory_episode = False
    self.exploratory_phase = False
    self.phase_elapsed = 0
    self.walk_elapsed = 0
    self.target = 0.0

  def begin_episode(self, u_episode, u_target):
    """marco_polo_exploration.py:53-62 + random_walk_agent.py:75-78 (target ~ U(6500, 11400) Pa)."""
    self.walk_elapsed = 0
    self.target = 6500.0 + (11400.0 - 6500.0) * u_target
    self.phase_elapsed = 0
    self.exploratory_episode = u_episode <= self.probability
    self.exploratory_phase = False

  def step(self, pressure_feature, rl_action, z):
    """marco_polo_exploration.py:75-93; z ~ N(0, 1) is consumed only in an exploratory phase."""
    self.phase_elapsed += AGENT_STEP_S
    if self.exploratory_episode:
      limit = EXPLORATORY_PHASE_S if self.exploratory_phase else RL_PHASE_S
      if self.phase_elapsed >= limit:
        self.exploratory_phase = not self.exploratory_phase
        self.phase_elapsed = 0
    if not self.exploratory_phase:
      return int(rl_action)
    self.walk_elapsed += AGENT_STEP_S                            # random_walk_agent.py:80-91
    self.target += self.walk_elapsed * 0.1666 * z
    p = float(pressure_feature) * (14000.0 - 5000.0) + 5000.0     # features.py:192-195
    if p - 100.0 > self.target:
      return 2
    if p + 100.0 < self.target:
      return 0
    return 1
