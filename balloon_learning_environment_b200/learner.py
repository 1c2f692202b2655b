"""QR-DQN learner surface for vectorised rollouts (BASELINE configs[4]).

Mirrors what the reference assembles in `acme_utils.create_dqn` (acme_utils.py:217-277) and
`train_acme_qrdqn.py:43-81` -- equivalently the Dopamine `QuantileAgent` of agents/quantile_agent.py with
agents/configs/quantile.gin -- for N balloons stepping in lockstep on the device:

  QuantileNetwork      agents/networks.py:63-98 (8 Dense layers x 600, 3 actions x 51 atoms)
  DeviceReplay         the replay table (max_replay_size 2,000,000, n_step 5, discount 0.993)
  MarcoPoloExploration agents/marco_polo_exploration.py:36-93 around RandomWalkAgent (acme_utils.py:161-214)
  QrDqnLearner         QrDqn(num_atoms=51, huber_param=1) loss, Adam(2e-6, eps 2e-5), target period 25
  TrainingLoop         the EnvironmentLoop of train_acme_qrdqn.py:72-81, one iteration = N env steps

The dense layers are cuBLAS GEMMs through torch; everything else on the update path is a hand-written
CUDA kernel behind the C ABI (csrc/ble_learner.cu).  Data-parallel training keeps one learner per GPU
and sums the flat gradient buffer with ONE all-reduce per learner step (NCCL; gloo in the CPU tests
of the host logic).  There is no CPU path for the kernels.
"""
import contextlib
import ctypes
import dataclasses
import math
import os
from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import nn

from balloon_learning_environment_b200 import _lib

NUM_FEATURES = 1099


def _ptr(t: Optional[torch.Tensor]):
  return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(device):
  return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, what: str):
  if not t.is_cuda:
    raise _lib.BleError(f'{what}: tensors must live on a CUDA device (no CPU fallback exists)')


def _check(rc: int, what: str):
  if rc != _lib.BLE_OK:
    raise _lib.BleError(f'{what} failed (code {rc})')


def _call(name: str, device, *args):
  """One stateless learner entry point of the C ABI, launched with `device` current (the entry points take a stream
  but no device ordinal; the stream belongs to `device`) and the caller's device restored afterwards."""
  with torch.cuda.device(device):
    _check(getattr(_lib.load(), name)(*args, _stream(device)), name)


@dataclasses.dataclass
class QrDqnConfig:
  """acme_utils.create_dqn (acme_utils.py:217-246) / agents/configs/quantile.gin."""
  num_actions: int = 3
  num_atoms: int = 51                    # acme_utils.py:234
  num_layers: int = 8                    # acme_utils.py:236
  hidden_units: int = 600
  num_features: int = NUM_FEATURES
  discount: float = 0.993                # acme_utils.py:223
  n_step: int = 5                        # acme_utils.py:224
  min_replay_size: int = 500             # acme_utils.py:225 (transitions)
  target_update_period: int = 25         # learner steps: 100 // 4, acme_utils.py:219-226
  learning_rate: float = 2e-6            # acme_utils.py:238
  adam_eps: float = 2e-5                 # acme_utils.py:227
  adam_b1: float = 0.9
  adam_b2: float = 0.999
  huber_param: float = 1.0               # acme_utils.py:245
  batch_size: int = 32                   # acme_utils.py:229 (per learner step, per GPU)
  samples_per_insert: float = 8.0        # batch_size / update_period, acme_utils.py:230
  max_replay_size: int = 2_000_000       # acme_utils.py:228 (transitions)
  epsilon: float = 0.05                  # acme_utils.py:225 builds dqn.DQNConfig(**params) without an epsilon, so the
                                         # behaviour policy (:256-262) runs at Acme's DQNConfig default (0.05 in
                                         # dm-acme 0.4, a dependency absent from the reference tree); quantile.gin's
                                         # epsilon_train = 0 belongs to the dopamine agent, not to this path
  exploratory_episode_probability: float = 0.8   # acme_utils.py:207
  max_episode_length: int = 960          # train_acme_qrdqn.py:30-33
  tf32_matmul: bool = True               # dense layers on the tensor cores (TF32 in, fp32 accumulate): what
                                         # jax's default matmul precision does on an Ampere+ GPU
  dense_backend: str = 'tcgen05'         # 'tcgen05': forward AND backward of the dense layers by the hand-written
                                         # ble_dense_tf32 kernel (DenseStack, no autograd); 'cublas': nn.Linear +
                                         # autograd on library GEMMs (A/B reference; the only choice for the fp32
                                         # audit setting tf32_matmul = False)
  cuda_graph: bool = True                # tcgen05 backend: the ~60 launches of one SGD step (target forward, online
                                         # forward, loss, backward) and the operand refresh are captured once per batch
                                         # size and replayed (the eager sequence is bound by host launch overhead)


class QuantileNetwork(nn.Module):
  """agents/networks.py:63-98 for a batch: obs [B, F] -> logits [B, A, N]; q_values = mean over atoms."""

  def __init__(self, config: QrDqnConfig = QrDqnConfig()):
    super().__init__()
    self.num_actions, self.num_atoms = config.num_actions, config.num_atoms
    dims = [config.num_features] + [config.hidden_units] * (config.num_layers - 1) + [config.num_actions * config.num_atoms]
    self.layers = nn.ModuleList(nn.Linear(i, o) for i, o in zip(dims[:-1], dims[1:]))
    for layer in self.layers:          # variance_scaling(1/sqrt(3), 'fan_in', 'uniform'), networks.py:79-82
      limit = math.sqrt(3.0 * (1.0 / math.sqrt(3.0)) / layer.in_features)
      nn.init.uniform_(layer.weight, -limit, limit)
      nn.init.zeros_(layer.bias)

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    h = x.to(self.layers[0].weight.dtype)           # float32 in production (networks.py:83)
    for i, layer in enumerate(self.layers):
      h = layer(h)
      if i + 1 < len(self.layers):
        h = torch.relu(h)
    return h.view(-1, self.num_actions, self.num_atoms)

  def load_flax_params(self, params: Dict) -> None:
    """'Dense_i': {'kernel' [in, out], 'bias'} as produced by QuantileAgent.load_perciatelli_weights
    (agents/quantile_agent.py:196-254) or a reference checkpoint."""
    with torch.no_grad():
      for i, layer in enumerate(self.layers):
        entry = params[f'Dense_{i}']
        layer.weight.copy_(torch.as_tensor(entry['kernel'], dtype=torch.float32).t())
        layer.bias.copy_(torch.as_tensor(entry['bias'], dtype=torch.float32))


def flatten_parameters(module: nn.Module, device) -> torch.Tensor:
  """Moves every parameter of `module` into ONE contiguous fp32 buffer (returned) of which the parameters
  become views, and gives every parameter a .grad that is a view of `module.flat_grad`: the optimiser
  kernel and the gradient all-reduce then work on two flat buffers."""
  params = list(module.parameters())
  total = sum(p.numel() for p in params)
  flat = torch.empty(total, dtype=torch.float32, device=device)
  grad = torch.zeros(total, dtype=torch.float32, device=device)
  offset = 0
  for p in params:
    n = p.numel()
    flat[offset:offset + n].copy_(p.detach().reshape(-1))
    p.data = flat[offset:offset + n].view(p.shape)
    p.grad = grad[offset:offset + n].view(p.shape)
    offset += n
  module.flat_params, module.flat_grad = flat, grad
  return flat


# ---------------------------------------------------------------------------------------------
# kernels behind the C ABI
# ---------------------------------------------------------------------------------------------
def greedy_actions(logits: torch.Tensor, with_q: bool = False):
  """argmax_a mean_j logits[b, a, j] -> int32 [B] (acme_utils.py:250-268 with epsilon = 0)."""
  _require_cuda(logits, 'greedy_actions')
  logits = logits.contiguous()
  b, a, n = logits.shape
  actions = torch.empty(b, dtype=torch.int32, device=logits.device)
  q = torch.empty(b, a, dtype=torch.float32, device=logits.device) if with_q else None
  _call('ble_qr_greedy', logits.device, _ptr(logits), b, a, n, _ptr(actions), _ptr(q))
  return (actions, q) if with_q else actions


def target_distribution(next_logits: torch.Tensor, reward: torch.Tensor, discount: torch.Tensor) -> torch.Tensor:
  """reward + discount * next_logits[greedy action] -> float32 [B, N] (dopamine target_distribution)."""
  _require_cuda(next_logits, 'target_distribution')
  next_logits = next_logits.contiguous()
  b, a, n = next_logits.shape
  target = torch.empty(b, n, dtype=torch.float32, device=next_logits.device)
  _call('ble_qr_target', next_logits.device, _ptr(next_logits), _ptr(reward.contiguous()), _ptr(discount.contiguous()),
        b, a, n, _ptr(target))
  return target


class _QuantileHuberLoss(torch.autograd.Function):
  """mean_b weight_b * loss_b with the gradient produced by the same kernel pass."""

  @staticmethod
  def forward(ctx, logits, actions, target, weight, kappa):
    _require_cuda(logits, 'quantile_huber_loss')
    logits = logits.contiguous()
    b, a, n = logits.shape
    loss = torch.empty(b, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits)
    _call('ble_qr_loss', logits.device, _ptr(logits), _ptr(actions.contiguous()), _ptr(target.contiguous()),
          _ptr(weight.contiguous() if weight is not None else None), float(kappa), b, a, n, 1.0 / b, _ptr(loss), _ptr(grad))
    ctx.save_for_backward(grad)
    ctx.mark_non_differentiable(loss)
    mean = (loss * weight).mean() if weight is not None else loss.mean()
    return mean, loss

  @staticmethod
  def backward(ctx, grad_mean, _grad_loss):
    (grad,) = ctx.saved_tensors
    return grad * grad_mean, None, None, None, None


def quantile_huber_loss(logits, actions, target, weight=None, kappa: float = 1.0):
  """Returns (mean loss, per-sample loss [B]); differentiable w.r.t. logits only (target is a constant)."""
  return _QuantileHuberLoss.apply(logits, actions, target, weight, kappa)


def adam_step(params, grads, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=2e-5, grad_scale=1.0):
  """optax.adam on flat fp32 buffers, in place; step counts from 1."""
  _require_cuda(params, 'adam_step')
  _call('ble_adam_step', params.device, _ptr(params), _ptr(grads), _ptr(m), _ptr(v), params.numel(), float(lr), float(b1),
        float(b2), float(eps), int(step), float(grad_scale))


def _pitch4(n: int) -> int:
  return (int(n) + 3) // 4 * 4


def _pitch32(n: int) -> int:
  """Row pitch of the dense path's own buffers: a multiple of 32 floats, so that every row starts on a 128-byte line and a
  TMA box row never straddles two (measured: 27 -> 23 us per 8,192 x 600 x 600 forward product against a 604-float pitch)."""
  return (int(n) + 31) // 32 * 32


def dense_tf32(a, lda, b, ldb, m, n, k, mode, aux=None, ld_aux=0, d=None, ldd=0, dt=None, ldt=0, split_k=1, relu_bits=None):
  """D[m, n] = A[m, k] . B[n, k]^T on the tcgen05 tensor cores (include/ble_b200.h: ble_dense_tf32).  relu_bits: int32
  [m, >= ceil(n / 32)], the packed ReLU mask mode 1 writes and mode 2 reads."""
  _require_cuda(a, 'dense_tf32')
  _call('ble_dense_tf32', a.device, _ptr(a), int(lda), _ptr(b), int(ldb), int(m), int(n), int(k), int(mode), _ptr(aux),
        int(ld_aux), _ptr(d), int(ldd), _ptr(dt), int(ldt), int(split_k), _ptr(relu_bits),
        0 if relu_bits is None else int(relu_bits.stride(0)))


def transpose_f32(src, ld_src, rows, cols, dst, ld_dst):
  _call('ble_transpose_f32', src.device, _ptr(src), int(ld_src), int(rows), int(cols), _ptr(dst), int(ld_dst))


def row_sum_f32(src, ld_src, rows, cols, out, accumulate=False):
  _call('ble_row_sum_f32', src.device, _ptr(src), int(ld_src), int(rows), int(cols), _ptr(out), int(bool(accumulate)))


class DenseStack:
  """Forward and backward pass of a QuantileNetwork's dense layers WITHOUT autograd: every product is one launch of
  ble_dense_tf32, with bias / ReLU (+ packed mask) / ReLU-mask / split-K accumulation fused into its epilogue.

    forward          H' = relu(H . W^T + b)      A = H [B, in], B = W [out, in]            both K-contiguous (mode 1 / 0)
    input gradient   dH = (dY . W) * mask        A = dY [B, out], B = W^T [in, out] (a copy kept by refresh())   (mode 2)
    weight gradient  dW^T = H^T . dY             A = H [B, in], B = dY [B, out] read as MN-MAJOR operands        (mode 4):
                     the row-major activations and output gradients as they are, split over K = batch and accumulated
                     through the kernel's transposed output so that a warp's reductions fall on consecutive floats of dW.

  Every activation buffer carries one extra column of ones, so the weight-gradient product's last output row is the
  bias gradient.  Pitches are rounded up to 4 floats (TMA needs 16-byte row pitches); `refresh()` re-derives the weight
  copies after the parameters changed.  Gradients land in the parameters' .grad views of the flat gradient buffer."""

  SPLIT_TARGET_CTAS = 296                               # 2 CTAs per SM x 148 SMs

  def __init__(self, net: QuantileNetwork, device):
    self.net, self.device = net, torch.device(device)
    self.dims = [(l.in_features, l.out_features) for l in net.layers]
    self.w_fwd, self.w_t = [], []
    for l, (fin, fout) in enumerate(self.dims):
      self.w_fwd.append(torch.zeros(fout, _pitch32(fin), dtype=torch.float32, device=self.device))      # line-aligned rows
      self.w_t.append(None if l == 0 else torch.zeros(fin, _pitch32(fout), dtype=torch.float32, device=self.device))
    self._work = {}
    self._ones_inputs = set()
    self.launches = 0
    self.refresh()

  def refresh(self) -> None:
    with torch.no_grad():
      for l, layer in enumerate(self.net.layers):
        fin, fout = self.dims[l]
        self.w_fwd[l][:, :fin].copy_(layer.weight)
        if self.w_t[l] is not None:
          transpose_f32(layer.weight, fin, fout, fin, self.w_t[l], self.w_t[l].shape[1])
          self.launches += 1

  def _with_ones(self, batch: int, width: int) -> torch.Tensor:
    """[batch, pitch] zeros with column `width` set to one (the bias-gradient column); use t[:, :width]."""
    t = torch.zeros(batch, _pitch32(width + 1), dtype=torch.float32, device=self.device)
    t[:, width].fill_(1.0)
    return t

  def input_buffer(self, batch: int) -> torch.Tensor:
    """A network input buffer forward(keep=True) can use IN PLACE: [batch, features] view of a row-padded allocation whose
    column `features` holds the ones the weight gradient of the first layer needs."""
    t = self._with_ones(batch, self.dims[0][0])
    self._ones_inputs.add(t.data_ptr())
    return t[:, :self.dims[0][0]]

  def _buffers(self, batch: int, keep: bool):
    key = (batch, keep)
    if key not in self._work:
      z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=self.device)
      last = len(self.dims) - 1
      w = {'x': self.input_buffer(batch),
           'h': [self._with_ones(batch, fout) if l < last else z(batch, fout) for l, (_, fout) in enumerate(self.dims)]}
      if keep:
        w['bits'] = [torch.zeros(batch, (fout + 31) // 32, dtype=torch.int32, device=self.device) for _, fout in self.dims[:-1]]
        w['g'] = [z(batch, _pitch32(fout)) for _, fout in self.dims]
      self._work[key] = w
    return self._work[key]

  @torch.no_grad()
  def forward(self, x: torch.Tensor, keep: bool = False) -> torch.Tensor:
    """x [B, F] -> logits [B, A * N] (a buffer owned by the stack, overwritten by the next call with the same B)."""
    batch, feat = x.shape
    assert feat == self.dims[0][0], (feat, self.dims[0][0])
    w = self._buffers(batch, keep)
    tma_able = (x.dtype == torch.float32 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.stride(0) >= feat and
                x.data_ptr() % 16 == 0)
    if tma_able and (not keep or x.data_ptr() in self._ones_inputs):
      a, lda = x, x.stride(0)                              # read in place (keep: one of input_buffer()'s allocations)
    else:
      w['x'].copy_(x)
      a, lda = w['x'], w['x'].stride(0)
    w['input'] = (a, lda)
    last = len(self.dims) - 1
    for l, (fin, fout) in enumerate(self.dims):
      layer = self.net.layers[l]
      b = self.w_fwd[l]
      h = w['h'][l]
      dense_tf32(a, lda, b, b.stride(0), batch, fout, fin, 0 if l == last else 1, aux=layer.bias, d=h, ldd=h.stride(0),
                 relu_bits=w['bits'][l] if keep and l < last else None)
      self.launches += 1
      a, lda = h, h.stride(0)
    return w['h'][last]

  @torch.no_grad()
  def backward(self, grad_logits: torch.Tensor) -> None:
    """grad_logits [B, A * N] = dLoss / dlogits of the LAST forward(keep=True) call with this batch size; ACCUMULATES the
    parameter gradients into the .grad views (the caller zeroes the flat gradient buffer)."""
    batch = grad_logits.shape[0]
    w = self._buffers(batch, True)
    last = len(self.dims) - 1
    w['g'][last][:, :self.dims[last][1]].copy_(grad_logits.reshape(batch, -1))
    for l in range(last, -1, -1):
      fin, fout = self.dims[l]
      layer = self.net.layers[l]
      g = w['g'][l]
      prev, ld_prev = w['input'] if l == 0 else (w['h'][l - 1], w['h'][l - 1].stride(0))
      # dW^T [in + 1, out] = [H, 1]^T . dY: rows 0 .. in-1 accumulate into dW (transposed output), row `in` into db
      tiles = ((fin + 1 + 127) // 128) * ((fout + 159) // 160)     # the kernel's 128 x 160 tiles
      split = max(1, min((batch + 31) // 32, self.SPLIT_TARGET_CTAS // tiles))
      dense_tf32(prev, ld_prev, g, g.stride(0), fin + 1, fout, batch, 4, aux=layer.bias.grad, dt=layer.weight.grad, ldt=fin,
                 split_k=split)
      self.launches += 1
      if l > 0:
        gp = w['g'][l - 1]
        dense_tf32(g, g.stride(0), self.w_t[l], self.w_t[l].shape[1], batch, fin, fout, 2, relu_bits=w['bits'][l - 1],
                   d=gp, ldd=gp.stride(0))
        self.launches += 1


# ---------------------------------------------------------------------------------------------
# replay
# ---------------------------------------------------------------------------------------------
class DeviceReplay:
  """Time-major ring of whole N-balloon steps kept in HBM; n-step transitions are assembled at sampling
  time by k_replay_sample.  capacity_steps * N transitions (the reference keeps 2,000,000).

  Known divergence (documented, not hidden): the ring stores the observation each action was CHOSEN on, so the
  last observation of an episode cut by the step limit (StepLimitWrapper, acme_utils.py:72-73) is not kept, and
  k_replay_sample rejects the n-step windows that cross such a truncation (`valid` = 0).  Acme's n-step adder
  still emits those windows, bootstrapping from the final observation with discount gamma^k: here the last
  n_step (5) of every 960 transitions of a truncated episode are not trained on (0.5 % of the data).  Windows that
  end in a TERMINAL step are kept (discount 0, no bootstrap observation needed)."""

  def __init__(self, num_envs: int, capacity_steps: int, *, num_features: int = NUM_FEATURES, n_step: int = 5,
               gamma: float = 0.993, device='cuda:0', seed: int = 0):
    self.device = torch.device(device)
    if self.device.type != 'cuda':
      raise _lib.BleError('DeviceReplay lives in GPU memory (no CPU fallback exists)')
    self.num_envs, self.capacity, self.num_features = int(num_envs), int(capacity_steps), int(num_features)
    self.n_step, self.gamma = int(n_step), float(gamma)
    self.obs = torch.zeros(self.capacity, self.num_envs, self.num_features, dtype=torch.float32, device=self.device)
    self.action = torch.zeros(self.capacity, self.num_envs, dtype=torch.int32, device=self.device)
    self.reward = torch.zeros(self.capacity, self.num_envs, dtype=torch.float32, device=self.device)
    self.terminal = torch.zeros(self.capacity, self.num_envs, dtype=torch.uint8, device=self.device)
    self.truncated = torch.zeros(self.capacity, self.num_envs, dtype=torch.uint8, device=self.device)
    self.count = 0
    self._seed = int(seed)
    self._draws = 0

  @property
  def num_transitions(self) -> int:
    return min(self.count, self.capacity) * self.num_envs

  def add(self, obs, action, reward, terminal, truncated) -> None:
    """One lockstep environment step: obs [N, F] the actions were chosen on, and what the step returned."""
    slot = self.count % self.capacity
    self.obs[slot].copy_(obs)
    self.action[slot].copy_(action)
    self.reward[slot].copy_(reward)
    self.terminal[slot].copy_(terminal)
    self.truncated[slot].copy_(truncated)
    self.count += 1

  def view(self, out_pitch: int = 0) -> _lib.BleReplayView:
    return _lib.BleReplayView(self.obs.data_ptr(), self.action.data_ptr(), self.reward.data_ptr(),
                              self.terminal.data_ptr(), self.truncated.data_ptr(), self.capacity, self.num_envs,
                              self.count, self.n_step, self.num_features, self.gamma, int(out_pitch))

  def sample(self, batch_size: int, indices: Optional[torch.Tensor] = None,
             out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """batch_size n-step transitions; indices (int64 [B, 2] = absolute step, balloon) forces the picks.  `out`
    (QrDqnLearner.sample_buffers): write the batch straight into the learner's own (row-padded) input buffers."""
    b, dev = int(batch_size), self.device
    if out is not None:
      pitch = out['state'].stride(0)
      assert out['state'].shape == (b, self.num_features) and out['next_state'].stride(0) == pitch
      out = dict(out)
    else:
      pitch = 0
      out = {'state': torch.empty(b, self.num_features, dtype=torch.float32, device=dev),
             'next_state': torch.empty(b, self.num_features, dtype=torch.float32, device=dev),
             'action': torch.empty(b, dtype=torch.int32, device=dev),
             'return': torch.empty(b, dtype=torch.float32, device=dev),
             'discount': torch.empty(b, dtype=torch.float32, device=dev),
             'valid': torch.empty(b, dtype=torch.uint8, device=dev)}
    out['indices'] = torch.empty(b, 2, dtype=torch.int64, device=dev)
    if indices is not None:
      indices = indices.to(dev, torch.int64).contiguous()
    view = self.view(pitch)
    self._draws += 1
    seed = (self._seed * 0x9E3779B97F4A7C15 + self._draws) & 0xFFFFFFFFFFFFFFFF
    _call('ble_replay_sample', dev, ctypes.byref(view), _ptr(indices), seed, b, _ptr(out['state']), _ptr(out['next_state']),
          _ptr(out['action']), _ptr(out['return']), _ptr(out['discount']), _ptr(out['valid']), _ptr(out['indices']))
    return out


# ---------------------------------------------------------------------------------------------
# exploration
# ---------------------------------------------------------------------------------------------
class MarcoPoloExploration:
  """agents/marco_polo_exploration.py:36-93 for N balloons, wrapped around RandomWalkAgent the way
  acme_utils.CombinedActor does (acme_utils.py:161-183)."""

  def __init__(self, num_envs: int, *, exploratory_episode_probability: float = 0.8, seed: int = 0, device='cuda:0'):
    self.device = torch.device(device)
    if self.device.type != 'cuda':
      raise _lib.BleError('MarcoPoloExploration runs on the GPU (no CPU fallback exists)')
    self.num_envs = int(num_envs)
    self.probability = float(exploratory_episode_probability)
    self.state = torch.zeros(4, self.num_envs, dtype=torch.int32, device=self.device)
    self.walk_target = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
    g = torch.Generator(device='cpu'); g.manual_seed(int(seed))
    self.seeds = torch.randint(0, 2**62, (self.num_envs,), dtype=torch.int64, generator=g).to(self.device)
    self._k = 0

  def step(self, obs: torch.Tensor, rl_actions: torch.Tensor, begin: Optional[torch.Tensor] = None) -> torch.Tensor:
    """obs [N, 1099], the learner's actions int32 [N], begin uint8 [N] (1 = first observation of an episode)."""
    actions = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
    if begin is not None:
      begin = begin.to(self.device, torch.uint8).contiguous()
    _call('ble_marco_polo_step', self.device, _ptr(obs.contiguous()), _ptr(rl_actions.to(torch.int32).contiguous()),
          _ptr(begin), self.num_envs, _ptr(self.state), _ptr(self.walk_target), _ptr(self.seeds), self._k, self.probability,
          _ptr(actions))
    self._k += 1
    return actions

  @property
  def exploratory_phase(self) -> torch.Tensor:
    return self.state[1].bool()


# ---------------------------------------------------------------------------------------------
# learner
# ---------------------------------------------------------------------------------------------
def allreduce_sum_(flat: torch.Tensor) -> int:
  """Sums a flat gradient buffer over the data-parallel group in place; returns the group size."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return dist.get_world_size()
  return 1


class QrDqnLearner:
  """One data-parallel replica of the QR-DQN learner (acme dqn learner with the QrDqn loss)."""

  def __init__(self, config: QrDqnConfig = QrDqnConfig(), *, device='cuda:0', seed: int = 0):
    self.config = config
    self.device = torch.device(device)
    if self.device.type != 'cuda':
      raise _lib.BleError('QrDqnLearner needs a CUDA device (no CPU fallback exists)')
    torch.manual_seed(int(seed))                       # same seed on every rank -> identical initial replicas
    self.online = QuantileNetwork(config).to(self.device)
    self.target = QuantileNetwork(config).to(self.device)
    self.flat = flatten_parameters(self.online, self.device)
    self.flat_target = flatten_parameters(self.target, self.device)
    self.flat_target.copy_(self.flat)
    for p in self.target.parameters():
      p.requires_grad_(False)
    self.m = torch.zeros_like(self.flat)
    self.v = torch.zeros_like(self.flat)
    self.steps = 0
    self.kernel_launches = 0
    if config.dense_backend not in ('tcgen05', 'cublas'):
      raise ValueError(f'unknown dense_backend {config.dense_backend!r}')
    self.hand_written_dense = config.dense_backend == 'tcgen05' and config.tf32_matmul
    if self.hand_written_dense:
      self.dense = DenseStack(self.online, self.device)
      self.dense_target = DenseStack(self.target, self.device)
      self._io = {}

  @contextlib.contextmanager
  def _matmul_precision(self):
    """config.tf32_matmul for THIS learner's GEMMs only: the process-wide torch flag is restored on exit."""
    previous = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = bool(self.config.tf32_matmul)
    try:
      yield
    finally:
      torch.backends.cuda.matmul.allow_tf32 = previous

  @torch.no_grad()
  def act(self, obs: torch.Tensor, epsilon: Optional[float] = None, generator: Optional[torch.Generator] = None):
    """Behaviour policy (acme_utils.py:250-258): epsilon-greedy on the mean over atoms."""
    if self.hand_written_dense:
      logits = self.dense.forward(obs.to(torch.float32))
      actions = greedy_actions(logits.view(-1, self.config.num_actions, self.config.num_atoms))
    else:
      with self._matmul_precision():
        actions = greedy_actions(self.online(obs))
    self.kernel_launches += 1
    eps = self.config.epsilon if epsilon is None else epsilon
    if eps > 0.0:
      n = actions.numel()
      explore = torch.rand(n, device=self.device, generator=generator) < eps
      random_actions = torch.randint(0, self.config.num_actions, (n,), dtype=torch.int32, device=self.device, generator=generator)
      actions = torch.where(explore, random_actions, actions)
    return actions

  def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    """One SGD step on a sampled batch; returns the mean loss (device scalar)."""
    cfg = self.config
    if self.hand_written_dense:
      return self._step_hand_written(batch)
    with self._matmul_precision():
      with torch.no_grad():
        target = target_distribution(self.target(batch['next_state']), batch['return'], batch['discount'])
      logits = self.online(batch['state'])
      weight = batch['valid'].to(torch.float32) if 'valid' in batch else None
      self.online.flat_grad.zero_()
      mean_loss, _ = quantile_huber_loss(logits, batch['action'], target, weight, cfg.huber_param)
      mean_loss.backward()
    world = allreduce_sum_(self.online.flat_grad)
    self.steps += 1
    adam_step(self.flat, self.online.flat_grad, self.m, self.v, self.steps, cfg.learning_rate, cfg.adam_b1, cfg.adam_b2,
              cfg.adam_eps, 1.0 / world)
    self.kernel_launches += 3
    if self.steps % cfg.target_update_period == 0:     # optax.periodically_update in acme's dqn learner
      self.flat_target.copy_(self.flat)
    return mean_loss.detach()

  def _sgd_pass(self, io: Dict[str, torch.Tensor]) -> None:
    """target forward -> target distribution -> online forward -> loss + dL/dlogits -> backward, on the buffers of `io`
    (every launch goes to the current stream: this is the body that is captured into a CUDA graph)."""
    cfg = self.config
    b = io['state'].shape[0]
    shape = (b, cfg.num_actions, cfg.num_atoms)
    target = target_distribution(self.dense_target.forward(io['next_state']).view(shape), io['return'], io['discount'])
    logits = self.dense.forward(io['state'], keep=True).view(shape)
    _call('ble_qr_loss', self.device, _ptr(logits), _ptr(io['action']), _ptr(target), _ptr(io['weight']),
          float(cfg.huber_param), b, cfg.num_actions, cfg.num_atoms, 1.0 / b, _ptr(io['loss']), _ptr(io['grad']))
    io['mean'].copy_((io['loss'] * io['weight']).mean())
    self.online.flat_grad.zero_()
    self.dense.backward(io['grad'].view(b, -1))

  def _step_io(self, b: int) -> Dict[str, torch.Tensor]:
    """Static buffers (and, with config.cuda_graph, the captured graphs) of the hand-written step for batch size b."""
    if b in self._io:
      return self._io[b]
    cfg, dev = self.config, self.device
    f32 = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
    # row-padded buffers the dense kernel reads in place (the online one with the column of ones the backward pass needs)
    io = {'state': self.dense.input_buffer(b), 'next_state': self.dense_target.input_buffer(b),
          'action': torch.zeros(b, dtype=torch.int32, device=dev), 'return': f32(b), 'discount': f32(b), 'weight': f32(b),
          'loss': f32(b), 'grad': f32(b, cfg.num_actions, cfg.num_atoms), 'mean': f32(()),
          'valid': torch.zeros(b, dtype=torch.uint8, device=dev)}
    if cfg.cuda_graph:
      snapshot = (self.online.flat_grad.clone(),)
      side = torch.cuda.Stream(device=dev)
      side.wait_stream(torch.cuda.current_stream(dev))
      with torch.cuda.stream(side):                       # warm-up off the capture: allocates every work buffer
        self._sgd_pass(io)
        self.dense.refresh()
      torch.cuda.current_stream(dev).wait_stream(side)
      io['pass_graph'], io['refresh_graph'] = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
      with torch.cuda.graph(io['pass_graph']):
        self._sgd_pass(io)
      with torch.cuda.graph(io['refresh_graph']):
        self.dense.refresh()
      self.online.flat_grad.copy_(snapshot[0])
    self._io[b] = io
    return io

  def sample_buffers(self, batch_size: int) -> Optional[Dict[str, torch.Tensor]]:
    """The learner's own input buffers for DeviceReplay.sample(out=...): the batch is then sampled where the dense kernel
    reads it (row pitch padded to 16 bytes) and step() copies nothing.  None on the library-GEMM path."""
    if not self.hand_written_dense:
      return None
    io = self._step_io(int(batch_size))
    out = {k: io[k] for k in ('state', 'next_state', 'action', 'return', 'discount')}
    out['valid'] = io['valid']
    return out

  @torch.no_grad()
  def _step_hand_written(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    """The same SGD step with every dense product on ble_dense_tf32 (no autograd graph)."""
    cfg = self.config
    b = batch['state'].shape[0]
    io = self._step_io(b)
    for k in ('state', 'next_state', 'action', 'return', 'discount'):
      if batch[k].data_ptr() != io[k].data_ptr():       # a batch sampled into sample_buffers() is already in place
        io[k].copy_(batch[k])
    if 'valid' in batch:
      io['weight'].copy_(batch['valid'])
    else:
      io['weight'].fill_(1.0)
    if cfg.cuda_graph:
      io['pass_graph'].replay()
    else:
      self._sgd_pass(io)
    world = allreduce_sum_(self.online.flat_grad)
    self.steps += 1
    adam_step(self.flat, self.online.flat_grad, self.m, self.v, self.steps, cfg.learning_rate, cfg.adam_b1, cfg.adam_b2,
              cfg.adam_eps, 1.0 / world)
    if cfg.cuda_graph:
      io['refresh_graph'].replay()
    else:
      self.dense.refresh()
    self.kernel_launches += 3
    if self.steps % cfg.target_update_period == 0:
      self.flat_target.copy_(self.flat)
      self.dense_target.refresh()
    return io['mean'].clone()

  def parameters_changed(self) -> None:
    """Call after writing the parameters from outside (load_flax_params, a manual copy into .flat): re-derives the
    operand copies the hand-written dense path keeps."""
    if self.hand_written_dense:
      self.dense.refresh(); self.dense_target.refresh()

  def state_dict(self) -> Dict[str, torch.Tensor]:
    return {'params': self.flat.clone(), 'target': self.flat_target.clone(), 'm': self.m.clone(), 'v': self.v.clone(),
            'steps': torch.tensor(self.steps)}

  def load_state_dict(self, state: Dict[str, torch.Tensor]) -> None:
    self.flat.copy_(state['params']); self.flat_target.copy_(state['target'])
    self.m.copy_(state['m']); self.v.copy_(state['v'])
    self.steps = int(state['steps'])
    if self.hand_written_dense:
      self.dense.refresh(); self.dense_target.refresh()


class TrainingLoop:
  """train_acme_qrdqn.py:72-81 for a BatchedBalloonEnv(observation='perciatelli'): every iteration steps all N
  balloons once, appends the step to the replay ring, restarts the balloons whose episode ended (terminal
  status or max_episode_length) and runs the learner.  learner_steps_per_iteration defaults to the
  reference's ratio (samples_per_insert = 8): N * 8 / batch_size SGD steps per N inserted transitions."""

  def __init__(self, env, learner: QrDqnLearner, *, replay: Optional[DeviceReplay] = None,
               exploration: Optional[MarcoPoloExploration] = None, learner_steps_per_iteration: Optional[int] = None,
               seed: int = 0):
    cfg = learner.config
    self.env, self.learner, self.exploration = env, learner, exploration
    n, dev = env.num_envs, env.device
    self.replay = replay if replay is not None else DeviceReplay(
        n, max(cfg.n_step + 2, cfg.max_replay_size // n), num_features=cfg.num_features, n_step=cfg.n_step,
        gamma=cfg.discount, device=dev, seed=seed)
    if learner_steps_per_iteration is None:
      learner_steps_per_iteration = max(1, round(n * cfg.samples_per_insert / cfg.batch_size))
    self.learner_steps_per_iteration = int(learner_steps_per_iteration)
    self.obs = env.reset(seed=seed)
    self.begin = torch.ones(n, dtype=torch.uint8, device=dev)
    self.episode_steps = torch.zeros(n, dtype=torch.int32, device=dev)
    self.last_loss = torch.zeros((), device=dev)
    self.iterations = 0
    self.phase_ms: Dict[str, float] = {}           # filled by run(profile=True)

  def run(self, num_iterations: int, log=None, profile: bool = False) -> Dict[str, float]:
    """profile=True brackets the phases of every iteration with CUDA events (sums land in self.phase_ms)."""
    env, learner, replay, cfg = self.env, self.learner, self.replay, self.learner.config
    n, dev = env.num_envs, env.device
    marks = []

    def mark(name):
      if profile:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, ev))
    reward_sum = torch.zeros((), dtype=torch.float64, device=dev)
    episodes = torch.zeros((), dtype=torch.int64, device=dev)
    steps_before = learner.steps
    for _ in range(num_iterations):
      mark('start')
      actions = learner.act(self.obs)
      if self.exploration is not None:
        actions = self.exploration.step(self.obs, actions, self.begin)
      acted_on = self.obs.clone()                  # env.step overwrites its observation buffer
      mark('act')
      obs, reward, done, _ = env.step(actions)
      mark('env_step_and_observation')
      self.episode_steps += 1
      terminal = done.ne(0)
      truncated = (self.episode_steps >= cfg.max_episode_length) & ~terminal     # StepLimitWrapper, acme_utils.py:72-73
      replay.add(acted_on, actions, reward, terminal.to(torch.uint8), truncated.to(torch.uint8))
      reward_sum += reward.sum(dtype=torch.float64)
      ended = terminal | truncated
      episodes += ended.sum()
      if bool(ended.any()):
        obs = env.reset_where(ended)
        self.episode_steps.masked_fill_(ended, 0)
      self.obs, self.begin = obs, ended.to(torch.uint8)
      mark('replay_add_and_resets')
      if replay.num_transitions >= cfg.min_replay_size and replay.count > cfg.n_step:
        for _ in range(self.learner_steps_per_iteration):
          batch = replay.sample(cfg.batch_size, out=learner.sample_buffers(cfg.batch_size))
          mark('replay_sample')
          self.last_loss = learner.step(batch)
          mark('learner_step')
      self.iterations += 1
      if log is not None:
        log(self.iterations, self.last_loss)
    if profile:
      torch.cuda.synchronize(dev)
      for (_, prev), (name, ev) in zip(marks[:-1], marks[1:]):
        if name != 'start':
          self.phase_ms[name] = self.phase_ms.get(name, 0.0) + prev.elapsed_time(ev)
    return {'env_steps': num_iterations * n, 'learner_steps': learner.steps - steps_before, 'episodes': int(episodes),
            'mean_reward': float(reward_sum) / max(1, num_iterations * n), 'last_loss': float(self.last_loss)}


def run_training(env, learner: QrDqnLearner, *, num_iterations: int, replay: Optional[DeviceReplay] = None,
                 exploration: Optional[MarcoPoloExploration] = None, learner_steps_per_iteration: Optional[int] = None,
                 seed: int = 0, log=None) -> Dict[str, float]:
  """One-shot form of TrainingLoop: reset, then num_iterations iterations."""
  loop = TrainingLoop(env, learner, replay=replay, exploration=exploration,
                      learner_steps_per_iteration=learner_steps_per_iteration, seed=seed)
  return loop.run(num_iterations, log=log)


class QuantileAgent:
  """agents/quantile_agent.py:37-160 behind the batched Agent interface (agents.py): begin_episode / step /
  end_episode take and return one entry per balloon.  In TRAIN mode every step appends the previous transition to
  the replay ring, runs `learner_steps_per_step` SGD steps once `min_replay_size` transitions are stored and lets
  MarcoPoloExploration override the greedy action (quantile.gin); in EVAL mode it is the greedy policy
  (epsilon_eval = 0).  Balloons whose episode already ended keep being stepped by the loops that drive this
  interface (their transitions carry reward 0 and are cut by the terminal flag recorded when they ended)."""

  def __init__(self, num_actions: int, observation_shape, arena, *, config: Optional[QrDqnConfig] = None,
               seed: Optional[int] = None, replay_steps: Optional[int] = None, learner_steps_per_step: int = 1):
    cfg = config if config is not None else QrDqnConfig()
    if num_actions != cfg.num_actions or tuple(observation_shape) != (cfg.num_features,):
      raise ValueError('QuantileAgent: the network configuration does not match the environment')
    self._arena = arena
    self.device, n = arena.device, arena.num_envs
    seed = 0 if seed is None else int(seed)
    self.learner = QrDqnLearner(cfg, device=self.device, seed=seed)
    self.exploration = MarcoPoloExploration(n, exploratory_episode_probability=cfg.exploratory_episode_probability,
                                            seed=seed + 1, device=self.device)
    steps = replay_steps if replay_steps is not None else max(cfg.n_step + 2, cfg.max_replay_size // n)
    self.replay = DeviceReplay(n, steps, num_features=cfg.num_features, n_step=cfg.n_step, gamma=cfg.discount,
                               device=self.device, seed=seed + 2)
    self.learner_steps_per_step = int(learner_steps_per_step)
    self.eval_mode = False
    self._last_obs = torch.empty(n, cfg.num_features, dtype=torch.float32, device=self.device)
    self._last_action = torch.zeros(n, dtype=torch.int32, device=self.device)
    self._alive = torch.ones(n, dtype=torch.bool, device=self.device)
    self.last_loss = torch.zeros((), device=self.device)

  def get_name(self) -> str:
    return self.__class__.__name__

  def set_mode(self, mode) -> None:                       # quantile_agent.py:152-157
    self.eval_mode = str(getattr(mode, 'value', mode)) == 'eval'

  def _act(self, observation: torch.Tensor, begin: bool) -> torch.Tensor:
    action = self.learner.act(observation, epsilon=0.0)
    if not self.eval_mode:
      flag = torch.full((self._arena.num_envs,), 1 if begin else 0, dtype=torch.uint8, device=self.device)
      action = self.exploration.step(observation, action, flag)
      self._last_obs.copy_(observation)
      self._last_action.copy_(action)
    return action

  def begin_episode(self, observation: torch.Tensor) -> torch.Tensor:
    self._alive.fill_(True)
    return self._act(observation, begin=True)

  def _store(self, reward: torch.Tensor, terminal: torch.Tensor, truncated: torch.Tensor) -> None:
    # a balloon that already ended contributes a terminal, zero-reward transition that no window can cross
    dead = ~self._alive
    self.replay.add(self._last_obs, self._last_action, torch.where(dead, torch.zeros_like(reward), reward),
                    (terminal | dead).to(torch.uint8), (truncated & ~dead).to(torch.uint8))
    self._alive &= ~terminal

  def _train(self) -> None:
    cfg = self.learner.config
    if self.replay.num_transitions >= cfg.min_replay_size and self.replay.count > cfg.n_step:
      for _ in range(self.learner_steps_per_step):
        self.last_loss = self.learner.step(self.replay.sample(cfg.batch_size, out=self.learner.sample_buffers(cfg.batch_size)))

  def step(self, reward: torch.Tensor, observation: torch.Tensor, done: Optional[torch.Tensor] = None) -> torch.Tensor:
    """done (uint8 / bool [N], optional): balloons whose episode ended with this reward (terminal status)."""
    if not self.eval_mode:
      terminal = done.ne(0) if done is not None else torch.zeros_like(self._alive)
      self._store(reward, terminal, torch.zeros_like(terminal))
      self._train()
    return self._act(observation, begin=False)

  def end_episode(self, reward: torch.Tensor, terminal: Optional[torch.Tensor] = None) -> None:
    """terminal False (or None) marks a step-limit truncation, as train_lib does at max_episode_length."""
    if self.eval_mode:
      return
    term = terminal.ne(0) if terminal is not None else torch.zeros_like(self._alive)
    self._store(reward, term, ~term)
    self._train()

  def save_checkpoint(self, checkpoint_dir: str, iteration_number: int) -> None:      # quantile_agent.py:159-170
    import os
    os.makedirs(checkpoint_dir, exist_ok=True)
    torch.save(self.learner.state_dict(), os.path.join(checkpoint_dir, f'ckpt.{int(iteration_number)}'))

  def load_checkpoint(self, checkpoint_dir: str, iteration_number: int) -> None:      # quantile_agent.py:172-176
    import os
    self.learner.load_state_dict(torch.load(os.path.join(checkpoint_dir, f'ckpt.{int(iteration_number)}'),
                                            map_location=self.device))

  def reload_latest_checkpoint(self, checkpoint_dir: str) -> int:                      # quantile_agent.py:178-190
    import os
    try:
      found = [int(f.split('.', 1)[1]) for f in os.listdir(checkpoint_dir) if f.startswith('ckpt.')]
    except (FileNotFoundError, ValueError):
      return -1
    if not found:
      return -1
    self.load_checkpoint(checkpoint_dir, max(found))
    return max(found)
