"""CPU check of the HOST logic of the autograd-free dense path (learner.DenseStack): which products are launched, with
which operands, pitches, epilogue modes, masks and the column of ones -- the CUDA entry points are replaced by a plain
torch restatement of their documented contract (include/ble_b200.h: ble_dense_tf32, ble_transpose_f32), so a wrong pitch,
a swapped operand or a missing mask shows up here without a GPU.  The kernels themselves are tested on the GPU
(tests/test_gpu_dense.py)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from balloon_learning_environment_b200 import learner as lrn


def view(t, rows, cols, pitch):
  return torch.as_strided(t, (rows, cols), (pitch, 1))


class Recorder:
  """torch restatement of the C-ABI contract of the dense entry points (fp32 arithmetic, no TF32 rounding)."""

  def __init__(self):
    self.calls = []

  def dense(self, a, lda, b, ldb, m, n, k, mode, aux=None, ld_aux=0, d=None, ldd=0, dt=None, ldt=0, split_k=1, relu_bits=None):
    assert lda % 4 == 0 and ldb % 4 == 0 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0     # TMA operand rules
    assert (mode >= 3) == (split_k >= 1 and d is None and dt is not None) or mode < 3
    self.calls.append(mode)
    if mode == 4:
      A, B = view(a, k, m, lda).t(), view(b, k, n, ldb).t()            # operands given as [k, m] / [k, n]
    else:
      A, B = view(a, m, k, lda), view(b, n, k, ldb)
    D = A.double() @ B.double().t()
    if mode in (0, 1):
      D = D + aux[:n].double()
      if mode == 1:
        D = D.clamp_min(0)
        if relu_bits is not None:
          words = (n + 31) // 32
          bits = torch.zeros(m, words * 32, dtype=torch.int64)
          bits[:, :n] = (D > 0).long()
          packed = (bits.view(m, words, 32) << torch.arange(32)).sum(-1)
          packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed)
          relu_bits[:, :words] = packed.to(torch.int32)
    elif mode == 2:
      if relu_bits is not None:
        words = (n + 31) // 32
        w = relu_bits[:, :words].long() & 0xFFFFFFFF
        mask = ((w.unsqueeze(-1) >> torch.arange(32)) & 1).reshape(m, -1)[:, :n]
      else:
        mask = (view(aux, m, n, ld_aux) > 0).long()
      D = D * mask
    if mode >= 3:
      T = view(dt, n, m - 1 if aux is not None else m, ldt)
      if aux is not None:                                              # last row of A^T = the caller's ones -> bias gradient
        aux[:n] += D[m - 1].float()
        T += D[:m - 1].t().float()
      else:
        T += D.t().float()
      return
    if d is not None:
      view(d, m, n, ldd).copy_(D.float())
    if dt is not None:
      view(dt, n, m, ldt).copy_(D.t().float())

  def transpose(self, src, ld_src, rows, cols, dst, ld_dst):
    view(dst, cols, rows, ld_dst).copy_(view(src, rows, cols, ld_src).t())


@pytest.mark.parametrize('layers,hidden,features,batch', [(3, 40, 19, 50), (4, 96, 1099, 33)])
def test_dense_stack_orchestration_matches_autograd(monkeypatch, layers, hidden, features, batch):
  rec = Recorder()
  monkeypatch.setattr(lrn, 'dense_tf32', rec.dense)
  monkeypatch.setattr(lrn, 'transpose_f32', rec.transpose)
  cfg = lrn.QrDqnConfig(num_layers=layers, hidden_units=hidden, num_features=features)
  torch.manual_seed(1)
  net = lrn.QuantileNetwork(cfg)
  with torch.no_grad():
    for layer in net.layers:
      layer.bias.uniform_(-0.2, 0.2)
  lrn.flatten_parameters(net, 'cpu')
  stack = lrn.DenseStack(net, 'cpu')
  x = torch.randn(batch, features)
  gl = torch.randn(batch, cfg.num_actions * cfg.num_atoms) / batch

  # forward without keep (target network / acting): K-major products only, the last one without ReLU
  logits0 = stack.forward(x).clone()
  assert rec.calls == [1] * (layers - 1) + [0]
  ref = lrn.QuantileNetwork(cfg).double()
  ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
  want = ref(x.double()).view(batch, -1)
  np.testing.assert_allclose(logits0.numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-6)

  # forward with keep + backward: per layer one MN-major weight / bias gradient product and, above the first layer, one
  # masked input-gradient product
  rec.calls.clear()
  logits = stack.forward(x, keep=True).clone()
  np.testing.assert_allclose(logits.numpy(), logits0.numpy(), rtol=1e-6, atol=1e-7)
  net.flat_grad.zero_()
  stack.backward(gl)
  assert rec.calls == [1] * (layers - 1) + [0] + [4, 2] * (layers - 1) + [4]
  want.backward(gl.double())
  for got, r in zip(net.parameters(), ref.parameters()):
    np.testing.assert_allclose(got.grad.numpy(), r.grad.numpy(), rtol=2e-4, atol=1e-7)
  w = stack._work[(batch, True)]
  for l in range(layers - 1):
    assert w['h'][l].stride(0) % 32 == 0 and (w['h'][l][:, stack.dims[l][1]] == 1).all()       # line-aligned, ones column
  assert w['input'][0].stride(0) % 32 == 0
  # a network input allocated by the stack is used in place (no copy), any other tensor is copied into the stack's buffer
  own = stack.input_buffer(batch)
  own.copy_(x)
  stack.forward(own, keep=True)
  assert w['input'][0].data_ptr() == own.data_ptr()
  stack.forward(x, keep=True)
  assert w['input'][0].data_ptr() == w['x'].data_ptr()
  # refresh() after a parameter change re-derives the operand copies
  with torch.no_grad():
    net.layers[1].weight.mul_(2.0)
  stack.refresh()
  assert torch.equal(stack.w_t[1][:, :stack.dims[1][1]], net.layers[1].weight.t())
  assert torch.equal(stack.w_fwd[1][:, :stack.dims[1][0]], net.layers[1].weight)
