"""GPU parity of the hand-written tcgen05 dense kernel (csrc/ble_dense.cu, through the C ABI) and of the autograd-free
DenseStack built on it.

Tolerances.  The tensor core reads fp32 operands as TF32 (10 explicit mantissa bits) and accumulates in fp32:
  * operands that are exactly representable in TF32 give products that are exact, so the kernel is held to fp32
    accumulation error (1e-5 relative to the row scale) against an fp64 reference;
  * general fp32 operands are held to the TF32 operand rounding, 2e-3 relative to |A| . |B|^T (truncation of both
    operands: 2 x 2^-10), and the full forward / backward pass to 1e-2 of each gradient's norm against fp64 autograd.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def lrn():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from balloon_learning_environment_b200 import _lib, learner
  _lib.load()
  return learner


def tf32_exact(t):
  """Rounds to values with 10 explicit mantissa bits (exactly representable in TF32)."""
  return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def alloc(rows, cols, pitch, gen, exact):
  buf = torch.zeros(rows, pitch, device='cuda')
  vals = torch.randn(rows, cols, device='cuda', generator=gen)
  buf[:, :cols] = tf32_exact(vals) if exact else vals
  buf[:, cols:] = float('nan')                      # the pitch padding must never be read
  return buf


@pytest.mark.parametrize('m,n,k', [(128, 128, 32), (256, 128, 64), (1000, 153, 600), (8192, 600, 600), (333, 600, 1099),
                                   (64, 8, 4), (130, 257, 36), (130, 260, 36), (7003, 700, 100)])   # the last: ragged 256-row tiles
@pytest.mark.parametrize('exact', [True, False])
def test_dense_forward_modes(lrn, m, n, k, exact):
  _dense_forward_modes(lrn, m, n, k, exact)


def test_dense_forward_modes_256_row_tiles():
  """The two-accumulator variant (BLE_DENSE_ROWS=256, read once per process) on the shapes that select it."""
  import os, subprocess, sys
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  code = ('import tests.test_gpu_dense as t; from balloon_learning_environment_b200 import learner as l;'
          '[t._dense_forward_modes(l, m, n, k, e) for (m, n, k) in ((8192, 600, 600), (7003, 700, 100)) for e in (True, False)]; print("ok256")')
  env = dict(os.environ, BLE_DENSE_ROWS='256')
  out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  assert 'ok256' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def _dense_forward_modes(lrn, m, n, k, exact):
  gen = torch.Generator(device='cuda'); gen.manual_seed(m * 7 + n * 3 + k)
  lda, ldb = lrn._pitch4(k) + 4, lrn._pitch4(k)
  a, b = alloc(m, k, lda, gen, exact), alloc(n, k, ldb, gen, exact)
  bias = torch.randn(n, device='cuda', generator=gen)
  ref = a[:, :k].double() @ b[:, :k].double().t()
  scale = (a[:, :k].double().abs() @ b[:, :k].double().abs().t()) + 1e-6
  tol = 1e-5 if exact else 2e-3
  mp = lrn._pitch4(m)
  for mode in (0, 1):
    d = torch.full((m, n + 3), -7.0, device='cuda'); dt = torch.full((n, mp), -7.0, device='cuda')
    lrn.dense_tf32(a, lda, b, ldb, m, n, k, mode, aux=bias, d=d, ldd=n + 3, dt=dt, ldt=mp)
    want = ref + bias.double()
    if mode == 1:
      want = want.clamp_min(0)
    err = ((d[:, :n].double() - want).abs() / scale).max().item()
    assert err < tol, (mode, err)
    assert torch.equal(d[:, :n].t().contiguous(), dt[:, :m].contiguous())        # the transposed copy is the same numbers
    assert (d[:, n:] == -7.0).all() and (dt[:, m:] == -7.0).all()                 # nothing written past the edges
  # mode 1 can leave the ReLU mask packed 32 columns per word; mode 2 reads it back instead of the activations
  words = (n + 31) // 32
  bits = torch.full((m, words + 1), 0x55555555, dtype=torch.int32, device='cuda')
  lrn.dense_tf32(a, lda, b, ldb, m, n, k, 1, aux=bias, d=d, ldd=n + 3, relu_bits=bits)
  cols = torch.arange(words * 32, device='cuda')
  got_bits = ((bits[:, :words].unsqueeze(-1) >> torch.arange(32, device='cuda', dtype=torch.int32)) & 1).reshape(m, -1)[:, :n]
  assert torch.equal(got_bits.bool(), d[:, :n] > 0) and (bits[:, words] == 0x55555555).all()
  d_bits = torch.empty(m, n, device='cuda')
  lrn.dense_tf32(a, lda, b, ldb, m, n, k, 2, d=d_bits, ldd=n, relu_bits=bits)
  d_act = torch.empty(m, n, device='cuda')
  lrn.dense_tf32(a, lda, b, ldb, m, n, k, 2, aux=d, ld_aux=n + 3, d=d_act, ldd=n)
  assert torch.equal(d_bits, d_act)
  # mode 2: ReLU mask of the backward pass; only the transposed output requested
  h = torch.randn(m, n, device='cuda', generator=gen)
  d = torch.empty(m, n, device='cuda')
  lrn.dense_tf32(a, lda, b, ldb, m, n, k, 2, aux=h, ld_aux=n, d=d, ldd=n)
  want = ref * (h > 0)
  assert ((d.double() - want).abs() / scale).max().item() < tol
  dt = torch.empty(n, mp, device='cuda')
  lrn.dense_tf32(a, lda, b, ldb, m, n, k, 2, aux=h, ld_aux=n, dt=dt, ldt=mp)
  assert torch.equal(dt[:, :m].t().contiguous(), d)


@pytest.mark.parametrize('m,n,k,split', [(600, 600, 8192, 11), (153, 600, 1000, 4), (600, 1099, 4096, 6), (40, 24, 100, 9),
                                         (128, 128, 64, 1)])
def test_dense_split_k_accumulates(lrn, m, n, k, split):
  gen = torch.Generator(device='cuda'); gen.manual_seed(k)
  ld = lrn._pitch4(k)
  a, b = alloc(m, k, ld, gen, True), alloc(n, k, ld, gen, True)
  start = torch.randn(n, m + 2, device='cuda', generator=gen)                   # the TRANSPOSED result is accumulated
  dt = start.clone()
  lrn.dense_tf32(a, ld, b, ld, m, n, k, 3, dt=dt, ldt=m + 2, split_k=split)
  want = start[:, :m].double() + (a[:, :k].double() @ b[:, :k].double().t()).t()
  scale = (a[:, :k].double().abs() @ b[:, :k].double().abs().t()).t() + 1.0
  assert ((dt[:, :m].double() - want).abs() / scale).max().item() < 1e-5
  assert torch.equal(dt[:, m:], start[:, m:])
  # with aux: the last row of A is the caller's row of ones and its results (column sums of B^T) go to aux
  a[m - 1, :k] = 1.0
  dt2 = start.clone(); bias = torch.full((n,), 0.5, device='cuda')
  lrn.dense_tf32(a, ld, b, ld, m, n, k, 3, aux=bias, dt=dt2, ldt=m + 2, split_k=split)
  assert ((dt2[:, :m - 1].double() - want[:, :m - 1]).abs() / scale[:, :m - 1]).max().item() < 1e-5
  assert torch.equal(dt2[:, m - 1:], start[:, m - 1:])                            # the last row went elsewhere
  col = b[:, :k].double().sum(1)
  assert ((bias.double() - 0.5 - col).abs() / (b[:, :k].double().abs().sum(1) + 1.0)).max().item() < 1e-5


@pytest.mark.parametrize('m,n,k,split', [(601, 600, 8192, 14), (154, 600, 1000, 4), (1100, 600, 4096, 6), (40, 24, 100, 3),
                                         (128, 160, 32, 1), (33, 153, 70, 2)])
def test_dense_mn_major_operands(lrn, m, n, k, split):
  """mode 4: D^T += A^T . B for ROW-MAJOR A [k, m], B [k, n] (MN-major tensor-core operands), the last column of A being
  the caller's column of ones whose products go to aux -- the weight + bias gradient straight from activations and
  output gradients."""
  gen = torch.Generator(device='cuda'); gen.manual_seed(k + m)
  lda, ldb = lrn._pitch4(m) + 4, lrn._pitch4(n)
  a, b = alloc(k, m, lda, gen, True), alloc(k, n, ldb, gen, True)
  a[:, m - 1] = 1.0
  start = torch.randn(n, m + 1, device='cuda', generator=gen)
  dt = start.clone(); bias = torch.full((n,), -2.0, device='cuda')
  lrn.dense_tf32(a, lda, b, ldb, m, n, k, 4, aux=bias, dt=dt, ldt=m + 1, split_k=split)
  full = b[:, :n].double().t() @ a[:, :m].double()                      # [n, m]
  scale = b[:, :n].double().abs().t() @ a[:, :m].double().abs() + 1.0
  err = ((dt[:, :m - 1].double() - start[:, :m - 1].double() - full[:, :m - 1]).abs() / scale[:, :m - 1]).max().item()
  assert err < 1e-5, err
  assert torch.equal(dt[:, m - 1:], start[:, m - 1:])
  assert ((bias.double() + 2.0 - full[:, m - 1]).abs() / scale[:, m - 1]).max().item() < 1e-5
  # without aux every row of A^T lands in dt
  dt2 = start.clone()
  lrn.dense_tf32(a, lda, b, ldb, m, n, k, 4, dt=dt2, ldt=m + 1, split_k=split)
  assert ((dt2[:, :m].double() - start[:, :m].double() - full).abs() / scale).max().item() < 1e-5


def test_transpose_and_row_sum(lrn):
  gen = torch.Generator(device='cuda'); gen.manual_seed(1)
  src = torch.randn(77, 1099 + 5, device='cuda', generator=gen)
  dst = torch.full((1099, 80), 3.0, device='cuda')
  lrn.transpose_f32(src, 1104, 77, 1099, dst, 80)
  assert torch.equal(dst[:, :77], src[:, :1099].t()) and (dst[:, 77:] == 3.0).all()
  out = torch.ones(77, device='cuda')
  lrn.row_sum_f32(src, 1104, 77, 1099, out, accumulate=True)
  np.testing.assert_allclose(out.cpu().numpy(), 1.0 + src[:, :1099].double().sum(1).cpu().numpy(), rtol=1e-5, atol=1e-4)
  lrn.row_sum_f32(src, 1104, 77, 1099, out)
  np.testing.assert_allclose(out.cpu().numpy(), src[:, :1099].double().sum(1).cpu().numpy(), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('layers,hidden,features,batch', [(3, 64, 19, 50), (8, 600, 1099, 1024), (8, 600, 1099, 333)])
def test_dense_stack_matches_autograd(lrn, layers, hidden, features, batch):
  """forward + backward of the autograd-free stack.
  (a) TIGHT, product by product: every activation, input gradient, weight gradient and bias gradient the stack produced
      against an fp64 evaluation of the same product on the stack's OWN operands truncated to TF32 (what the tensor core
      reads) -- what is left is fp32 accumulation order, 5e-6 of the result's norm.  (A chained emulation cannot be held
      this tight on the 8-layer network: a 1e-7 accumulation difference that flips one TF32 truncation or one ReLU mask
      downstream moves that element by 1e-3 / 100 %, measured 2.6 % on the first layer's gradient.)
  (b) LOOSE, end to end: logits and every parameter gradient against plain fp64 autograd on the same parameters; that
      difference is the TF32 operand precision itself (ReLU masks that flip included): 0.15 of each gradient's norm."""
  cfg = lrn.QrDqnConfig(num_layers=layers, hidden_units=hidden, num_features=features)
  torch.manual_seed(3)
  net = lrn.QuantileNetwork(cfg).cuda()
  with torch.no_grad():
    for layer in net.layers:
      layer.bias.uniform_(-0.1, 0.1)
  lrn.flatten_parameters(net, 'cuda')
  stack = lrn.DenseStack(net, 'cuda')
  gen = torch.Generator(device='cuda'); gen.manual_seed(5)
  x = torch.randn(batch, features, device='cuda', generator=gen)
  gl = torch.randn(batch, cfg.num_actions * cfg.num_atoms, device='cuda', generator=gen) / batch
  logits = stack.forward(x, keep=True).clone()
  net.flat_grad.zero_()
  stack.backward(gl)

  q = lambda t: (t.float().contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32).double()
  rel = lambda got, want: ((got.double() - want).norm() / (want.norm() + 1e-30)).item()
  w = stack._work[(batch, True)]
  ws = [l.weight.detach() for l in net.layers]
  fouts = [o for _, o in stack.dims]
  tight = 5e-6
  acts = [w['h'][l][:, :fouts[l]] for l in range(layers)]         # views without the column of ones
  inputs = [x] + acts[:-1]
  for l in range(layers):                                       # forward products, the ones columns, the packed masks
    z = q(inputs[l]) @ q(ws[l]).t() + net.layers[l].bias.detach().double()
    assert rel(acts[l], z.clamp_min(0) if l + 1 < layers else z) < tight, ('forward', l)
    if l + 1 < layers:
      assert (w['h'][l][:, fouts[l]] == 1).all()
      bits = ((w['bits'][l].unsqueeze(-1) >> torch.arange(32, device='cuda', dtype=torch.int32)) & 1).reshape(batch, -1)
      assert torch.equal(bits[:, :fouts[l]].bool(), acts[l] > 0)
  assert torch.equal(w['input'][0][:, :features], x) and (w['x'].storage_offset() == 0)
  assert torch.equal(w['g'][-1][:, :fouts[-1]], gl)
  for l in range(layers - 1, -1, -1):                           # backward products on the stack's own gradients
    g = w['g'][l][:, :fouts[l]]
    assert rel(net.layers[l].weight.grad, q(g.t()) @ q(inputs[l].t()).t()) < tight, ('weight gradient', l)
    assert rel(net.layers[l].bias.grad, q(g).sum(0)) < tight, ('bias gradient', l)     # TF32 g x exact ones
    if l > 0:
      want = (q(g) @ q(ws[l].t()).t()) * (inputs[l].double() > 0)
      assert rel(w['g'][l - 1][:, :fouts[l - 1]], want) < tight, ('input gradient', l)

  ref = lrn.QuantileNetwork(cfg).double().cuda()
  ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
  want = ref(x.double()).view(batch, -1)
  want.backward(gl.double())
  assert rel(logits, want.detach()) < 5e-3
  loose = max(rel(got.grad, r.grad) for got, r in zip(net.parameters(), ref.parameters()))
  print(f'DenseStack {layers} x {hidden}, batch {batch}: worst gradient vs fp64 autograd {loose:.3g}')
  assert loose < 0.15, loose                                   # measured 0.067 on the 8 x 600 network
  # a second backward ACCUMULATES (the learner zeroes the buffer once per step)
  before = net.flat_grad.clone()
  stack.backward(gl)
  assert ((net.flat_grad - 2 * before).norm() / before.norm()).item() < 1e-4


def test_learner_step_hand_written_vs_library(lrn):
  """One SGD step of the learner on the tcgen05 path lands where the library-GEMM path (cuBLAS TF32 + autograd) does."""
  outs = {}
  rng = np.random.default_rng(2)
  b = 512
  batch = {'state': rng.normal(0, 1, (b, 1099)).astype(np.float32), 'next_state': rng.normal(0, 1, (b, 1099)).astype(np.float32),
           'action': rng.integers(0, 3, b).astype(np.int32), 'return': rng.uniform(0, 2, b).astype(np.float32),
           'discount': np.full(b, 0.993 ** 5, np.float32), 'valid': (rng.uniform(size=b) < 0.9).astype(np.uint8)}
  for backend in ('tcgen05', 'tcgen05-eager', 'cublas'):
    cfg = lrn.QrDqnConfig(dense_backend=backend.split('-')[0], cuda_graph=backend == 'tcgen05', learning_rate=1e-4)
    learner = lrn.QrDqnLearner(cfg, seed=4)
    assert learner.hand_written_dense == backend.startswith('tcgen05')
    start = learner.flat.clone()
    losses = [float(learner.step({k: torch.as_tensor(v).cuda() for k, v in batch.items()})) for _ in range(3)]
    outs[backend] = (losses, (learner.flat - start).clone(), learner.online.flat_grad.clone())
  (l0, d0, g0), (l2, d2, g2) = outs['tcgen05'], outs['tcgen05-eager']
  # the captured graph replays the eager sequence; the split-K atomics leave the summation order of the weight gradient
  # free, and over three steps a last-bit difference can flip a TF32 truncation or a ReLU mask downstream
  np.testing.assert_allclose(l0, l2, rtol=1e-4)
  assert ((g0 - g2).norm() / g2.norm()).item() < 5e-3
  assert ((d0 - d2).norm() / d2.norm()).item() < 5e-3
  (l1, d1, g1) = outs['cublas']
  np.testing.assert_allclose(l0, l1, rtol=2e-3)
  assert ((g0 - g1).norm() / g1.norm()).item() < 2e-2
  assert ((d0 - d1).norm() / d1.norm()).item() < 5e-2 and d1.abs().max().item() > 0
