"""Small end-to-end exercise of the CUDA paths for compute-sanitizer (memcheck / racecheck / synccheck):
generation, reset, steps of every kernel shape, rollout, features, auto-reset, wind / atmosphere queries, and the
learner's tcgen05 dense path (forward, backward and one SGD step at ragged sizes)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import batched_env, learner, models


def main():
  n = 100                                              # ragged: not a multiple of 32
  dev = torch.device('cuda:0')
  for layout in (() if '--dense-only' in sys.argv else ('x64', 'x128')):
    a = batched_env.BatchedBalloonArena(n, precision='fp32', enable_noise=True, enable_features=True, field_layout=layout)
    a.set_decoder(models.load_decoder(''))
    a.alloc_wind_fields(n)
    seeds = torch.arange(n, dtype=torch.int64) + 5
    a.sample_wind_fields(seeds)
    a.set_field_map(torch.arange(n, dtype=torch.int32))
    a.reset(seeds)
    acts = torch.randint(0, 3, (8, n), dtype=torch.int32, device=dev)
    for shape in ('0', '4', '8', '14'):
      os.environ['BLE_STEP_WARPS'] = shape
      for t in range(3):
        a.step(acts[t])
        obs = a.features()
    os.environ.pop('BLE_STEP_WARPS')
    a.features_track(False)
    a.rollout(acts[:4])
    a.wind_query(torch.tensor([[0.0, 0.0, 9000.0, 100.0]] * 5, dtype=torch.float64), torch.arange(5, dtype=torch.int32), True)
    a.atmosphere_query('pressure', torch.tensor([9000.0, 5000.0], dtype=torch.float64), torch.tensor([0, 1], dtype=torch.int32))
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all()
    a.close()
  if '--dense-only' not in sys.argv:
    b = batched_env.BatchedBalloonArena(n, precision='fp32', wind_model='simple_static', enable_noise=True, auto_reset=True)
    b.reset(torch.arange(n, dtype=torch.int64))
    for t in range(3):
      b.step(torch.randint(0, 3, (n,), dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    b.close()
  # learner: every epilogue mode of ble_dense_tf32 at ragged sizes, DenseStack forward / backward, one eager SGD step
  cfg = learner.QrDqnConfig(num_layers=3, hidden_units=200, num_features=1099, cuda_graph=False)
  lrn = learner.QrDqnLearner(cfg, seed=0)
  bsz = 70
  batch = {'state': torch.rand(bsz, 1099, device=dev), 'next_state': torch.rand(bsz, 1099, device=dev),
           'action': torch.randint(0, 3, (bsz,), dtype=torch.int32, device=dev), 'return': torch.rand(bsz, device=dev),
           'discount': torch.full((bsz,), 0.96, device=dev), 'valid': torch.ones(bsz, dtype=torch.uint8, device=dev)}
  loss = lrn.step(batch)
  acts2 = lrn.act(torch.rand(n, 1099, device=dev))
  torch.cuda.synchronize()
  assert torch.isfinite(loss) and acts2.shape == (n,)
  print('sanitizer probe ok')


if __name__ == '__main__':
  main()
