"""Small end-to-end exercise of the CUDA paths for compute-sanitizer (memcheck / racecheck / synccheck):
generation, reset, steps of every kernel shape, rollout, features, auto-reset, wind / atmosphere queries."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import batched_env, models


def main():
  n = 100                                              # ragged: not a multiple of 32
  dev = torch.device('cuda:0')
  for layout in ('x64', 'x128'):
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
  b = batched_env.BatchedBalloonArena(n, precision='fp32', wind_model='simple_static', enable_noise=True, auto_reset=True)
  b.reset(torch.arange(n, dtype=torch.int64))
  for t in range(3):
    b.step(torch.randint(0, 3, (n,), dtype=torch.int32, device=dev))
  torch.cuda.synchronize()
  b.close()
  print('sanitizer probe ok')


if __name__ == '__main__':
  main()
