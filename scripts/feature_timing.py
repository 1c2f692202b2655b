"""Timing of the observation kernels with a FULL GP window (120 measurements).

    python scripts/feature_timing.py [--num-envs 65536] [--fill-steps 125]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from balloon_learning_environment_b200 import batched_env  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--num-envs', type=int, default=65536)
  ap.add_argument('--fill-steps', type=int, default=125)
  ap.add_argument('--fields', type=int, default=4096)
  args = ap.parse_args()
  dev = torch.device('cuda:0')
  n = args.num_envs
  arena = batched_env.BatchedBalloonArena(n, precision='fp32', enable_noise=True, enable_features=True)
  bench.upload_synthetic_fields(torch, arena, args.fields, dev, seed=3)
  arena.set_field_map(torch.arange(n, dtype=torch.int32, device=dev) % args.fields)
  arena.reset(torch.arange(n, dtype=torch.int64) + 17)
  g = torch.Generator(device=dev); g.manual_seed(0)
  obs = torch.empty(n, 1099, dtype=torch.float32, device=dev)
  for t in range(args.fill_steps):
    arena.step(torch.randint(0, 3, (n,), dtype=torch.int32, device=dev, generator=g))
  torch.cuda.synchronize()
  for _ in range(2):
    arena.features(obs)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 3
  e0.record()
  for _ in range(reps):
    arena.features(obs)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / reps
  # sliding window: every step drops the oldest measurement and appends a new one
  slide = []
  for t in range(6):
    arena.step(torch.randint(0, 3, (n,), dtype=torch.int32, device=dev, generator=g))
    e0.record(); arena.features(obs); e1.record(); torch.cuda.synchronize()
    slide.append(e0.elapsed_time(e1))
  live = float((arena.get_state_dict()['status'] == 0).float().mean())
  print(json.dumps({'num_envs': n, 'gp_window': min(args.fill_steps + 1, 120), 'features_ms': ms,
                    'features_per_s': n / ms * 1e3, 'live_fraction': live,
                    'features_ms_after_a_step': sorted(slide)[len(slide) // 2],
                    'obs_checksum': float(obs.double().sum())}), flush=True)
  arena.close()


if __name__ == '__main__':
  main()
