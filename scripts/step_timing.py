"""Times BalloonEnv.step kernels over batch sizes / CTA shapes (A/B data for DESIGN.md; not the bench).

    python scripts/step_timing.py [--sizes 8192,16384,65536] [--variants fused0,fused4,fused8,fused14,fusedauto,thread,ws]
Prints one JSON line per (size, variant): ms per ble_step and ms per step inside ble_rollout (32 steps / launch).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import batched_env  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--sizes', default='8192,16384,32768,65536')
  ap.add_argument('--variants', default='fused0,fused4,fused8,fused14,thread,ws')
  ap.add_argument('--fields', type=int, default=2048, help='size of the shared field pool (per-balloon fields: 0)')
  ap.add_argument('--steps', type=int, default=200)
  args = ap.parse_args()
  dev = torch.device('cuda:0')
  for n in [int(v) for v in args.sizes.split(',')]:
    nf = n if args.fields == 0 else min(n, args.fields)
    arena = batched_env.BatchedBalloonArena(n, precision='fp32', enable_noise=True, field_layout='x128' if nf * 3686400 < 60e9 else 'x64')
    g = torch.Generator(device=dev); g.manual_seed(1)
    arena.alloc_wind_fields(nf)
    for s in range(0, nf, 1024):
      e = min(nf, s + 1024)
      arena.write_wind_fields(torch.randn(e - s, 21, 21, 10, 9, 2, generator=g, device=dev) * 5.0, s)
    arena.set_field_map(torch.arange(n, dtype=torch.int32, device=dev) % nf)
    seeds = torch.arange(n, dtype=torch.int64) * 7 + 1
    actions = torch.randint(0, 3, (64, n), dtype=torch.int32, device=dev, generator=g)
    for variant in args.variants.split(','):
      if variant.startswith('fused'):
        os.environ['BLE_STEP_KERNEL'] = 'fused'; os.environ['BLE_STEP_WARPS'] = variant[5:] if variant[5:].isdigit() else ''
      else:
        os.environ['BLE_STEP_KERNEL'] = variant; os.environ.pop('BLE_STEP_WARPS', None)
      arena.reset(seeds)
      for t in range(5):
        arena.step(actions[t])
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for t in range(args.steps):
        arena.step(actions[t % 64])
      e1.record(); torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / args.steps
      row = {'n': n, 'variant': variant, 'ms_per_step': ms, 'M_env_steps_per_s': n / ms / 1e3}
      if variant.startswith('fused'):
        arena.rollout(actions[:32]); torch.cuda.synchronize()
        reps = max(1, args.steps // 32)
        e0.record()
        for r in range(reps):
          arena.rollout(actions[:32] if r % 2 == 0 else actions[32:])
        e1.record(); torch.cuda.synchronize()
        row['rollout_ms_per_step'] = e0.elapsed_time(e1) / (reps * 32)
      row['live'] = float((arena.get_state_dict()['status'] == 0).float().mean())
      print(json.dumps(row), flush=True)
    arena.close()
    torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
