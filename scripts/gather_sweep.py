"""Sweep of the wind-gather kernel: achieved algorithmic GB/s vs number of resident fields / layout.

    python scripts/gather_sweep.py [--layouts x64 x128] [--fields 4096 16384 32768 65536]
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
  ap.add_argument('--layouts', nargs='+', default=['x64'])
  ap.add_argument('--fields', nargs='+', type=int, default=[4096, 16384, 32768, 65536])
  ap.add_argument('--lookups', type=int, default=1 << 24)
  args = ap.parse_args()
  dev = torch.device('cuda:0')
  m = args.lookups
  g = torch.Generator(device=dev); g.manual_seed(1)
  xyzt = torch.empty(m, 4, dtype=torch.float32, device=dev)
  xyzt[:, 0].uniform_(-500, 500, generator=g); xyzt[:, 1].uniform_(-500, 500, generator=g)
  xyzt[:, 2].uniform_(5000, 14000, generator=g); xyzt[:, 3].uniform_(0, 48, generator=g)
  perm = torch.randperm(m, device=dev, generator=g)
  for layout in args.layouts:
    for nf in args.fields:
      per_field = 1935360 if layout == 'x64' else 3686400
      if nf * per_field > 150e9:
        continue
      arena = batched_env.BatchedBalloonArena(8, precision='fp32', enable_noise=False, field_layout=layout)
      bench.upload_synthetic_fields(torch, arena, nf, dev, seed=5)
      for order in ('random', 'sorted_by_field'):
        fidx = (perm % nf).to(torch.int32)
        if order == 'sorted_by_field':
          fidx = torch.sort(fidx).values
        for _ in range(3):
          arena.wind_forecast(xyzt, fidx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
          arena.wind_forecast(xyzt, fidx)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({'layout': layout, 'fields': nf, 'footprint_gb': nf * per_field / 1e9, 'order': order,
                          'ms': ms, 'algorithmic_gbs': 156 * m / ms / 1e6}), flush=True)
      arena.close()
      torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
